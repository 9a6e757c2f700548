"""Static checks on the compiled kernels (cuobjdump, no GPU): no hot kernel may index a thread-local array at run time.
That is what put the 5 x 5 UniaxialStress system into local memory (346 LDL/STL, 47.7 instead of 37.4 ms per 1e8 points)
until `small_solve` got a compile-time elimination column; the general pivoted crystal path is the one place where it is the design."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_runtime_indexed_local_memory_outside_the_general_crystal_path():
    import sass_scan

    objs = sorted(glob.glob(os.path.join(ROOT, "materialmodels.jl_b200", "build", "*.o")))
    if not objs:                                   # the library was shipped prebuilt (GPU box): rebuild objects is not this test's job
        pytest.skip("no object files next to the library")
    allowed = ("CrystalGeneralOp", "WrappedCrystalOp", "crystal_fixup_kernel")
    seen = 0
    for obj in objs:
        for name, r in sass_scan.scan(obj).items():
            seen += 1
            if any(a in name for a in allowed):
                continue
            assert r["local_dynamic"] == 0, "%s: %d run-time indexed local-memory accesses" % (name, r["local_dynamic"])
            assert r["local"] <= 120, "%s: %d local-memory instructions (spills)" % (name, r["local"])
    assert seen > 100
