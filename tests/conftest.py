import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build libmmb200.so (nvcc cross-compiles without a GPU) before
    the first test needs it.  Only when the library is MISSING — a present one is used as it is (file times do not survive every copy)."""
    import importlib.util
    import shutil

    if os.path.exists(os.path.join(ROOT, "materialmodels.jl_b200", "libmmb200.so")):
        return
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return
    spec = importlib.util.spec_from_file_location("_mm_build", os.path.join(ROOT, "materialmodels.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mm():
    import materialmodels_jl_b200 as mm

    return mm


@pytest.fixture(scope="session")
def cuda_lib(mm):
    """The native library on a GPU box; fails (does not skip) if it is missing or sees no device."""
    import materialmodels_jl_b200._lib as L

    lib = L.load()
    assert lib.mm_device_count() >= 1, "no CUDA device visible"
    return lib
