"""Build libmmb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmmb200.so")
SOURCES = ["mm_api.cu", "mm_ops.cu", "mm_crystal.cu", "mm_wrap.cu", "mm_cells.cu", "mm_host.cu"]
HEADERS = ["mm_math.cuh", "mm_crystal_math.cuh", "mm_wrap_math.cuh", "mm_kernels.cuh", os.path.join("..", "..", "include", "mmb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++",
]


def nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libmmb200.so")
    return p


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=()):
    extra = list(extra) + os.environ.get("MM_NVCC_EXTRA", "").split()      # e.g. -DMM_PLASTIC_MINBLOCKS=3 for tuning sweeps
    if not force and not stale():
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [nvcc(), "-shared", "-o", SO] + objs + ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    ex = ["-Xptxas", "-v"] if "-v" in sys.argv else []
    print(build(force=True, verbose=True, extra=ex))
