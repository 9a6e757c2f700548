"""Build libmmb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmmb200.so")
SOURCES = ["mm_api.cu", "mm_ops.cu", "mm_compact.cu", "mm_crystal.cu", "mm_wrap.cu", "mm_cells.cu", "mm_host.cu", "mm_expand.cpp"]
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


import re

_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def deps_of(path, seen=None):
    """the source and every header it includes by a quoted path, transitively"""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path, encoding="utf-8") as f:
        for inc in _INC.findall(f.read()):
            deps_of(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _obj(s):
    return os.path.join(HERE, "build", os.path.splitext(s)[0] + ".o")


def _src_stale(s, extra_key):
    o = _obj(s)
    if not os.path.exists(o):
        return True
    try:
        with open(o + ".flags") as f:
            if f.read() != extra_key:
                return True
    except OSError:
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in list(deps_of(os.path.join(CSRC, s))) + [os.path.abspath(__file__)])


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = set()
    for s in SOURCES:
        deps |= deps_of(os.path.join(CSRC, s))
    return any(os.path.getmtime(d) > t for d in list(deps) + [os.path.abspath(__file__)])


def build(force=False, verbose=False, extra=()):
    """Recompiles only the translation units whose source, headers or flags changed (one nvcc process per unit, in parallel)."""
    extra = list(extra) + os.environ.get("MM_NVCC_EXTRA", "").split()      # e.g. -DMM_PLASTIC_MINBLOCKS=3 for tuning sweeps
    if not force and not stale() and not extra:
        return SO
    key = " ".join(NVCC_FLAGS + extra)
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = _obj(s)
        objs.append(o)
        if not force and not _src_stale(s, key):
            continue
        cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for s, o, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out)
        if p.returncode:
            failed = failed or s
            if os.path.exists(o):
                os.remove(o)
        else:
            with open(o + ".flags", "w") as f:
                f.write(key)
    if failed:
        raise RuntimeError("nvcc failed on %s" % failed)
    cmd = [nvcc(), "-shared", "-o", SO] + objs + ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    ex = ["-Xptxas", "-v"] if "-v" in sys.argv else []
    print(build(force=True, verbose=True, extra=ex))
