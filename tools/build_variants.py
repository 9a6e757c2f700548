#!/usr/bin/env python3
"""Tuning helper: build variants of ONE translation unit locally (nvcc cross-compiles without a GPU) and link each with the
already-built objects of the others into variants/libmmb200_<tag>.so, so that a GPU call can time them all without rebuilding
on the box (`cp variants/libmmb200_<tag>.so materialmodels.jl_b200/libmmb200.so` between runs).
usage: build_variants.py mm_ops.cu  tagA="-DMM_XU_PPT=1"  tagB="-DMM_XU_PPT=2 -DMM_XU_SOA_BS=128" ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "materialmodels.jl_b200"))
import build as B

unit = sys.argv[1]
B.build()                                                   # make sure every object exists and is current
out = os.path.join(ROOT, "variants"); os.makedirs(out, exist_ok=True)
procs = []
for spec in sys.argv[2:]:
    tag, flags = spec.split("=", 1)
    o = os.path.join(out, "%s_%s.o" % (unit.replace(".cu", ""), tag))
    cmd = [B.nvcc()] + B.NVCC_FLAGS + flags.split() + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, unit), "-o", o]
    procs.append((tag, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for tag, o, p in procs:
    log, _ = p.communicate()
    open(os.path.join(out, tag + ".ptxas.log"), "w").write(log)
    if p.returncode:
        sys.exit("nvcc failed for %s:\n%s" % (tag, log[-2000:]))
    objs = [o if s == unit else B._obj(s) for s in B.SOURCES]
    so = os.path.join(out, "libmmb200_%s.so" % tag)
    subprocess.check_call([B.nvcc(), "-shared", "-o", so] + objs + ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a"])
    print(so)
