#!/usr/bin/env python3
"""H2D / D2H bandwidth of successive page-locked allocations: on the pool's VMs (one visible NUMA node, two physical sockets) the rate of a
pinned block depends on which socket its pages landed on.  usage: pinned_probe.py [blocks] [GiB per block]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import materialmodels_jl_b200 as mm

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
n = int(gib * 2**30 / 8)
dev = torch.empty(n, dtype=torch.float64, device="cuda")
keep = []
for b in range(nb):
    a = mm.PinnedArray((n,)).array
    a[...] = 1.0
    t = torch.from_numpy(a)
    res = {}
    for name, fn in (("h2d", lambda: dev.copy_(t, non_blocking=True)), ("d2h", lambda: t.copy_(dev, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[name] = round(8 * n / min(ts) / 1e9, 1)
    keep.append(a)
    print(json.dumps({"block": b, "GiB": gib, "is_pinned": bool(t.is_pinned()), **res}), flush=True)

# both directions at once (two streams, 1 GiB copies in flight for ~1 s): what the link gives a full-duplex pipeline
if os.environ.get("PROBE_BIDIR", "1") == "1":
    a, b = keep[0], keep[1]
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    dev2 = torch.empty_like(dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for chunk_mib in (4, 16, 64, 256, 1024):
        m = chunk_mib * 2**20 // 8
        reps = max(2, int(16 * 1024 / chunk_mib) // 8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            for r in range(reps):
                o = (r * m) % (n - m)
                dev[o:o + m].copy_(ta[o:o + m], non_blocking=True)
        with torch.cuda.stream(s2):
            for r in range(reps):
                o = (r * m) % (n - m)
                tb[o:o + m].copy_(dev2[o:o + m], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"bidirectional_chunk_MiB": chunk_mib, "each_direction_GBps": round(reps * m * 8 / dt / 1e9, 1)}), flush=True)
