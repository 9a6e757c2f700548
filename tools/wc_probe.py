#!/usr/bin/env python3
"""Aggregate PCIe rate of N ranks copying both ways at once, with default pinned vs write-combined pinned host buffers
(cudaHostAllocWriteCombined for the H2D source).  torchrun --nproc-per-node N tools/wc_probe.py"""
import ctypes as C, json, os, time
import torch
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
rt = C.CDLL("libcudart.so.12")
n = 1 << 30                                    # bytes per direction per rank
def host_alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    return p
dsrc = torch.empty(n, dtype=torch.uint8, device="cuda"); ddst = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, fl_in, fl_out in (("default/default", 0, 0), ("wc/default", 4, 0), ("wc/wc", 4, 4), ("portable/portable", 1, 1)):
    hin, hout = host_alloc(fl_in), host_alloc(fl_out)
    best = 0
    for rep in range(4):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        rt.cudaMemcpyAsync(C.c_void_p(ddst.data_ptr()), hin, C.c_size_t(n), 1, C.c_void_p(s1.cuda_stream))
        rt.cudaMemcpyAsync(hout, C.c_void_p(dsrc.data_ptr()), C.c_size_t(n), 2, C.c_void_p(s2.cuda_stream))
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = time.perf_counter() - t0
        best = max(best, n / dt / 1e9)
    res[name] = round(best * world, 1)
    rt.cudaFreeHost(hin); rt.cudaFreeHost(hout)
if rank == 0:
    print(json.dumps({"world": world, "GBps_per_direction_all_ranks": res}), flush=True)
