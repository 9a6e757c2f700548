#!/usr/bin/env python3
"""Host-side probe for the e2e (host-buffer) path on a multi-GPU box: where do pinned buffers land, and does the NUMA
placement of the staging memory bound the aggregate PCIe rate?  Run under torchrun (one rank per GPU).
For each policy the rank allocates fresh pinned buffers, then all ranks time `mmh_plastic` together.
policies: default | interleave (MPOL_INTERLEAVE over all nodes) | local (MPOL_BIND to the GPU's node + CPU affinity)"""
import ctypes, glob, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import sharding, synth as sy

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238           # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND, MPOL_INTERLEAVE = 0, 1, 2, 3


def nodes():
    return sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))


def set_mempolicy(mode, nodelist):
    mask = 0
    for n in nodelist:
        mask |= 1 << n
    m = ctypes.c_ulong(mask)
    rc = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(m) if nodelist else None, 64 if nodelist else 0)
    return rc, ctypes.get_errno()


def gpu_numa_node(local):
    bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        busid = busid.lower()
        if len(busid.split(":")[0]) == 8:
            busid = busid[4:]
        return int(open("/sys/bus/pci/devices/%s/numa_node" % busid).read()), busid
    except Exception as e:
        return -1, repr(e)


def node_cpus(n):
    s = open("/sys/devices/system/node/node%d/cpulist" % n).read().strip()
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def main():
    rank, world, local = sharding.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    Ne = int(float(os.environ.get("PROBE_POINTS", "2e7")))
    node, busid = gpu_numa_node(local)
    allowed = sorted(os.sched_getaffinity(0))
    info = {"rank": rank, "gpu": local, "busid": busid, "gpu_numa_node": node, "nodes": nodes(), "allowed_cpus": "%d..%d (%d)" % (allowed[0], allowed[-1], len(allowed))}
    print(json.dumps(info), flush=True)
    if rank == 0:
        os.system("nvidia-smi topo -m 2>&1 | head -20; lscpu | grep -i -E 'numa|socket|model name' ; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null")
    if os.environ.get("PROBE_CHUNK"):
        import materialmodels_jl_b200._lib as L
        L.load().mmh_set_chunk_points(int(float(os.environ["PROBE_CHUNK"])))
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = sy.host_uniform(Ne, nc, seed, lo, hi)
    for policy in os.environ.get("PROBE_POLICIES", "default,interleave,local").split(","):
        note = ""
        if policy == "interleave":
            rc, er = set_mempolicy(MPOL_INTERLEAVE, nodes()); note = "set_mempolicy rc=%d errno=%d" % (rc, er)
        elif policy == "local" and node >= 0:
            cpus = [c for c in node_cpus(node) if c in allowed]
            if cpus:
                os.sched_setaffinity(0, cpus)
            rc, er = set_mempolicy(MPOL_BIND, [node]); note = "bind node %d, %d cpus, rc=%d errno=%d" % (node, len(cpus), rc, er)
        else:
            set_mempolicy(MPOL_DEFAULT, [])
        h_eps, h_st = mm.PinnedArray((6, Ne)), mm.PinnedArray((14, Ne))
        h_out = {"stress": mm.PinnedArray((6, Ne)), "tangent": mm.PinnedArray((36, Ne)), "state": mm.PinnedArray((14, Ne)),
                 "flags": mm.PinnedArray((Ne,), np.uint8), "iters": mm.PinnedArray((Ne,), np.uint16)}
        h_eps.array[...] = eps
        h_st.array[...] = 0.0
        outs = {k: v.array for k, v in h_out.items()}
        st = mm.PlasticState(h_st.array)
        mm.material_response(m, h_eps.array, st, out=outs)
        sharding.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            mm.material_response(m, h_eps.array, st, out=outs)
        dt = time.perf_counter() - t0
        dtm = sharding.allreduce_max(dt, device=dev)
        if rank == 0:
            print(json.dumps({"policy": policy, "note": note, "chunk": os.environ.get("PROBE_CHUNK", "default"), "world": world, "updates_per_s": world * Ne * 3 / dtm, "GBps_total": world * Ne * 3 * 610 / dtm / 1e9}), flush=True)
        for v in [h_eps, h_st] + list(h_out.values()):
            v.free()
        set_mempolicy(MPOL_DEFAULT, [])
        os.sched_setaffinity(0, allowed)
    sharding.barrier()


if __name__ == "__main__":
    main()
