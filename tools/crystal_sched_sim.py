"""Scheduling model of the lane-persistent crystal Newton kernel on the CPU (no GPU needed).

The host twin (tests/hosttwin) records, for every point of the configs[3] workload, the candidate count of each
Newton pass (sparse pass: systems with x_i != 0 or Phi_i > 0).  This script replays those traces through a model
of one warp — 32 persistent lanes, a chunk of points, refill in batches of >= THRESH finished lanes — and counts
warp-level issue slots under a cost model  fixed + per_cand * (longest list in the warp)  for
  * the order the points come in,
  * the chunk re-ordered by the candidate count of the first pass (what a warp can compute for its chunk up front),
  * an oracle order (by total work) as the bound of what any ordering can give.
Usage: python tools/crystal_sched_sim.py [npoints] [step]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import helpers as H  # noqa: E402
import oracle as O  # noqa: E402
from materialmodels_jl_b200 import synth as sy  # noqa: E402


def traces(N, step):
    tw = H.hosttwin()
    cp = H.crystal_c(O, q=0.0, structure=0)
    st = np.zeros((42, N))
    out = None
    for s in range(step + 1):
        nc, lo, hi, seed = sy.c4_strain_increment(s)
        de = sy.host_uniform(N, nc, seed, lo, hi)
        sig, tan, so = np.zeros((6, N)), np.zeros((36, N)), np.zeros((42, N))
        ac, it = np.zeros(N, np.uint16), np.zeros(N, np.uint16)
        buf = np.zeros(N * 120, np.uint8)
        tw.twin_crystal_trace.argtypes = [C.c_void_p, C.c_longlong]
        tw.twin_crystal_trace_len.restype = C.c_longlong
        tw.twin_crystal_trace(buf.ctypes.data, buf.size)
        nfb = C.c_int64(0)
        tw.twin_crystal_mode(C.byref(cp), H.P(de), C.c_double(1.0), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(ac), H.P(it),
                             C.c_double(1e-8), 100, C.c_int64(N), 2, C.byref(nfb))
        n = tw.twin_crystal_trace_len()
        tw.twin_crystal_trace(None, 0)
        t = buf[:n]
        starts = np.flatnonzero(t == 255)
        out = [t[a + 1:b] for a, b in zip(starts, list(starts[1:]) + [n])]
        st = so
    return out


def simulate(tr, order, chunk, thresh, fixed, per_cand, dense=False):
    """returns (issue slots, lane-slots useful / lane-slots issued in the per-candidate part)"""
    total = 0.0
    useful = 0
    issued = 0
    for c0 in range(0, len(order), chunk):
        pts = [tr[j] for j in order[c0:c0 + chunk]]
        nxt = 0
        lanes = [None] * 32          # (trace, position)
        done = [False] * 32
        while True:
            idle = [l for l in range(32) if lanes[l] is None]
            if idle and nxt < len(pts):
                for l in idle:
                    if nxt < len(pts):
                        lanes[l] = [pts[nxt], 0]
                        done[l] = False
                        nxt += 1
            work = [l for l in range(32) if lanes[l] is not None]
            if not work:
                break
            run = [l for l in work if not done[l]]
            if run:
                longest = 12 if dense else max(int(lanes[l][0][lanes[l][1]]) for l in run)
                total += fixed + per_cand * longest
                issued += 32 * longest
                useful += sum(int(lanes[l][0][lanes[l][1]]) for l in run)
                for l in run:
                    lanes[l][1] += 1
                    if lanes[l][1] >= len(lanes[l][0]):
                        done[l] = True
            fin = [l for l in work if done[l]]
            iterating = len(work) - len(fin)
            if fin and (len(fin) >= thresh or iterating == 0 or nxt >= len(pts)):
                for l in fin:
                    lanes[l] = None
    return total, useful / max(issued, 1)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    tr = traces(N, step)
    npass = np.array([len(t) for t in tr])
    first = np.array([int(t[0]) for t in tr])
    work = np.array([int(t.sum()) for t in tr])
    print("points %d  passes/point %.2f  cand/pass %.2f  first-pass cand mean %.2f" % (N, npass.mean(), work.sum() / npass.sum(), first.mean()))
    print("corr(first-pass count, mean count over passes) = %.3f" % np.corrcoef(first, work / npass)[0, 1])
    # cost model in FP64 warp-instructions per warp-pass: sparse fixed part (sigma from nz, yield test, LDL, solve) and per candidate
    # (chain + rank-1 update of M + y_r + update);  dense pass = everything for all 12 systems
    FIX, PER = 300.0, 85.0
    for chunk in (256, 512, 1024):
        nat = list(range(N))
        by_first = []
        by_work = []
        by_two = []
        for c0 in range(0, N, chunk):
            idx = np.arange(c0, min(c0 + chunk, N))
            by_first += list(idx[np.argsort(-first[idx], kind="stable")])
            by_work += list(idx[np.argsort(-(work[idx] / npass[idx]), kind="stable")])
            by_two += list(idx[np.lexsort((-npass[idx], -first[idx]))])
        d, _ = simulate(tr, nat, chunk, 4, 150.0, 90.0, dense=True)
        for name, order in (("natural", nat), ("by first-pass count", by_first), ("by count, then passes (oracle)", by_two), ("by mean count (oracle)", by_work)):
            s, u = simulate(tr, order, chunk, 4, FIX, PER)
            print("chunk %4d  %-32s sparse slots/point %7.0f  lanes useful %.2f   (dense model %7.0f)" % (chunk, name, s / N * 1.0, u, d / N))


if __name__ == "__main__":
    main()
