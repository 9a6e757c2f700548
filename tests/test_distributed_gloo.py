"""N>1 path on CPU: two gloo ranks shard the point range exactly as bench.py does on GPUs (contiguous
shards, per-rank generator offset, no data-path collective; only the control-plane reductions).  The
per-shard compute is done by the oracle here (tests may use it) — what is under test is the sharding,
the seed arithmetic of the per-rank input streams and the sum / max reductions."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

import helpers as H

WORKER = textwrap.dedent(
    """
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
    import torch.distributed as dist
    import oracle as O
    import materialmodels_jl_b200 as mm
    from materialmodels_jl_b200 import sharding, synth as sy, _abi
    rank, world, local = sharding.init_process_group(backend="gloo")
    assert dist.get_world_size() == 2 and world == 2
    N = 30001                                   # odd on purpose: shards of 15000 and 15001 points
    lo, hi = sharding.shard_range(N, rank, world)
    nc, a, b, seed = sy.c2_strain(0)
    # bench.py's per-rank stream: same generator, seed advanced by lo*nc  ==  points [lo, hi) of the global stream
    eps = sy.host_uniform(hi - lo, nc, seed + lo * nc, a, b)
    assert np.array_equal(eps, sy.host_uniform(hi - lo, nc, seed, a, b, i0=lo))
    p = _abi.MMPlastic(*[sy.PLASTIC_PARAMS[k] for k in ("E", "nu", "sigma_y", "H", "r", "kappa_inf", "alpha_inf")])
    O.set_threads(2)
    r = O.plastic(p, eps, np.zeros((14, hi - lo)), max_iter={max_iter})
    n_fail = sharding.allreduce_sum(int(r["n_fail"]))
    n_plastic = sharding.allreduce_sum(int(r["flags"].astype(bool).sum()))
    checksum = sharding.allreduce_sum(float(r["sig"].sum()))
    npts = sharding.allreduce_sum(hi - lo)
    tmax = sharding.allreduce_max(1.0 + rank)
    sharding.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(dict(n_fail=n_fail, n_plastic=n_plastic, checksum=checksum, npts=npts, tmax=tmax)))
    dist.destroy_process_group()
    """
)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_two_ranks(tmp_path, max_iter):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=H.ROOT, max_iter=max_iter))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    import json

    return json.loads(line[len("RESULT "):])


def test_two_gloo_ranks_match_single_process(tmp_path, oracle):
    import materialmodels_jl_b200.synth as sy

    res = run_two_ranks(tmp_path, 1000)
    N = 30001
    nc, a, b, seed = sy.c2_strain(0)
    r = oracle.plastic(H.plastic_c(), sy.host_uniform(N, nc, seed, a, b), np.zeros((14, N)))
    assert res["npts"] == N and res["tmax"] == 2.0
    assert res["n_fail"] == 0 == r["n_fail"]
    assert res["n_plastic"] == int(r["flags"].astype(bool).sum())
    assert np.isclose(res["checksum"], float(r["sig"].sum()), rtol=1e-12)


def test_two_gloo_ranks_sum_the_nonconverged_counters(tmp_path, oracle):
    import materialmodels_jl_b200.synth as sy

    res = run_two_ranks(tmp_path, 1)        # one Newton update is not enough: both ranks report failures
    N = 30001
    nc, a, b, seed = sy.c2_strain(0)
    r = oracle.plastic(H.plastic_c(), sy.host_uniform(N, nc, seed, a, b), np.zeros((14, N)), max_iter=1)
    assert r["n_fail"] > 0 and res["n_fail"] == r["n_fail"]
