"""Multi-GPU plumbing: points are independent, so the point index range is split into contiguous
shards, one process per GPU, with NO collective on the data path.  The only
exchanges are tiny control-plane reductions: the non-converged counter (sum) and timings (max)."""
import os


def shard_range(N, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`: lo = floor(rank*N/world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return (rank * N) // world, ((rank + 1) * N) // world


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Join the torchrun rendezvous if WORLD_SIZE > 1; returns (rank, world, local_rank)."""
    rank, world, local = env_rank_world()
    if world > 1:
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                # NCCL writes its banner / debug lines to stdout unless told otherwise; the bench's stdout is one JSON line
                os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/mmb200_nccl.%h.%p.log")
                torch.cuda.set_device(local)
                dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend=backend)
        # the host-buffer entry points expand / fill dense arrays with the host cores: one process per GPU shares them
        share_host_cores(int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    return rank, world, local


def share_host_cores(processes_on_this_host):
    """Give this process its share of the host cores for the library's host-side work (mmh_set_host_threads): `cores // processes`,
    at least 2.  Called by init_process_group under torchrun; returns the thread count set (0 if the library is not loaded)."""
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return 0
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    n = max(2, cores // max(1, int(processes_on_this_host)))
    lib.mmh_set_host_threads(n)
    return n


def allreduce_sum(value, device="cpu"):
    """Sum of a python int/float over ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64 if isinstance(value, float) else torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


def allreduce_max(value, device="cpu"):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
