"""Pair-level data parallelism: pairs are independent units, so ranks shard the batch with
no collective in the data path; torch.distributed is only used for the barrier and for the max-over-ranks
of the device time.  The reference has no multi-GPU path at all (single GPU, single stream)."""
import os


def shard_pairs(n_pairs, world, rank):
    """Contiguous, balanced shard of pair indices [0, n_pairs) for `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    """Initialise torch.distributed when launched under torchrun with WORLD_SIZE > 1."""
    import torch.distributed as dist

    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(pairs_per_rank_per_step, world, ms_per_step_max):
    """Whole-job pairs/s: all ranks' pairs divided by the slowest rank's time."""
    return world * pairs_per_rank_per_step * 1e3 / ms_per_step_max
