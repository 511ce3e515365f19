"""Multi-GPU plumbing: scenarios are independent, so the batch is split contiguously over ranks with no collective
on the data path; the only exchange is an optional gather of the fixed-size per-scenario results afterwards
(SURVEY.md §8e).  Backend-agnostic (`nccl` on the box, `gloo` in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(B, world, rank):
    """Rank r owns scenarios [lo, hi): contiguous, sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, B, group=None):
    """all_gather of per-scenario rows: local[hi-lo, ...] on every rank -> [B, ...] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
