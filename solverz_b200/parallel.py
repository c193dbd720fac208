"""Multi-GPU plumbing: scenarios are independent, so the batch is block-partitioned over ranks with no
collective inside the iteration; NCCL (gloo in CPU tests) only gathers solutions, iteration counts
and flags at the end."""
from __future__ import annotations

from typing import Tuple


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of scenarios of `rank`: the first B % world ranks get one extra."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(y, iters, flags, B: int):
    """all_gather of ragged shards [b_r, n], [b_r], [b_r] -> full [B, n], [B], [B] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = [shard_range(B, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    n = y.shape[1]

    def pad(t, shape):
        out = torch.zeros(shape, dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        return out
    ys = [torch.zeros((mx, n), dtype=y.dtype, device=y.device) for _ in range(world)]
    its = [torch.zeros(mx, dtype=iters.dtype, device=y.device) for _ in range(world)]
    fls = [torch.zeros(mx, dtype=flags.dtype, device=y.device) for _ in range(world)]
    dist.all_gather(ys, pad(y, (mx, n)))
    dist.all_gather(its, pad(iters, (mx,)))
    dist.all_gather(fls, pad(flags, (mx,)))
    Y = torch.cat([ys[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)])
    IT = torch.cat([its[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)])
    FL = torch.cat([fls[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)])
    return Y, IT, FL
