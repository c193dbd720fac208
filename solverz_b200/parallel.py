"""Multi-GPU: one process per GPU (``torchrun``), scenarios block-partitioned over the ranks, no collective inside the
iteration; NCCL over NVLink (gloo in the CPU tests) only gathers solutions, iteration counts and flags at the end
(SURVEY.md section 8e).

    # torchrun --nproc-per-node 8 my_study.py
    dist.init_process_group("nccl"); torch.cuda.set_device(local_rank)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q               # the FULL batch on every rank (host arrays, [B, len])
    pend = nr_method_distributed(mdl, Y0, Opt(ite_tol=1e-8))     # solves this rank's shard, starts the gather
    ... next batch ...                                 # the all_gather runs on NCCL's stream meanwhile
    sol = pend.wait()                                  # aesol of the whole batch, on every rank

The gather is issued asynchronously from private send buffers, so it overlaps whatever the rank does next (the next
batch of a parameter study, as bench.py does); ``wait()`` is the only synchronisation point.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


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


class PendingGather:
    """Handle of an asynchronous gather (``gather_results_async``); ``wait()`` returns the full arrays."""

    def __init__(self, works, ys, its, fls, sizes):
        self._works, self._ys, self._its, self._fls, self._sizes = works, ys, its, fls, sizes
        self._done = None

    def wait(self):
        import torch
        if self._done is None:
            for w in self._works:
                w.wait()
            sz = self._sizes
            self._done = (torch.cat([self._ys[r][:hi - lo] for r, (lo, hi) in enumerate(sz)]),
                          torch.cat([self._its[r][:hi - lo] for r, (lo, hi) in enumerate(sz)]),
                          torch.cat([self._fls[r][:hi - lo] for r, (lo, hi) in enumerate(sz)]))
        return self._done


def gather_results_async(y, iters, flags, B: int) -> PendingGather:
    """Like ``gather_results``, but the three all_gathers are only ISSUED (``async_op=True``) from private copies of
    the shards: the caller may overwrite ``y`` / ``iters`` / ``flags`` at once (the next batch) and calls ``wait()``
    when it needs the whole result."""
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
    ys = [torch.empty((mx, n), dtype=y.dtype, device=y.device) for _ in range(world)]
    its = [torch.empty(mx, dtype=iters.dtype, device=y.device) for _ in range(world)]
    fls = [torch.empty(mx, dtype=flags.dtype, device=y.device) for _ in range(world)]
    works = [dist.all_gather(ys, pad(y, (mx, n)), async_op=True),
             dist.all_gather(its, pad(iters, (mx,)), async_op=True),
             dist.all_gather(fls, pad(flags, (mx,)), async_op=True)]
    return PendingGather(works, ys, its, fls, sizes)


class PendingSolution:
    def __init__(self, pending: PendingGather, B: int):
        self._p, self._B = pending, B

    def wait(self):
        """``aesol`` of the whole batch (``y`` a device tensor ``[B, n]``; ``Stats`` counters int arrays ``[B]``)."""
        from .solvers.solution import aesol
        from .solvers.stats import Stats
        Y, it, fl = self._p.wait()
        it = it.cpu().numpy().astype(np.int64)
        fl = fl.cpu().numpy().astype(np.int64)
        st = Stats('Newton')
        st.nstep, st.nfeval = it, it + 1
        st.nJeval, st.ndecomp, st.nsolve = it.copy(), it.copy(), it.copy()
        st.succeed, st.flag = fl == 0, fl
        return aesol(Y, st)


def nr_method_distributed(eqn, y, opt=None) -> PendingSolution:
    """``nr_method`` over all ranks of the default process group: this rank solves scenarios
    ``shard_range(B, rank, world)`` of the batch ``y[B, n]`` (per-scenario parameters in ``eqn.p`` carry the full
    ``B`` axis on every rank) on its own GPU and starts the asynchronous gather; see the module docstring."""
    import torch
    import torch.distributed as dist
    from .solvers.nr import nr_method
    from .solvers.option import Opt
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("nr_method_distributed needs an initialised torch.distributed process group "
                           "(one process per GPU, launched with torchrun)")
    if opt is None:
        opt = Opt()
    rank, world = dist.get_rank(), dist.get_world_size()
    y_arr = y.array if hasattr(y, "array") else y
    if y_arr.ndim != 2:
        raise ValueError("nr_method_distributed takes a batch y[B, n]")
    B = int(y_arr.shape[0])
    lo, hi = shard_range(B, rank, world)
    full = dict(eqn.p)
    try:
        for pm in eqn.params_meta:
            v = full.get(pm["name"])
            if pm["per_scenario"] and v is not None and np.ndim(v) == 2 and np.shape(v)[0] == B:
                eqn.p[pm["name"]] = v[lo:hi]
        ys = y_arr[lo:hi]
        if not isinstance(ys, torch.Tensor):
            ys = torch.as_tensor(np.asarray(ys, dtype=float)).to(torch.device("cuda", torch.cuda.current_device()))
        sol = nr_method(eqn, ys, opt)
    finally:
        eqn.p.update(full)
    dev = sol.y.device
    it = torch.as_tensor(np.asarray(sol.stats.nstep, dtype=np.int32)).to(dev)
    fl = torch.as_tensor(np.asarray(sol.stats.flag, dtype=np.int32)).to(dev)
    return PendingSolution(gather_results_async(sol.y, it, fl, B), B)
