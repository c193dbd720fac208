"""Batched Newton-Raphson with the reference's ``nr_method(eqn, y, opt)`` surface.

Reference: Solverz/solvers/nlaesolver/nr.py:4-66 (loop 45-61, counters 44-61, ``succeed`` 63-64) and
the ``ae_io_parser`` decorator (solvers/parser.py:12-38: a ``Vars`` in gives a ``Vars`` out).  The
loop itself runs inside ``szb_newton`` (csrc/szb_core.cu); this file only moves data.

``y`` may be a 1-D array / ``Vars`` (one scenario), a ``[B, n]`` array or a CUDA ``torch.Tensor``
(``B`` scenarios).  Per-scenario parameters in ``eqn.p`` may have a leading ``B`` axis.
"""
from __future__ import annotations

import numpy as np

from ..num_api.model import CudaModel
from ..variables import Vars
from .option import Opt
from .solution import aesol
from .stats import Stats


def _pivot_fallback(eqn: CudaModel, s, y_start, y_res, it, fl, opt, rounds: int = 3):
    """Scenarios that did not converge are solved again with a pivot order of their OWN: the static order of the
    model comes from the Jacobian at (y0, default parameters), while the reference pivots inside every call
    (laesolver.py:11-15 -> SuperLU, partial pivoting), so a scenario far from the representative point may have
    failed on the pivots rather than on the problem.  The symbolic analysis is redone on the Jacobian of the first
    failed scenario at ITS start, the failed ones are re-run from their starts on that schedule (still on the
    GPU); up to ``rounds`` times.  A scenario that also fails there keeps its original result and flag."""
    import torch
    from ..lu.symbolic import analyze
    from ..num_api.model import BatchSession
    import scipy.sparse as sps
    for _ in range(rounds):
        bad = np.flatnonzero(fl != 0)
        if len(bad) == 0:
            break
        idx = torch.as_tensor(bad, device=y_start.device)
        one = eqn.session(1, s.device, with_lu=False)
        one.set_params({k: (v[int(bad[0])] if eqn_param_is_batched(eqn, k, v, s.B) else v) for k, v in s.P.items()})
        one.set_y(y_start[int(bad[0])])
        jv = one.eval_J(0.0)[0].cpu().numpy()
        if not np.all(np.isfinite(jv)):
            break
        A = sps.csc_matrix((jv, eqn.indices, eqn.indptr), shape=(eqn.n, eqn.n))
        try:
            sym = analyze(A, forced_pairs=eqn.mass, **eqn.lu_options)
        except ValueError:
            break
        fb = BatchSession(eqn, len(bad), s.device, sym=sym)
        try:
            fb.set_params({k: (v[idx] if eqn_param_is_batched(eqn, k, v, s.B) else v) for k, v in s.P.items()})
            fb.set_y(y_start[idx])
            it2, fl2 = fb.newton(opt.ite_tol, opt.max_it)
            it2 = it2.cpu().numpy().astype(np.int64)
            fl2 = fl2.cpu().numpy().astype(np.int64)
            ok = fl2 == 0
            if not ok.any():
                break
            good = torch.as_tensor(bad[ok], device=y_res.device)
            y_res[good] = fb.Y[torch.as_tensor(np.flatnonzero(ok), device=y_res.device)]
            it[bad[ok]] = it2[ok]
            fl[bad[ok]] = 0
        finally:
            fb.close()
    return y_res, it, fl


def eqn_param_is_batched(eqn: CudaModel, name: str, tns, B: int) -> bool:
    pm = next((p for p in eqn.params_meta if p["name"] == name), None)
    return bool(pm and pm["per_scenario"] and tns.ndim == 2 and tns.shape[0] == B)


def nr_method(eqn: CudaModel, y, opt: Opt = None):
    if opt is None:
        opt = Opt()
    if not isinstance(eqn, CudaModel):
        raise TypeError("nr_method of solverz_b200 needs a model rendered with lang='cuda' "
                        "(module_printer(...).render() or made_numerical)")
    import torch
    vars_in = y if isinstance(y, Vars) else None
    y_arr = y.array if isinstance(y, Vars) else y
    on_device = isinstance(y_arr, torch.Tensor)
    if not on_device:
        y_arr = np.asarray(y_arr, dtype=float)
    single = y_arr.ndim == 1
    B = 1 if single else int(y_arr.shape[0])
    for pm in eqn.params_meta:
        v = eqn.p.get(pm["name"])
        if pm["per_scenario"] and v is not None and np.ndim(v) == 2 and single and np.shape(v)[0] != 1:
            raise ValueError(f"parameter {pm['name']} has a batch axis of {np.shape(v)[0]} but y is a single "
                             f"scenario")
    s = eqn.session(B)
    s.set_params(eqn.p)
    s.set_y(y_arr)
    y_start = s.Y.clone() if getattr(opt, "pivot_fallback", True) else None
    iters, flags = s.newton(opt.ite_tol, opt.max_it)
    stats = Stats('Newton')
    it = iters.cpu().numpy().astype(np.int64)
    fl = flags.cpu().numpy().astype(np.int64)
    y_res = s.Y
    if y_start is not None and (fl != 0).any():
        y_res, it, fl = _pivot_fallback(eqn, s, y_start, y_res.clone(), it, fl, opt)
    y_out = y_res.clone() if on_device else y_res.cpu().numpy()
    if single:
        stats.nstep = int(it[0])
        stats.nfeval = stats.nstep + 1
        stats.nJeval = stats.ndecomp = stats.nsolve = stats.nstep
        stats.succeed = bool(fl[0] == 0)
        stats.flag = int(fl[0])
        if fl[0] == 1:
            print(f"Cannot converge within 100 iterations. Deviation: {float(s.eval_F()[0].abs().max())}!")
        y_out = y_out[0]
    else:
        stats.nstep = it
        stats.nfeval = it + 1
        stats.nJeval = it.copy()
        stats.ndecomp = it.copy()
        stats.nsolve = it.copy()
        stats.succeed = fl == 0
        stats.flag = fl
    if vars_in is not None:
        y_out = Vars(vars_in.a, y_out if not on_device else y_out.cpu().numpy())
    return aesol(y_out, stats)
