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
    iters, flags = s.newton(opt.ite_tol, opt.max_it)
    stats = Stats('Newton')
    it = iters.cpu().numpy().astype(np.int64)
    fl = flags.cpu().numpy().astype(np.int64)
    y_out = s.Y.clone() if on_device else s.Y.cpu().numpy()
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
