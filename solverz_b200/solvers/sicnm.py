"""Batched semi-implicit continuous Newton method with the reference's ``sicnm(ae, y0, opt)`` surface
(SURVEY.md §8f item 3).

Reference: Solverz/solvers/nlaesolver/sicnm.py:13-369.  ``0 = g(y)`` becomes the DAE ``y' = z``,
``0 = J(y) z + g(y)`` (:19-31), integrated by Rodas from ``z0 = -J(y0)^-1 g(y0)`` (:97) over ``[0, 10000]`` (:81)
until ``max |g(y)| < ite_tol`` (:333); the controller is the Rodas one (:120-140, :219-233, :344-350).  Every step
the reference factors the 2n x 2n matrix ``[[I, -a I], [-a (H + J), -a J]]``, ``a = dt gamma``, ``H = HVP(y, z)``
(:157-186).

Here the first block row is eliminated by hand -- ``k1 = r1 + a k2`` and ``(a H + (1 + a) J) k2 = -(r2 / a +
(H + J) r1)`` -- so one n x n refactorisation on the model's own pivot schedule serves all ``s`` stage solves
(``H`` has the pattern of ``J``: row ``i`` of ``J v`` depends on the variables row ``i`` of ``g`` does).  ``g``, ``J``
and ``H`` come from the generated row kernels (``H`` from the ``make_hvp`` twin model, solverz_b200/hvp.py), the
factorisation and the triangular sweeps from libszb; the step controller runs here as elementwise torch
operations over the scenario axis, every scenario with its own ``t``, ``dt``, reject counter and ``facmax`` -- the
arithmetic of one scenario does not depend on its neighbours, so a batch reproduces the reference's
single-scenario runs step for step.  Nothing is computed on the CPU.

``y0``: 1-D array / ``Vars`` (one scenario), ``[B, n]`` array or CUDA tensor.  ``Stats`` counters are the
reference's (``nstep``, ``nfeval`` = stage evaluations, ``ndecomp``, ``nreject``), arrays ``[B]`` for a batch."""
from __future__ import annotations

import numpy as np

from ..num_api.model import CudaModel
from ..variables import Vars
from .option import Opt
from .rodas_param import Rodas_param
from .solution import aesol
from .stats import Stats

T_END = 10000.0          # sicnm.py:81


def sicnm(ae: CudaModel, y0, opt: Opt = None):
    if opt is None:
        opt = Opt()
    if not isinstance(ae, CudaModel):
        raise TypeError("sicnm of solverz_b200 needs a model rendered with lang='cuda' and make_hvp=True")
    if ae.kind != "AE":
        raise TypeError("sicnm solves algebraic equations (AE)")
    if opt.event is not None:
        raise NotImplementedError("sicnm: events are not supported by this backend")
    hm = ae.hvp_model                                    # raises when the module was rendered without make_hvp
    import torch
    vars_in = y0 if isinstance(y0, Vars) else None
    y_arr = y0.array if isinstance(y0, Vars) else y0
    on_device = isinstance(y_arr, torch.Tensor)
    if not on_device:
        y_arr = np.asarray(y_arr, dtype=float)
    single = y_arr.ndim == 1
    B = 1 if single else int(y_arr.shape[0])
    n = ae.n
    s = ae.session(B)
    sh = hm.session(B, with_lu=False)
    s.set_params(ae.p)
    sh.set_params({k: v for k, v in ae.p.items() if k in hm.p})
    s.set_y(y_arr)
    dev = s.Y.device
    f64 = dict(dtype=torch.float64, device=dev)
    rp = Rodas_param(opt.scheme)
    ns = rp.s
    alpha = np.asarray(rp.alpha, dtype=float)
    gt = np.asarray(rp.gammatilde, dtype=float)
    b_w = torch.tensor(np.asarray(rp.b, dtype=float), **f64)
    bd_w = torch.tensor(np.asarray(rp.bd, dtype=float), **f64)

    # pattern plumbing: CSC slot -> row / column, and the slots of H inside the pattern of J
    cols_j = np.repeat(np.arange(n), np.diff(ae.indptr))
    rows_t = torch.as_tensor(ae.indices, device=dev)
    cols_t = torch.as_tensor(cols_j, device=dev)
    key_j = cols_j * n + ae.indices
    cols_h = np.repeat(np.arange(n), np.diff(hm.indptr))
    key_h = cols_h * n + hm.indices
    pos = np.searchsorted(key_j, key_h)
    if len(key_h) and (pos.max(initial=0) >= len(key_j) or np.any(key_j[np.minimum(pos, len(key_j) - 1)] != key_h)):
        raise NotImplementedError("sicnm: the Hessian-vector product has entries outside the pattern of J")
    h2j = torch.as_tensor(pos, device=dev)
    vslices = {k: ae.a[k] for k in ae.a.names}

    def spmv(vals, x):                                   # [B, nnz] on the pattern of J times [B, n]
        return torch.zeros(B, n, **f64).index_add_(1, rows_t, vals * x[:, cols_t])

    def eval_gJ(y):
        s.Y.copy_(y)
        return s.eval_F().clone(), s.eval_J().clone()

    def eval_H(y, v):
        sh.Y.copy_(y)
        for k, sl in vslices.items():
            sh.P["v__" + k].copy_(v[:, sl])
        hv = torch.zeros(B, ae.nnz_j, **f64)
        hv[:, h2j] = sh.eval_J()                         # the slots of H are distinct slots of J
        return hv

    def lin_solve(rhs):
        s.RHS.copy_(rhs)
        return s.lu_solve().clone()

    y = s.Y.clone()
    g0, J0 = eval_gJ(y)
    s.lu_refactor(0.0)                                   # v0 = solve(J(y0), -F(y0))  (:97)
    v = lin_solve(-g0)

    t = torch.zeros(B, **f64)
    hmax = T_END if opt.hmax is None else float(opt.hmax)
    hmin = 16 * np.spacing(0.0)
    uround = np.spacing(1.0)
    dt = torch.full((B,), 1e-6 * T_END if opt.hinit is None else float(opt.hinit), **f64).clamp(hmin, hmax)
    facmax = torch.full((B,), float(opt.facmax), **f64)
    reject = torch.zeros(B, dtype=torch.int64, device=dev)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    failed = torch.zeros(B, dtype=torch.bool, device=dev)
    nstep = torch.zeros(B, dtype=torch.int64, device=dev)
    nfeval = torch.zeros_like(nstep)
    ndecomp = torch.zeros_like(nstep)
    nreject = torch.zeros_like(nstep)
    Hv = torch.zeros(B, ae.nnz_j, **f64)
    Jv = J0
    t_hist = [t.clone()] if single else None
    y_hist = [y.clone()] if single else None
    max_rows = 10000                                     # sicnm.py:325-328

    while not bool(done.all()):
        act = ~done
        small = act & (dt.abs() < uround)                # :123-127
        many = act & (reject > 100)                      # :129-133
        if bool((small | many).any()):
            failed |= small | many
            done |= small | many
            act = ~done
            if not bool(act.any()):
                break
        dt = torch.where(t + dt >= T_END, T_END - t, torch.minimum(dt, 0.5 * (T_END - t)))   # :136-139
        fresh = act & (reject == 0)                      # :157-159: H and J are kept over rejected attempts
        g_y, J_y = eval_gJ(y)
        if bool(fresh.any()):
            Hn = eval_H(y, v)
            m = fresh[:, None]
            Hv = torch.where(m, Hn, Hv)
            Jv = torch.where(m, J_y, Jv)
        a = (dt * rp.gamma)[:, None]
        s.JV.copy_(a * Hv + (1.0 + a) * Jv)
        s.lu_refactor(0.0)
        ndecomp += act
        HJ = Hv + Jv

        def stage(r1, r2):
            k2 = lin_solve(-(r2 / a + spmv(HJ, r1)))
            return r1 + a * k2, k2

        K1 = torch.zeros(ns, B, n, **f64)
        K2 = torch.zeros(ns, B, n, **f64)
        K1[0], K2[0] = stage(v, spmv(J_y, v) + g_y)      # F_tilda(y0, v0), :72-73, :162
        nfeval += act
        for j in range(1, ns):
            al = torch.tensor(alpha[:, j], **f64)[:, None, None]
            ga = torch.tensor(gt[:, j], **f64)[:, None, None]
            s1y, s1v = (K1 * al).sum(0), (K2 * al).sum(0)
            s2y, s2v = (K1 * ga).sum(0), (K2 * ga).sum(0)
            y1 = y + dt[:, None] * s1y
            v1 = v + dt[:, None] * s1v
            g1, J1 = eval_gJ(y1)
            k1, k2 = stage(v1 + s2y, spmv(J1, v1) + g1)  # rhs = F_tilda(y1, v1) + M @ sum_2, :204
            K1[j], K2[j] = k1 - s2y, k2 - s2v            # :219
            nfeval += act
        bw = b_w[:, None, None]
        bdw = bd_w[:, None, None]
        dtc = dt[:, None]
        s1y, s1v = dtc * (K1 * bw).sum(0), dtc * (K2 * bw).sum(0)
        s2y, s2v = dtc * (K1 * bdw).sum(0), dtc * (K2 * bdw).sum(0)
        ynew, vnew = y + s1y, v + s1v
        yt = torch.cat([ynew, vnew], dim=1)
        sk = opt.atol + opt.rtol * yt.abs()
        err = (torch.cat([s1y - s2y, s1v - s2v], dim=1) / sk).abs().amax(dim=1)
        bad = ~torch.isfinite(yt).all(dim=1)
        err = torch.where(bad | ~torch.isfinite(err), torch.full_like(err, 1.0e6), err).clamp_min(1.0e-6)
        fac = torch.minimum(facmax, (opt.f_savety / err ** (1.0 / rp.pord)).clamp_min(opt.fac1))
        dtnew = dt * fac
        acc = act & (err <= 1.0)
        rej = act & ~acc
        if bool(acc.any()):
            m = acc[:, None]
            t = torch.where(acc, t + dt, t)
            nstep += acc
            y = torch.where(m, ynew, y)
            v = torch.where(m, vnew, v)
            s.Y.copy_(y)
            conv = s.eval_F().abs().amax(dim=1) < opt.ite_tol           # :333
            done |= acc & (conv | ((T_END - t).abs() < uround) | (nstep >= max_rows))
            failed |= acc & (nstep >= max_rows)
            if single:
                t_hist.append(t.clone())
                y_hist.append(y.clone())
        reject = torch.where(acc, torch.zeros_like(reject), reject + rej)
        nreject += rej
        facmax = torch.where(acc, torch.full_like(facmax, float(opt.fac2)), torch.where(rej, torch.ones_like(facmax), facmax))
        dt = torch.where(act, dtnew.clamp(hmin, hmax), dt)

    stats = Stats('Sicnm based on ' + opt.scheme)
    cnt = [c.cpu().numpy() for c in (nstep, nfeval, ndecomp, nreject)]
    ok = (~failed).cpu().numpy()
    y_out = y.clone() if on_device else y.cpu().numpy()
    if single:
        stats.nstep, stats.nfeval, stats.ndecomp, stats.nreject = (int(c[0]) for c in cnt)
        stats.succeed = bool(ok[0])
        if not ok[0]:
            stats.ret = 'failed'
        stats.T = torch.stack(t_hist)[:, 0].cpu().numpy()
        stats.y = torch.stack(y_hist)[:, 0].cpu().numpy()
        y_out = y_out[0]
    else:
        stats.nstep, stats.nfeval, stats.ndecomp, stats.nreject = cnt
        stats.succeed = ok
        stats.T = t.cpu().numpy()                        # the pseudo-time each scenario stopped at
    if vars_in is not None:
        y_out = Vars(vars_in.a, y_out if not on_device else y_out.cpu().numpy())
    return aesol(y_out, stats)
