"""Oracle: the reference's Rodas loop restated with numpy + SciPy splu.

TEST INFRASTRUCTURE ONLY.  Follows Solverz/solvers/daesolver/rodas/rodas.py:66-344 line by line
(step-size safeguards 127-146, stages 148-174, error estimate and controller 175-192, acceptance and
dense output 195-327), ``dfdt`` 347-352, ``ntrp1`` 383-386, and ``DaeIc`` (daeic.py:24-68).  The
linear solves are ``lu_decomposition`` -> ``scipy.sparse.linalg.splu`` (laesolver.py:18-45).  Events
(rodas.py:200-271) are not restated: the CUDA path does not implement them in this round.

``dfdt_mode``: 0 reproduces the reference's rendered-module path, where ``f0`` and ``f1`` alias one
buffer and dF/dt evaluates to exactly 0; 1 is the literal formula, which
is what the inline ``made_numerical`` path computes.
Pinned by tests/golden/dae_test_rodas.npz = the reference's own tests/dae_test.xlsx sheets
``rodas`` / ``rodas_dense`` and by outputs of the reference run in the build container.
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as sla


def daeic(F, J, M, y0, t0, rtol):
    """daeic.py:24-68."""
    M = sps.csc_matrix(M)
    F0 = F(t0, y0)
    de, dv = M.nonzero()
    n = y0.shape[0]
    alg_var = np.setdiff1d(np.arange(n), dv)
    alg_eqn = np.setdiff1d(np.arange(n), de)
    if np.linalg.norm(F0[alg_eqn]) <= 1e-6:
        return y0
    y = y0.copy()
    for _ in range(15):
        Jn = sps.csc_matrix(J(t0, y))[np.ix_(alg_eqn, alg_var)]
        Fn = F(t0, y)
        dY = sla.spsolve(sps.csc_matrix(Jn), -Fn[alg_eqn])
        nz = np.abs(y[alg_var]) > np.spacing(t0)
        res = np.linalg.norm(dY[nz] / y[alg_var][nz])
        lam = 1.0
        ynew = y.copy()
        for _probe in range(3):
            ynew[alg_var] = y[alg_var] + lam * dY
            Fnew = F(t0, ynew)
            if np.linalg.norm(Fnew[alg_eqn]) <= 1e-5 * rtol:
                return ynew
            dYnew = sla.spsolve(sps.csc_matrix(sps.csc_matrix(J(t0, ynew))[np.ix_(alg_eqn, alg_var)]),
                                Fnew[alg_eqn])
            resnew = np.linalg.norm(dYnew[nz] / y[alg_var][nz])
            if resnew < 0.9 * res:
                break
            lam = 0.5 * lam
        y = ynew
        if resnew <= 1e-3 * rtol:
            return ynew
    raise ValueError("Need Better y0")


def rodas(F, J, M, tspan, y0, rp, rtol=1e-3, atol=1e-6, f_savety=0.9, facmax=6, fac1=0.2, fac2=6,
          hinit=None, hmax=None, fix_h=False, dfdt_mode=0):
    """F(t,y)->ndarray, J(t,y)->sparse, M sparse, rp = tableau object (alpha, gammatilde, a, g, b, bd,
    c, d, e, s, pord, gamma).  Returns dict(T, Y, stats)."""
    M = sps.csc_matrix(M)
    tspan = np.array(tspan, dtype=float)
    tend, t0 = tspan[-1], tspan[0]
    if hmax is None:
        hmax = np.abs(tend - t0)
    st = dict(nstep=0, nfeval=0, nJeval=0, ndecomp=0, nreject=0, nsolve=0, ret=None)
    n = y0.shape[0]
    nt = 0
    t = t0
    hmin = 16 * np.spacing(t0)
    uround = np.spacing(1.0)
    T = [t0]
    y0 = daeic(F, J, M, np.array(y0, dtype=float), t0, rtol)
    Y = [y0.copy()]
    dense = len(tspan) > 2
    told = t0
    if dense:
        inext = 1
        tnext = tspan[inext]
    dt = 1e-6 * (tend - t0) if hinit is None else hinit
    dt = np.maximum(dt, hmin)
    dt = np.minimum(dt, hmax)
    done = False
    reject = 0
    Jm = None
    while not done:
        if np.abs(dt) < uround:
            st["ret"] = "failed"
            break
        if reject > 100:
            st["ret"] = "failed"
            break
        if t + dt >= tend:
            dt = tend - t
        else:
            dt = np.minimum(dt, 0.5 * (tend - t))
        if fix_h:
            dt = hinit
        K = np.zeros((n, rp.s))
        if reject == 0:
            Jm = sps.csc_matrix(J(t, y0))
            st["nJeval"] += 1
        if dfdt_mode:
            tscale = np.maximum(0.1 * np.abs(t), 1e-8)
            ddt = t + np.sqrt(np.spacing(1)) * tscale - t
            dfdt0 = dt * (F(t + ddt, y0) - F(t, y0)) / ddt
        else:
            dfdt0 = np.zeros(n)
        rhs = F(t, y0) + rp.g[0] * dfdt0
        st["nfeval"] += 1
        try:
            lu = sla.splu(sps.csc_matrix(M - dt * rp.gamma * Jm))
        except RuntimeError:
            break
        st["ndecomp"] += 1
        K[:, 0] = lu.solve(rhs)
        for j in range(1, rp.s):
            sum_1 = K @ rp.alpha[:, j]
            sum_2 = K @ rp.gammatilde[:, j]
            y1 = y0 + dt * sum_1
            rhs = F(t + dt * rp.a[j], y1) + M @ sum_2 + rp.g[j] * dfdt0
            st["nfeval"] += 1
            sol = lu.solve(rhs)
            K[:, j] = sol - sum_2
        st["nsolve"] += rp.s
        sum_1 = K @ (dt * rp.b)
        ynew = y0 + sum_1
        if not fix_h:
            sum_2 = K @ (dt * rp.bd)
            SK = atol + rtol * np.abs(ynew)
            err = np.max(np.abs((sum_1 - sum_2) / SK))
            if np.any(np.isinf(ynew)) or np.any(np.isnan(ynew)):
                err = 1.0e6
            err = np.maximum(err, 1.0e-6)
            fac = f_savety / (err ** (1 / rp.pord))
            fac = np.minimum(facmax, np.maximum(fac1, fac))
            dtnew = dt * fac
        else:
            err = 1.0
            dtnew = dt
        if err <= 1.0:
            reject = 0
            told = t
            t = t + dt
            st["nstep"] += 1
            if dense:
                while t >= tnext > told:
                    tau = (tnext - told) / dt
                    ynext = y0 + tau * dt * K @ (rp.b + (tau - 1) * (rp.c + tau * (rp.d + tau * rp.e)))
                    T.append(tnext)
                    Y.append(ynext)
                    inext += 1
                    tnext = tspan[inext] if inext <= len(tspan) - 1 else tend + dt
            else:
                T.append(t)
                Y.append(ynew.copy())
            if np.abs(tend - t) < uround:
                done = True
            y0 = ynew
            facmax = fac2
        else:
            reject += 1
            st["nreject"] += 1
            facmax = 1
        dt = np.min([hmax, np.max([hmin, dtnew])])
    return dict(T=np.array(T), Y=np.array(Y), stats=st)
