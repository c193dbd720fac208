"""Synthetic workloads of BASELINE.json's configs.

Only generators live here: a grid, the polar power-flow model written with this package's
front-end exactly in the form the survey probed on the reference (``P_eqn`` over ``PVPQ``,
``Q_eqn`` over ``PQ``, ``Va[ref] = 0``, ``Vm[ref u pv] = 1``; G and B walked in separate ``Sum``s),
and the per-scenario load draws.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import scipy.sparse as sps

from .equations import Eqn, LoopEqn, Model
from .sym import Param, Set, Sum, Var, cos, sin

__all__ = ["lattice_grid", "polar_pf_model", "pf_scenarios", "swing_dae_model", "swing_fault_scenarios"]


def lattice_grid(nx: int, ny: int, chord_frac: float = 0.02, seed: int = 0):
    """nx*ny lattice plus ``chord_frac * nb`` random chords; r~U(.01,.05), x~U(.05,.2), no shunts.
    Returns (nb, Ybus as complex csc)."""
    rng = np.random.default_rng(seed)
    nb = nx * ny
    edges = []
    for i in range(nx):
        for j in range(ny):
            if i + 1 < nx:
                edges.append((i * ny + j, (i + 1) * ny + j))
            if j + 1 < ny:
                edges.append((i * ny + j, i * ny + j + 1))
    nch = int(round(chord_frac * nb))
    es = set(edges)
    while nch > 0:
        a, b = (int(v) for v in rng.integers(0, nb, 2))
        e = (min(a, b), max(a, b))
        if a == b or e in es:
            continue
        es.add(e)
        edges.append(e)
        nch -= 1
    edges = np.array(edges)
    r = rng.uniform(.01, .05, len(edges))
    x = rng.uniform(.05, .2, len(edges))
    y = 1.0 / (r + 1j * x)
    a, b = edges[:, 0], edges[:, 1]
    rows = np.concatenate([a, b, a, b])
    cols = np.concatenate([a, b, b, a])
    vals = np.concatenate([y, y, -y, -y])
    Y = sps.coo_matrix((vals, (rows, cols)), shape=(nb, nb)).tocsc()
    Y.sum_duplicates()
    Y.sort_indices()
    return nb, Y


def polar_pf_model(nx: int, ny: int, chord_frac: float = 0.02, seed: int = 0):
    """Returns (eqs, y0, info) for the polar AC power flow of the lattice; n = 2 * nb unknowns."""
    nb, Y = lattice_grid(nx, ny, chord_frac, seed)
    G = sps.csc_matrix(Y.real)
    Bm = sps.csc_matrix(Y.imag)
    npv = max(1, int(0.1 * nb))
    pv = np.arange(1, 1 + npv)
    pq = np.arange(1 + npv, nb)
    rng = np.random.default_rng(seed + 1)
    Pb = rng.uniform(-.05, .05, nb)
    Qb = rng.uniform(-.02, .02, nb)
    m = Model()
    m.Va = Var('Va', np.zeros(nb))
    m.Vm = Var('Vm', np.ones(nb))
    m.G = Param('G', G, dim=2, sparse=True)
    m.B = Param('B', Bm, dim=2, sparse=True)
    m.Pinj = Param('Pinj', Pb, per_scenario=True)
    m.Qinj = Param('Qinj', Qb, per_scenario=True)
    m.Bus = Set('Bus', nb)
    m.PVPQ = Set('PVPQ', np.concatenate([pv, pq]))
    m.PQ = Set('PQ', pq)
    i = m.PVPQ.idx('i')
    j = m.Bus.idx('j')
    body_p = (m.Vm[i] * Sum(m.Vm[j] * m.G[i, j] * cos(m.Va[i] - m.Va[j]), j)
              + m.Vm[i] * Sum(m.Vm[j] * m.B[i, j] * sin(m.Va[i] - m.Va[j]), j) - m.Pinj[i])
    m.P_eqn = LoopEqn('P_eqn', outer_index=i, body=body_p, model=m)
    k = m.PQ.idx('k')
    body_q = (m.Vm[k] * Sum(m.Vm[j] * m.G[k, j] * sin(m.Va[k] - m.Va[j]), j)
              - m.Vm[k] * Sum(m.Vm[j] * m.B[k, j] * cos(m.Va[k] - m.Va[j]), j) - m.Qinj[k])
    m.Q_eqn = LoopEqn('Q_eqn', outer_index=k, body=body_q, model=m)
    m.ref_eqn = Eqn('Va_ref', m.Va[0])
    m.vm_eqn = Eqn('Vm_fix', m.Vm[0:1 + npv] - 1)
    eqs, y0 = m.create_instance()
    info = dict(nb=nb, Ybus=Y, pv=pv, pq=pq, ref=np.array([0]), Pbase=Pb, Qbase=Qb, npv=npv)
    return eqs, y0, info


def pf_scenarios(info: Dict, B: int, first: int = 0, sigma: float = 0.2) -> Tuple[np.ndarray, np.ndarray]:
    """Load scenarios b = first .. first+B-1: ``P,Q[b] = base * (1 + sigma * N(0,1))``, seed 1000+b."""
    nb = info["nb"]
    P = np.empty((B, nb))
    Q = np.empty((B, nb))
    for b in range(B):
        rng = np.random.default_rng(1000 + first + b)
        P[b] = info["Pbase"] * (1.0 + sigma * rng.standard_normal(nb))
        Q[b] = info["Qbase"] * (1.0 + sigma * rng.standard_normal(nb))
    return P, Q


# ------------------------------------------------------------------------------------------------
# config #4: stiff DAE ensemble -- one machine (swing + flux decay + exciter) per bus of a lattice grid,
# rectangular-coordinate stator and network current balance, bus-fault scenarios
# ------------------------------------------------------------------------------------------------
def swing_dae_model(nx: int, ny: int, ns=None, seed: int = 0, chord_frac: float = 0.02):
    """Builds the model with the DSL in namespace ``ns`` (this package by default; the golden generator
    passes the reference's ``Solverz`` module -- the two front-ends take the same statements, which is
    the point of a drop-in).  Per machine i (at bus i):

        delta' = wb (omega - 1)                        omega' = (Pm - (Ux Ix + Uy Iy) - D (omega - 1)) / Tj
        Eq'    = (Efd - Eq - (xd - xdp) Id) / Td0      Efd'   = (Ka (Vref - |U|) - Efd) / Ta
        0 = xdp Ix - Eq sin(delta) + Uy                0 = xdp Iy + Eq cos(delta) - Ux
        0 = G Ux - B Uy + Fx - Ix                      0 = G Uy + B Ux + Fy - Iy
        0 = Fx - gf(t) fmask Ux                        0 = Fy - gf(t) fmask Uy

    with Id = Ix sin(delta) - Iy cos(delta), constant-admittance loads folded into Y, and the bus fault a
    shunt conductance gf(t) (a TimeSeriesParam) at the buses selected by the 0/1 vector ``fmask``.  The fault
    current (Fx, Fy) is a variable of its own: the reference's Jacobian drops d/dUx of ``gf*fmask*Ux`` when that
    product shares an equation with ``Mat_Mul`` terms (found while generating the goldens), and a model the
    reference differentiates correctly is what a parity fixture needs.
    Returns (model, info); the model starts in equilibrium (gf = 0)."""
    if ns is None:
        import solverz_b200 as ns
    own = ns.__name__ == "solverz_b200"
    nb, Y = lattice_grid(nx, ny, chord_frac, seed)
    rng = np.random.default_rng(seed + 7)
    yload = rng.uniform(0.3, 0.8, nb) - 1j * rng.uniform(0.05, 0.2, nb)      # constant-admittance loads
    Yt = (Y + sps.diags(yload)).tocsc()
    G = sps.csc_array(Yt.real)
    Bm = sps.csc_array(Yt.imag)
    # equilibrium: pick the bus voltages, everything else follows
    th = 0.05 * rng.standard_normal(nb)
    U = (1.0 + 0.02 * rng.standard_normal(nb)) * np.exp(1j * th)
    I = Yt @ U
    xd, xdp = rng.uniform(0.8, 1.2, nb), rng.uniform(0.2, 0.4, nb)
    E = U + 1j * xdp * I
    Eq0, d0 = np.abs(E), np.angle(E)
    Id0 = I.real * np.sin(d0) - I.imag * np.cos(d0)
    Efd0 = Eq0 + (xd - xdp) * Id0
    Pm = (U * np.conj(I)).real
    Ka, Ta = rng.uniform(20.0, 50.0, nb), rng.uniform(0.05, 0.2, nb)
    Vref = np.abs(U) + Efd0 / Ka
    Tj, D, Td0 = rng.uniform(4.0, 10.0, nb), rng.uniform(1.0, 3.0, nb), rng.uniform(4.0, 8.0, nb)
    wb = 2 * np.pi * 50.0
    kw = dict(per_scenario=True) if own else {}
    m = ns.Model()
    m.delta = ns.Var('delta', d0)
    m.omega = ns.Var('omega', np.ones(nb))
    m.Eq = ns.Var('Eq', Eq0)
    m.Efd = ns.Var('Efd', Efd0)
    m.Ux = ns.Var('Ux', U.real)
    m.Uy = ns.Var('Uy', U.imag)
    m.Ix = ns.Var('Ix', I.real)
    m.Iy = ns.Var('Iy', I.imag)
    m.Fx = ns.Var('Fx', np.zeros(nb))
    m.Fy = ns.Var('Fy', np.zeros(nb))
    m.G = ns.Param('G', G, dim=2, sparse=True)
    m.B = ns.Param('B', Bm, dim=2, sparse=True)
    for nm, v in (("Pm", Pm), ("Dm", D), ("Tj", Tj), ("Td0", Td0), ("xd", xd), ("xdp", xdp), ("Ka", Ka),
                  ("Ta", Ta), ("Vref", Vref)):
        setattr(m, nm, ns.Param(nm, v))
    m.fmask = ns.Param('fmask', np.zeros(nb), **kw)
    m.gf = ns.TimeSeriesParam('gf', v_series=[0.0, 0.0, 5.0, 5.0, 0.0, 0.0],
                              time_series=[0.0, 0.1, 0.101, 0.2, 0.201, 10.0], **kw)
    sin, cos, Mat_Mul = ns.sin, ns.cos, ns.Mat_Mul
    m.f_delta = ns.Ode('f_delta', wb * (m.omega - 1), diff_var=m.delta)
    m.f_omega = ns.Ode('f_omega', (m.Pm - (m.Ux * m.Ix + m.Uy * m.Iy) - m.Dm * (m.omega - 1)) / m.Tj,
                       diff_var=m.omega)
    m.f_Eq = ns.Ode('f_Eq', (m.Efd - m.Eq - (m.xd - m.xdp) * (m.Ix * sin(m.delta) - m.Iy * cos(m.delta))) / m.Td0,
                    diff_var=m.Eq)
    m.f_Efd = ns.Ode('f_Efd', (m.Ka * (m.Vref - (m.Ux ** 2 + m.Uy ** 2) ** 0.5) - m.Efd) / m.Ta, diff_var=m.Efd)
    m.g_sx = ns.Eqn('g_sx', m.xdp * m.Ix - m.Eq * sin(m.delta) + m.Uy)
    m.g_sy = ns.Eqn('g_sy', m.xdp * m.Iy + m.Eq * cos(m.delta) - m.Ux)
    m.g_nx = ns.Eqn('g_nx', Mat_Mul(m.G, m.Ux) - Mat_Mul(m.B, m.Uy) + m.Fx - m.Ix)
    m.g_ny = ns.Eqn('g_ny', Mat_Mul(m.G, m.Uy) + Mat_Mul(m.B, m.Ux) + m.Fy - m.Iy)
    m.g_fx = ns.Eqn('g_fx', m.Fx - m.gf * m.fmask * m.Ux)
    m.g_fy = ns.Eqn('g_fy', m.Fy - m.gf * m.fmask * m.Uy)
    info = dict(nb=nb, n=10 * nb, Ybus=Yt)
    return m, info


def swing_fault_scenarios(nb: int, B: int):
    """Scenario b: fault at bus ``b mod nb`` from t_f = 0.1 + 0.01 (b mod 5) for 0.08 + 0.01 (b mod 3) s with
    conductance 3 + (b mod 4).  Returns (fmask[B, nb], ts_t[B, 6], ts_v[B, 6])."""
    fmask = np.zeros((B, nb))
    tst = np.zeros((B, 6))
    tsv = np.zeros((B, 6))
    for b in range(B):
        fmask[b, b % nb] = 1.0
        t0 = 0.1 + 0.01 * (b % 5)
        t1 = t0 + 0.08 + 0.01 * (b % 3)
        g = 3.0 + (b % 4)
        tst[b] = [0.0, t0, t0 + 1e-3, t1, t1 + 1e-3, 10.0]
        tsv[b] = [0.0, 0.0, g, g, 0.0, 0.0]
    return fmask, tst, tsv


# ------------------------------------------------------------------------------------------------
# config #5: district-heating hydraulic-thermal FDAE -- supply network on a lattice, hydraulics by mass
# continuity + loop pressure drops (mutable-matrix Jacobian), upwind pipe-outlet temperatures with an
# AliasVar history node, time-stepped load profiles
# ------------------------------------------------------------------------------------------------
def dhs_fdae_model(nx: int, ny: int, ns=None, seed: int = 0, dt: float = 60.0):
    """Builds the model with the DSL in namespace ``ns`` (this package by default; the golden generator
    passes the reference's ``Solverz`` module).  Nodes on an nx x ny lattice, the source at node 0, pipes
    oriented away from it (+x, +y); with loads at every other node all flows stay positive, which is what
    the fixed upwind direction of the thermal part assumes.

        0 = V_ns m - sc(t) m_inj                                  mass continuity at the non-source nodes
        0 = L (K m |m|)                                           pressure drop around every lattice cell
        0 = Tin - Vf Tn                                           pipe inlet = temperature of the node it leaves
        0 = Tout - Tout_prev - cth m (Tin - Tout) + lam (Tout - Ta)        pipe outlet, one history node
        0 = Vin (m Tout) - Tn[1:] (Vin m)                         mixing at the non-source nodes
        0 = Tn[0] - Ts                                            supply temperature at the source

    (Tin is a variable of its own: the reference's matrix calculus cannot differentiate ``m * (Vf Tn - Tout)``,
    an element-wise product around a Mat_Mul -- found while generating the goldens.)
    ``sc`` (a TimeSeriesParam) scales the loads over time and is the per-scenario input.  Returns
    (model, info); the model starts from the steady state of the nominal load (sc = 1)."""
    if ns is None:
        import solverz_b200 as ns
    own = ns.__name__ == "solverz_b200"
    rng = np.random.default_rng(seed + 11)
    n_node = nx * ny
    node = lambda ix, iy: ix * ny + iy
    pipes = []
    for ix in range(nx):
        for iy in range(ny):
            if ix + 1 < nx:
                pipes.append((node(ix, iy), node(ix + 1, iy)))
            if iy + 1 < ny:
                pipes.append((node(ix, iy), node(ix, iy + 1)))
    n_pipe = len(pipes)
    pidx = {p: k for k, p in enumerate(pipes)}
    # incidence / loop / selection matrices as sparse triplets (50 000 pipes: the dense forms would need 10 GB each)
    pa = np.array([p[0] for p in pipes])
    pb = np.array([p[1] for p in pipes])
    ar = np.arange(n_pipe)
    V = sps.csr_matrix((np.concatenate([-np.ones(n_pipe), np.ones(n_pipe)]),
                        (np.concatenate([pa, pb]), np.concatenate([ar, ar]))), shape=(n_node, n_pipe))
    lr, lc, lv = [], [], []
    nl = 0
    for ix in range(nx - 1):
        for iy in range(ny - 1):                         # around the cell: +x, +y, -x, -y
            lc += [pidx[(node(ix, iy), node(ix + 1, iy))], pidx[(node(ix + 1, iy), node(ix + 1, iy + 1))],
                   pidx[(node(ix, iy + 1), node(ix + 1, iy + 1))], pidx[(node(ix, iy), node(ix, iy + 1))]]
            lv += [1.0, 1.0, -1.0, -1.0]
            lr += [nl] * 4
            nl += 1
    Lm = sps.csr_matrix((lv, (lr, lc)), shape=(nl, n_pipe))
    K = rng.uniform(0.5, 2.0, n_pipe)
    load = rng.uniform(0.5, 1.5, n_node - 1)            # withdrawals at the non-source nodes (flow into the node)
    V_ns = V[1:, :].tocsr()
    Vin = V_ns.maximum(0.0).tocsr()                     # pipes that end at the node
    Vin.eliminate_zeros()
    Vf = sps.csr_matrix((np.ones(n_pipe), (ar, pa)), shape=(n_pipe, n_node))   # the node a pipe starts from
    # steady state of the nominal load: hydraulic Newton, then temperatures node by node (lattice order is topological)
    rhs0 = np.concatenate([load, np.zeros(nl)])
    if n_pipe <= 500:                                   # the small fixtures keep their dense start bit for bit
        mflow = np.linalg.lstsq(np.vstack([V_ns.toarray(), Lm.toarray()]), rhs0, rcond=None)[0]
    else:
        import scipy.sparse.linalg as sla
        mflow = sla.spsolve(sps.vstack([V_ns, Lm]).tocsc(), rhs0)
    for _ in range(50):
        res = np.concatenate([V_ns @ mflow - load, Lm @ (K * mflow * np.abs(mflow))])
        if n_pipe <= 500:
            Jh = np.vstack([V_ns.toarray(), Lm.toarray() * (2 * K * np.abs(mflow))[None, :]])
            mflow = mflow - np.linalg.solve(Jh, res)
        else:
            Jh = sps.vstack([V_ns, Lm @ sps.diags(2 * K * np.abs(mflow))]).tocsc()
            dm = sla.spsolve(Jh, res)
            mflow = mflow - dm
            if np.abs(dm).max() <= 1e-14 * max(1.0, np.abs(mflow).max()):
                break
    # The thermal part fixes the upwind direction of every pipe, so every pipe must carry its flow from its first to
    # its second node: pipes whose nominal flow came out negative (loop flows on a large lattice) are turned round.
    # The loads of a scenario are scaled uniformly, and with quadratic resistances m(sc) = sc m(1) exactly, so the
    # orientation holds for every scenario and time.  (The small fixtures have no such pipe: nothing changes there.)
    flip = mflow < 0
    if flip.any():
        sgn = np.where(flip, -1.0, 1.0)
        mflow = np.abs(mflow)
        pa, pb = np.where(flip, pb, pa), np.where(flip, pa, pb)
        V = sps.csr_matrix(V @ sps.diags(sgn))
        Lm = sps.csr_matrix(Lm @ sps.diags(sgn))
        V_ns = V[1:, :].tocsr()
        Vin = V_ns.maximum(0.0).tocsr()
        Vin.eliminate_zeros()
        Vf = sps.csr_matrix((np.ones(n_pipe), (ar, pa)), shape=(n_pipe, n_node))
        scale = max(1.0, float(np.abs(K * mflow * mflow).max()))
        assert np.abs(V_ns @ mflow - load).max() < 1e-8 * scale and \
            np.abs(Lm @ (K * mflow * np.abs(mflow))).max() < 1e-8 * scale, "re-oriented network no longer balanced"
    cth, lam, Ta, Ts = 0.08 * dt / 60.0, 0.01 * dt / 60.0, 10.0, 90.0
    Tn = np.full(n_node, Ts)
    Tout = np.full(n_pipe, Ts)
    for _ in range(max(200, 2 * (nx + ny))):              # fixed point of the steady thermal equations
        Tin = Vf @ Tn
        Tout = (cth * mflow * Tin + lam * Ta) / (cth * mflow + lam)
        Tn[1:] = (Vin @ (mflow * Tout)) / (Vin @ mflow)
    kw = dict(per_scenario=True) if own else {}
    m = ns.Model()
    m.m = ns.Var('m', mflow)
    m.Tout = ns.Var('Tout', Tout)
    m.Tn = ns.Var('Tn', Tn)
    m.Tin = ns.Var('Tin', Vf @ Tn)
    m.Tout0 = ns.AliasVar('Tout', value=Tout, step=1) if own else ns.AliasVar('Tout', step=1, value=Tout)
    for nm, A in (("V_ns", V_ns), ("L", Lm), ("Vin", Vin), ("Vf", Vf)):
        setattr(m, nm, ns.Param(nm, sps.csc_array(A), dim=2, sparse=True))   # structural zeros are not stored
    m.K = ns.Param('K', K)
    m.m_inj = ns.Param('m_inj', load)
    m.sc = ns.TimeSeriesParam('sc', v_series=[1.0, 1.0, 1.2, 1.2], time_series=[0.0, dt, 3 * dt, 1e6], **kw)
    Mat_Mul, Abs = ns.Mat_Mul, ns.Abs
    m.mass = ns.Eqn('mass', Mat_Mul(m.V_ns, m.m) - m.sc * m.m_inj)
    m.loop = ns.Eqn('loop', Mat_Mul(m.L, m.K * m.m * Abs(m.m)))
    m.inlet = ns.Eqn('inlet', m.Tin - Mat_Mul(m.Vf, m.Tn))
    m.pipe = ns.Eqn('pipe', m.Tout - m.Tout0 - cth * m.m * (m.Tin - m.Tout) + lam * (m.Tout - Ta))
    m.mix = ns.Eqn('mix', Mat_Mul(m.Vin, m.m * m.Tout) - m.Tn[1:n_node] * Mat_Mul(m.Vin, m.m))
    m.src = ns.Eqn('src', m.Tn[0] - Ts)
    info = dict(n_node=n_node, n_pipe=n_pipe, n=3 * n_pipe + n_node, dt=dt)
    return m, info


def dhs_load_profiles(B: int, dt: float = 60.0):
    """Scenario b: the loads ramp from 1 to 1 + 0.05 (b mod 7 + 1) between t = dt (1 + b mod 3) and two steps
    later, then stay.  Returns (ts_t[B, 4], ts_v[B, 4])."""
    tst = np.zeros((B, 4))
    tsv = np.zeros((B, 4))
    for b in range(B):
        t0 = dt * (1 + b % 3)
        a = 1.0 + 0.05 * (b % 7 + 1)
        tst[b] = [0.0, t0, t0 + 2 * dt, 1e6]
        tsv[b] = [1.0, 1.0, a, a]
    return tst, tsv
