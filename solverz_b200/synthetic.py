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

__all__ = ["lattice_grid", "polar_pf_model", "pf_scenarios"]


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
