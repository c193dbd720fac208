"""The same Solverz model object on both back-ends, on the GPU box: built with `from Solverz import ...` (private copy
baseline/_ref, which travels with the snapshot), solved by the reference's own numba module + nr_method / Rodas on
the host and by this package's module_printer(..., lang='cuda') + nr_method / Rodas on the device."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
if not (REF / "Solverz" / "__init__.py").exists():
    pytest.skip("baseline/_ref (private copy of the reference) is not present", allow_module_level=True)
if str(REF) not in sys.path:
    sys.path.insert(0, str(REF))


def test_power_flow_same_model_object_both_backends(libszb, tmp_path):
    import Solverz as SZ
    import solverz_b200 as SB
    from solverz_b200.synthetic import lattice_grid
    nx = 6
    nb, Y = lattice_grid(nx, nx)
    npv = max(1, int(0.1 * nb))
    pv, pq = list(range(1, 1 + npv)), list(range(1 + npv, nb))
    rng = np.random.default_rng(1)
    Pb, Qb = rng.uniform(-.05, .05, nb), rng.uniform(-.02, .02, nb)
    m = SZ.Model()
    m.Va = SZ.Var('Va', np.zeros(nb))
    m.Vm = SZ.Var('Vm', np.ones(nb))
    m.G = SZ.Param('G', sps.csc_array(Y.real), dim=2, sparse=True)
    m.B = SZ.Param('B', sps.csc_array(Y.imag), dim=2, sparse=True)
    m.Pinj = SZ.Param('Pinj', Pb)
    m.Qinj = SZ.Param('Qinj', Qb)
    m.Bus = SZ.Set('Bus', nb)
    m.PVPQ = SZ.Set('PVPQ', np.array(pv + pq))
    m.PQ = SZ.Set('PQ', np.array(pq))
    i, j, k = m.PVPQ.idx('i'), m.Bus.idx('j'), m.PQ.idx('k')
    sin, cos, Sum = SZ.sin, SZ.cos, SZ.Sum
    m.P_eqn = SZ.LoopEqn('P_eqn', outer_index=i,
                         body=m.Vm[i] * Sum(m.Vm[j] * m.G[i, j] * cos(m.Va[i] - m.Va[j]), j)
                         + m.Vm[i] * Sum(m.Vm[j] * m.B[i, j] * sin(m.Va[i] - m.Va[j]), j) - m.Pinj[i], model=m)
    m.Q_eqn = SZ.LoopEqn('Q_eqn', outer_index=k,
                         body=m.Vm[k] * Sum(m.Vm[j] * m.G[k, j] * sin(m.Va[k] - m.Va[j]), j)
                         - m.Vm[k] * Sum(m.Vm[j] * m.B[k, j] * cos(m.Va[k] - m.Va[j]), j) - m.Qinj[k], model=m)
    m.ref_eqn = SZ.Eqn('Va_ref', m.Va[0])
    m.vm_eqn = SZ.Eqn('Vm_fix', m.Vm[0:1 + npv] - 1)
    spf, y0 = m.create_instance()
    os.environ.setdefault("NUMBA_CACHE_DIR", str(tmp_path / "numba"))
    # back-end 1: the reference's renderer
    SZ.module_printer(spf, y0, "ga_pf_ref", directory=str(tmp_path), jit=True).render()
    # back-end 2: this package's renderer, same objects, lang='cuda'
    SB.module_printer(spf, y0, "ga_pf_cuda", lang='cuda', directory=str(tmp_path),
                      per_scenario=("Pinj", "Qinj")).render()     # the scenario inputs of this batch
    sys.path.insert(0, str(tmp_path))
    ref = importlib.import_module("ga_pf_ref")
    dev = importlib.import_module("ga_pf_cuda")
    yp = y0.array + 0.05 * np.random.default_rng(7).standard_normal(2 * nb)
    Fr, Jr = ref.mdl.F(yp, ref.mdl.p).copy(), sps.csc_matrix(ref.mdl.J(yp, ref.mdl.p))
    Fd, Jd = dev.mdl.F(yp, dev.mdl.p), sps.csc_matrix(dev.mdl.J(yp, dev.mdl.p))
    assert np.abs(Fd - Fr).max() <= 1e-12 and abs(Jd - Jr).max() <= 1e-12 * abs(Jr).max()
    assert np.array_equal(Jd.indices, Jr.indices) and np.array_equal(Jd.indptr, Jr.indptr)
    B = 8
    srng = np.random.default_rng(11)
    P = Pb[None, :] * (1 + 0.2 * srng.standard_normal((B, nb)))
    Q = Qb[None, :] * (1 + 0.2 * srng.standard_normal((B, nb)))
    dev.mdl.p["Pinj"], dev.mdl.p["Qinj"] = P, Q
    sol = SB.nr_method(dev.mdl, np.tile(y0.array, (B, 1)), SB.Opt(ite_tol=1e-8))
    for b in range(B):
        ref.mdl.p["Pinj"], ref.mdl.p["Qinj"] = P[b], Q[b]
        rs = SZ.nr_method(ref.mdl, ref.y, SZ.Opt(ite_tol=1e-8))
        assert int(sol.stats.nstep[b]) == int(rs.stats.nstep)
        assert np.abs(sol.y[b] - rs.y.array).max() <= 1e-10 * np.abs(rs.y.array).max()


def test_dae_same_model_object_rodas(libszb):
    import Solverz as SZ
    import solverz_b200 as SB
    m = SZ.Model()
    m.x = SZ.Var('x', 1)
    m.y = SZ.Var('y', 1)
    m.f = SZ.Ode(name='f', f=-m.x ** 3 + 0.5 * m.y ** 2, diff_var=m.x)
    m.g = SZ.Eqn(name='g', eqn=m.x ** 2 + m.y ** 2 - 2)
    dae, y0 = m.create_instance()
    ref = SZ.made_numerical(dae, y0, sparse=True)
    dev = SB.made_numerical(dae, y0, sparse=True)          # the reference's DAE object, lowered by the adapter
    s1 = SZ.Rodas(ref, [0, 20], y0, SZ.Opt(hinit=0.1))
    s2 = SB.Rodas(dev, [0, 20], y0.array, SB.Opt(hinit=0.1))
    assert s2.stats.nstep == s1.stats.nstep and s2.stats.nreject == s1.stats.nreject
    np.testing.assert_allclose(s2.T, s1.T, rtol=1e-10)
    assert np.abs(np.asarray(s2.Y) - s1.Y.array).max() < 1e-8


def test_triggerable_parameter_on_the_device(libszb, tmp_path):
    """Param(triggerable=True) (equation/param.py:16-59) lowered as frozen(f(vars)): F, J on the device and the whole
    Newton iteration (which sees a Jacobian WITHOUT the derivative through f, like the reference's) against the
    reference's made_numerical + nr_method: same solution, same iteration count."""
    import Solverz as SZ
    import solverz_b200 as SB
    m = SZ.Model()
    m.x = SZ.Var('x', [0.3, 0.7, 1.1])
    m.y = SZ.Var('y', [1.0, 2.0, 0.5])
    m.ax = SZ.Param('ax', [0.0, 0.0, 0.0], triggerable=True, trigger_var=['x'], trigger_fun=np.sin)
    m.c = SZ.Param('c', [0.5, 1.5, 2.5])
    m.e1 = SZ.Eqn('e1', 0.1 * m.ax * m.y + 2 * m.x - m.c)
    m.e2 = SZ.Eqn('e2', 3 * m.y - 0.2 * m.ax + m.x * m.y - m.c)
    spf, y0 = m.create_instance()
    ref = SZ.made_numerical(spf, y0, sparse=True)
    dev = SB.made_numerical(spf, y0, sparse=True)
    yp = y0.array + 0.1 * np.random.default_rng(2).standard_normal(6)
    Fr, Jr = np.asarray(ref.F(yp, ref.p)).copy(), sps.csc_matrix(ref.J(yp, ref.p))
    Fd, Jd = dev.F(yp, dev.p), sps.csc_matrix(dev.J(yp, dev.p))
    assert np.abs(Fd - Fr).max() <= 1e-13 and abs(Jd - Jr).max() <= 1e-13
    rs = SZ.nr_method(ref, y0, SZ.Opt(ite_tol=1e-10))
    ds = SB.nr_method(dev, y0.array, SB.Opt(ite_tol=1e-10))
    assert bool(rs.stats.succeed) and bool(np.all(ds.stats.succeed))
    np.testing.assert_allclose(np.asarray(ds.y).ravel(), rs.y.array, rtol=0, atol=1e-12)
    assert int(np.asarray(ds.stats.nstep).ravel()[0]) == int(rs.stats.nstep)
