"""`solverz_b200.adapters.from_solverz`: the SAME Python model object, built with `from Solverz import ...`, goes
through the reference's own numerical back-end and through this package's lowering + Jacobian derivation + CUDA
printer (checked on the CPU by the host-check build of the generated source).  F at 1e-12, J as matrices at 1e-12
relative (the reference's three-way parity bar, tests/test_matrix_equation.py:111-158).

Needs the private reference copy baseline/_ref (baseline/make_ref.sh; it travels to the GPU box with the snapshot);
skipped when it is not there.  The GPU twin of these tests is tests/test_gpu_adapter.py."""
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
if not (REF / "Solverz" / "__init__.py").exists():
    pytest.skip("baseline/_ref (private copy of the reference) is not present", allow_module_level=True)
if str(REF) not in sys.path:
    sys.path.insert(0, str(REF))

from host_check import host_eval, host_lib  # noqa: E402


def _SZ():
    import Solverz
    return Solverz


def _compare(ref_mdl, mine, name, tmp_path, points, params=None, hist=None, rtol=1e-12):
    ir, pat, lib, _ = host_lib(mine, name, tmp_path)
    for t, y in points:
        args = (y, ref_mdl.p) if t is None else (t, y, ref_mdl.p)
        if hist is not None:
            args = args + (hist,)
        Fr = np.asarray(ref_mdl.F(*args)).copy()
        Jr = sps.csc_matrix(ref_mdl.J(*args))
        F, JV = host_eval(lib, ir, pat, y, params or {}, 0.0 if t is None else t)
        J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(ir.n, ir.n))
        assert np.abs(F[0] - Fr).max() <= rtol * max(1.0, np.abs(Fr).max())
        assert abs(J - Jr).max() <= rtol * max(1.0, abs(Jr).max())
    return ir, pat


def test_polar_power_flow_loop_eqn(tmp_path):
    """LoopEqn bodies over IndexedBase / Idx / Sum with subset-Set gathers (the configs[1] model, 6 x 6)."""
    SZ = _SZ()
    from solverz_b200.adapters import from_solverz
    from solverz_b200.synthetic import lattice_grid
    nx = 6
    nb, Y = lattice_grid(nx, nx)
    npv = max(1, int(0.1 * nb))
    pv, pq = list(range(1, 1 + npv)), list(range(1 + npv, nb))
    rng = np.random.default_rng(1)
    m = SZ.Model()
    m.Va = SZ.Var('Va', np.zeros(nb))
    m.Vm = SZ.Var('Vm', np.ones(nb))
    m.G = SZ.Param('G', sps.csc_array(Y.real), dim=2, sparse=True)
    m.B = SZ.Param('B', sps.csc_array(Y.imag), dim=2, sparse=True)
    m.Pinj = SZ.Param('Pinj', rng.uniform(-.05, .05, nb))
    m.Qinj = SZ.Param('Qinj', rng.uniform(-.02, .02, nb))
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
    # the reference evaluates LoopEqn only through its rendered numba module (equations.py:925-938)
    import importlib
    import os
    os.environ.setdefault("NUMBA_CACHE_DIR", str(tmp_path / "numba"))
    SZ.module_printer(spf, y0, "ad_pf_ref", directory=str(tmp_path), jit=True).render()
    sys.path.insert(0, str(tmp_path))
    ref = importlib.import_module("ad_pf_ref").mdl
    mine, y = from_solverz(spf, y0)
    assert mine.kind == "AE" and np.array_equal(y.array, y0.array)
    yp = y0.array + 0.05 * np.random.default_rng(7).standard_normal(2 * nb)
    ir, pat = _compare(ref, mine, "ad_pf", tmp_path, [(None, y0.array), (None, yp)])
    Jr = sps.csc_matrix(ref.J(yp, ref.p))
    assert np.array_equal(pat.indices, Jr.indices) and np.array_equal(pat.indptr, Jr.indptr)   # same CSC slots


def test_swing_dae_mat_mul_time_series(tmp_path):
    """Vector Ode / Eqn, Mat_Mul with frozen sparse matrices, TimeSeriesParam, a DAE's equation order and mass."""
    SZ = _SZ()
    from solverz_b200.adapters import from_solverz
    from solverz_b200.synthetic import swing_dae_model
    m, info = swing_dae_model(3, 3, ns=SZ)
    dae, y0 = m.create_instance()
    ref = SZ.made_numerical(dae, y0, sparse=True)
    mine, y = from_solverz(dae, y0)
    assert mine.kind == "DAE"
    yp = y0.array + 0.01 * np.random.default_rng(0).standard_normal(len(y0.array))
    ir, pat = _compare(ref, mine, "ad_swing", tmp_path, [(0.0, y0.array), (0.15, yp)])
    assert abs(sps.csc_matrix(mine.M) - sps.csc_matrix(ref.M)).max() == 0.0


def test_dhs_fdae_alias_var_mutable_matrix(tmp_path):
    """FDAE with an AliasVar history node, x * Mat_Mul(A, x)-type mutable-matrix Jacobians, Abs, a sliced variable."""
    SZ = _SZ()
    from solverz_b200.adapters import from_solverz
    from solverz_b200.synthetic import dhs_fdae_model
    m, info = dhs_fdae_model(3, 3, ns=SZ)
    f, y0 = m.create_instance()
    ref = SZ.made_numerical(f, y0, sparse=True)
    mine, y = from_solverz(f, y0)
    assert mine.kind == "FDAE" and mine.nstep == 1
    ir = mine.to_ir()
    yp = y0.array * (1 + 0.01 * np.random.default_rng(0).standard_normal(len(y0.array)))
    o, ln = ir.var_off["Tout"]
    hist = {"Tout_tag_0": np.atleast_2d(y0.array[o:o + ln])}
    _compare(ref, mine, "ad_dhs", tmp_path, [(0.0, y0.array), (150.0, yp)], params=hist, hist=y0.array)


@pytest.mark.parametrize("form", ["nested", "neg_outer", "three_arg", "diag_left", "diag_both", "scaled"])
def test_mat_mul_chains(tmp_path, form):
    """Nested / multi-argument Mat_Mul of frozen matrices (reference tests/test_matrix_equation.py:482-587, 648-680)
    and Diag factors around one matrix."""
    SZ = _SZ()
    from Solverz.sym_algebra.functions import Diag
    from solverz_b200.adapters import from_solverz
    Mat_Mul = SZ.Mat_Mul
    m = SZ.Model()
    m.x = SZ.Var('x', [0.5, 0.7, -0.3])
    m.b = SZ.Param('b', [3.0, 4.0, 1.0])
    m.c = SZ.Param('c', [1.5, -0.5, 2.0])
    m.A = SZ.Param('A', [[1.0, 0.5, 0.0], [0.5, 1.0, 0.25], [0.0, 0.3, 2.0]], dim=2, sparse=True)
    m.B = SZ.Param('B', [[2.0, 0.0, 1.0], [0.0, 2.0, 0.0], [0.5, 0.0, 1.0]], dim=2, sparse=True)
    body = {
        "nested": lambda: Mat_Mul(m.A, Mat_Mul(m.B, m.x)) - m.b,
        "neg_outer": lambda: Mat_Mul(-m.A, Mat_Mul(m.B, m.x)) - m.b,
        "three_arg": lambda: Mat_Mul(m.A, m.B, m.x) - m.b,
        "diag_left": lambda: Mat_Mul(Diag(m.x), m.A, m.x) - m.b,
        "diag_both": lambda: Mat_Mul(Diag(m.x), m.A, Diag(m.c), m.x) - m.b,
        "scaled": lambda: m.x * Mat_Mul(2 * m.A, m.x) + Mat_Mul(m.B, m.c * m.x) - m.b,
    }[form]()
    m.eqn = SZ.Eqn('f', body)
    spf, y0 = m.create_instance()
    ref = SZ.made_numerical(spf, y0, sparse=True)
    mine, y = from_solverz(spf, y0)
    yp = y0.array + 0.3 * np.random.default_rng(5).standard_normal(3)
    _compare(ref, mine, f"ad_mm_{form}", tmp_path, [(None, y0.array), (None, yp)])


def test_control_vocabulary(tmp_path):
    """Saturation / In / Min / heaviside / Sign / AntiWindUp / switch / logic through the adapter."""
    SZ = _SZ()
    import Solverz.sym_algebra.functions as RF
    from solverz_b200.adapters import from_solverz
    m = SZ.Model()
    m.x = SZ.Var('x', [0.3, -1.2, 2.5, 0.8])
    m.z = SZ.Var('z', [0.1, 0.2, 0.3, 0.4])
    m.lo = SZ.Param('lo', [-1.0, -1.0, -1.0, -1.0])
    m.hi = SZ.Param('hi', [1.0, 1.0, 1.0, 1.0])
    m.e1 = SZ.Eqn('e1', RF.Saturation(m.x, m.lo, m.hi) + RF.heaviside(m.x) * m.z - RF.Min(m.x, m.z) + RF.Sign(m.x))
    m.e2 = SZ.Eqn('e2', m.z - RF.In(m.x, m.lo, m.hi) * RF.exp(-m.x * m.x) - RF.Abs(m.x) * RF.ln(m.z + 2)
                  + RF.atan2(m.x, m.z + 1))
    spf, y0 = m.create_instance()
    ref = SZ.made_numerical(spf, y0, sparse=True)
    mine, y = from_solverz(spf, y0)
    yp = y0.array + 0.2 * np.random.default_rng(3).standard_normal(8)
    _compare(ref, mine, "ad_ctrl", tmp_path, [(None, y0.array), (None, yp)])


def test_module_printer_accepts_reference_objects(tmp_path):
    """module_printer / made_numerical of this package take the reference's objects directly (lang='cuda'): the
    rendered package carries the lowered pattern; only the nvcc build + load need no GPU."""
    SZ = _SZ()
    from solverz_b200 import module_printer
    m = SZ.Model()
    m.x = SZ.Var('x', [0.0, 0.0])
    m.b = SZ.Param('b', [0.5, 1.0])
    m.A = SZ.Param('A', [[1, 3], [-1, 2]], dim=2, sparse=True)
    m.eqnf = SZ.Eqn('eqnf', m.b - SZ.Mat_Mul(m.A, m.x))
    spf, y0 = m.create_instance()
    pkg = module_printer(spf, y0, "ad_pkg", lang='cuda', directory=str(tmp_path)).render()
    import pickle
    st = pickle.load(open(Path(pkg) / "param_and_setting.pkl", "rb"))
    assert st["n"] == 2 and st["nnz_j"] == 4 and (Path(pkg) / st["so"]).exists()


def test_triggerable_parameters(tmp_path):
    """Param(triggerable=True, trigger_var, trigger_fun) (equation/param.py:16-59): the reference's generated F_ / J_
    overwrite the parameter with trigger_fun(variables) at every call and differentiate as if it were a constant
    (module test expectations: code_printer/python/module/test/test_module_printer.py:20-60).  The adapter writes
    it as frozen(f(vars)): same F, same J."""
    SZ = _SZ()
    from solverz_b200.adapters import from_solverz, frozen
    import sympy as sp
    m = SZ.Model()
    m.x = SZ.Var('x', [0.3, -0.7, 1.1])
    m.y = SZ.Var('y', [1.0, 2.0, -0.5])
    m.ax = SZ.Param('ax', [0.0, 0.0, 0.0], triggerable=True, trigger_var=['x'], trigger_fun=np.sin)
    m.bq = SZ.Param('bq', [0.0, 0.0, 0.0], triggerable=True, trigger_var=['x', 'y'],
                    trigger_fun=lambda x, y: 2 * x * y + 1)
    m.c = SZ.Param('c', [0.5, 1.5, 2.5])
    m.e1 = SZ.Eqn('e1', m.ax * m.y + m.x ** 2 - m.c)
    m.e2 = SZ.Eqn('e2', m.y * m.bq - SZ.exp(m.x) * m.ax + m.c * m.y)
    spf, y0 = m.create_instance()
    ref = SZ.made_numerical(spf, y0, sparse=True)
    mine, y = from_solverz(spf, y0)
    assert "ax" not in mine.PARAM and "bq" not in mine.PARAM and "c" in mine.PARAM
    x = sp.Symbol('x')
    assert sp.diff(frozen(sp.sin(x)) * x, x) == frozen(sp.sin(x))
    rng = np.random.default_rng(8)
    pts = [(None, y0.array)] + [(None, y0.array + 0.3 * rng.standard_normal(6)) for _ in range(3)]
    _compare(ref, mine, "ad_trig", tmp_path, pts)


def test_rejections():
    SZ = _SZ()
    from solverz_b200.adapters import from_solverz
    m = SZ.Model()
    m.x = SZ.Var('x', [1.0, 2.0])
    m.p = SZ.Param('p', [1.0, 1.0], triggerable=True, trigger_var='x', trigger_fun=lambda x: np.where(x > 0, x, 0.0))
    m.e = SZ.Eqn('e', m.x - m.p)
    spf, y0 = m.create_instance()
    with pytest.raises(NotImplementedError, match="opaque host callback|not an expression"):
        from_solverz(spf, y0)
    m2 = SZ.Model()
    m2.x = SZ.Var('x', [1.0, 2.0])
    m2.p = SZ.Param('p', [1.0, 1.0], triggerable=True, trigger_var='x', trigger_fun=np.sin)
    m2.e = SZ.Eqn('e', m2.x - m2.p[0:1] * 2)
    spf2, y02 = m2.create_instance()
    with pytest.raises(NotImplementedError, match="used with an index"):
        from_solverz(spf2, y02)
    with pytest.raises(TypeError):
        from_solverz(object(), None)
