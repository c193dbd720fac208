"""CPU tests of the host-side logic: symbolic LU analysis (walked by a numpy emulator of the kernels'
schedule), front-end lowering / Jacobian derivation / CUDA printer (through the -DSZBM_HOST_CHECK build
of the generated source), the C-ABI library, and the multi-rank sharding plan (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as sla

ROOT = Path(__file__).resolve().parents[1]
GOLD = Path(__file__).parent / "golden"


# ----------------------------------------------------------------------------- symbolic LU
def _pf_jacobian(nx, seed=0):
    from oracle.pf import PFOracle
    from solverz_b200.synthetic import lattice_grid
    nb, Y = lattice_grid(nx, nx)
    npv = max(1, int(0.1 * nb))
    orc = PFOracle(Y, np.arange(1, 1 + npv), np.arange(1 + npv, nb))
    rng = np.random.default_rng(seed)
    y = np.concatenate([0.05 * rng.standard_normal(nb), 1 + 0.05 * rng.standard_normal(nb)])
    J = sps.csc_matrix(orc.J(y))
    J.sort_indices()
    return J, orc


@pytest.mark.parametrize("nx,opts", [(6, {}), (12, {}), (12, dict(warp_front_max=8, group_front_max=20, n_warps=4, n_groups=2)),
                                     (10, dict(orderings=("nd",))), (10, dict(orderings=("natural",)))])
def test_symbolic_schedule_solves(nx, opts):
    import mf_emulate as E
    from solverz_b200.lu.symbolic import analyze
    J, orc = _pf_jacobian(nx)
    sym = analyze(J, **opts)
    assert sorted(sym.order.tolist()) == list(range(sym.ns))
    assert sorted(sym.rowp.tolist()) == list(range(sym.n)) and sorted(sym.colp.tolist()) == list(range(sym.n))
    rng = np.random.default_rng(1)
    b = rng.standard_normal(sym.n)
    # a *different* scenario with the same pattern: the static pivots must still work
    J2, _ = _pf_jacobian(nx, seed=5)
    assert np.array_equal(J.indices, J2.indices)
    for A in (J, J2):
        lu = E.factor(sym, A.data)
        x = E.solve(sym, lu, b)
        xr = sla.spsolve(A, b)
        assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()


@pytest.mark.parametrize("nx,opts", [(8, {}), (12, dict(warp_front_max=8, group_front_max=20, n_warps=4, n_groups=2))])
def test_device_tables_walk(nx, opts):
    """The tables the kernels read (front headers, O-block gather entries, child records + rel vectors) walked in
    numpy exactly as szb_lu.cu does: same factors as the plain supernode walk, and the fused right-hand side
    column equals the forward substitution."""
    import mf_emulate as E
    from solverz_b200.lu.schedule import build_tables, plan_threads
    from solverz_b200.lu.symbolic import analyze
    J, _ = _pf_jacobian(nx)
    sym = analyze(J, **opts)
    plan = plan_threads(sym)
    tabs = build_tables(sym, plan["threads"], plan["wpg"])
    rng = np.random.default_rng(2)
    b = rng.standard_normal(sym.n)
    lu_t, y = E.factor_tables(sym, tabs, J.data, rhs=b)
    lu = E.factor(sym, J.data)
    assert np.abs(lu_t - lu).max() <= 1e-11 * np.abs(lu).max()
    x = E.solve(sym, lu_t, b)
    xr = sla.spsolve(J, b)
    assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()
    # y = L^{-1} P b: finishing with the backward sweep alone must give the same x
    lu_n, _ = E.factor_tables(sym, tabs, J.data, rhs=None)
    assert np.abs(lu_n - lu).max() <= 1e-11 * np.abs(lu).max()


def test_symbolic_shifted_matrix_and_singular_pattern():
    import mf_emulate as E
    from solverz_b200.lu.symbolic import analyze
    J, _ = _pf_jacobian(6)
    n = J.shape[0]
    mv = np.zeros(J.nnz)
    d = sps.csc_matrix((np.ones(n // 2), (np.arange(n // 2), np.arange(n // 2))), shape=(n, n))
    U = (abs(J) + d).tocsc()
    U.sort_indices()
    sym = analyze(U)
    # M - h*gamma*J assembled as mvals + scale*jv, both given on the union pattern's CSC slots
    def on_pattern(A):
        A = sps.coo_matrix(A)
        slot = {}
        for c in range(n):
            for p in range(U.indptr[c], U.indptr[c + 1]):
                slot[(U.indices[p], c)] = p
        out = np.zeros(U.nnz)
        for r, c, v in zip(A.row, A.col, A.data):
            out[slot[(r, c)]] += v
        return out
    jd, md = on_pattern(J), on_pattern(d)
    lu = E.factor(sym, jd, mvals=md, scale=-0.01)
    b = np.arange(n, dtype=float)
    x = E.solve(sym, lu, b)
    A = (d - 0.01 * J).tocsc()
    assert np.abs(A @ x - b).max() <= 1e-11 * abs(A).max() * np.abs(x).max()   # backward stable
    with pytest.raises(ValueError):
        analyze(sps.csc_matrix(np.array([[1.0, 1.0], [0.0, 0.0]])))


# ----------------------------------------------------------------------------- printer via host check
from host_check import host_lib as _host_lib, host_eval as _eval  # noqa: E402


def test_printer_pf_matches_reference_golden(tmp_path):
    """Generated row functions vs the reference's numba module output (golden) at 1e-12 -- the
    three-way parity of the reference's tests/test_matrix_equation.py:111-158."""
    from solverz_b200.synthetic import polar_pf_model
    g = np.load(GOLD / "pf_ref_8x8.npz")
    eqs, y0, info = polar_pf_model(8, 8)
    ir, pat, lib, src = _host_lib(eqs, "pf8", tmp_path)
    assert np.array_equal(pat.indices, g["J_indices"]) and np.array_equal(pat.indptr, g["J_indptr"])
    for y, F, Jd in ((g["y0"], g["F0"], g["J0_data"]), (g["yp"], g["Fp"], g["Jp_data"])):
        Fh, JVh = _eval(lib, ir, pat, y, {})
        assert np.abs(Fh[0] - F).max() <= 1e-13
        assert np.abs(JVh[0] - Jd).max() <= 1e-12 * np.abs(Jd).max()
    assert "sincos(" in src and "szbm_launch_F" in src
    # trig hoisting: the F rows read the per-scenario sin / cos table of Va, the Jacobian rows do not
    f0 = src[src.index("szbm_F0(const int i"):src.index("szbm_F1(const int i")]
    j0 = src[src.index("szbm_J0(const int i"):src.index("szbm_J1(const int i")]
    assert "c.ts_Va[j]" in f0 and "c.tc_Va[j]" in f0 and f0.count("sincos(") == 1
    assert "c.ts_Va" not in j0 and "sincos(" in j0
    assert "#define SZBM_TRIG_N 64" in src


def test_printer_vector_eqn_matmul_and_known_values(tmp_path):
    """F0 = [3, 1 + sin 1], J0 = [[2, 1], [2, cos 1]] (reference test_module_generator.py:64-127);
    x * (A @ x) - b: J0 = [[2.5, 0.5], [0.5, 3.5]] (reference golden known_answers.npz)."""
    from solverz_b200 import Eqn, Mat_Mul, Model, Param, Var, sin
    m = Model()
    m.x = Var('x', [1, 1])
    m.f1 = Eqn('f1', 2 * m.x[0] + m.x[1])
    m.f2 = Eqn('f2', m.x[0] ** 2 + sin(m.x[1]))
    eqs, y0 = m.create_instance()
    ir, pat, lib, _ = _host_lib(eqs, "ae1", tmp_path)
    F, JV = _eval(lib, ir, pat, y0.array, {})
    np.testing.assert_allclose(F[0], [3, 1 + np.sin(1)], rtol=1e-14)
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(2, 2)).toarray()
    np.testing.assert_allclose(J, [[2, 1], [2, np.cos(1)]], rtol=1e-14)

    k = np.load(GOLD / "known_answers.npz")
    m = Model()
    m.x = Var('x', [0.5, 0.5])
    m.A = Param('A', [[2, 1], [1, 3]], dim=2, sparse=True)
    m.b = Param('b', [4.0, 5.0])
    m.eqn = Eqn('f', m.x * Mat_Mul(m.A, m.x) - m.b)
    eqs, y0 = m.create_instance()
    ir, pat, lib, _ = _host_lib(eqs, "mm1", tmp_path)
    F, JV = _eval(lib, ir, pat, y0.array, {})
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(2, 2)).toarray()
    np.testing.assert_allclose(J, k["mutable_J0"], rtol=1e-14)
    np.testing.assert_allclose(F[0], 0.5 * (np.array([[2, 1], [1, 3.]]) @ [0.5, 0.5]) - [4, 5], rtol=1e-14)


def test_printer_dae_time_series_and_functions(tmp_path):
    """DAE ordering (Ode rows first, equations.py:672), mass matrix, TimeSeriesParam end clamp
    (param.py:187-212), Abs / heaviside / exp / ln vocabulary; J against complex-step-free central
    differences of the generated F."""
    from solverz_b200 import Abs, Eqn, Model, Ode, Param, TimeSeriesParam, Var, exp, heaviside, ln, t_sym, cos
    m = Model()
    m.u = Var('u', [0.3, -0.2, 0.8])
    m.w = Var('w', [1.5])
    m.c = Param('c', [2.0, 3.0, 4.0], per_scenario=True)
    m.load = TimeSeriesParam('load', v_series=[1.0, 2.0, 0.5], time_series=[0.0, 1.0, 3.0])
    m.alg = Eqn('alg', m.w * m.w - m.load - exp(-m.w) + ln(m.w + 2) * heaviside(m.w))
    m.ode = Ode('ode', -m.c * m.u + Abs(m.u) * cos(t_sym) + m.w, diff_var=m.u)
    eqs, y0 = m.create_instance()
    assert eqs.kind == "DAE"
    ir, pat, lib, _ = _host_lib(eqs, "dae1", tmp_path)
    assert [e.name for e in ir.eqs] == ["ode", "alg"]          # Ode first
    r, c = ir.mass
    assert r.tolist() == [0, 1, 2] and c.tolist() == [0, 1, 2]
    cs = np.array([[2.0, 3.0, 4.0], [1.0, 1.5, 2.5]])
    Y = np.array([[0.3, -0.2, 0.8, 1.5], [0.1, 0.4, -0.6, 0.7]])
    for t, load in ((0.5, 1.5), (2.0, 1.25), (3.0, 0.5), (10.0, 0.5)):
        F, JV = _eval(lib, ir, pat, Y, {"c": cs}, t)
        for b in range(2):
            u, w = Y[b, :3], Y[b, 3]
            ref = np.concatenate([-cs[b] * u + np.abs(u) * np.cos(t) + w,
                                  [w * w - load - np.exp(-w) + np.log(w + 2) * (w >= 0)]])
            np.testing.assert_allclose(F[b], ref, rtol=1e-14, atol=1e-15)
            J = sps.csc_matrix((JV[b], pat.indices, pat.indptr), shape=(4, 4)).toarray()
            Jfd = np.zeros((4, 4))
            for j in range(4):
                h = 1e-6
                yp, ym = Y.copy(), Y.copy()
                yp[b, j] += h
                ym[b, j] -= h
                Jfd[:, j] = (_eval(lib, ir, pat, yp, {"c": cs}, t)[0][b] - _eval(lib, ir, pat, ym, {"c": cs}, t)[0][b]) / (2 * h)
            np.testing.assert_allclose(J, Jfd, atol=1e-8)


def test_printer_control_vocabulary_and_custom_function(tmp_path):
    """Saturation / AntiWindUp / switch / Min / atan2 / Sign (functions.py:307-606) against numpy, and a
    user-defined function through the ``_cudacode`` + ``_cuda_device_code`` hook -- the CUDA analogue of
    the reference's custom-function recipe (docs/src/advanced.md:1-80, num_api/test/test_udm.py): symbolic
    derivatives from ``fdiff``, numerical code supplied by the function class."""
    import sympy as sp
    from solverz_b200 import AntiWindUp, Eqn, Model, Param, Saturation, Sign, Var, atan2, switch

    class SoftMin(sp.Function):           # the documentation's Min example, smooth variant to exercise fdiff
        nargs = 2
        is_real = True
        _cuda_device_code = """
SZB_HD static inline double user_softmin(double x, double y) { return x <= y ? x : y; }
SZB_HD static inline double user_dsoftmin(double x, double y, int k) { return (k == 0) == (x <= y) ? 1.0 : 0.0; }
"""

        def fdiff(self, argindex=1):
            return DSoftMin(self.args[0], self.args[1], sp.Integer(argindex - 1))

        def _cudacode(self, printer):
            return f"user_softmin({printer._print(self.args[0])}, {printer._print(self.args[1])})"

    class DSoftMin(sp.Function):
        nargs = 3
        is_real = True
        _cuda_device_code = SoftMin._cuda_device_code

        def fdiff(self, argindex=1):
            return sp.Integer(0)

        def _cudacode(self, printer):
            return (f"user_dsoftmin({printer._print(self.args[0])}, {printer._print(self.args[1])}, "
                    f"{int(self.args[2])})")

    m = Model()
    m.u = Var('u', [0.3, 1.7, -0.4])
    m.e = Var('e', [0.5, 0.2, -0.1])
    m.flag = Param('flag', [0.0, 1.0, 2.0])
    m.f1 = Eqn('f1', Saturation(m.u, -0.2, 1.0) + AntiWindUp(m.u, -0.2, 1.0, m.e))
    m.f2 = Eqn('f2', switch(m.u ** 2, 3 * m.e, m.flag) + atan2(m.u, m.e) + Sign(m.u) * SoftMin(m.u * m.u, m.e))
    eqs, y0 = m.create_instance()
    ir, pat, lib, src = _host_lib(eqs, "ctl", tmp_path)
    assert src.count("double user_softmin(") == 1
    for Y in (y0.array, np.array([1.5, -0.5, 0.1, 0.4, -0.3, 0.2]), np.array([-0.9, 0.2, 2.0, -0.6, 0.8, 0.05])):
        F, JV = _eval(lib, ir, pat, Y, {})
        u, e = Y[:3], Y[3:]
        flag = np.array([0.0, 1.0, 2.0])
        gate = ~(((u > 1.0) & (e > 0)) | ((u < -0.2) & (e < 0)))
        f1 = np.clip(u, -0.2, 1.0) + e * gate
        f2 = np.select([flag == 0, flag == 1], [u ** 2, 3 * e], 0) + np.arctan2(u, e) + np.sign(u) * np.minimum(u * u, e)
        np.testing.assert_allclose(F[0], np.concatenate([f1, f2]), rtol=1e-14, atol=1e-15)
        J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(6, 6)).toarray()
        r2 = u * u + e * e
        d1u = ((u >= -0.2) & (u <= 1.0)).astype(float)
        d2u = np.select([flag == 0, flag == 1], [2 * u, 0 * u], 0) + e / r2 + np.sign(u) * np.where(u * u <= e, 2 * u, 0)
        d2e = np.select([flag == 0, flag == 1], [0 * u, 3 + 0 * u], 0) - u / r2 + np.sign(u) * np.where(u * u <= e, 0, 1.0)
        Jref = np.block([[np.diag(d1u), np.diag(gate.astype(float))], [np.diag(d2u), np.diag(d2e)]])
        np.testing.assert_allclose(J, Jref, rtol=1e-13, atol=1e-15)


def test_printer_swing_dae_matches_reference_golden(tmp_path):
    """Config #4 in miniature (swing / flux decay / exciter per bus, Mat_Mul network, fault through a
    TimeSeriesParam): the generated F and J at (t = 0, y0) against the reference's rendered module."""
    from solverz_b200.synthetic import swing_dae_model, swing_fault_scenarios
    g = np.load(GOLD / "swing_ref_3x3.npz")
    m, info = swing_dae_model(3, 3)
    eqs, y0 = m.create_instance()
    assert eqs.kind == "DAE" and len(y0.array) == info["n"]
    np.testing.assert_allclose(y0.array, g["y0"], rtol=0, atol=1e-15)
    ir, pat, lib, _ = _host_lib(eqs, "swing9", tmp_path)
    F, JV = _eval(lib, ir, pat, y0.array, {})
    assert np.abs(F[0] - g["F0"]).max() <= 1e-12
    Jo = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(ir.n, ir.n))
    Jr = sps.csc_matrix((g["J0_data"], g["J0_indices"], g["J0_indptr"]), shape=(ir.n, ir.n))
    assert abs(Jo - Jr).max() <= 1e-12 * abs(Jr).max()
    # with a fault on, against central differences of the generated F (the reference's J is only pinned at t = 0)
    fmask, tst, tsv = swing_fault_scenarios(info["nb"], 2)
    prm = {"fmask": fmask[1:2], "gf__ts_t": tst[1:2], "gf__ts_v": tsv[1:2]}
    rng = np.random.default_rng(0)
    y = y0.array + 0.01 * rng.standard_normal(ir.n)
    t = 0.15
    F, JV = _eval(lib, ir, pat, y, prm, t)
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(ir.n, ir.n)).toarray()
    Jfd = np.zeros_like(J)
    for j in range(ir.n):
        e = np.zeros(ir.n)
        e[j] = 1e-6
        Jfd[:, j] = (_eval(lib, ir, pat, y + e, prm, t)[0][0] - _eval(lib, ir, pat, y - e, prm, t)[0][0]) / 2e-6
    np.testing.assert_allclose(J, Jfd, atol=2e-7 * max(1.0, np.abs(J).max()))
    assert J[8 * 9 + 1, 4 * 9 + 1] != 0.0       # d(Fx - gf fmask Ux)/dUx at the faulted bus


def test_printer_dhs_fdae_matches_reference_golden(tmp_path):
    """Config #5 in miniature (Mat_Mul hydraulics with the mutable-matrix Jacobian of L (K m |m|), node
    mixing Vin (m Tout) - Tn[1:] (Vin m), AliasVar history, TimeSeriesParam load profile): generated F and J
    against the reference's own evaluation of the same model (tests/golden/make_golden.py::golden_dhs)."""
    from solverz_b200.synthetic import dhs_fdae_model, dhs_load_profiles
    g = np.load(GOLD / "dhs_ref_3x3.npz")
    m, info = dhs_fdae_model(3, 3)
    eqs, y0 = m.create_instance()
    assert eqs.kind == "FDAE"
    ir, pat, lib, _ = _host_lib(eqs, "dhs3", tmp_path)
    assert ir.nstep == 1 and ir.n == info["n"]
    np.testing.assert_allclose(y0.array, g["y0"], rtol=0, atol=1e-13)
    tst, tsv = dhs_load_profiles(6, info["dt"])
    tout = slice(*ir.var_off["Tout"][:1], ir.var_off["Tout"][0] + ir.var_off["Tout"][1])
    hist = {"Tout_tag_0": np.atleast_2d(g["y0"][tout])}
    F, JV = _eval(lib, ir, pat, g["y0"], hist, 0.0)
    assert np.abs(F[0] - g["F0"]).max() <= 1e-12
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(ir.n, ir.n))
    Jr = sps.csc_matrix((g["J0_data"], g["J_indices"], g["J_indptr"]), shape=(ir.n, ir.n))
    assert abs(J - Jr).max() <= 1e-12 * abs(Jr).max()
    par = dict(hist)
    par["sc__ts_t"], par["sc__ts_v"] = np.atleast_2d(tst[1]), np.atleast_2d(tsv[1])
    F, JV = _eval(lib, ir, pat, g["yp"], par, 2.5 * info["dt"])
    assert np.abs(F[0] - g["Fp"]).max() <= 1e-11 * max(1.0, np.abs(g["Fp"]).max())
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(ir.n, ir.n)).toarray()
    assert np.abs(J - g["Jp_dense"]).max() <= 1e-12 * np.abs(g["Jp_dense"]).max()


def test_front_end_errors():
    from solverz_b200 import Eqn, LoopEqn, Model, Param, Set, Sum, Var
    m = Model()
    m.x = Var('x', [1.0, 2.0, 3.0])
    m.y = Var('y', [1.0, 2.0])
    m.e = Eqn('e', m.x + m.y)
    with pytest.raises(ValueError, match="incompatible operand lengths"):
        m.create_instance()
    with pytest.raises(ValueError):
        Set('Bad', [1, 1, 2])
    with pytest.raises(ValueError):
        Set('Neg', [-1, 2])
    with pytest.raises(TypeError):
        m.x[[0, 1]]
    m2 = Model()
    m2.x = Var('x', [1.0, 2.0])
    m2.e = Eqn('e', m2.x[0])
    with pytest.raises(ValueError, match="not equal"):
        m2.create_instance()


# ----------------------------------------------------------------------------- C-ABI
def test_c_abi_exports_every_declared_symbol(libszb):
    from solverz_b200.num_api import lib as L
    hdr = (ROOT / "include" / "szb.h").read_text()
    declared = set(re.findall(r"SZB_API\s+(?:const\s+)?\w+\*?\s+(szb_\w+)\(", hdr))
    assert len(declared) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (szb_\w+)", out))
    assert declared <= exported, declared - exported
    assert libszb.szb_abi_version() == L.SZB_ABI_VERSION == 4
    # error convention: NULL handle -> negative code, never a crash
    assert libszb.szb_set_batch(None, 4) < 0
    assert libszb.szb_destroy(None) == 0


def test_model_library_exports(tmp_path):
    """The nvcc-built model library exports the szbm_* entry points of include/szb_model.h."""
    from solverz_b200 import module_printer
    from solverz_b200.synthetic import polar_pf_model
    eqs, y0, _ = polar_pf_model(4, 4)
    pkg = module_printer(eqs, y0, "pf4_abi", directory=str(tmp_path)).render()
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg / "libpf4_abi.so")], capture_output=True, text=True).stdout
    hdr = (ROOT / "include" / "szb_model.h").read_text()
    declared = set(re.findall(r"SZBM_API\s+(?:const\s+)?\w+\*?\s+(szbm_\w+)\(", hdr))
    assert declared <= set(re.findall(r" T (szbm_\w+)", out))
    assert sorted(p.name for p in pkg.iterdir() if not p.name.startswith(".")) == \
        ["__init__.py", "dependency.py", "libpf4_abi.so", "num_func.cu", "param_and_setting.pkl"]
    sass = subprocess.run(["cuobjdump", "-lelf", str(pkg / "libpf4_abi.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from solverz_b200 import Eqn, Model, Var, made_numerical, nr_method
    from solverz_b200.num_api.lib import SzbError
    m = Model()
    m.x = Var('x', 2.0)
    m.e = Eqn('e', m.x ** 2 - 1)
    g, y0 = m.create_instance()
    mdl = made_numerical(g, y0, sparse=True)      # printing + nvcc work without a GPU
    with pytest.raises(SzbError, match="no CPU fallback"):
        nr_method(mdl, y0)
    # the other product entry points: time steppers, events, the linear-solver seam
    from solverz_b200 import (BatchedLU, ComponentEvents, Ode, Opt, Rodas, backward_euler, implicit_trapezoid,
                              lu_decomposition, solve)
    m = Model()
    m.x = Var('x', [0.0, 20.0])
    m.f1 = Ode('f1', m.x[1], m.x[0])
    m.f2 = Ode('f2', -9.8, m.x[1])
    dae, yd = m.create_instance()
    nd = made_numerical(dae, yd, sparse=True)
    for call in (lambda: Rodas(nd, [0, 1], yd),
                 lambda: Rodas(nd, [0, 1], yd, Opt(event=ComponentEvents([0], 0.0, -1, 1))),
                 lambda: Rodas(nd, [0, 1], yd, Opt(event=lambda t, y: (np.array([y[0]]), np.array([1]), np.array([-1])))),
                 lambda: backward_euler(nd, [0, 1], yd, Opt(step_size=0.1)),
                 lambda: implicit_trapezoid(nd, [0, 1], yd, Opt(step_size=0.1)),
                 lambda: BatchedLU(sps.identity(4, format="csc")),
                 lambda: lu_decomposition(sps.identity(4, format="csc")),
                 lambda: solve(sps.identity(4, format="csc"), np.ones(4))):
        with pytest.raises(SzbError, match="no CPU fallback"):
            call()
    # the Hessian-vector product and sicnm
    from solverz_b200 import sicnm
    mh = made_numerical(g, y0, sparse=True, make_hvp=True)
    for call in (lambda: sicnm(mh, y0), lambda: mh.HVP(y0.array, mh.p, np.ones(1))):
        with pytest.raises(SzbError, match="no CPU fallback"):
            call()


def test_component_events_is_a_reference_style_callback():
    """ComponentEvents doubles as the reference's events(t, y) callback (rodas.py:102-103, 203)."""
    from solverz_b200 import ComponentEvents
    ev = ComponentEvents([0, 2, 2], [0.0, 1.5, -1.0], [1, -1, 0], [0, 0, 1])
    value, isterminal, direction = ev(0.3, np.array([0.25, 9.0, 1.0]))
    np.testing.assert_allclose(value, [0.25, -0.5, 2.0])
    assert isterminal.tolist() == [0, 0, 1] and direction.tolist() == [1, -1, 0]
    with pytest.raises(ValueError):
        ComponentEvents(list(range(9)))


# ----------------------------------------------------------------------------- multi-rank plan
def test_shard_plan_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from solverz_b200.parallel import shard_range, gather_results, gather_results_async
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
B, n = 11, 3
lo, hi = shard_range(B, r, w)
y = torch.arange(lo, hi, dtype=torch.float64)[:, None].repeat(1, n) + 0.5
it = torch.arange(lo, hi, dtype=torch.int32) %% 4
fl = torch.zeros(hi - lo, dtype=torch.int32)
Y, IT, FL = gather_results(y, it, fl, B)
assert Y.shape == (B, n) and torch.equal(Y[:, 0], torch.arange(B, dtype=torch.float64) + 0.5)
assert torch.equal(IT, torch.arange(B, dtype=torch.int32) %% 4) and int(FL.sum()) == 0
# the asynchronous form: the shard buffers may be overwritten as soon as the gather is issued
pend = gather_results_async(y, it, fl, B)
y.fill_(-1.0); it.fill_(-1)
Y2, IT2, FL2 = pend.wait()
assert torch.equal(Y2, Y) and torch.equal(IT2, IT) and torch.equal(FL2, FL)
dist.destroy_process_group()
sys.stdout.write("rank " + str(r) + " ok " + str(lo) + " " + str(hi) + chr(10)); sys.stdout.flush()
''' % str(ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok 0 6" in r.stdout and "rank 1 ok 6 11" in r.stdout


def test_hvp_twin_model_matches_reference_golden(tmp_path):
    """SURVEY.md §8f item 3 on the host: the twin model of solverz_b200/hvp.py evaluates ``J(y) v`` as its residual
    and the reference's ``HVP(y, p, v)`` matrix (equation/hvp.py:12-74, golden from the reference's rendered
    module) as its Jacobian; on the LoopEqn power-flow model (which the reference's Hvp rejects) the same
    Jacobian agrees with a central difference of ``J v`` and stays inside the pattern of ``J``."""
    import scipy.sparse as sps
    import solverz_b200 as szb
    from solverz_b200.hvp import HvpEquations
    from solverz_b200.synthetic import polar_pf_model
    sys.path.insert(0, str(GOLD))
    from make_golden import sicnm_model
    g = np.load(GOLD / "sicnm_ref.npz")
    ae, y0 = sicnm_model(szb)
    ir, pat, lib, _ = _host_lib(ae, "sic", tmp_path)
    hir, hpat, hlib, _ = _host_lib(HvpEquations(ir), "sich", tmp_path)
    vp = {f"v__{k}": g["v"][o:o + s] for k, (o, s) in ir.var_off.items()}
    G, HV = _eval(hlib, hir, hpat, g["yp"], vp)
    H = sps.csc_matrix((HV[0], hpat.indices, hpat.indptr), shape=(ir.n, ir.n)).toarray()
    np.testing.assert_allclose(H, g["H"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(G[0], g["J"] @ g["v"], rtol=0, atol=1e-13)

    eqs, y0, info = polar_pf_model(4, 4)
    ir, pat, lib, _ = _host_lib(eqs, "pf4", tmp_path)
    hir, hpat, hlib, _ = _host_lib(HvpEquations(ir), "pf4h", tmp_path)
    rng = np.random.default_rng(0)
    y = y0.array + 0.05 * rng.normal(size=ir.n)
    v = rng.normal(size=ir.n)
    vp = {f"v__{k}": v[o:o + s] for k, (o, s) in ir.var_off.items()}

    def Jv(yy):
        return sps.csc_matrix((_eval(lib, ir, pat, yy, {})[1][0], pat.indices, pat.indptr), shape=(ir.n, ir.n)) @ v
    G, HV = _eval(hlib, hir, hpat, y, vp)
    assert np.abs(G[0] - Jv(y)).max() < 1e-12
    H = sps.csc_matrix((HV[0], hpat.indices, hpat.indptr), shape=(ir.n, ir.n)).toarray()
    e = 1e-6
    for k in range(ir.n):
        d = np.zeros(ir.n)
        d[k] = e
        assert np.abs(H[:, k] - (Jv(y + d) - Jv(y - d)) / (2 * e)).max() < 1e-6
    hp = sps.csc_matrix((np.ones(hpat.nnz), hpat.indices, hpat.indptr), shape=(ir.n, ir.n))
    jp = sps.csc_matrix((np.ones(pat.nnz), pat.indices, pat.indptr), shape=(ir.n, ir.n))
    assert (hp - hp.multiply(jp)).nnz == 0
    m = szb.Model()
    m.h = szb.Var('h', 0)
    m.f1 = szb.Ode('f1', f=-m.h, diff_var=m.h)
    with pytest.raises(NotImplementedError):
        HvpEquations(m.create_instance()[0].to_ir())
