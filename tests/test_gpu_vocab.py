"""GPU twins of host-only checks: the control vocabulary (Saturation / AntiWindUp / switch /
heaviside / In / Min / atan2 / Sign, functions.py:307-606) and a user function with the _cudacode hook compiled by
NVCC and evaluated by the device kernels, and the failure exits of the batched Rodas (dt underflow, more than 100
consecutive rejections: rodas.py:127-135) as per-scenario flags."""
import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu


def test_control_vocabulary_on_the_device(libszb):
    import sympy as sp
    from solverz_b200 import (AntiWindUp, Eqn, In, Min, Model, Param, Saturation, Sign, Var, atan2, heaviside,
                              made_numerical, switch)

    class SoftMin(sp.Function):
        nargs = 2
        is_real = True
        _cuda_device_code = """
SZB_HD static inline double user_softmin2(double x, double y) { return x <= y ? x : y; }
SZB_HD static inline double user_dsoftmin2(double x, double y, int k) { return (k == 0) == (x <= y) ? 1.0 : 0.0; }
"""

        def fdiff(self, argindex=1):
            return DSoftMin(self.args[0], self.args[1], sp.Integer(argindex - 1))

        def _cudacode(self, printer):
            return f"user_softmin2({printer._print(self.args[0])}, {printer._print(self.args[1])})"

    class DSoftMin(sp.Function):
        nargs = 3
        is_real = True
        _cuda_device_code = SoftMin._cuda_device_code

        def fdiff(self, argindex=1):
            return sp.Integer(0)

        def _cudacode(self, printer):
            return (f"user_dsoftmin2({printer._print(self.args[0])}, {printer._print(self.args[1])}, "
                    f"{int(self.args[2])})")

    m = Model()
    m.u = Var('u', [0.3, 1.7, -0.4])
    m.e = Var('e', [0.5, 0.2, -0.1])
    m.flag = Param('flag', [0.0, 1.0, 2.0])
    m.f1 = Eqn('f1', Saturation(m.u, -0.2, 1.0) + AntiWindUp(m.u, -0.2, 1.0, m.e) + heaviside(m.e) * m.u
               - In(m.u, -0.2, 1.0) * m.e)
    m.f2 = Eqn('f2', switch(m.u ** 2, 3 * m.e, m.flag) + atan2(m.u, m.e) + Sign(m.u) * SoftMin(m.u * m.u, m.e)
               + Min(m.u, m.e))
    eqs, y0 = m.create_instance()
    mdl = made_numerical(eqs, y0, sparse=True)
    pts = np.array([y0.array, [1.5, -0.5, 0.1, 0.4, -0.3, 0.2], [-0.9, 0.2, 2.0, -0.6, 0.8, 0.05]])
    s = mdl.session(len(pts))
    s.set_y(pts)
    Fd = s.eval_F().cpu().numpy()
    JV = s.eval_J().cpu().numpy()
    flag = np.array([0.0, 1.0, 2.0])
    for k, Y in enumerate(pts):
        u, e = Y[:3], Y[3:]
        gate = ~(((u > 1.0) & (e > 0)) | ((u < -0.2) & (e < 0)))
        inn = ((u >= -0.2) & (u <= 1.0)).astype(float)
        hv = (e >= 0).astype(float)
        f1 = np.clip(u, -0.2, 1.0) + e * gate + hv * u - inn * e
        f2 = (np.select([flag == 0, flag == 1], [u ** 2, 3 * e], 0) + np.arctan2(u, e)
              + np.sign(u) * np.minimum(u * u, e) + np.minimum(u, e))
        np.testing.assert_allclose(Fd[k], np.concatenate([f1, f2]), rtol=1e-14, atol=1e-15)
        J = sps.csc_matrix((JV[k], mdl.indices, mdl.indptr), shape=(6, 6)).toarray()
        r2 = u * u + e * e
        d1u = inn + hv
        d1e = gate.astype(float) - inn
        d2u = (np.select([flag == 0, flag == 1], [2 * u, 0 * u], 0) + e / r2
               + np.sign(u) * np.where(u * u <= e, 2 * u, 0) + (u <= e))
        d2e = (np.select([flag == 0, flag == 1], [0 * u, 3 + 0 * u], 0) - u / r2
               + np.sign(u) * np.where(u * u <= e, 0, 1.0) + (e < u))
        Jref = np.block([[np.diag(d1u), np.diag(d1e)], [np.diag(d2u), np.diag(d2e)]])
        np.testing.assert_allclose(J, Jref, rtol=1e-13, atol=1e-15)


def test_rodas_failure_flags(libszb):
    """rodas.py:127-135 per scenario: a right-hand side that blows up in finite time drives the step size below
    machine precision (flag 3, 'failed'); a discontinuous right-hand side that no step size satisfies collects more
    than 100 rejections (flag 4); a benign scenario of the same batch finishes."""
    from solverz_b200 import Model, Ode, Opt, Param, Rodas, Var, made_numerical
    m = Model()
    m.x = Var('x', [1.0])
    m.k = Param('k', [1.0], per_scenario=True)
    m.f = Ode('f', m.k * m.x * m.x, diff_var=m.x)           # x' = k x^2: blows up at t = 1 / k
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    nd.p["k"] = np.array([[-1.0], [2.0]])                    # decays / blows up at t = 0.5
    sol = Rodas(nd, [0, 1], np.tile(y0.array, (2, 1)), Opt(rtol=1e-6, atol=1e-9))
    assert sol.stats.succeed[0] and int(sol.stats.flag[0]) == 0
    assert abs(sol.Y[0, -1, 0] - 0.5) < 1e-5                 # x(1) = 1 / (1 + t)
    assert not sol.stats.succeed[1] and int(sol.stats.flag[1]) in (3, 4, 5)
    assert sol.T[1, -1] < 0.51                                # stopped at the singularity
