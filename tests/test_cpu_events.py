"""CPU tests of the event-mode Rodas controller (solvers/rodas_event.py; reference rodas.py:200-303).

The controller is host logic around four device operations.  Here those four are supplied by a stub
built from the SZBM_HOST_CHECK build of the generated model source (F, J values) and SciPy's SuperLU
(the reference's own linear solver), so the event detection, bisection, dense-output clipping and the
``te / ye / ie`` bookkeeping are checked against the reference's fixtures without a GPU.  The GPU
tests (tests/test_gpu_events.py) run the same cases through ``Rodas`` and the device library."""
import ctypes
from pathlib import Path

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as sla

from test_cpu_host import _eval, _host_lib

GOLD = Path(__file__).parent / "golden"


class HostStubOps:
    """Same four operations as solvers.rodas_event.DeviceOps, on the host-check library."""

    def __init__(self, ir, pat, lib, M):
        self.ir, self.pat, self.lib, self.M = ir, pat, lib, sps.csc_matrix(M)
        self.Jm = None
        self.lu = None

    def F(self, t, y):
        return _eval(self.lib, self.ir, self.pat, y, {}, float(t))[0][0]

    def J(self, t, y):
        jv = _eval(self.lib, self.ir, self.pat, y, {}, float(t))[1][0]
        n = self.ir.n
        self.Jm = sps.csc_matrix((jv, self.pat.indices, self.pat.indptr), shape=(n, n))

    def factor(self, dt_gamma):
        self.lu = sla.splu(sps.csc_matrix(self.M - dt_gamma * self.Jm))

    def solve(self, rhs):
        return self.lu.solve(np.asarray(rhs, dtype=float))


def _mass(ir):
    r, c = ir.mass
    return sps.csc_array((np.ones(len(r)), (r, c)), shape=(ir.n, ir.n))


def ball_events(t, y):
    return np.array([y[0]]), np.array([1]), np.array([-1])


def test_readme_bouncing_ball_config1(tmp_path):
    """BASELINE.json configs[0]: README.md:28-51 -- lands at t = 2*20/9.8, 8 steps, 48 F evaluations."""
    from solverz_b200 import Model, Ode, Opt, Var
    from solverz_b200.solvers.rodas_event import rodas_with_events
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.h = Var('h', 0)
    m.v = Var('v', 20)
    m.f1 = Ode('f1', f=m.v, diff_var=m.h)
    m.f2 = Ode('f2', f=-9.8, diff_var=m.v)
    bball, y0 = m.create_instance()
    ir, pat, lib, _ = _host_lib(bball, "bball", tmp_path)
    F, JV = _eval(lib, ir, pat, y0.array, {})
    np.testing.assert_allclose(F[0], [20, -9.8])             # reference test_inline_printer: nbball F / J
    J = sps.csc_matrix((JV[0], pat.indices, pat.indptr), shape=(2, 2)).toarray()
    np.testing.assert_allclose(J, [[0, 1], [0, 0]])
    opt = Opt(event=ball_events)
    opt.hmax = 30.0
    sol = rodas_with_events(HostStubOps(ir, pat, lib, _mass(ir)), _mass(ir), np.linspace(0, 30, 100), y0.array, opt)
    assert abs(sol.T[-1] - 2 * 20 / 9.8) < 1e-9
    st = sol.stats
    assert [st.nstep, st.nfeval, st.nJeval, st.ndecomp, st.nreject, st.nsolve] == g["readme_stats"].tolist()
    assert len(sol.T) == len(g["readme_T"])
    np.testing.assert_allclose(sol.T, g["readme_T"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sol.Y, g["readme_Y"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(sol.te, g["readme_te"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sol.ye, g["readme_ye"], rtol=0, atol=1e-10)
    assert sol.ie.tolist() == g["readme_ie"].tolist()


def test_orbit_terminal_and_nonterminal_events(tmp_path):
    """reference Solverz/solvers/test/test_rodas_event.py:49-89 (rodas4 -> ntrp1, rodas3 -> ntrp2)."""
    from solverz_b200 import Model, Ode, Opt, Var
    from solverz_b200.solvers.rodas_event import rodas_with_events
    g = np.load(GOLD / "events_ref.npz")
    mu = 1 / 82.45
    mustar = 1 - mu
    m = Model()
    m.y = Var('y', [1.2, 0, 0, -1.04935750983031990726])
    m.f1 = Ode('f1', m.y[2], m.y[0])
    m.f2 = Ode('f2', m.y[3], m.y[1])
    r13 = ((m.y[0] + mu) ** 2 + m.y[1] ** 2) ** 1.5
    r23 = ((m.y[0] - mustar) ** 2 + m.y[1] ** 2) ** 1.5
    m.f3 = Ode('f3', 2 * m.y[3] + m.y[0] - mustar * ((m.y[0] + mu) / r13) - mu * ((m.y[0] - mustar) / r23), m.y[2])
    m.f4 = Ode('f4', -2 * m.y[2] + m.y[1] - mustar * (m.y[1] / r13) - mu * (m.y[1] / r23), m.y[3])
    orbit, y0 = m.create_instance()
    ir, pat, lib, _ = _host_lib(orbit, "orbit", tmp_path)

    def events(t, y):
        d = 2 * (y[0:2] - np.array([1.2, 0])).dot(y[2:4])
        return np.array([d, d]), np.array([1, 0]), np.array([1, -1])

    known = {"rodas4": [3.09609516, 6.19215951], "rodas3": [3.095607, 6.192964]}   # the reference test's own numbers
    for sch in ("rodas4", "rodas3"):
        opt = Opt(event=events, rtol=1e-5, atol=1e-4, scheme=sch)
        opt.hmax = 7.0
        sol = rodas_with_events(HostStubOps(ir, pat, lib, _mass(ir)), _mass(ir), [0, 7], y0.array, opt)
        np.testing.assert_allclose(sol.te, known[sch], rtol=1e-5, atol=0)
        st = sol.stats
        assert [st.nstep, st.nfeval, st.nJeval, st.ndecomp, st.nreject, st.nsolve] == g[f"orbit_{sch}_stats"].tolist()
        np.testing.assert_allclose(sol.te, g[f"orbit_{sch}_te"], rtol=1e-9)
        np.testing.assert_allclose(sol.ye, g[f"orbit_{sch}_ye"], rtol=1e-7, atol=1e-9)
        assert sol.ie.tolist() == g[f"orbit_{sch}_ie"].tolist()
        np.testing.assert_allclose(sol.T, g[f"orbit_{sch}_T"], rtol=1e-9)
        np.testing.assert_allclose(sol.Y, g[f"orbit_{sch}_Y"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(g["orbit_rodas4_ye"], [[-1.262468e+00, 1.162918e-05, 4.956687e-06, 1.049574e+00],
                                                      [1.200024e+00, 1.455118e-09, 6.495375e-05, -1.049349e+00]],
                               rtol=1e-5, atol=0)
