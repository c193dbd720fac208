"""GPU tests of ``Rodas(..., Opt(event=callable))`` through the device library: BASELINE.json configs[0]
(the README bouncing ball) and the reference's own event tests
(Solverz/solvers/test/test_rodas_event.py), against fixtures produced by the reference itself
(tests/golden/make_golden.py::golden_events)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def ball_events(t, y):
    return np.array([y[0]]), np.array([1]), np.array([-1])


def _stats(st):
    return [st.nstep, st.nfeval, st.nJeval, st.ndecomp, st.nreject, st.nsolve]


def test_readme_bouncing_ball(libszb):
    """README.md:28-51 verbatim (made_numerical(sparse=True) + Rodas + terminal event)."""
    from solverz_b200 import Model, Ode, Opt, Rodas, Var, made_numerical
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.h = Var('h', 0)
    m.v = Var('v', 20)
    m.f1 = Ode('f1', f=m.v, diff_var=m.h)
    m.f2 = Ode('f2', f=-9.8, diff_var=m.v)
    bball, y0 = m.create_instance()
    nbball = made_numerical(bball, y0, sparse=True)
    sol = Rodas(nbball, np.linspace(0, 30, 100), y0, Opt(event=ball_events))
    assert abs(sol.T[-1] - 4.08163265) < 1e-8
    assert _stats(sol.stats) == g["readme_stats"].tolist()          # nstep 8, nfeval 48
    np.testing.assert_allclose(sol.T, g["readme_T"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sol.Y.array, g["readme_Y"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(sol.Y['h'][:, 0], g["readme_Y"][:, 0], rtol=0, atol=1e-10)
    np.testing.assert_allclose(sol.te, g["readme_te"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sol.ye.array, g["readme_ye"], rtol=0, atol=1e-10)


def test_ten_bounces(libszb):
    """test_rodas_event.py:7-46: restart after every impact with 0.9 restitution."""
    from solverz_b200 import Model, Ode, Opt, Rodas, Var, made_numerical
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.x = Var('x', [0, 20])
    m.f1 = Ode('f1', m.x[1], m.x[0])
    m.f2 = Ode('f2', -9.8, m.x[1])
    bball, y0 = m.create_instance()
    nbball = made_numerical(bball, y0, sparse=True)
    opt = Opt(event=ball_events)
    teout = np.array([])
    tstart = 0
    for _ in range(10):
        sol = Rodas(nbball, np.linspace(tstart, 30, 100), y0, opt)
        teout = np.concatenate([teout, sol.te])
        y0['x'][0] = 0
        y0['x'][1] = -0.9 * sol.Y[-1]['x'][1]
        tstart = sol.T[-1]
    np.testing.assert_allclose(teout, [4.081633047365558, 7.755103268847604, 11.061226369910107, 14.03673684129459,
                                       16.714695668619328, 19.124858513456523, 21.294005283692695,
                                       23.246237159791765, 25.003245923211466, 26.584554164220997],
                               rtol=1e-5, atol=0)              # the reference test's own numbers
    np.testing.assert_allclose(teout, g["bounce_te"], rtol=1e-10)


@pytest.mark.parametrize("scheme", ["rodas4", "rodas3"])
def test_orbit(libszb, scheme):
    """test_rodas_event.py:49-89: one terminal, one non-terminal event with directions."""
    from solverz_b200 import Model, Ode, Opt, Rodas, Var, made_numerical
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
    norbit = made_numerical(orbit, y0, sparse=True)

    def events(t, y):
        d = 2 * (y[0:2] - np.array([1.2, 0])).dot(y[2:4])
        return np.array([d, d]), np.array([1, 0]), np.array([1, -1])

    sol = Rodas(norbit, [0, 7], y0, Opt(event=events, rtol=1e-5, atol=1e-4, scheme=scheme))
    known = {"rodas4": [3.09609516, 6.19215951], "rodas3": [3.095607, 6.192964]}
    np.testing.assert_allclose(sol.te, known[scheme], rtol=1e-5, atol=0)
    assert _stats(sol.stats) == g[f"orbit_{scheme}_stats"].tolist()
    np.testing.assert_allclose(sol.te, g[f"orbit_{scheme}_te"], rtol=1e-8)
    np.testing.assert_allclose(sol.ye.array, g[f"orbit_{scheme}_ye"], rtol=1e-6, atol=1e-8)
    assert sol.ie.tolist() == g[f"orbit_{scheme}_ie"].tolist()
    np.testing.assert_allclose(sol.Y.array, g[f"orbit_{scheme}_Y"], rtol=1e-5, atol=1e-7)


def test_callback_events_on_a_batch(libszb):
    """The reference's callback interface on MORE than one trajectory: every scenario is asked after each of its
    accepted steps (szb_rodas one accepted step per call); five launch speeds, each against the single-trajectory
    run of the same call path and the closed form t = 2 v0 / g."""
    from solverz_b200 import Model, Ode, Opt, Rodas, Var, made_numerical
    m = Model()
    m.x = Var('x', [0, 20])
    m.f1 = Ode('f1', m.x[1], m.x[0])
    m.f2 = Ode('f2', -9.8, m.x[1])
    bball, y0 = m.create_instance()
    nb = made_numerical(bball, y0, sparse=True)
    v0 = np.array([20.0, 10.0, 35.0, 5.0, 27.5])
    Y0 = np.stack([np.zeros_like(v0), v0], axis=1)
    tsp = np.linspace(0, 30, 100)
    sol = Rodas(nb, tsp, Y0, Opt(event=ball_events))
    assert (sol.stats.flag == 8).all() and (sol.stats.nevent == 1).all()
    for b in range(len(v0)):
        one = Rodas(nb, tsp, Y0[b], Opt(event=ball_events))
        assert abs(sol.te[b][0] - 2 * v0[b] / 9.8) < 1e-8
        assert sol.te[b][0] == one.te[0] and int(sol.stats.nstep[b]) == one.stats.nstep
        np.testing.assert_array_equal(sol.T[b], one.T)
        np.testing.assert_array_equal(sol.Y[b], one.Y)
        assert sol.T[b][-1] == sol.te[b][0]                      # the event point closes T (rodas.py:291-299)


def test_device_events_readme_ball(libszb):
    """The README case again, with the event given as a ComponentEvents object: detection, bisection and the
    roll-back run inside the step kernel (szb_rodas) instead of the host loop; same fixture."""
    from solverz_b200 import ComponentEvents, Model, Ode, Opt, Rodas, Var, made_numerical
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.h = Var('h', 0)
    m.v = Var('v', 20)
    m.f1 = Ode('f1', f=m.v, diff_var=m.h)
    m.f2 = Ode('f2', f=-9.8, diff_var=m.v)
    bball, y0 = m.create_instance()
    nbball = made_numerical(bball, y0, sparse=True)
    sol = Rodas(nbball, np.linspace(0, 30, 100), y0, Opt(event=ComponentEvents([0], 0.0, direction=-1, terminal=1)))
    assert _stats(sol.stats) == g["readme_stats"].tolist()
    np.testing.assert_allclose(sol.T, g["readme_T"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(sol.Y.array, g["readme_Y"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(sol.te, g["readme_te"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(sol.ye.array, g["readme_ye"], rtol=0, atol=1e-9)
    assert sol.ie.tolist() == g["readme_ie"].tolist()


def test_device_events_batch_of_balls(libszb):
    """Five launch speeds as ONE batch: every scenario stops at its own impact time."""
    from solverz_b200 import ComponentEvents, Model, Ode, Opt, Rodas, Var, made_numerical
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.x = Var('x', [0, 20])
    m.f1 = Ode('f1', m.x[1], m.x[0])
    m.f2 = Ode('f2', -9.8, m.x[1])
    bball, y0 = m.create_instance()
    nb = made_numerical(bball, y0, sparse=True)
    v0 = g["batch_v0"]
    Y0 = np.stack([np.zeros_like(v0), v0], axis=1)
    sol = Rodas(nb, np.linspace(0, 8, 41), Y0, Opt(event=ComponentEvents([0], 0.0, -1, 1)))
    for b in range(len(v0)):
        st = [int(getattr(sol.stats, k)[b]) for k in ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")]
        assert st == g["batch_stats"][b].tolist(), b
        if np.isnan(g["batch_te"][b]):                  # the ball is still in the air at t = 8
            assert sol.stats.nevent[b] == 0 and sol.stats.flag[b] == 0
            rows = 41
        else:
            assert sol.stats.nevent[b] == 1 and sol.stats.flag[b] == 8
            assert abs(sol.te[b, 0] - g["batch_te"][b]) < 1e-11 and abs(sol.te[b, 0] - 2 * v0[b] / 9.8) < 1e-9
            rows = int(sol.stats.n_rows[b])
            assert rows in (g["batch_nT"][b] - 1, g["batch_nT"][b])   # the reference appends the event row to T
            np.testing.assert_allclose(sol.ye[b, 0], g["batch_Y"][b][g["batch_nT"][b] - 1], rtol=0, atol=1e-9)
        np.testing.assert_allclose(sol.Y[b, :rows], g["batch_Y"][b][:rows], rtol=0, atol=1e-9)
    assert sol.stats.succeed.all()


@pytest.mark.parametrize("tag", ["nt", "term"])
@pytest.mark.parametrize("kind", ["steps", "dense"])
def test_device_events_oscillator(libszb, tag, kind):
    """Damped oscillator with four component events (rising / falling zero crossings of x, two thresholds),
    all non-terminal or the last one terminal; two-point tspan (every step) and dense output."""
    from solverz_b200 import ComponentEvents, Model, Ode, Opt, Rodas, Var, made_numerical
    g = np.load(GOLD / "events_ref.npz")
    m = Model()
    m.x = Var('x', 1.0)
    m.v = Var('v', 0.0)
    m.f1 = Ode('f1', f=m.v, diff_var=m.x)
    m.f2 = Ode('f2', f=-m.x - 0.1 * m.v, diff_var=m.v)
    osc, yo = m.create_instance()
    nosc = made_numerical(osc, yo, sparse=True)
    ev = ComponentEvents([0, 0, 1, 0], [0.0, 0.0, 0.3, -0.5], [1, -1, 1, -1], [0, 0, 0, 1 if tag == "term" else 0])
    tsp = [0, 12] if kind == "steps" else np.linspace(0, 12, 61)
    sol = Rodas(nosc, tsp, yo, Opt(event=ev, rtol=1e-6, atol=1e-8))
    k = f"osc_{tag}_{kind}"
    assert _stats(sol.stats) == g[f"{k}_stats"].tolist()
    assert sol.ie.tolist() == g[f"{k}_ie"].tolist()
    np.testing.assert_allclose(sol.te, g[f"{k}_te"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(sol.ye.array, g[f"{k}_ye"], rtol=0, atol=1e-8)
    assert len(sol.T) == len(g[f"{k}_T"])
    # the step sizes after an event inherit its last-bit differences through the controller's err**(1/4)
    np.testing.assert_allclose(sol.T, g[f"{k}_T"], rtol=0, atol=2e-8)
    np.testing.assert_allclose(sol.Y.array, g[f"{k}_Y"], rtol=0, atol=2e-8)
