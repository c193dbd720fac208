"""GPU tests of the batched fixed-step one-step methods: backward_euler and
implicit_trapezoid against the reference's own runs (tests/golden/make_golden.py::golden_implicit) on the
DAE of the reference's tests/test_dae.py, one scenario and a batch with a per-scenario parameter."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _model(per_scenario):
    from solverz_b200 import Eqn, Model, Ode, Param, Var, made_numerical
    m = Model()
    m.x = Var('x', 1)
    m.y = Var('y', 1)
    m.a = Param('a', [0.5], per_scenario=per_scenario)
    m.f = Ode(name='f', f=-m.x ** 3 + m.a * m.y ** 2, diff_var=m.x)
    m.g = Eqn(name='g', eqn=m.x ** 2 + m.y ** 2 - 2)
    dae, y0 = m.create_instance()
    return made_numerical(dae, y0, sparse=True), y0


def _stats(st, b=None):
    v = [st.nstep, st.nfeval, st.nJeval, st.ndecomp, 0, st.nsolve]
    return [int(x if np.isscalar(x) else x[b]) for x in v]


@pytest.mark.parametrize("name", ["trap", "beuler"])
def test_one_scenario_matches_reference(libszb, name):
    from solverz_b200 import Opt, backward_euler, implicit_trapezoid
    g = np.load(GOLD / "implicit_ref.npz")
    nd, y0 = _model(False)
    fn = implicit_trapezoid if name == "trap" else backward_euler
    sol = fn(nd, [0, 20], y0, Opt(step_size=0.1))
    np.testing.assert_allclose(sol.T, g[f"{name}_05_T"], rtol=0, atol=1e-12)
    assert np.max(np.abs(sol.Y.array - g[f"{name}_05_Y"])) < 1e-10
    assert _stats(sol.stats) == g[f"{name}_05_stats"].tolist()
    if name == "trap":     # the reference's own criterion, tests/test_dae.py:47
        assert np.max(np.abs(g["rodas_bench_Y"] - sol.Y.array)) < 7e-4


@pytest.mark.parametrize("name", ["trap", "beuler"])
def test_batch_with_per_scenario_parameter(libszb, name):
    from solverz_b200 import Opt, backward_euler, implicit_trapezoid
    g = np.load(GOLD / "implicit_ref.npz")
    nd, y0 = _model(True)
    nd.p["a"] = np.array([[0.5], [0.8], [0.5]])
    fn = implicit_trapezoid if name == "trap" else backward_euler
    sol = fn(nd, [0, 20], np.tile(y0.array, (3, 1)), Opt(step_size=0.1))
    for b, tag in enumerate(("05", "08", "05")):
        assert np.max(np.abs(sol.Y[b] - g[f"{name}_{tag}_Y"])) < 1e-10
        assert _stats(sol.stats, b) == g[f"{name}_{tag}_stats"].tolist()
    assert sol.stats.succeed.all()
