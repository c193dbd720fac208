"""CPU tests of the host half of ``Opt(event=callable)`` (solvers/rodas_event.py::EventLocator): what happens to ONE
accepted step once the callback has been asked -- bracketing, direction filter, root search on the step's dense
output, the event_duration rule, terminal events, and which values are remembered for the next step (reference
semantics: rodas.py:200-271).  The steps themselves are taken by ``szb_rodas`` on the device; the GPU tests
(tests/test_gpu_events.py) run the reference's own event cases end to end against its fixtures."""
import numpy as np


def _fall(h0, v0):
    """Exact trajectory of a falling ball and its dense output over a step [told, told + dt]."""
    def state(t):
        return np.array([h0 + v0 * t - 4.9 * t * t, v0 - 9.8 * t])

    def interp_of(told, dt):
        return lambda tau: state(told + tau * dt)
    return state, interp_of


def ball_events(t, y):
    return np.array([y[0]]), np.array([1]), np.array([-1])


def test_terminal_event_is_located_and_the_step_ends_there():
    from solverz_b200.solvers.rodas_event import EventLocator
    state, interp_of = _fall(0.0, 20.0)
    loc = EventLocator(ball_events, 1e-8, 0.0, state(1e-3))        # just after the launch: h > 0
    t, y, moved = loc.after_step(1e-3, 2.0, 2.001, state(2.001), interp_of(1e-3, 2.0))
    assert not moved and not loc.te and t == 2.001                 # still in the air: nothing happens
    t, y, moved = loc.after_step(2.001, 3.0, 5.001, state(5.001), interp_of(2.001, 3.0))
    assert moved and loc.terminal and len(loc.te) == 1 and loc.ie == [0]
    assert abs(t - 2 * 20 / 9.8) < 1e-12 and t == loc.te[0]
    assert abs(y[0]) < 1e-9 and abs(y[1] + 20.0) < 1e-9             # lands with the launch speed
    # the values kept for the next step are those at the ORIGINAL step end (rodas.py:266), not at the event
    assert loc.value[0] == ball_events(5.001, state(5.001))[0][0]


def test_direction_filter_and_non_terminal_events():
    from solverz_b200.solvers.rodas_event import EventLocator

    def ev(t, y):      # component 0 rising only (non-terminal), component 1 falling only (non-terminal)
        return np.array([y[0], y[0]]), np.array([0, 0]), np.array([1, -1])
    traj = lambda t: np.array([np.sin(t), np.cos(t)])
    loc = EventLocator(ev, 1e-8, 0.5, traj(0.5))
    interp = lambda told, dt: (lambda tau: traj(told + tau * dt))
    t, y, moved = loc.after_step(0.5, 3.0, 3.5, traj(3.5), interp(0.5, 3.0))     # sin crosses 0 downwards at pi
    assert moved and loc.ie == [1] and abs(loc.te[0] - np.pi) < 1e-12 and not loc.terminal
    assert t == loc.te[0]                                           # the step ends at the event even if not terminal
    t, y, moved = loc.after_step(t, 3.5, t + 3.5, traj(t + 3.5), interp(t, 3.5))  # upwards at 2 pi: component 0
    assert moved and loc.ie == [1, 0] and abs(loc.te[1] - 2 * np.pi) < 1e-11


def test_event_too_close_to_the_step_start_is_dropped():
    from solverz_b200.solvers.rodas_event import EventLocator
    state, interp_of = _fall(1e-12, -1.0)                            # reaches h = 0 after 1e-12 s
    loc = EventLocator(ball_events, 1e-8, 0.0, state(0.0))
    t, y, moved = loc.after_step(0.0, 0.1, 0.1, state(0.1), interp_of(0.0, 0.1))
    assert not moved and loc.te == [] and not loc.terminal           # |tevent - told| < event_duration (rodas.py:257-259)
    # the reference leaves `value` at its last evaluation in that case: NOT the value at the step end
    assert loc.value[0] != ball_events(0.1, state(0.1))[0][0]


def test_two_components_same_step_second_sees_the_shortened_step():
    from solverz_b200.solvers.rodas_event import EventLocator

    def ev(t, y):
        return np.array([y[0] - 1.0, y[0] - 2.0]), np.array([0, 0]), np.array([0, 0])
    traj = lambda t: np.array([t])
    loc = EventLocator(ev, 1e-8, 0.0, traj(0.0))
    t, y, moved = loc.after_step(0.0, 3.0, 3.0, traj(3.0), lambda tau: traj(3.0 * tau))
    # component 0 crosses at t = 1 and ends the step there.  Component 1 is still searched with the end values of
    # the ORIGINAL step (rodas.py:207-209 reads value_save), so its secant guess 2.0 -- an exact root -- is taken
    # although it lies behind the shortened step: the reference's behaviour, reproduced on purpose
    assert loc.ie == [0, 1] and abs(loc.te[0] - 1.0) < 1e-12 and loc.te[1] == 2.0 and t == 2.0
