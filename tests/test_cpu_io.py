"""Result files: the reference's ``save_result`` layout (Solverz/utilities/io.py:19-90) with a
leading scenario axis for batched solutions."""
import numpy as np
import pytest

from solverz_b200 import Address, Stats, TimeVars, Vars, aesol, daesol, load, load_result, save, save_result


def _addr():
    a = Address()
    a.add("h", 1)
    a.add("x", 2)
    return a


def test_batched_daesol_npz_and_matlab(tmp_path):
    a, B, nt = _addr(), 5, 7
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(B, nt, 3))
    st = Stats("rodas4")
    st.nstep, st.nfeval, st.succeed = np.arange(B), 6 * np.arange(B), np.ones(B, bool)
    sol = daesol(np.linspace(0, 1, nt), TimeVars(a, Y), te=rng.normal(size=(B, 2)), ye=rng.normal(size=(B, 2, 3)),
                 ie=np.zeros((B, 2)), stats=st)
    p = save_result(sol, "run", str(tmp_path))
    d = load_result(p)
    assert sorted(d) == ["Time", "Y", "ie", "stats", "te", "ye"]
    np.testing.assert_array_equal(d["Y"]["x"], Y[..., 1:3])
    np.testing.assert_array_equal(d["Y"]["h"], Y[..., 0:1])
    np.testing.assert_array_equal(d["stats"]["nfeval"], 6 * np.arange(B))
    np.testing.assert_array_equal(d["ye"], sol.ye)                  # a plain array stays flat
    # plain-array solution with the rendered model's slices
    sol2 = daesol(np.linspace(0, 1, nt), Y, stats=st)
    d2 = load_result(save_result(sol2, "run2", str(tmp_path), var_slices={"h": slice(0, 1), "x": slice(1, 3)}))
    np.testing.assert_array_equal(d2["Y"]["x"], Y[..., 1:3])
    assert "te" not in d2
    from scipy.io import loadmat
    m = loadmat(str(save_result(sol, "run", str(tmp_path), f="matlab")), simplify_cells=True)["sol"]
    np.testing.assert_array_equal(m["Y"]["x"], Y[..., 1:3])
    np.testing.assert_array_equal(m["Time"], sol.T)


def test_aesol_and_pickle_round_trip(tmp_path):
    a = _addr()
    y = Vars(a, np.arange(12.0).reshape(4, 3))
    st = Stats()
    st.nstep = np.array([3, 3, 4, 3])
    d = load_result(save_result(aesol(y, st), "pf", str(tmp_path)))
    np.testing.assert_array_equal(d["y"]["x"], y["x"])
    np.testing.assert_array_equal(d["stats"]["nstep"], st.nstep)
    save({"y": y.array}, str(tmp_path / "obj.pkl"))
    np.testing.assert_array_equal(load(str(tmp_path / "obj.pkl"))["y"], y.array)
    with pytest.raises(TypeError):
        save_result(object(), "bad", str(tmp_path))
    with pytest.raises(ValueError):
        save_result(aesol(y, st), "bad", str(tmp_path), f="csv")


def test_xlsx_needs_openpyxl_or_one_trajectory(tmp_path):
    a = _addr()
    sol = daesol(np.linspace(0, 1, 4), TimeVars(a, np.zeros((4, 3))), stats=Stats())
    try:
        import openpyxl  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            save_result(sol, "one", str(tmp_path), f="xlsx")
    else:
        assert save_result(sol, "one", str(tmp_path), f="xlsx").exists()


def test_daesol_window_and_append():
    """daesol[i], daesol[a:b] (events inside the window) and daesol.append (restart chains), the contract of the
    reference's container (solvers/solution.py:40-99)."""
    from solverz_b200.solvers.solution import daesol
    from solverz_b200.solvers.stats import Stats
    T = np.linspace(0.0, 1.0, 11)
    Y = np.stack([T, T ** 2], axis=1)
    st = Stats("rodas4")
    st.nstep, st.nfeval = 10, 60
    sol = daesol(T, Y, te=np.array([0.25, 0.75]), ye=np.array([[0.25, 0.0625], [0.75, 0.5625]]), ie=np.array([0.0, 1.0]),
                 stats=st)
    w = sol[2:6]
    assert np.array_equal(w.T, T[2:6]) and np.array_equal(w.Y, Y[2:6])
    assert w.te.tolist() == [0.25] and w.ie.tolist() == [0.0] and w.ye.shape == (1, 2)
    assert sol[8:].te.size == 0 and sol[3].T == T[3]
    with pytest.raises(NotImplementedError):
        sol["x"]
    st2 = Stats("rodas4")
    st2.nstep, st2.nfeval = 5, 30
    more = daesol(T[:6] + 1.0, Y[:6], te=np.array([1.3]), ye=np.array([[1.3, 1.69]]), ie=np.array([0.0]), stats=st2)
    sol.append(more)
    assert len(sol.T) == 17 and sol.Y.shape == (17, 2) and sol.te.tolist() == [0.25, 0.75, 1.3]
    assert sol.stats.nstep == 15 and sol.stats.nfeval == 90
    empty = daesol()
    empty.append(more)
    assert empty.stats.scheme == "rodas4" and np.array_equal(empty.T, more.T) and empty.te.tolist() == [1.3]
