"""CPU tests of lu/amd.py (the approximate-minimum-degree ordering behind the analysis of large patterns): it returns
a permutation for every kind of graph, members merged into one vertex leave together, trees are eliminated without
fill, and on a lattice the fill is in the range SuperLU's multiple-minimum-degree reaches."""
import numpy as np
import scipy.sparse as sp


def _sym(n, rows, cols):
    S = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    S = ((S + S.T) != 0).astype(float).tocsr()
    S.setdiag(0)
    S.eliminate_zeros()
    S.sort_indices()
    return S


def _fill(S, order):
    """nnz(L) of the symbolic Cholesky factor of pattern(S) + I under ``order`` (plain set elimination)."""
    n = S.shape[0]
    pos = np.empty(n, dtype=int)
    pos[order] = np.arange(n)
    adj = [set(pos[S.indices[S.indptr[v]:S.indptr[v + 1]]]) for v in order]
    nnz = 0
    for k in range(n):
        later = {j for j in adj[k] if j > k}
        nnz += len(later)
        if later:
            m = min(later)
            adj[m] |= later - {m}
    return nnz


def test_permutation_on_all_kinds_of_graphs():
    from solverz_b200.lu.amd import amd_order
    rng = np.random.default_rng(0)
    for n, m in ((1, 0), (7, 0), (40, 30), (60, 400), (25, 25 * 24 // 2)):
        r, c = rng.integers(0, n, m), rng.integers(0, n, m)
        S = _sym(n, r, c)
        for w in (None, rng.integers(1, 4, n)):
            o = amd_order(S.indptr, S.indices, w)
            assert sorted(o.tolist()) == list(range(n))


def test_trees_have_no_fill_and_the_hub_of_a_star_goes_last():
    from solverz_b200.lu.amd import amd_order
    n = 50
    path = _sym(n, np.arange(n - 1), np.arange(1, n))
    assert _fill(path, amd_order(path.indptr, path.indices)) == n - 1
    star = _sym(n, np.zeros(n - 1, dtype=int), np.arange(1, n))
    o = amd_order(star.indptr, star.indices)
    assert _fill(star, o) == n - 1 and 0 in o[-2:]          # the hub and the last leaf are interchangeable
    rng = np.random.default_rng(3)
    parent = np.array([rng.integers(0, v) for v in range(1, 200)])
    tree = _sym(200, np.arange(1, 200), parent)
    assert _fill(tree, amd_order(tree.indptr, tree.indices)) == 199


def test_indistinguishable_vertices_leave_together():
    from solverz_b200.lu.amd import amd_order
    # a clique of 6 hanging off a path: once the path is gone the clique members are indistinguishable
    n = 16
    r = list(range(9)) + [i for i in range(10, 16) for j in range(10, 16) if i < j]
    c = list(range(1, 10)) + [j for i in range(10, 16) for j in range(10, 16) if i < j]
    S = _sym(n, np.array(r + [9]), np.array(c + [10]))
    o = amd_order(S.indptr, S.indices)
    assert sorted(o.tolist()) == list(range(n))
    assert set(o[-6:].tolist()) == set(range(10, 16))
    assert _fill(S, o) == S.nnz // 2                         # a chordal graph: a perfect elimination order exists


def test_lattice_fill_is_in_the_range_of_superlu_mmd_and_weights_are_respected():
    from solverz_b200.lu.amd import amd_order
    from solverz_b200.lu.symbolic import _order_mmd_superlu
    nx = 24
    idx = np.arange(nx * nx).reshape(nx, nx)
    S = _sym(nx * nx, np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]),
             np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]))
    f_amd = _fill(S, amd_order(S.indptr, S.indices))
    f_mmd = _fill(S, _order_mmd_superlu(S))
    f_nat = _fill(S, np.arange(nx * nx))
    assert f_amd < 1.25 * f_mmd and f_amd < 0.6 * f_nat, (f_amd, f_mmd, f_nat)
    # the interpreted and the compiled routine are the same algorithm
    from solverz_b200.lu import amd as A
    small = _sym(36, np.concatenate([idx[:6, :5].ravel() % 36]), np.concatenate([idx[:6, 1:6].ravel() % 36]))
    w = np.arange(36) % 3 + 1
    o1 = A.amd_order(small.indptr, small.indices, w)
    o2 = A._amd_py(36, small.indptr.astype(np.int64), small.indices.astype(np.int64), w.astype(np.int64))
    assert np.array_equal(o1, o2)


def test_analysis_takes_the_amd_candidates_on_request_and_factorises_correctly():
    """orderings=('amd',) through lu.symbolic.analyze: the schedule it produces factorises a matrix (numpy walk of the
    multifrontal schedule, tests/mf_emulate.py) to the same solution as a dense solve."""
    from solverz_b200.lu.symbolic import analyze
    import mf_emulate as E
    rng = np.random.default_rng(5)
    nx = 9
    n = nx * nx
    idx = np.arange(n).reshape(nx, nx)
    S = _sym(n, np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]),
             np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]))
    A = (S.multiply(rng.uniform(-1, 1, S.shape)) + sp.diags(rng.uniform(4, 5, n))).tocsc()
    sym = analyze(A, orderings=("amd",))
    assert sym.ordering == "amd"
    b = rng.standard_normal(n)
    A.sort_indices()
    x = E.solve(sym, E.factor(sym, A.data), b)
    np.testing.assert_allclose(x, np.linalg.solve(A.toarray(), b), rtol=1e-10, atol=1e-12)


def test_large_pattern_is_analysed_without_superlu_or_python_dissection():
    """Above MMD_SUPERLU_MAX / ND_MAX vertices the analysis must not detour through SuperLU's numeric factorisation
    (54 minutes at n = 173 800 before lu/amd.py) nor through the interpreted BFS dissection: a 182 x 182 lattice
    (33 124 unknowns) is analysed by the AMD candidates alone, in seconds, and the schedule it yields is consistent."""
    import time
    from solverz_b200.lu import symbolic as SY
    nx = 182
    n = nx * nx
    idx = np.arange(n).reshape(nx, nx)
    S = _sym(n, np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]),
             np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]))
    rng = np.random.default_rng(9)
    A = (S.multiply(sp.csr_matrix((rng.uniform(-1, 1, S.nnz), S.indices, S.indptr), shape=S.shape))
         + sp.diags(rng.uniform(4, 5, n))).tocsc()
    assert n > SY.MMD_SUPERLU_MAX and n > SY.ND_MAX
    called = []
    orig_mmd, orig_nd = SY._order_mmd_superlu, SY._nested_dissection
    SY._order_mmd_superlu = lambda *a, **k: called.append("mmd") or orig_mmd(*a, **k)
    SY._nested_dissection = lambda *a, **k: called.append("nd") or orig_nd(*a, **k)
    try:
        t0 = time.perf_counter()
        sym = SY.analyze(A)
        dt = time.perf_counter() - t0
    finally:
        SY._order_mmd_superlu, SY._nested_dissection = orig_mmd, orig_nd
    assert called == [] and sym.ordering == "amd"
    assert dt < 120.0, dt
    assert sorted(sym.colp.tolist()) == list(range(n)) and sorted(sym.rowp.tolist()) == list(range(n))
    assert int(sym.sn_s.sum()) == n and sym.nnz_lu < 40 * A.nnz
