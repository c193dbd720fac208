"""Host-side matching used for the static pivots (lu/matching.py, lu/symbolic.py::static_row_matching)."""
import numpy as np
import pytest
import scipy.sparse as sps
from scipy.sparse.csgraph import maximum_bipartite_matching


@pytest.mark.parametrize("n,density,seed", [(50, 0.05, 0), (200, 0.02, 1), (400, 0.01, 2)])
def test_hopcroft_karp_is_a_maximum_matching(n, density, seed):
    from solverz_b200.lu.matching import max_matching
    A = sps.random(n, n, density=density, random_state=seed, format="coo")
    match = max_matching(A.row.astype(np.int64), A.col.astype(np.int64), n)      # match[j] = row of column j
    got = match >= 0
    assert len(np.unique(match[got])) == int(got.sum())                          # rows used once
    present = set(zip(A.row.tolist(), A.col.tolist()))
    assert all((int(match[j]), int(j)) in present for j in np.flatnonzero(got))  # only existing edges
    ref = maximum_bipartite_matching(sps.csr_matrix((np.ones(A.nnz), (A.row, A.col)), shape=(n, n)), perm_type='row')
    assert int(got.sum()) == int((ref >= 0).sum())                               # same cardinality as SciPy's


def test_bottleneck_matching_on_an_incidence_type_matrix():
    """A matrix without a natural diagonal (node-pipe incidence block next to a loop block): the greedy pass leaves
    most columns free, the bottleneck matching returns a perfect matching whose smallest entry is as large as any
    perfect matching allows, forced (mass) pairs are kept."""
    from solverz_b200.lu.symbolic import _bottleneck_matching, static_row_matching
    rng = np.random.default_rng(3)
    n = 300
    # a permuted strongly diagonal matrix plus weak fill: the best bottleneck is the permuted diagonal
    perm = rng.permutation(n)
    R = sps.random(n, n, density=0.03, random_state=4, format="lil") * 0.1
    for j in range(n):
        R[perm[j], j] = 0.6 + 0.4 * rng.random()
    R = sps.csr_matrix(R)
    match = _bottleneck_matching(R)
    assert sorted(match.tolist()) == list(range(n))
    vals = np.asarray(R[match, np.arange(n)]).ravel()
    assert vals.min() >= 0.6                                   # the permuted diagonal is the only way to stay >= 0.6
    forced = (np.array([perm[0], perm[1]]), np.array([0, 1]))
    m2 = _bottleneck_matching(R, forced)
    assert m2[0] == perm[0] and m2[1] == perm[1] and sorted(m2.tolist()) == list(range(n))
    # the public entry: a DAE-like call with forced pairs on a large-enough matrix takes the same route
    A = sps.csc_matrix(R)
    rp = static_row_matching(A, exact_below=10, forced=forced)
    assert sorted(rp.tolist()) == list(range(n)) and rp[0] == perm[0]
