"""Maximum bipartite matching (Hopcroft-Karp) for the static row matching of lu/symbolic.py.

SciPy's ``maximum_bipartite_matching`` needs minutes on incidence-type graphs of 10^5 vertices (district-heating
Jacobians: no natural diagonal, long augmenting paths); this layered BFS + iterative DFS runs them in tens of
milliseconds once numba has compiled it (host-side integer work of the one-time analysis; numba is what the reference
itself depends on).  Without numba the same code runs interpreted for small graphs and SciPy's routine is used for
large ones."""
from __future__ import annotations

import numpy as np


def _hopcroft_karp_py(indptr, indices, n_left, n_right):
    """Maximum bipartite matching; left vertices u with neighbours indices[indptr[u]:indptr[u+1]] (right vertices).
    Returns match_l (right partner of each left vertex or -1)."""
    INF = 1 << 60
    match_l = np.full(n_left, -1, np.int64)
    match_r = np.full(n_right, -1, np.int64)
    # greedy start
    for u in range(n_left):
        for k in range(indptr[u], indptr[u + 1]):
            v = indices[k]
            if match_r[v] < 0:
                match_l[u] = v
                match_r[v] = u
                break
    dist = np.empty(n_left, np.int64)
    queue = np.empty(n_left, np.int64)
    it = np.empty(n_left, np.int64)
    stack = np.empty(n_left, np.int64)
    while True:
        # BFS from free left vertices
        qh = 0
        qt = 0
        for u in range(n_left):
            if match_l[u] < 0:
                dist[u] = 0
                queue[qt] = u
                qt += 1
            else:
                dist[u] = INF
        found = False
        while qh < qt:
            u = queue[qh]
            qh += 1
            for k in range(indptr[u], indptr[u + 1]):
                w = match_r[indices[k]]
                if w < 0:
                    found = True
                elif dist[w] == INF:
                    dist[w] = dist[u] + 1
                    queue[qt] = w
                    qt += 1
        if not found:
            break
        # DFS (iterative) along the layers
        for u in range(n_left):
            it[u] = indptr[u]
        for root in range(n_left):
            if match_l[root] >= 0:
                continue
            sp = 0
            stack[0] = root
            while sp >= 0:
                u = stack[sp]
                advanced = False
                while it[u] < indptr[u + 1]:
                    v = indices[it[u]]
                    it[u] += 1
                    w = match_r[v]
                    if w < 0:
                        # augment along the stack
                        vv = v
                        for q in range(sp, -1, -1):
                            uu = stack[q]
                            prev = match_l[uu]
                            match_l[uu] = vv
                            match_r[vv] = uu
                            vv = prev
                        sp = -1
                        advanced = True
                        break
                    if dist[w] == dist[u] + 1:
                        sp += 1
                        stack[sp] = w
                        advanced = True
                        break
                if not advanced:
                    dist[u] = INF
                    sp -= 1
    return match_l


try:                                   # pragma: no cover - depends on the environment
    from numba import njit
    hopcroft_karp = njit(cache=True)(_hopcroft_karp_py)
    HAVE_NUMBA = True
except Exception:                      # pragma: no cover
    hopcroft_karp = _hopcroft_karp_py
    HAVE_NUMBA = False


def max_matching(rows: np.ndarray, cols: np.ndarray, n: int) -> np.ndarray:
    """``match[j]`` = row matched to column ``j`` (or -1) of a maximum matching of the bipartite graph with edges
    ``(rows[k], cols[k])``."""
    import scipy.sparse as sp
    if not HAVE_NUMBA and len(rows) > 20000:
        from scipy.sparse.csgraph import maximum_bipartite_matching
        Gt = sp.csr_matrix((np.ones(len(rows)), (cols, rows)), shape=(n, n))
        col_of = maximum_bipartite_matching(Gt, perm_type='row')
        match = np.full(n, -1, dtype=np.int64)
        ok = col_of >= 0
        match[col_of[ok]] = np.flatnonzero(ok)
        return match
    G = sp.csr_matrix((np.ones(len(rows)), (cols, rows)), shape=(n, n))       # left = columns, right = rows
    return hopcroft_karp(G.indptr.astype(np.int64), G.indices.astype(np.int64), n, n)
