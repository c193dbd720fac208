"""Host-side symbolic analysis for the batched fixed-pattern multifrontal sparse LU.

This replaces what SuperLU does *inside every call* of the reference's linear-solve
seam -- ``solve(A, b)`` -> ``scipy.sparse.linalg.spsolve`` (reference
``Solverz/solvers/laesolver.py:11-15``) and ``lu_decomposition(A)`` -> ``splu``
(``laesolver.py:18-45``) -- with an analysis that is done ONCE per model, because every
scenario of a batch shares one Jacobian sparsity pattern:

1. static pivoting : maximum-product row matching on a representative numeric Jacobian
   (the batched device kernels never pivot; a pivot-growth flag reports trouble),
2. fill-reducing symmetric ordering on pattern(PA + (PA)^T) (nested dissection by BFS
   level structures with local minimum degree at the leaves, and SuperLU's MMD as a
   second candidate; the cheaper of the two in factor flops wins),
3. elimination tree, column structures, relaxed supernodes,
4. the multifrontal schedule consumed by ``csrc/szb_lu.cu``: per supernode a front index
   list, the scatter map J-slot -> front position, the child -> parent extend-add maps,
   panel offsets of the L / U factors, and the split of the tree into independent
   warp-level subtrees and a block-level top.

Everything here is integer work on the host; all numeric work happens on the GPU.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import os

import numpy as np
import scipy.sparse as sp

__all__ = ["LUSymbolic", "analyze", "ANALYSIS_VERSION"]

# bumped whenever the analysis or the tables derived from it change meaning: analyses cached next to a
# rendered module (param_and_setting.pkl) with another version are recomputed
ANALYSIS_VERSION = 3

# the "mmd" candidate is read off SuperLU (perm_c of a throw-away factorisation) up to this many compressed vertices;
# above it that factorisation costs minutes (20 at n = 69 400) and lu/amd.py orders the graph instead
MMD_SUPERLU_MAX = 25000
# the multi-start AMD candidates (lu/amd.py) compete only where the factorisation is flop-dominated.  Measured on the
# B200 (profiles/r2_ordering_experiment.md): cfg#3 (6.0e8 flops under MMD) 380 -> 339 ms per refactor with the best AMD
# start (4.8e8 flops); cfg#2 (5.2e6 flops) 10.42 -> 10.78 ms although AMD has 12.6 % fewer flops -- at that size the
# kernel time is assembly and write-out of many small fronts, and AMD's top fronts are larger (154 against 128)
AMD_MIN_FLOPS = 5.0e7
ND_MAX = 30000


# --------------------------------------------------------------------------------------
# 1. static pivoting
# --------------------------------------------------------------------------------------
def _heavy_matching(A: sp.csc_matrix, forced=None) -> np.ndarray:
    """Greedy heaviest-edge matching on |a_ij| / max_i |a_ij|; when it leaves columns unmatched, a bottleneck
    matching takes over.  Not guaranteed optimal for the product objective, but on diagonally strong
    matrices (power-flow, network, DAE Jacobians) the greedy pass reproduces the optimum, and it costs
    O(nnz log nnz) where SciPy's assignment solver needs minutes for n ~ 10^4."""
    n = A.shape[0]
    W = sp.coo_matrix(A)
    absd = np.abs(W.data)
    colmax = np.zeros(n)
    np.maximum.at(colmax, W.col, absd)
    colmax[colmax == 0.0] = 1.0
    ratio = absd / colmax[W.col]
    order = np.argsort(-ratio, kind="stable")
    row_of = np.full(n, -1, dtype=np.int64)      # column -> matched row
    col_of = np.full(n, -1, dtype=np.int64)      # row -> matched column
    if forced is not None:                       # pre-assigned (row, column) pairs: never displaced below
        row_of[forced[1]] = forced[0]
        col_of[forced[0]] = forced[1]
    for i, j, r in zip(W.row[order].tolist(), W.col[order].tolist(), ratio[order].tolist()):
        if r <= 0.0:
            break
        if row_of[j] < 0 and col_of[i] < 0:
            row_of[j] = i
            col_of[i] = j
    free = np.flatnonzero(row_of < 0)
    if len(free) > 64:
        # many columns without a heavy partner (blocks with no natural diagonal, e.g. incidence matrices, where the
        # scaled weights are flat anyway): bottleneck matching -- the largest threshold tau such that the entries
        # with ratio >= tau still admit a perfect matching, found by bisection with Hopcroft-Karp probes
        # (lu/matching.py, O(E sqrt(V)) each); forced pairs stay.  The augmenting-path completion below is O(free * nnz) in
        # Python: fine for the handful of columns a diagonally strong matrix leaves over, minutes for thousands.
        return _bottleneck_matching(sp.csr_matrix((ratio, (W.row, W.col)), shape=(n, n)), forced)
    is_forced_row = None
    if forced is not None:
        is_forced_row = np.zeros(n, dtype=bool)
        is_forced_row[forced[0]] = True
    if len(free):
        # a few columns left: augmenting paths from the greedy matching that try heavy edges first (keeps the heavy
        # pairs; a bottleneck matching from scratch would only protect the SMALLEST entry and was measured to give
        # the power-flow lattices pivot orders Newton does not survive)
        Ac = sp.csc_matrix((ratio, (W.row, W.col)), shape=(n, n))
        indptr, indices, data = Ac.indptr, Ac.indices, Ac.data

        def nbrs(j):
            ks = range(indptr[j], indptr[j + 1])
            return iter(sorted(ks, key=lambda k: -data[k]))

        for j0 in free.tolist():
            visited = set()
            reached_from: Dict[int, int] = {}
            stack = [(j0, nbrs(j0))]
            found = -1
            while stack and found < 0:
                j, it = stack[-1]
                advanced = False
                for k in it:
                    i = int(indices[k])
                    if i in visited:
                        continue
                    visited.add(i)
                    reached_from[i] = j
                    if col_of[i] < 0:
                        found = i
                    elif is_forced_row is not None and is_forced_row[i]:
                        continue                     # a forced pair is not re-matched
                    else:
                        j2 = int(col_of[i])
                        stack.append((j2, nbrs(j2)))
                    advanced = True
                    break
                if not advanced:
                    stack.pop()
            if found < 0:
                raise ValueError("Jacobian pattern is structurally singular: no full matching")
            i = found
            while True:
                j = reached_from[i]
                prev = int(row_of[j])
                row_of[j] = i
                col_of[i] = j
                if j == j0:
                    break
                i = prev
    return row_of


def _bottleneck_matching(R: sp.csr_matrix, forced=None) -> np.ndarray:
    """rowperm (column -> row) of a perfect matching of the bipartite graph of ``R`` (rows x columns, entries = scaled
    magnitudes in (0, 1]) whose SMALLEST entry is as large as possible."""
    n = R.shape[0]
    R = sp.coo_matrix(R)
    keep = R.data > 0
    rows, cols, vals = R.row[keep], R.col[keep], R.data[keep]
    row_of = np.full(n, -1, dtype=np.int64)
    if forced is not None:
        row_of[forced[1]] = forced[0]
        fr = np.zeros(n, dtype=bool)
        fc = np.zeros(n, dtype=bool)
        fr[forced[0]] = True
        fc[forced[1]] = True
        k2 = ~fr[rows] & ~fc[cols]
        rows, cols, vals = rows[k2], cols[k2], vals[k2]
    need = int((row_of < 0).sum())

    from .matching import max_matching

    def probe(tau):
        m = vals >= tau
        match = max_matching(rows[m], cols[m], n)                   # match[j] = row of column j, or -1
        return match if int((match >= 0).sum()) == need else None

    best = probe(0.0)
    if best is None:
        raise ValueError("Jacobian pattern is structurally singular: no full matching")
    cand = np.unique(vals)
    lo, hi = 0, len(cand) - 1                      # invariant: cand[lo] is feasible (the smallest entry always is)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        m_ = probe(cand[mid])
        if m_ is not None:
            best, lo = m_, mid
        else:
            hi = mid - 1
    free = row_of < 0
    row_of[free] = best[free]
    return row_of


def static_row_matching(A: sp.csc_matrix, exact_below: int = 1500, forced=None) -> np.ndarray:
    """Return ``rowperm`` with ``rowperm[c]`` = the original row matched to column ``c``.

    ``forced = (rows, cols)``: pairs that must be matched to each other whatever their values.  A DAE passes
    the unit entries of its mass matrix: the matrices factored on this order are ``M - h*gamma*J`` for every
    step size and the row-masked ``M + diag(alg rows) J`` of the consistent initialisation, and the only
    entries of a differential row that survive in ALL of them are the mass entries.

    Maximises prod |a_{rowperm[c], c}| (MC64 job-5 style objective), so the permuted
    matrix has a zero-free, large diagonal.  Structural zeros that are stored explicitly
    get a huge cost and are only used when the pattern leaves no alternative.  Above
    ``exact_below`` unknowns the exact assignment solver is replaced by ``_heavy_matching``.
    """
    n = A.shape[0]
    if forced is not None:
        forced = (np.asarray(forced[0], dtype=np.int64), np.asarray(forced[1], dtype=np.int64))
        if len(forced[0]) == 0:
            forced = None
    if n >= exact_below:
        return _heavy_matching(A, forced)
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching

    W = sp.coo_matrix(A)
    if forced is not None:                       # keep only the forced entry in its row and in its column
        fr = np.zeros(n, dtype=bool)
        fc = np.zeros(n, dtype=bool)
        fr[forced[0]] = True
        fc[forced[1]] = True
        pair_col = np.full(n, -1, dtype=np.int64)
        pair_col[forced[0]] = forced[1]
        keep = (~fr[W.row] & ~fc[W.col]) | (pair_col[W.row] == W.col)
        W = sp.coo_matrix((W.data[keep], (W.row[keep], W.col[keep])), shape=(n, n))
    absd = np.abs(W.data)
    colmax = np.zeros(n)
    np.maximum.at(colmax, W.col, absd)
    colmax[colmax == 0.0] = 1.0
    with np.errstate(divide="ignore"):
        cost = np.log(colmax[W.col]) - np.log(absd)
    cost[~np.isfinite(cost)] = 1e6  # explicit zeros: last resort
    cost = cost + 1.0  # strictly positive weights (the matcher drops explicit 0 weights)
    G = sp.csr_matrix((cost, (W.row, W.col)), shape=(n, n))
    try:
        r, c = min_weight_full_bipartite_matching(G)
    except ValueError as e:  # structurally singular
        raise ValueError("Jacobian pattern is structurally singular: no full matching") from e
    rowperm = np.empty(n, dtype=np.int64)
    rowperm[c] = r
    return rowperm


# --------------------------------------------------------------------------------------
# 2. ordering
# --------------------------------------------------------------------------------------
def _sym_pattern(B: sp.spmatrix) -> sp.csr_matrix:
    n = B.shape[0]
    coo = sp.coo_matrix(B)            # stored entries, INCLUDING explicit zeros: they are structural
    P = sp.csr_matrix((np.ones(coo.nnz), (coo.row, coo.col)), shape=(n, n))
    S = (P + P.T).tocsr()
    S.setdiag(0)
    S.eliminate_zeros()
    S.data[:] = 1.0
    S.sort_indices()
    return S


def _compress(S: sp.csr_matrix) -> Tuple[np.ndarray, int]:
    """Group indistinguishable vertices (identical closed neighbourhoods)."""
    n = S.shape[0]
    indptr, indices = S.indptr, S.indices
    keys: Dict[bytes, int] = {}
    label = np.empty(n, dtype=np.int64)
    for v in range(n):
        nb = indices[indptr[v]:indptr[v + 1]]
        closed = np.union1d(nb, [v]).astype(np.int64)
        k = closed.tobytes()
        g = keys.get(k)
        if g is None:
            g = len(keys)
            keys[k] = g
        label[v] = g
    return label, len(keys)


def _local_min_degree(adj: Dict[int, set], weight: Dict[int, int]) -> List[int]:
    """Exact minimum (weighted external) degree on an explicit elimination graph.

    Used only on small leaf sub-graphs of the dissection, where forming fill cliques
    explicitly is cheap.
    """
    adj = {v: set(a) for v, a in adj.items()}
    order = []
    alive = set(adj)
    while alive:
        v = min(alive, key=lambda x: (sum(weight[y] for y in adj[x]), x))
        nb = adj[v]
        for a in nb:
            adj[a] |= nb
            adj[a].discard(a)
            adj[a].discard(v)
        alive.discard(v)
        del adj[v]
        order.append(v)
    return order


def _bfs_levels(adj: Sequence[np.ndarray], nodes_mask: np.ndarray, start: int) -> List[List[int]]:
    levels = [[start]]
    seen = {start}
    while True:
        nxt = []
        for v in levels[-1]:
            for a in adj[v]:
                a = int(a)
                if nodes_mask[a] and a not in seen:
                    seen.add(a)
                    nxt.append(a)
        if not nxt:
            break
        levels.append(nxt)
    return levels


def _nested_dissection(adj: Sequence[np.ndarray], weight: np.ndarray, leaf: int = 24) -> List[int]:
    """Nested dissection with BFS level-structure separators; returns elimination order."""
    n = len(adj)
    mask = np.zeros(n, dtype=bool)
    out: List[int] = []

    def order_leaf(nodes: List[int]) -> List[int]:
        ns = set(nodes)
        sub = {v: {int(a) for a in adj[v] if int(a) in ns} for v in nodes}
        return _local_min_degree(sub, {v: int(weight[v]) for v in nodes})

    # explicit stack of work items; an item is either ("nodes", list) or ("emit", list)
    work: List[Tuple[str, List[int]]] = [("nodes", list(range(n)))]
    while work:
        kind, nodes = work.pop()
        if kind == "emit":
            out.extend(nodes)
            continue
        if not nodes:
            continue
        if len(nodes) <= leaf:
            out.extend(order_leaf(nodes))
            continue
        mask[:] = False
        mask[nodes] = True
        # connected components first
        start = nodes[0]
        lv = _bfs_levels(adj, mask, start)
        comp = [v for l in lv for v in l]
        if len(comp) < len(nodes):
            cs = set(comp)
            rest = [v for v in nodes if v not in cs]
            work.append(("nodes", rest))
            work.append(("nodes", comp))
            continue
        # pseudo-peripheral start: repeat BFS from a min-degree vertex of the last level
        for _ in range(3):
            last = lv[-1]
            cand = min(last, key=lambda v: len(adj[v]))
            lv2 = _bfs_levels(adj, mask, cand)
            if len(lv2) <= len(lv):
                break
            lv = lv2
        if len(lv) < 3:
            out.extend(order_leaf(nodes))
            continue
        wl = np.array([sum(int(weight[v]) for v in l) for l in lv], dtype=np.float64)
        cum = np.cumsum(wl)
        tot = cum[-1]
        best, best_cost = None, None
        for m in range(1, len(lv) - 1):
            below = cum[m - 1]
            above = tot - cum[m]
            bal = min(below, above) / max(below, above, 1.0)
            if bal < 0.4:
                continue
            cost = wl[m] / (0.5 + bal)
            if best_cost is None or cost < best_cost:
                best, best_cost = m, cost
        if best is None:
            best = int(np.argmin(np.abs(cum - wl / 2 - tot / 2)))
            best = min(max(best, 1), len(lv) - 2)
        sep = lv[best]
        lower = [v for l in lv[:best] for v in l]
        upper = [v for l in lv[best + 1:] for v in l]
        # thin the separator: a separator vertex with no neighbour in `upper` can move to `lower`
        upset = set(upper)
        keep, moved = [], []
        for v in sep:
            if any(int(a) in upset for a in adj[v]):
                keep.append(v)
            else:
                moved.append(v)
        if keep and moved:
            sep, lower = keep, lower + moved
        # elimination order: lower part, upper part, separator last (stack => push reversed)
        work.append(("emit", sep))
        work.append(("nodes", upper))
        work.append(("nodes", lower))
    return out


def _order_mmd_superlu(S: sp.csr_matrix) -> np.ndarray:
    """Multiple minimum degree on pattern(S) borrowed from SuperLU's MMD_AT_PLUS_A."""
    import scipy.sparse.linalg as sla

    n = S.shape[0]
    M = (S + sp.identity(n, format="csr") * (float(n) + 1.0)).tocsc()
    lu = sla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                  options=dict(SymmetricMode=True))
    return np.argsort(lu.perm_c).astype(np.int64)


def _order_amd_candidates(Sc: sp.csr_matrix, wc: np.ndarray, starts: int = 6):
    """Approximate-minimum-degree orderings of the compressed graph (lu/amd.py) from several vertex numberings.  On
    lattice-like patterns the fill of a minimum-degree ordering is decided by how ties are broken (measured on the
    district-heating graph, n = 33 880: 1.36e8 ... 1.93e8 flops over six numberings, SuperLU's MMD 1.59e8), and one
    ordering costs milliseconds, so the caller keeps the best of a few."""
    from .amd import amd_order
    nc = Sc.shape[0]
    rng = np.random.default_rng(20250101)
    out = []
    for trial in range(max(1, starts)):
        perm = np.arange(nc) if trial == 0 else (np.arange(nc)[::-1].copy() if trial == 1 else rng.permutation(nc))
        Sp = Sc[perm][:, perm].tocsr()
        Sp.sort_indices()
        out.append(perm[amd_order(Sp.indptr, Sp.indices, wc[perm])])
    return out


# --------------------------------------------------------------------------------------
# 3. elimination tree, structures, supernodes
# --------------------------------------------------------------------------------------
def _symbolic_structs(S: sp.csr_matrix, order: np.ndarray):
    """Column structures of the Cholesky-like factor of pattern(S) under ``order``."""
    n = S.shape[0]
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.arange(n)
    Sp = S[order][:, order].tocsc()
    Sp.sort_indices()
    indptr, indices = Sp.indptr, Sp.indices
    parent = np.full(n, -1, dtype=np.int64)
    children: List[List[int]] = [[] for _ in range(n)]
    struct: List[Optional[set]] = [None] * n
    for j in range(n):
        col = indices[indptr[j]:indptr[j + 1]]
        s = set(int(i) for i in col[col > j])
        for c in children[j]:
            s |= struct[c]
        s.discard(j)
        struct[j] = s
        if s:
            p = min(s)
            parent[j] = p
            children[p].append(j)
    return parent, children, struct


def _col_counts_py(n, indptr, indices):
    """len(struct[j]) of ``_symbolic_structs`` for the already permuted pattern (CSC, sorted), without the Python
    sets: the structures live in one append-only pool, a marker array removes duplicates.  O(nnz(L)); used to rank
    the ordering candidates, where only the counts matter."""
    cnt = np.zeros(n, np.int64)
    mark = np.full(n, -1, np.int64)
    head = np.full(n, -1, np.int64)          # children of j as a linked list
    nxt = np.full(n, -1, np.int64)
    s0 = np.zeros(n, np.int64)
    s1 = np.zeros(n, np.int64)
    pool = np.empty(max(4 * len(indices), 1024), np.int64)
    top = 0
    for j in range(n):
        start = top
        mark[j] = j
        mn = n
        need = indptr[j + 1] - indptr[j]
        c = head[j]
        while c >= 0:
            need += s1[c] - s0[c]
            c = nxt[c]
        if top + need > len(pool):
            grown = np.empty(max(2 * len(pool), top + need), np.int64)
            grown[:top] = pool[:top]
            pool = grown
        for k in range(indptr[j], indptr[j + 1]):
            i = indices[k]
            if i > j and mark[i] != j:
                mark[i] = j
                pool[top] = i
                top += 1
                if i < mn:
                    mn = i
        c = head[j]
        while c >= 0:
            for q in range(s0[c], s1[c]):
                i = pool[q]
                if mark[i] != j:
                    mark[i] = j
                    pool[top] = i
                    top += 1
                    if i < mn:
                        mn = i
            c = nxt[c]
        s0[j] = start
        s1[j] = top
        cnt[j] = top - start
        if top > start:
            nxt[j] = head[mn]
            head[mn] = j
    return cnt


try:  # pragma: no cover - exercised wherever numba is installed
    import numba as _nb

    _col_counts = _nb.njit(cache=True)(_col_counts_py)
except Exception:  # pragma: no cover
    _col_counts = _col_counts_py


def _struct_sizes(S: sp.csr_matrix, order: np.ndarray) -> np.ndarray:
    """Column counts of the factor of pattern(S) under ``order`` (same numbers as ``_symbolic_structs``)."""
    Sp = S[order][:, order].tocsc()
    Sp.sort_indices()
    return _col_counts(S.shape[0], Sp.indptr.astype(np.int64), Sp.indices.astype(np.int64))


def _postorder(parent: np.ndarray, children: List[List[int]], key) -> np.ndarray:
    n = len(parent)
    post = []
    roots = [j for j in range(n) if parent[j] < 0]
    for r in sorted(roots, key=key):
        stack = [(r, False)]
        while stack:
            v, done = stack.pop()
            if done:
                post.append(v)
                continue
            stack.append((v, True))
            # last child in postorder = largest key => pushed first popped last... push in
            # reverse so that the smallest-key child is visited first
            for c in sorted(children[v], key=key, reverse=True):
                stack.append((c, False))
    return np.array(post, dtype=np.int64)


def front_ld(f, level):
    """Leading dimension of a front of order ``f`` scheduled at ``level`` (0 = warp, else multi-warp);
    mirrors ld_warp / ld_blk of csrc/szb_lu.cu."""
    f = np.asarray(f)
    fp = f + 7      # multi-warp fronts carry 7 rows of zero padding (kPad): unmasked 8x8 tensor-core tiles
    return np.where(np.asarray(level) == 0, f | 1, fp + ((2 - fp) & 3))


@dataclasses.dataclass
class LUSymbolic:
    """Frozen result of the analysis; the integer arrays are uploaded to the device once."""
    n: int
    nnzA: int
    rowp: np.ndarray          # pivot k takes original row rowp[k]
    colp: np.ndarray          # pivot k takes original column colp[k]
    ns: int                   # number of supernodes (postorder = index order)
    sn_piv0: np.ndarray       # first pivot of supernode
    sn_s: np.ndarray          # pivots in supernode
    sn_f: np.ndarray          # front order (s + u)
    sn_ld: np.ndarray         # leading dimension of the front buffer (front_ld)
    sn_parent: np.ndarray     # parent supernode or -1
    idx_ptr: np.ndarray       # front index lists (permuted indices; first s are the pivots)
    idx: np.ndarray
    rel_ptr: np.ndarray       # for supernode c: positions of its update rows in parent's front
    rel: np.ndarray
    child_ptr: np.ndarray
    child: np.ndarray
    a_ptr: np.ndarray         # per supernode: J slots scattered into the front
    a_src: np.ndarray         # J slot (CSC position in the original matrix)
    a_dst: np.ndarray         # r + c*f position in the column-major front
    lp_off: np.ndarray        # offset of the f x s column-major [L11\U11; L21] panel
    up_off: np.ndarray        # offset of the s x u column-major U12 panel
    lu_size: int              # doubles per scenario of factor storage (with padding)
    cb_off: np.ndarray        # offset of the u x (u+1) contribution block in the workspace
    ea_ptr: np.ndarray        # extend-add destination tables, one run of u(u+1) entries per supernode
    ea_dst: np.ndarray
    cb_size: int              # doubles per resident scenario of contribution workspace
    nnz_lu: int               # true nnz(L+U) of the symbolic factor (no supernode padding)
    flops: float              # factor flops per scenario (2 per multiply-add)
    f_max: int
    # schedule: warp-level subtrees, then group-level (a few warps) subtrees, then the CTA-level top
    n_warp_lists: int
    wl_ptr: np.ndarray        # warp work lists (supernode ids, each list in postorder)
    wl: np.ndarray
    n_group_lists: int
    gl_ptr: np.ndarray        # group work lists
    gl: np.ndarray
    top: np.ndarray           # block-level supernodes in postorder
    f_warp_max: int
    f_group_max: int
    f_top_max: int
    ordering: str = ""
    version: int = ANALYSIS_VERSION

    @property
    def order(self) -> np.ndarray:
        """A valid sequential elimination order (children before parents)."""
        return np.concatenate([self.wl, self.gl, self.top]).astype(np.int64)

    def stats(self) -> Dict[str, float]:
        return dict(n=self.n, nnzA=self.nnzA, nnz_lu=self.nnz_lu, lu_size=self.lu_size,
                    cb_size=self.cb_size, flops=self.flops, ns=self.ns, f_max=self.f_max,
                    f_warp_max=self.f_warp_max, f_group_max=self.f_group_max, f_top_max=self.f_top_max,
                    n_warp_sn=len(self.wl), n_group_sn=len(self.gl), n_top=len(self.top),
                    ordering=self.ordering)


def _flops_of(struct_sizes: np.ndarray) -> float:
    s = struct_sizes.astype(np.float64)
    return float(np.sum(2.0 * s * s + s))


def _build(A: sp.csc_matrix, rowperm: np.ndarray, S: sp.csr_matrix, order: np.ndarray,
           relax_frac: float, relax_small: int, warp_front_max: int, n_warps: int,
           group_front_max: int, n_groups: int, name: str, warp_fixed: float = 3000.0,
           group_fixed: float = 60000.0) -> LUSymbolic:
    n = A.shape[0]
    parent, children, struct = _symbolic_structs(S, order)
    size = np.array([len(s) for s in struct], dtype=np.int64)
    # postorder with the child of largest structure last (it is the one a parent can absorb)
    post = _postorder(parent, children, key=lambda v: (size[v], v))
    invpost = np.empty(n, dtype=np.int64)
    invpost[post] = np.arange(n)
    order = order[post]
    parent2 = np.full(n, -1, dtype=np.int64)
    struct2: List[np.ndarray] = [None] * n  # type: ignore
    for newj in range(n):
        oldj = post[newj]
        if parent[oldj] >= 0:
            parent2[newj] = invpost[parent[oldj]]
        struct2[newj] = np.sort(invpost[np.fromiter(struct[oldj], dtype=np.int64,
                                                    count=len(struct[oldj]))])
    parent, struct = parent2, struct2
    size = np.array([len(s) for s in struct], dtype=np.int64)
    nnz_lu = int(np.sum(2 * size + 1))
    flops = _flops_of(size)

    # ---- relaxed supernodes along last-child chains
    sn_first = [0]
    zeros_acc = 0
    for j in range(n - 1):
        j0 = sn_first[-1]
        s_now = j - j0 + 1
        merge = False
        if parent[j] == j + 1:
            extra = int(size[j + 1] - (size[j] - 1))  # rows the earlier columns must pad
            if extra == 0:
                merge = True
            else:
                new_zeros = zeros_acc + extra * s_now
                s_new = s_now + 1
                total = s_new * (s_new + 2 * size[j + 1])
                if s_new <= relax_small or 2 * new_zeros <= relax_frac * total:
                    merge = True
                    zeros_acc = new_zeros
        if not merge:
            sn_first.append(j + 1)
            zeros_acc = 0
    ns = len(sn_first)
    sn_piv0 = np.array(sn_first, dtype=np.int64)
    sn_last = np.append(sn_piv0[1:], n) - 1
    sn_s = sn_last - sn_piv0 + 1
    sn_of = np.repeat(np.arange(ns), sn_s)
    upd = [struct[sn_last[k]] for k in range(ns)]
    sn_u = np.array([len(u) for u in upd], dtype=np.int64)
    sn_f = sn_s + sn_u
    sn_parent = np.array([sn_of[parent[sn_last[k]]] if parent[sn_last[k]] >= 0 else -1
                          for k in range(ns)], dtype=np.int64)

    idx_ptr = np.zeros(ns + 1, dtype=np.int64)
    idx_ptr[1:] = np.cumsum(sn_f)
    idx = np.empty(idx_ptr[-1], dtype=np.int64)
    for k in range(ns):
        idx[idx_ptr[k]:idx_ptr[k] + sn_s[k]] = np.arange(sn_piv0[k], sn_last[k] + 1)
        idx[idx_ptr[k] + sn_s[k]:idx_ptr[k + 1]] = upd[k]

    # children lists and relative indices
    ch: List[List[int]] = [[] for _ in range(ns)]
    for k in range(ns):
        if sn_parent[k] >= 0:
            ch[sn_parent[k]].append(k)
    child_ptr = np.zeros(ns + 1, dtype=np.int64)
    child_ptr[1:] = np.cumsum([len(c) for c in ch])
    child = np.array([c for l in ch for c in l], dtype=np.int64)
    rel_ptr = np.zeros(ns + 1, dtype=np.int64)
    rel_ptr[1:] = np.cumsum(sn_u)
    rel = np.empty(rel_ptr[-1], dtype=np.int64)
    for k in range(ns):
        p = sn_parent[k]
        if p < 0:
            assert sn_u[k] == 0
            continue
        pidx = idx[idx_ptr[p]:idx_ptr[p + 1]]
        # parent's list = contiguous pivots followed by sorted update rows: searchsorted on
        # the whole (sorted) list works because pivots < every update row
        pos = np.searchsorted(pidx, upd[k])
        assert np.all(pidx[pos] == upd[k]), "child update rows must lie inside the parent front"
        rel[rel_ptr[k]:rel_ptr[k + 1]] = pos

    # ---- scatter map of the original matrix entries
    rowp = rowperm[order]      # pivot k <- original row
    colp = order.copy()        # pivot k <- original column
    inv_r = np.empty(n, dtype=np.int64)
    inv_r[rowp] = np.arange(n)
    inv_c = np.empty(n, dtype=np.int64)
    inv_c[colp] = np.arange(n)
    Ac = sp.csc_matrix(A)
    cols = np.repeat(np.arange(n), np.diff(Ac.indptr))
    kk = inv_r[Ac.indices]
    ll = inv_c[cols]
    owner = sn_of[np.minimum(kk, ll)]
    loc_r = np.empty_like(kk)
    loc_c = np.empty_like(ll)
    by_owner = np.argsort(owner, kind="stable")
    a_ptr = np.zeros(ns + 1, dtype=np.int64)
    np.add.at(a_ptr, owner + 1, 1)
    a_ptr = np.cumsum(a_ptr)
    for k in range(ns):
        sel = by_owner[a_ptr[k]:a_ptr[k + 1]]
        fidx = idx[idx_ptr[k]:idx_ptr[k + 1]]
        pr = np.searchsorted(fidx, kk[sel])
        pc = np.searchsorted(fidx, ll[sel])
        assert np.all(fidx[pr] == kk[sel]) and np.all(fidx[pc] == ll[sel]), \
            "matrix entry outside the symbolic front"
        loc_r[sel] = pr
        loc_c[sel] = pc
    a_src = by_owner.astype(np.int64)

    # ---- factor storage
    lp_sz = sn_f * sn_s
    up_sz = sn_u * sn_s
    lp_off = np.zeros(ns, dtype=np.int64)
    up_off = np.zeros(ns, dtype=np.int64)
    off = 0
    for k in range(ns):
        lp_off[k] = off
        off += lp_sz[k]
        up_off[k] = off
        off += up_sz[k]
    lu_size = int(off)

    # ---- schedule: three levels of the tree.  A supernode is warp-level when every front in
    # its subtree fits a warp-private buffer, group-level when every front fits a group buffer,
    # CTA-level otherwise.  Each level's forest is split into independent subtrees that are
    # assigned to work lists by longest-processing-time first.
    sub_fmax = sn_f.copy()
    sub_cost = (sn_s * sn_f * sn_f).astype(np.float64)
    own_cost = sub_cost.copy()
    # warp level: a front costs a few thousand cycles of latency (descriptor, gather, write-out) whatever its
    # size; with s f^2 alone the lists held 25 to 44 fronts and the longest set the level's time (measured:
    # factor 11.25 -> 11.13 ms, sweeps 2.18 -> 2.14 ms with 3000 per front)
    warp_cost = own_cost + warp_fixed
    for k in range(ns):  # postorder: children before parents
        p = sn_parent[k]
        if p >= 0:
            sub_fmax[p] = max(sub_fmax[p], sub_fmax[k])
            sub_cost[p] += sub_cost[k]
    level = np.where(sub_fmax <= warp_front_max, 0, np.where(sub_fmax <= group_front_max, 1, 2))
    # leading dimension of the column-major front in shared memory (csrc/szb_lu.cu: ld_warp / ld_blk):
    # warp-level fronts (lane = row, rank-1 steps): odd, so rows AND columns are conflict-free;
    # multi-warp fronts (8x8 tensor-core tiles): ld >= f + 7 and ld = 2 (mod 4), which puts the four column
    # pairs of an mma.m8n8k4 accumulator fragment on disjoint banks.
    sn_ld = front_ld(sn_f, level)
    a_dst = (loc_r[by_owner] + loc_c[by_owner] * sn_ld[owner[by_owner]]).astype(np.int64)
    # (measured: lists balanced on s * f^2 finish within 5 % of each other; cost models with a per-front fixed
    # part or per-panel barrier costs did not do better, profiles/r1_v6_factor_ncu.md)

    def make_lists(lv: int, n_max: int):
        mine = level == lv
        roots = [k for k in range(ns) if mine[k] and (sn_parent[k] < 0 or not mine[sn_parent[k]])]
        root_of = np.full(ns, -1, dtype=np.int64)
        cost = np.zeros(ns)
        for k in range(ns - 1, -1, -1):
            if mine[k]:
                p = sn_parent[k]
                root_of[k] = k if (p < 0 or not mine[p]) else root_of[p]
        # (group level: 60000 per front on the same scale, measured 11.13 -> 11.06 ms)
        per_front = warp_cost if lv == 0 else (own_cost + group_fixed if lv == 1 else own_cost)
        for k in range(ns):
            if mine[k]:
                cost[root_of[k]] += per_front[k]
        n_lists = max(1, min(n_max, len(roots))) if roots else 0
        loads = np.zeros(max(n_lists, 1))
        assign = {}
        for r in sorted(roots, key=lambda k: -cost[k]):
            w = int(np.argmin(loads))
            assign[r] = w
            loads[w] += cost[r]
        lists: List[List[int]] = [[] for _ in range(n_lists)]
        for k in range(ns):
            if mine[k]:
                lists[assign[int(root_of[k])]].append(k)
        ptr = np.zeros(n_lists + 1, dtype=np.int64)
        if n_lists:
            ptr[1:] = np.cumsum([len(l) for l in lists])
        flat = np.array([k for l in lists for k in l], dtype=np.int64)
        return n_lists, lists, ptr, flat

    n_lists, lists, wl_ptr, wl = make_lists(0, n_warps)
    n_glists, glists, gl_ptr, gl = make_lists(1, n_groups)
    top = np.array([k for k in range(ns) if level[k] == 2], dtype=np.int64)
    is_warp = level == 0

    # ---- contribution-block workspace.  Each work list (and the top list) is a sequential
    # postorder walk, so the classic multifrontal stack is valid per list: when a supernode
    # starts, the blocks of its children (of the same list) are the top of the stack; they are
    # absorbed into the front (which lives in shared memory) and popped, and the supernode's own
    # block is pushed in their place once its pivots are eliminated.  Blocks that cross from a
    # warp list to the top list live in a separate, never-reused region.
    cb_off = np.zeros(ns, dtype=np.int64)
    cb_need = sn_u * (sn_u + 1)   # u x u Schur block + one right-hand-side column
    cursor = 0

    def alloc_list(seq: Sequence[int], base: int) -> int:
        inseq = set(int(s) for s in seq)
        sp_ = 0
        peak = 0
        start_of: Dict[int, int] = {}
        for k in seq:
            mine = [c for c in ch[k] if c in inseq]
            if mine:
                sp_ = min(start_of[c] for c in mine)      # pop the absorbed children
            if sn_parent[k] < 0 or int(sn_parent[k]) not in inseq:
                continue                                   # crosses lists: persistent region
            start_of[k] = sp_
            cb_off[k] = base + sp_
            sp_ += int(cb_need[k])
            peak = max(peak, sp_)
        return peak

    # Blocks that cross lists go one level up or more (a list holds whole subtrees of its level), and the
    # levels are separated by CTA-wide barriers in the factor kernel, so regions of different lifetimes
    # share addresses.  The workspace of a resident CTA is what decides whether the contribution blocks
    # stay in L2 (148 CTAs x cb_size: 114 MB without this sharing on the 1024-bus benchmark, 88 MB with it,
    # against 126 MB of L2).
    #   A  warp -> CTA level blocks      live: warp level .. CTA level
    #   B  warp -> group level blocks    live: warp level .. group level
    #   C  group -> CTA level blocks     live: group level .. CTA level
    #   S0 / S1 / S2  stacks of the warp lists / group lists / CTA-level list, live within their level
    # layout:  [A][B][S0 ...]            warp level
    #          [A][B][S1][C]             group level  (S1 and C over S0)
    #          [A][S2 ......][C]         CTA level    (S2 over B and S1 when it fits, else behind C)
    all_lists = [list(map(int, l)) for l in lists] + [list(map(int, l)) for l in glists] + \
                [list(map(int, top))]
    list_level = [0] * len(lists) + [1] * len(glists) + [2]
    lvl_of = {}
    for seq, lv in zip(all_lists, list_level):
        for k in seq:
            lvl_of[k] = lv
    cross = {"A": [], "B": [], "C": []}
    for seq, lv in zip(all_lists, list_level):
        inseq = set(seq)
        for k in seq:
            p = int(sn_parent[k])
            if p >= 0 and p not in inseq:
                kind = {(0, 2): "A", (0, 1): "B", (1, 2): "C"}.get((lv, lvl_of[p]))
                if kind is None:
                    raise AssertionError("a contribution block crosses lists of one level")
                cross[kind].append(k)
    size_of = lambda ks: int(sum(int(cb_need[k]) for k in ks))
    a_sz, b_sz, c_sz = size_of(cross["A"]), size_of(cross["B"]), size_of(cross["C"])

    def place(ks, base):
        for k in ks:
            cb_off[k] = base
            base += int(cb_need[k])

    def stacks(seqs, base):
        top_ = base
        for seq in seqs:
            top_ += alloc_list(seq, top_)
        return top_ - base

    place(cross["A"], 0)
    place(cross["B"], a_sz)
    zone = a_sz + b_sz
    s0_sz = stacks(all_lists[:len(lists)], zone)
    s1_sz = stacks(all_lists[len(lists):len(lists) + len(glists)], zone)
    place(cross["C"], zone + s1_sz)
    c_end = zone + s1_sz + c_sz
    # the CTA-level stack: measure, then put it over [B][S1] if it fits there
    s2_sz = alloc_list(all_lists[-1], 0)
    s2_base = a_sz if s2_sz <= b_sz + s1_sz else c_end
    alloc_list(all_lists[-1], s2_base)
    cursor = max(zone + s0_sz, c_end, s2_base + s2_sz)
    cb_size = int(cursor)

    # ---- extend-add destinations: element t = i + j*u of the u x (u+1) block of child c goes to
    # rel[i] + col*ld_parent with col = rel[j] (Schur part) or f_parent (right-hand-side column)
    ea_ptr = np.zeros(ns + 1, dtype=np.int64)
    ea_ptr[1:] = np.cumsum(cb_need)
    ea_dst = np.zeros(ea_ptr[-1], dtype=np.int64)
    for k in range(ns):
        p = sn_parent[k]
        uk = int(sn_u[k])
        if p < 0 or uk == 0:
            continue
        rl = rel[rel_ptr[k]:rel_ptr[k + 1]]
        ldp = int(sn_ld[p])
        cols = np.concatenate([rl, [sn_f[p]]])
        ea_dst[ea_ptr[k]:ea_ptr[k + 1]] = (rl[None, :] + cols[:, None] * ldp).reshape(-1)

    return LUSymbolic(
        n=n, nnzA=int(Ac.nnz), rowp=rowp, colp=colp, ns=ns, sn_piv0=sn_piv0, sn_s=sn_s,
        sn_f=sn_f, sn_ld=sn_ld, sn_parent=sn_parent, idx_ptr=idx_ptr, idx=idx, rel_ptr=rel_ptr, rel=rel,
        child_ptr=child_ptr, child=child, a_ptr=a_ptr, a_src=a_src, a_dst=a_dst, lp_off=lp_off,
        up_off=up_off, lu_size=lu_size, cb_off=cb_off, ea_ptr=ea_ptr, ea_dst=ea_dst, cb_size=cb_size,
        nnz_lu=nnz_lu,
        flops=flops, f_max=int(sn_f.max()), n_warp_lists=n_lists, wl_ptr=wl_ptr, wl=wl,
        n_group_lists=n_glists, gl_ptr=gl_ptr, gl=gl, top=top,
        f_warp_max=int(sn_f[level == 0].max()) if (level == 0).any() else 0,
        f_group_max=int(sn_f[level == 1].max()) if (level == 1).any() else 0,
        f_top_max=int(sn_f[level == 2].max()) if (level == 2).any() else 0, ordering=name)


def analyze(A: sp.spmatrix, *, relax_frac: float = 0.3, relax_small: int = 4,
            warp_front_max: int = 32, n_warps: int = 16, group_front_max: int = 72, n_groups: int = 4,
            orderings: Sequence[str] = ("nd", "mmd", "amd?"), warp_fixed: float = 3000.0, group_fixed: float = 60000.0,
            forced_pairs=None) -> LUSymbolic:
    """Analyse the pattern (and representative values) of ``A``; see the module docstring."""
    A = sp.csc_matrix(A)
    A.sort_indices()
    n = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise ValueError("square matrix expected")
    rowperm = static_row_matching(A, forced=forced_pairs)
    B = A.tocsr()[rowperm]
    S = _sym_pattern(B)
    label, nc = _compress(S)
    # compressed graph
    members: List[List[int]] = [[] for _ in range(nc)]
    for v in range(n):
        members[label[v]].append(v)
    coo = S.tocoo()
    Sc = sp.csr_matrix((np.ones(coo.nnz), (label[coo.row], label[coo.col])), shape=(nc, nc))
    Sc.setdiag(0)
    Sc.eliminate_zeros()
    Sc.sort_indices()
    adjc = [Sc.indices[Sc.indptr[v]:Sc.indptr[v + 1]] for v in range(nc)]
    wc = np.array([len(m) for m in members], dtype=np.int64)

    def expand(oc: Sequence[int]) -> np.ndarray:
        return np.array([v for g in oc for v in members[g]], dtype=np.int64)

    cands: List[Tuple[str, np.ndarray]] = []
    maybe_amd = False
    for o in orderings:
        if o == "nd":
            # the BFS dissection is interpreted Python (18 s at 173 800 vertices) and has not won on any pattern of
            # that size (2x the flops of AMD on the district-heating graphs): above ND_MAX it only runs when it is
            # the sole candidate asked for
            if nc <= ND_MAX or all(x == "nd" for x in orderings):
                cands.append(("nd", expand(_nested_dissection(adjc, wc))))
        elif o == "mmd" and nc <= MMD_SUPERLU_MAX:
            try:
                cands.append(("mmd", expand(_order_mmd_superlu(Sc))))
            except Exception:  # pragma: no cover - SuperLU hack unavailable
                pass
        elif o in ("amd", "mmd"):         # "mmd" above the size SuperLU's detour is affordable for
            cands.extend(("amd", expand(oc)) for oc in _order_amd_candidates(Sc, wc))
        elif o == "amd?":                 # AMD only where flops decide (see AMD_MIN_FLOPS); evaluated below
            maybe_amd = True
        elif o == "natural":
            cands.append(("natural", np.arange(n, dtype=np.int64)))
        else:
            raise ValueError(f"unknown ordering {o!r}")
    best = None

    def compete(some):
        nonlocal best
        for name, order in some:
            sizes = _struct_sizes(S, order)
            fl = _flops_of(sizes)
            if os.environ.get("SZB_ANALYZE_VERBOSE"):
                print(f"[analyze] candidate {name}: flops {fl:.4g}, widest column structure {int(sizes.max())}")
            if best is None or fl < best[0]:
                best = (fl, name, order)

    compete(cands)
    if maybe_amd and not any(nm == "amd" for nm, _ in cands) and (best is None or best[0] >= AMD_MIN_FLOPS):
        compete([("amd", expand(oc)) for oc in _order_amd_candidates(Sc, wc)])
    assert best is not None
    _, name, order = best
    return _build(A, rowperm, S, order, relax_frac, relax_small, warp_front_max, n_warps,
                  group_front_max, n_groups, name, float(warp_fixed), float(group_fixed))
