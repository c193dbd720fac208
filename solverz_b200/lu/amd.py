"""Approximate minimum degree ordering on a quotient graph (Amestoy, Davis, Duff: "An approximate minimum degree
ordering algorithm", SIAM J. Matrix Anal. Appl. 17, 1996), for the large patterns of lu/symbolic.py.

Why it is here: the minimum-degree candidate of the analysis used to be read off SuperLU (``splu(...,
permc_spec='MMD_AT_PLUS_A').perm_c``), which also runs SuperLU's numeric factorisation.  On the district-heating
Jacobians (n = 69 400: long chains of pipe cells hanging off a lattice) that factorisation alone takes 20 minutes, 48 at
n = 173 800, for an ordering that costs seconds.  This routine needs O(nnz(L)) integer work and no values.

The graph is the compressed symmetric pattern: vertex ``i`` stands for ``weight[i]`` indistinguishable unknowns, and
degrees are weighted.  Eliminated pivots become elements; a vertex keeps the list of its still uneliminated neighbours
that no element covers, and the list of elements it belongs to.  Degrees are the AMD upper bound
``|A_i| + |L_p \\ i| + sum_e |L_e \\ L_p|``, elements that end up inside the new one are absorbed, and vertices of the
new element with identical lists (hash + comparison) merge into one weighted vertex that is eliminated as a whole --
without that merge the ordering loses 20-70 % in flops on the lattice-like patterns here (measured against SuperLU's
MMD; with it: within -12 ... +6 % from one numbering, better than it from the best of a few numberings).

numba compiles the loop (the reference itself depends on numba); without it the same code runs interpreted."""
from __future__ import annotations

import numpy as np


def _amd_py(n, indptr, indices, weight0):
    weight = weight0.copy()
    """Returns the elimination order (array of n vertex numbers).  ``indptr/indices``: symmetric adjacency without
    the diagonal; ``weight``: positive vertex weights."""
    # variable -> variable adjacency: the caller's CSR rows, compacted in place as entries get covered by elements
    adj = indices.copy()
    alen = np.empty(n, np.int64)
    for i in range(n):
        alen[i] = indptr[i + 1] - indptr[i]
    # element -> variables: append-only pool
    cap = max(4 * len(indices), 1024)
    pool = np.empty(cap, np.int64)
    estart = np.zeros(n, np.int64)
    elen = np.zeros(n, np.int64)             # live entries of element e (by count)
    ewt = np.zeros(n, np.int64)              # their total weight
    ptop = 0
    # variable -> elements: linked cells
    ccap = max(2 * len(indices), 1024)
    cell_e = np.empty(ccap, np.int64)
    cell_n = np.empty(ccap, np.int64)
    ehead = np.full(n, -1, np.int64)
    ctop = 0
    cfree = -1                               # free list of cells
    state = np.zeros(n, np.int8)             # 0 variable, 1 live element, 2 dead (absorbed) element
    total = 0
    for i in range(n):
        total += weight[i]
    # weighted degrees and their bucket lists
    deg = np.zeros(n, np.int64)
    for i in range(n):
        d = 0
        for k in range(indptr[i], indptr[i + 1]):
            d += weight[indices[k]]
        deg[i] = d
    nb = total + 1
    bhead = np.full(nb, -1, np.int64)
    bnext = np.full(n, -1, np.int64)
    bprev = np.full(n, -1, np.int64)
    for i in range(n - 1, -1, -1):
        d = deg[i]
        bnext[i] = bhead[d]
        if bhead[d] >= 0:
            bprev[bhead[d]] = i
        bhead[d] = i
    mindeg = 0
    w = np.zeros(n, np.int64)                # element marks for the |L_e \ L_p| computation
    wflg = 1
    mark = np.zeros(n, np.int64)             # vertex marks: mark[i] == tag  <=>  i in L_p
    tag = 0
    seen = np.zeros(n, np.int64)             # marks of the supervariable comparison (vertices and elements)
    stag = 0
    order = np.empty(n, np.int64)
    left = total
    lp = np.empty(n, np.int64)
    mnext = np.full(n, -1, np.int64)         # chain of the vertices merged into a principal one
    mtail = np.arange(n)
    hhead = np.full(n, -1, np.int64)         # hash buckets of the supervariable detection
    hnext = np.full(n, -1, np.int64)
    hval = np.zeros(n, np.int64)
    ecnt = np.zeros(n, np.int64)
    k = 0
    while k < n:
        while bhead[mindeg] < 0:
            mindeg += 1
        p = bhead[mindeg]
        # unlink p
        bhead[mindeg] = bnext[p]
        if bnext[p] >= 0:
            bprev[bnext[p]] = -1
        j = p
        while j >= 0:                        # p and everything merged into it leave together
            order[k] = j
            k += 1
            j = mnext[j]
        left -= weight[p]
        tag += 1
        mark[p] = tag
        # ---- L_p = A_p  u  union of L_e over e in E_p, without p
        nlp = 0
        for q in range(indptr[p], indptr[p] + alen[p]):
            j = adj[q]
            if state[j] == 0 and mark[j] != tag:
                mark[j] = tag
                lp[nlp] = j
                nlp += 1
        c = ehead[p]
        while c >= 0:
            e = cell_e[c]
            cn = cell_n[c]
            cell_n[c] = cfree                # the cell is released
            cfree = c
            if state[e] == 1:
                for q in range(estart[e], estart[e] + elen[e]):
                    j = pool[q]
                    if state[j] == 0 and mark[j] != tag:
                        mark[j] = tag
                        lp[nlp] = j
                        nlp += 1
                state[e] = 2                 # absorbed into p
            c = cn
        ehead[p] = -1
        state[p] = 1
        if ptop + nlp > len(pool):           # compact the pool (dead elements out), grow if still short
            live = 0
            for e in range(n):
                if state[e] == 1 and e != p:
                    live += elen[e]
            newcap = len(pool)
            while live + nlp > newcap // 2:
                newcap *= 2
            npool = np.empty(newcap, np.int64)
            t = 0
            for e in range(n):
                if state[e] == 1 and e != p:
                    s = estart[e]
                    estart[e] = t
                    for q in range(elen[e]):
                        npool[t] = pool[s + q]
                        t += 1
            pool = npool
            ptop = t
        estart[p] = ptop
        elen[p] = nlp
        wl = 0
        for q in range(nlp):
            pool[ptop + q] = lp[q]
            wl += weight[lp[q]]
        ptop += nlp
        ewt[p] = wl
        # ---- first pass over L_p: unlink from the buckets, prune A_i, count |L_e \ L_p| per neighbouring element
        for q in range(nlp):
            i = lp[q]
            d = deg[i]
            if bprev[i] >= 0:
                bnext[bprev[i]] = bnext[i]
            else:
                bhead[d] = bnext[i]
            if bnext[i] >= 0:
                bprev[bnext[i]] = bprev[i]
            bnext[i] = -1
            bprev[i] = -1
            s = indptr[i]
            t = s
            for r in range(s, s + alen[i]):
                j = adj[r]
                if state[j] == 0 and mark[j] != tag:       # neither eliminated nor covered by the new element
                    adj[t] = j
                    t += 1
            alen[i] = t - s
            c = ehead[i]
            while c >= 0:
                e = cell_e[c]
                if state[e] == 1:
                    if w[e] < wflg:
                        w[e] = ewt[e] + wflg
                    w[e] -= weight[i]
                c = cell_n[c]
        # ---- second pass: rebuild E_i (dead and swallowed elements out, p in), new degree bound, back to the buckets
        for q in range(nlp):
            i = lp[q]
            d = 0
            for r in range(indptr[i], indptr[i] + alen[i]):
                d += weight[adj[r]]
            c = ehead[i]
            keep = -1
            while c >= 0:
                e = cell_e[c]
                cn = cell_n[c]
                drop = True
                if state[e] == 1 and e != p:
                    ext = w[e] - wflg
                    if ext > 0:
                        d += ext
                        drop = False
                    else:
                        state[e] = 2         # L_e inside L_p: aggressive absorption
                if drop:
                    cell_n[c] = cfree
                    cfree = c
                else:
                    cell_n[c] = keep
                    keep = c
                c = cn
            # link p in
            if cfree >= 0:
                c = cfree
                cfree = cell_n[c]
            else:
                if ctop == len(cell_e):
                    ne = np.empty(2 * len(cell_e), np.int64)
                    nn = np.empty(2 * len(cell_e), np.int64)
                    ne[:ctop] = cell_e[:ctop]
                    nn[:ctop] = cell_n[:ctop]
                    cell_e = ne
                    cell_n = nn
                c = ctop
                ctop += 1
            cell_e[c] = p
            cell_n[c] = keep
            ehead[i] = c
            d += wl - weight[i]
            d = min(d, deg[i] + wl - weight[i], left - weight[i])
            deg[i] = d
            h = 0
            for r in range(indptr[i], indptr[i] + alen[i]):
                h += adj[r]
            ne = 0
            c = ehead[i]
            while c >= 0:
                h += cell_e[c]
                ne += 1
                c = cell_n[c]
            ecnt[i] = ne
            h = h % n
            hval[i] = h
            hnext[i] = hhead[h]
            hhead[h] = i
        # ---- indistinguishable vertices of L_p (same A_i, same E_i) merge into one weighted vertex
        for q in range(nlp):
            i = lp[q]
            h = hval[i]
            if hhead[h] < 0:
                continue
            a = hhead[h]
            hhead[h] = -1
            while a >= 0:
                an = hnext[a]
                if state[a] == 0 and an >= 0:
                    stag += 1
                    for r in range(indptr[a], indptr[a] + alen[a]):
                        seen[adj[r]] = stag
                    c = ehead[a]
                    while c >= 0:
                        seen[cell_e[c]] = stag
                        c = cell_n[c]
                    b = an
                    while b >= 0:
                        if state[b] == 0 and alen[b] == alen[a] and ecnt[b] == ecnt[a]:
                            same = True
                            for r in range(indptr[b], indptr[b] + alen[b]):
                                if seen[adj[r]] != stag:
                                    same = False
                                    break
                            if same:
                                c = ehead[b]
                                while c >= 0:
                                    if seen[cell_e[c]] != stag:
                                        same = False
                                        break
                                    c = cell_n[c]
                            if same:
                                deg[a] -= weight[b]
                                weight[a] += weight[b]
                                weight[b] = 0
                                state[b] = 3
                                mnext[mtail[a]] = b
                                mtail[a] = mtail[b]
                        b = hnext[b]
                a = an
        for q in range(nlp):
            i = lp[q]
            if state[i] != 0:
                continue
            d = deg[i]
            if d < 0:
                d = 0
                deg[i] = 0
            bnext[i] = bhead[d]
            bprev[i] = -1
            if bhead[d] >= 0:
                bprev[bhead[d]] = i
            bhead[d] = i
            if d < mindeg:
                mindeg = d
        wflg += total + 1
    return order


try:  # pragma: no cover - exercised wherever numba is installed
    import numba as _nb

    _amd = _nb.njit(cache=True)(_amd_py)
except Exception:  # pragma: no cover
    _amd = _amd_py


def amd_order(indptr, indices, weight=None) -> np.ndarray:
    """Elimination order of the symmetric pattern (CSR, no diagonal); ``weight[i]`` unknowns per vertex."""
    n = len(indptr) - 1
    if weight is None:
        weight = np.ones(n, np.int64)
    return _amd(n, np.ascontiguousarray(indptr, np.int64), np.ascontiguousarray(indices, np.int64),
                np.ascontiguousarray(weight, np.int64))

