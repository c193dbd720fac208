"""Device schedule of the batched multifrontal LU: per-front descriptors and gather tables.

``solverz_b200.lu.symbolic.analyze`` fixes *what* is computed (pivots, ordering, supernodes, the
three-level split of the elimination tree).  This module lowers that result into the two flat int32
tables the kernels of ``csrc/szb_lu.cu`` actually read, so that a front costs one descriptor load
plus three streamed tables instead of a chain of dependent index look-ups:

``fh``  front headers, 16 int32 each, in *schedule order* (warp lists, then group lists, then the
        CTA-level list; list ``w`` of a level is the contiguous run ``ptr[w]:ptr[w+1]``):
        ``[s, f, piv0, lp_off, up_off, cb_off, gaO, itO, gaC, itC, gaR, itR, k, n_src, idx_ptr, 0]``

``ga``  gather table, ``(src, dst)`` int32 pairs.  A front owns ONE block laid out ``[iterations][G]`` (G = threads
        of the group that owns the front; offset / iterations in ``gaO / itO``, the C and R fields of the header
        are zero): in iteration ``it`` thread ``r`` adds the value addressed by ``src`` to front position ``dst``
        (``src == -1``: no-op).  Bits 30-31 of ``src`` name the address space, which makes the value fetch a
        pointer select instead of a branch:
          1  Jacobian slot (scaled / masked / mass-shifted on the fly),
          2  row of the right-hand side          } ignored when no right-hand side rides along
          3  contribution workspace, a child's right-hand-side column }
          0  contribution workspace, Schur part of a child.
        Every destination is owned by exactly one thread, and a thread meets the sources of a destination in a
        fixed order (matrix entry, children in schedule order), so the assembly needs no barrier between
        children, no atomics, and is bitwise reproducible.  Destinations are dealt to the threads by source-count
        class (see ``_schedule_front``), so the per-thread entry counts of the whole stream are balanced: the
        iteration count is within a few per cent of entries / G (three separately padded blocks in first-touch
        order left 35 % of the slots empty).

This is what SuperLU's ``dgstrf`` re-derives inside every call of the reference's ``solve`` /
``lu_decomposition`` (Solverz/solvers/laesolver.py:11-45); here it is done once per model.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

__all__ = ["plan_threads", "build_tables", "NOOP", "TAG_CB", "TAG_J", "TAG_RHS", "TAG_CBR", "OFF_MASK"]

NOOP = 0xFFFFFFFF
TAG_CB, TAG_J, TAG_RHS, TAG_CBR = 0, 1, 2, 3
OFF_MASK = (1 << 30) - 1
HDR_WORDS = 16


def plan_threads(sym) -> Dict[str, int]:
    """Threads per CTA and warps per group for this schedule (one warp per warp list, a whole
    number of warps per group, at least 8 warps when CTA-level fronts exist)."""
    warps = max(int(sym.n_warp_lists), 1)
    wpg = 0
    ngl = int(sym.n_group_lists)
    if ngl > 0:
        per = -(-warps // ngl)
        wpg = max(2, per)
        warps = wpg * ngl
    if len(sym.top) > 0 and warps < 8:
        if ngl > 0:
            wpg = -(-8 // ngl)
            warps = wpg * ngl
        else:
            warps = 8
    if warps > 32:
        raise ValueError(f"schedule needs {warps} warps per CTA (max 32)")
    return dict(threads=32 * warps, wpg=wpg)


def _front_sources(sym, k: int):
    """Sources of front k in assembly order: (dst, src, kind) with kind 0 = O, 1 = C, 2 = R."""
    s, f = int(sym.sn_s[k]), int(sym.sn_f[k])
    ld = int(sym.sn_ld[k])
    a0, a1 = int(sym.a_ptr[k]), int(sym.a_ptr[k + 1])
    dst = [np.asarray(sym.a_dst[a0:a1], dtype=np.int64)]
    src = [np.asarray(sym.a_src[a0:a1], dtype=np.int64) | (TAG_J << 30)]
    kind = [np.zeros(a1 - a0, dtype=np.int64)]
    p0 = int(sym.sn_piv0[k])
    dst.append(np.arange(s, dtype=np.int64) + f * ld)
    src.append(np.asarray(sym.rowp[p0:p0 + s], dtype=np.int64) | (TAG_RHS << 30))
    kind.append(np.zeros(s, dtype=np.int64))
    for c in sym.child[sym.child_ptr[k]:sym.child_ptr[k + 1]]:
        c = int(c)
        uc = int(sym.sn_f[c] - sym.sn_s[c])
        if uc == 0:
            continue
        e0 = int(sym.ea_ptr[c])
        cnt = uc * (uc + 1)
        if int(sym.cb_off[c]) + cnt > OFF_MASK:
            raise ValueError("contribution workspace exceeds 2^30 doubles per resident scenario")
        dst.append(np.asarray(sym.ea_dst[e0:e0 + cnt], dtype=np.int64))
        off = int(sym.cb_off[c]) + np.arange(cnt, dtype=np.int64)
        is_rhs = np.arange(cnt) >= uc * uc
        src.append(np.where(is_rhs, off | (TAG_CBR << 30), off | (TAG_CB << 30)))
        kind.append(np.zeros(cnt, dtype=np.int64))      # one block: the whole stream is balanced together
    return np.concatenate(dst), np.concatenate(src), np.concatenate(kind)


def _schedule_front(dst: np.ndarray, src: np.ndarray, kind: np.ndarray, G: int):
    """Assign sources to (block, iteration, thread): unique owner thread per destination, first-touch
    round-robin so that consecutive sources of the first toucher land on consecutive threads."""
    uniq, first, inv, cnt = np.unique(dst, return_index=True, return_inverse=True, return_counts=True)
    # destinations with the same number of sources are dealt round-robin in first-touch order: every thread gets
    # the same number (+-1) of 3-source, 2-source, 1-source destinations, so the per-thread entry counts -- whose
    # maximum is the iteration count every thread walks -- are balanced (first-touch order alone left 35 % of the
    # slots as padding), and neighbours in first-touch order still sit on neighbouring threads (coalesced sources)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.lexsort((first, -cnt))] = np.arange(len(uniq))
    thread = (rank % G)[inv]
    tabs = []
    for kd in range(3):
        sel = np.flatnonzero(kind == kd)
        if len(sel) == 0:
            tabs.append(np.zeros((0, G, 2), dtype=np.uint32))
            continue
        th = thread[sel]
        order = np.argsort(th, kind="stable")          # sources of one thread stay in source order
        th_sorted = th[order]
        start = np.searchsorted(th_sorted, np.arange(G))
        it_sorted = np.arange(len(sel)) - start[th_sorted]
        iters = int(it_sorted.max()) + 1
        tab = np.empty((iters, G, 2), dtype=np.uint32)
        tab[:, :, 0] = NOOP
        tab[:, :, 1] = 0
        tab[it_sorted, th_sorted, 0] = src[sel][order].astype(np.uint32)
        tab[it_sorted, th_sorted, 1] = dst[sel][order].astype(np.uint32)
        tabs.append(tab)
    return tabs


def build_tables(sym, threads: int, wpg: int) -> Dict[str, np.ndarray]:
    """Returns ``dict(fh=int32[n_sched,16], ga=int32[n_entries,2], n_sched, n_entries, G_of)``."""
    seqs = []
    for w in range(int(sym.n_warp_lists)):
        seqs += [(int(k), 32) for k in sym.wl[sym.wl_ptr[w]:sym.wl_ptr[w + 1]]]
    for g in range(int(sym.n_group_lists)):
        seqs += [(int(k), 32 * wpg) for k in sym.gl[sym.gl_ptr[g]:sym.gl_ptr[g + 1]]]
    seqs += [(int(k), threads) for k in sym.top]
    fh = np.zeros((max(len(seqs), 1), HDR_WORDS), dtype=np.int32)
    blocks = []
    cursor = 0
    for q, (k, G) in enumerate(seqs):
        dst, src, kind = _front_sources(sym, k)
        tabs = _schedule_front(dst, src, kind, G)
        fh[q, :6] = (int(sym.sn_s[k]), int(sym.sn_f[k]), int(sym.sn_piv0[k]), int(sym.lp_off[k]),
                     int(sym.up_off[k]), int(sym.cb_off[k]))
        for b, tab in enumerate(tabs):
            fh[q, 6 + 2 * b] = cursor
            fh[q, 7 + 2 * b] = tab.shape[0]
            blocks.append(tab.reshape(-1, 2))
            cursor += tab.shape[0] * G
        fh[q, 12] = k
        fh[q, 13] = len(dst)
        fh[q, 14] = int(sym.idx_ptr[k])
        if cursor >= 2 ** 31:
            raise ValueError("gather table exceeds int32 offsets")
    ga = np.concatenate(blocks) if blocks else np.zeros((0, 2), dtype=np.uint32)
    if len(ga) == 0:
        ga = np.full((1, 2), NOOP, dtype=np.uint32)
    return dict(fh=fh, ga=np.ascontiguousarray(ga).view(np.int32), n_sched=len(seqs), n_entries=int(cursor),
                G_of=np.array([g for _, g in seqs], dtype=np.int64),
                ld_of=np.array([int(sym.sn_ld[k]) for k, _ in seqs], dtype=np.int64))
