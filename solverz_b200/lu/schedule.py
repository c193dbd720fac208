"""Device schedule of the batch-interleaved multifrontal LU: front descriptors, child table, tile tables.

``solverz_b200.lu.symbolic.analyze`` fixes *what* is computed (pivots, ordering, supernodes, the
two-level split of the elimination tree, the contribution-workspace layout).  This module lowers that
result into the flat int32 tables the kernels of ``csrc/szb_lu.cu`` read.  Every table is the same
for all scenarios, and since one warp instruction works on 32 scenarios (lane = scenario) each entry is
decoded once per 32 scenarios:

``fh``  front descriptors, 16 int32 each, in *schedule order*: the level-0 lists one after the other (list
        ``w`` = the run ``wl_ptr[w]:wl_ptr[w+1]``), the level-1 lists (``n_warp_sn + ml_ptr[w]`` ..), then the
        level-2 (CTA-level) list:
        ``[s, f, piv0, lp_off, up_off, cb_off, a0, a1, c0, c1, idx_ptr, cv_off, atp0, tcp0, k, 0]``
        ``a0:a1`` is the front's run of matrix entries in ``a_src`` / ``a_dst``, ``c0:c1`` its run of
        rows in the child table, ``atp0`` / ``tcp0`` the starts of its tile pointers (-1: not a tiled front).
``ct``  child table, 8 int32 per (parent, child) edge, children in schedule order:
        ``[cb_off(child), u(child), rel_ptr(child), cv_off(child), 0, 0, 0, 0]``
``a_src / a_dst``  matrix entries grouped by front: Jacobian slot and destination.  For a front that is
        assembled in a shared-memory buffer (level 0) the destination is the offset in the factor storage;
        for a TILED front (levels 1 and 2) the entries are sorted by the 8 x 8 tile they fall in and the
        destination is the position inside that tile, ``r + 8 c``.
``atp`` tiled fronts: ``atp[atp0 + tr + tc * nt]`` = first entry of tile (tr, tc); one extra pointer closes
        the last tile.  Tile rows / columns 0..Tt-1 cover the pivots in steps of 8 (the last may be short),
        Tt..nt-1 the update rows from ``s`` on.
``tcp / tce``  tiled fronts: the CONTRIBUTORS of every tile.  ``tcp[tcp0 + tr + tc * nt]`` (``tc = nt`` is the
        right-hand-side column) = first row of ``tce`` for that tile; a row of ``tce`` is one int4
        ``[src, u_c, rowpack, colpack]``: the piece of a child's contribution block that lands in the tile starts
        at workspace offset ``src`` (leading dimension ``u_c``), and nibble ``q`` of ``rowpack`` / ``colpack`` is
        the position inside the tile of the piece's row / column ``q`` (0xF: the piece is shorter).  One load
        per contributor replaces the walk over the child's relative indices.

This is what SuperLU's ``dgstrf`` re-derives inside every call of the reference's ``solve`` /
``lu_decomposition`` (Solverz/solvers/laesolver.py:11-45); here it is done once per model.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

__all__ = ["build_tables", "HDR_WORDS", "WARPS", "WARP_LINES", "TILE"]

HDR_WORDS = 16
WARPS = 8          # warps per CTA of the factor / sweep kernels (csrc/szb_lu.cu: kWarps)
WARP_LINES = 110   # lines of 32 doubles in a warp's shared-memory front buffer (csrc/szb_lu.cu: kWarpLines)
TILE = 8           # csrc/szb_lu_front.cuh: kNB


def _tile_bounds(s: int, f: int) -> np.ndarray:
    """Tile boundaries of a front in front-index space: 0, 8, .., (pivot tiles), s, s + 8, .., f."""
    top = np.arange(0, s, TILE)
    bot = np.arange(s, f, TILE)
    return np.concatenate([top, bot, [f]]).astype(np.int64)


def _pack(pos: np.ndarray) -> int:
    """Nibble q = pos[q] for q < len(pos), 0xF beyond (two's-complement int32)."""
    v = 0
    for q in range(TILE):
        v |= (int(pos[q]) if q < len(pos) else 0xF) << (4 * q)
    return v - (1 << 32) if v >= (1 << 31) else v


def build_tables(sym) -> Dict[str, np.ndarray]:
    """Returns fh, ct, a_src, a_dst, atp, tcp, tce (int32) and the counts n_sched, n_child, n_warp_sn, n_mid_sn."""
    if max(int(sym.n_warp_lists), int(sym.n_mid_lists)) > WARPS:
        raise ValueError(f"schedule has {sym.n_warp_lists} / {sym.n_mid_lists} warp lists, the kernels run "
                         f"{WARPS} warps per CTA")
    if int(sym.f_warp_max) * (int(sym.f_warp_max) + 1) > WARP_LINES:
        raise ValueError(f"level-0 fronts of order {sym.f_warp_max} do not fit the {WARP_LINES}-line warp buffer")
    seq = [int(k) for k in sym.wl] + [int(k) for k in sym.ml] + [int(k) for k in sym.top]
    n_l0 = len(sym.wl)
    fh = np.zeros((max(len(seq), 1), HDR_WORDS), dtype=np.int64)
    a_src = np.asarray(sym.a_src, dtype=np.int64).copy()
    a_dst = np.asarray(sym.a_dst, dtype=np.int64).copy()
    ct, atp, tcp, tce = [], [], [], []
    n_atp = n_tcp = n_tce = 0
    for q, k in enumerate(seq):
        s, f = int(sym.sn_s[k]), int(sym.sn_f[k])
        tiled = q >= n_l0
        kids = [int(c) for c in sym.child[sym.child_ptr[k]:sym.child_ptr[k + 1]]
                if int(sym.sn_f[c] - sym.sn_s[c]) > 0]
        c0 = len(ct)
        for c in kids:
            ct.append((int(sym.cb_off[c]), int(sym.sn_f[c] - sym.sn_s[c]), int(sym.rel_ptr[c]),
                       int(sym.rel_ptr[c]), 0, 0, 0, 0))
        atp0 = tcp0 = -1
        a0, a1 = int(sym.a_ptr[k]), int(sym.a_ptr[k + 1])
        if tiled:
            tb = _tile_bounds(s, f)
            nt = len(tb) - 1
            lp, up = int(sym.lp_off[k]), int(sym.up_off[k])
            d = a_dst[a0:a1]
            in_p = d < up
            off = np.where(in_p, d - lp, d - up)
            c_ = np.where(in_p, off // f, s + off // max(s, 1))
            r_ = np.where(in_p, off % f, off % max(s, 1))
            tr = np.searchsorted(tb, r_, side="right") - 1
            tc = np.searchsorted(tb, c_, side="right") - 1
            tid = tr + tc * nt
            order = np.argsort(tid, kind="stable")
            a_src[a0:a1] = a_src[a0:a1][order]
            a_dst[a0:a1] = ((r_ - tb[tr]) + TILE * (c_ - tb[tc]))[order]
            cnt = np.bincount(tid, minlength=nt * nt)
            atp0 = n_atp
            ptr = a0 + np.concatenate([[0], np.cumsum(cnt)])
            atp.append(ptr)
            n_atp += len(ptr)
            # contributors of every tile (rhs column = tile column nt), children in schedule order
            per_tile = [[] for _ in range(nt * (nt + 1))]
            for c in kids:
                rl = np.asarray(sym.rel[sym.rel_ptr[c]:sym.rel_ptr[c + 1]], dtype=np.int64)
                uc = len(rl)
                cut = np.searchsorted(rl, tb, side="left")     # child rows below each tile boundary
                cbo = int(sym.cb_off[c])
                pieces = [(t, int(cut[t]), int(cut[t + 1])) for t in range(nt) if cut[t + 1] > cut[t]]
                for (t_r, r_lo, r_hi) in pieces:
                    rp = _pack(rl[r_lo:r_hi] - tb[t_r])
                    for (t_c, c_lo, c_hi) in pieces:
                        per_tile[t_r + t_c * nt].append((cbo + r_lo + c_lo * uc, uc, rp, _pack(rl[c_lo:c_hi] - tb[t_c])))
                    per_tile[t_r + nt * nt].append((cbo + r_lo + uc * uc, uc, rp, _pack(np.zeros(1, dtype=np.int64))))
            tcp0 = n_tcp
            ptr = n_tce + np.concatenate([[0], np.cumsum([len(x) for x in per_tile])])
            tcp.append(ptr)
            n_tcp += len(ptr)
            for x in per_tile:
                tce.extend(x)
                n_tce += len(x)
        fh[q, :15] = (s, f, int(sym.sn_piv0[k]), int(sym.lp_off[k]), int(sym.up_off[k]), int(sym.cb_off[k]),
                      a0, a1, c0, len(ct), int(sym.idx_ptr[k]), int(sym.rel_ptr[k]), atp0, tcp0, k)
    if fh.max(initial=0) >= 2 ** 31 or max(n_atp, n_tcp, n_tce) >= 2 ** 29:
        raise ValueError("LU schedule exceeds int32 offsets")

    def flat(parts):
        return np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(1), dtype=np.int32)
    cta = np.array(ct, dtype=np.int64).reshape(-1, 8) if ct else np.zeros((1, 8), dtype=np.int64)
    tcea = np.array(tce, dtype=np.int64).reshape(-1, 4) if tce else np.zeros((1, 4), dtype=np.int64)
    return dict(fh=np.ascontiguousarray(fh, dtype=np.int32), ct=np.ascontiguousarray(cta, dtype=np.int32),
                a_src=np.ascontiguousarray(a_src, dtype=np.int32), a_dst=np.ascontiguousarray(a_dst, dtype=np.int32),
                atp=flat(atp), tcp=flat(tcp), tce=np.ascontiguousarray(tcea, dtype=np.int32),
                n_sched=len(seq), n_child=len(ct), n_warp_sn=len(sym.wl), n_mid_sn=len(sym.ml))
