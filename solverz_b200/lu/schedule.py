"""Device schedule of the batch-interleaved multifrontal LU: front descriptors and the child table.

``solverz_b200.lu.symbolic.analyze`` fixes *what* is computed (pivots, ordering, supernodes, the
two-level split of the elimination tree, the contribution-workspace layout).  This module lowers that
result into the flat int32 tables the kernels of ``csrc/szb_lu.cu`` read.  Every table is the same
for all scenarios, and since one warp instruction works on 32 scenarios (lane = scenario) each entry is
decoded once per 32 scenarios:

``fh``  front descriptors, 16 int32 each, in *schedule order* (the warp lists one after the other,
        list ``w`` = the run ``wl_ptr[w]:wl_ptr[w+1]``, then the CTA-level list):
        ``[s, f, piv0, lp_off, up_off, cb_off, a0, a1, c0, c1, idx_ptr, cv_off, k, 0, 0, 0]``
        ``a0:a1`` is the front's run of matrix entries in ``a_src`` / ``a_dst``, ``c0:c1`` its run of
        rows in the child table.
``ct``  child table, 4 int32 per (parent, child) edge, children in schedule order:
        ``[cb_off(child), u(child), rel_ptr(child), cv_off(child)]``

This is what SuperLU's ``dgstrf`` re-derives inside every call of the reference's ``solve`` /
``lu_decomposition`` (Solverz/solvers/laesolver.py:11-45); here it is done once per model.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

__all__ = ["build_tables", "HDR_WORDS", "WARPS"]

HDR_WORDS = 16
WARPS = 8   # warps per CTA of the factor / sweep kernels (csrc/szb_lu.cu: kWarps)


def build_tables(sym) -> Dict[str, np.ndarray]:
    """Returns ``dict(fh=int32[n_sched,16], ct=int32[n_child,4], n_sched, n_child, n_warp_sn)``."""
    if int(sym.n_warp_lists) > WARPS:
        raise ValueError(f"schedule has {sym.n_warp_lists} warp lists, the kernels run {WARPS} warps per CTA")
    seq = [int(k) for k in sym.wl] + [int(k) for k in sym.top]
    fh = np.zeros((max(len(seq), 1), HDR_WORDS), dtype=np.int64)
    ct = []
    for q, k in enumerate(seq):
        kids = [int(c) for c in sym.child[sym.child_ptr[k]:sym.child_ptr[k + 1]]
                if int(sym.sn_f[c] - sym.sn_s[c]) > 0]
        c0 = len(ct)
        for c in kids:
            ct.append((int(sym.cb_off[c]), int(sym.sn_f[c] - sym.sn_s[c]), int(sym.rel_ptr[c]), int(sym.rel_ptr[c])))
        fh[q, :13] = (int(sym.sn_s[k]), int(sym.sn_f[k]), int(sym.sn_piv0[k]), int(sym.lp_off[k]),
                      int(sym.up_off[k]), int(sym.cb_off[k]), int(sym.a_ptr[k]), int(sym.a_ptr[k + 1]),
                      c0, len(ct), int(sym.idx_ptr[k]), int(sym.rel_ptr[k]), k)
    if fh.max(initial=0) >= 2 ** 31:
        raise ValueError("LU schedule exceeds int32 offsets")
    cta = np.array(ct, dtype=np.int64).reshape(-1, 4) if ct else np.zeros((1, 4), dtype=np.int64)
    return dict(fh=np.ascontiguousarray(fh, dtype=np.int32), ct=np.ascontiguousarray(cta, dtype=np.int32),
                n_sched=len(seq), n_child=len(ct), n_warp_sn=len(sym.wl))
