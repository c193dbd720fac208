"""numpy walk of the multifrontal schedule produced by ``solverz_b200.lu.symbolic.analyze``.

TEST INFRASTRUCTURE ONLY.  It executes, one supernode at a time and in exactly the order
and with exactly the index maps the CUDA kernels in ``solverz_b200/csrc/szb_lu.cu`` use, the
factorisation and the two triangular sweeps, so that the host-side symbolic analysis can be
checked on a machine without a GPU.  Nothing in the product imports this file.
"""
import numpy as np


def factor(sym, jv, mvals=None, scale=1.0):
    """jv: values of the original matrix in CSC slot order. Returns (lu, ok)."""
    lu = np.zeros(sym.lu_size)
    cb = np.zeros(max(sym.cb_size, 1))
    order = list(sym.order)
    done = np.zeros(sym.ns, dtype=bool)
    for k in order:
        s, f = int(sym.sn_s[k]), int(sym.sn_f[k])
        u = f - s
        ld = int(sym.sn_ld[k])
        F = np.zeros(ld * f)
        a0, a1 = sym.a_ptr[k], sym.a_ptr[k + 1]
        v = scale * jv[sym.a_src[a0:a1]]
        if mvals is not None:
            v = v + mvals[sym.a_src[a0:a1]]
        np.add.at(F, sym.a_dst[a0:a1], v)
        F = F.reshape(ld, f, order="F")[:f, :].copy()
        for c in sym.child[sym.child_ptr[k]:sym.child_ptr[k + 1]]:
            assert done[c]
            uc = int(sym.sn_f[c] - sym.sn_s[c])
            r = sym.rel[sym.rel_ptr[c]:sym.rel_ptr[c + 1]]
            C = cb[sym.cb_off[c]:sym.cb_off[c] + uc * uc].reshape(uc, uc, order="F")
            F[np.ix_(r, r)] += C
        for p in range(s):
            piv = F[p, p]
            F[p + 1:, p] *= 1.0 / piv
            F[p + 1:, p + 1:] -= np.outer(F[p + 1:, p], F[p, p + 1:])
        lu[sym.lp_off[k]:sym.lp_off[k] + f * s] = F[:, :s].reshape(-1, order="F")
        lu[sym.up_off[k]:sym.up_off[k] + u * s] = F[:s, s:].reshape(-1, order="F")
        cb[sym.cb_off[k]:sym.cb_off[k] + u * u] = F[s:, s:].reshape(-1, order="F")
        done[k] = True
    return lu


def solve(sym, lu, b):
    n = sym.n
    xb = b[sym.rowp].astype(float).copy()
    cv = np.zeros(max(int(sym.rel_ptr[-1]), 1))
    order = list(sym.order)
    for k in order:
        s, f = int(sym.sn_s[k]), int(sym.sn_f[k])
        u = f - s
        p0 = int(sym.sn_piv0[k])
        r = np.zeros(f)
        r[:s] = xb[p0:p0 + s]
        for c in sym.child[sym.child_ptr[k]:sym.child_ptr[k + 1]]:
            rc = sym.rel[sym.rel_ptr[c]:sym.rel_ptr[c + 1]]
            r[rc] += cv[sym.rel_ptr[c]:sym.rel_ptr[c + 1]]
        Lp = lu[sym.lp_off[k]:sym.lp_off[k] + f * s].reshape(f, s, order="F")
        for p in range(s):
            r[p + 1:] -= Lp[p + 1:, p] * r[p]
        xb[p0:p0 + s] = r[:s]
        cv[sym.rel_ptr[k]:sym.rel_ptr[k + 1]] = r[s:]
    for k in reversed(order):
        s, f = int(sym.sn_s[k]), int(sym.sn_f[k])
        u = f - s
        p0 = int(sym.sn_piv0[k])
        Lp = lu[sym.lp_off[k]:sym.lp_off[k] + f * s].reshape(f, s, order="F")
        U12 = lu[sym.up_off[k]:sym.up_off[k] + u * s].reshape(s, u, order="F")
        ids = sym.idx[sym.idx_ptr[k] + s:sym.idx_ptr[k + 1]]
        r = xb[p0:p0 + s] - U12 @ xb[ids]
        for p in range(s - 1, -1, -1):
            r[p] /= Lp[p, p]
            r[:p] -= Lp[:p, p] * r[p]
        xb[p0:p0 + s] = r
    x = np.empty(n)
    x[sym.colp] = xb
    return x


# ---------------------------------------------------------------------------------------------
# table-driven walk (solverz_b200/lu/schedule.py): assembly through the gather table, blocked
# elimination in the order of szb_lu.cu::eliminate_blocked
# ---------------------------------------------------------------------------------------------
def _row_solve(a, Lb, Ub):
    """a: [m, nb] rows of A21; returns (L rows = a U11^{-1}, Lt rows = L L11^{-1})."""
    nb = a.shape[1]
    l = np.zeros_like(a)
    for j in range(nb):
        l[:, j] = (a[:, j] - l[:, :j] @ Ub[:j, j]) / Ub[j, j]
    x = l.copy()
    for j in range(nb - 1, -1, -1):
        x[:, j] = l[:, j] - x[:, j + 1:] @ Lb[j + 1:, j]
    return l, x


def _lu_small(D):
    D = D.copy()
    nb = D.shape[0]
    for k in range(nb):
        D[k + 1:, k] *= 1.0 / D[k, k]
        D[k + 1:, k + 1:] -= np.outer(D[k + 1:, k], D[k, k + 1:])
    return D


def eliminate_blocked(F, s, f, ncol, NB=4):
    """F: [f, ncol] front; on return the leading s columns hold [L11\\U11] blocks and *unscaled* A21
    below them, rows < s hold U (solved), the trailing block holds the Schur complement; returns L21 rows
    (what the workers stream to the factor storage), shape [f, s]."""
    Lout = np.zeros((f, s))
    npan = (s + NB - 1) // NB
    nb0 = min(NB, s)
    F[:nb0, :nb0] = _lu_small(F[:nb0, :nb0])
    for t in range(npan):
        p0 = t * NB
        nb = min(NB, s - p0)
        r0 = p0 + nb
        nb1 = min(NB, s - r0) if t + 1 < npan else 0
        D = F[p0:r0, p0:r0]
        Lb = np.tril(D, -1) + np.eye(nb)
        Ub = np.triu(D)
        l, x = _row_solve(F[r0:, p0:r0], Lb, Ub)
        Lout[r0:, p0:r0] = l
        A11n_old = F[r0:r0 + nb1, r0:r0 + nb1].copy()
        top = F[p0:r0, r0:ncol].copy()
        F[r0:, r0:ncol] -= x @ top
        if nb1:
            # owned by the diagonal warp: same update, then factor in place
            N = A11n_old - x[:nb1] @ top[:, :nb1]
            assert np.allclose(N, F[r0:r0 + nb1, r0:r0 + nb1], rtol=1e-12, atol=1e-300)
            F[r0:r0 + nb1, r0:r0 + nb1] = _lu_small(N)
    # finalize_u: U12 rows of every panel
    for t in range(npan):
        p0 = t * NB
        nb = min(NB, s - p0)
        Lb = np.tril(F[p0:p0 + nb, p0:p0 + nb], -1) + np.eye(nb)
        F[p0:p0 + nb, p0 + nb:ncol] = np.linalg.solve(Lb, F[p0:p0 + nb, p0 + nb:ncol])
    return Lout


def factor_tables(sym, tabs, jv, rhs=None, mvals=None, scale=1.0, store_l=True):
    """Walks the schedule of ``build_tables``; returns (lu, y) with y = L^{-1} P rhs in pivot order."""
    from solverz_b200.lu.schedule import NOOP, OFF_MASK, TAG_CB, TAG_CBR, TAG_J, TAG_RHS
    lu = np.zeros(sym.lu_size)
    cb = np.full(max(sym.cb_size, 1), np.nan)
    y = np.zeros(sym.n)
    fh, ga = tabs["fh"], tabs["ga"].view(np.uint32)
    with_rhs = rhs is not None
    G_of = tabs["G_of"]
    for q in range(tabs["n_sched"]):
        s, f, p0, lp_off, up_off, cb_off = (int(v) for v in fh[q, :6])
        G = int(G_of[q])
        u = f - s
        ld = int(tabs["ld_of"][q])
        ncol = f + 1
        Fm = np.zeros(ld * ncol)
        owner = {}
        for b in range(1):
            g0, iters = int(fh[q, 6 + 2 * b]), int(fh[q, 7 + 2 * b])
            blk = ga[g0:g0 + iters * G].reshape(iters, G, 2)
            for it in range(iters):
                src, dst = blk[it, :, 0], blk[it, :, 1]
                live = src != NOOP
                for r in np.flatnonzero(live):
                    assert owner.setdefault(int(dst[r]), int(r)) == int(r), "two threads share a destination"
                tag = src >> 30
                off = (src & OFF_MASK).astype(np.int64)
                v = np.zeros(G)
                m = live & (tag == TAG_CB)
                v[m] = cb[off[m]]
                m = live & (tag == TAG_J)
                v[m] = scale * jv[off[m]] + (mvals[off[m]] if mvals is not None else 0.0)
                m = live & (tag == TAG_RHS)
                if with_rhs:
                    v[m] = rhs[off[m]]
                else:
                    live = live & ~m
                m = live & (tag == TAG_CBR)
                if with_rhs:
                    v[m] = cb[off[m]]
                else:
                    live = live & ~m
                Fm[dst[live]] += v[live]
        F = Fm.reshape(ld, ncol, order="F")[:f, :].copy()
        assert np.all(np.isfinite(F)), "front read an unwritten contribution"
        if G == 32:
            Lout = None
            for p in range(s):
                F[p + 1:, p] *= 1.0 / F[p, p]
                F[p + 1:, p + 1:] -= np.outer(F[p + 1:, p], F[p, p + 1:])
        else:
            Lout = eliminate_blocked(F, s, f, ncol)
        panel = F[:, :s].copy()
        if Lout is not None:
            for p in range(s):
                pend = min(s, (p // 4) * 4 + 4)
                panel[pend:, p] = Lout[pend:, p] if store_l else 0.0
        lu[lp_off:lp_off + f * s] = panel.reshape(-1, order="F")
        lu[up_off:up_off + u * s] = F[:s, s:f].reshape(-1, order="F")
        cb[cb_off:cb_off + u * (u + 1)] = F[s:, s:f + 1].reshape(-1, order="F")
        y[p0:p0 + s] = F[:s, f]
    return lu, y
