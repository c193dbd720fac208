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
        ld = f | 1
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
        lu[sym.up_off[k]:sym.up_off[k] + u * s] = F[:s, s:].T.reshape(-1, order="F")
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
        Ut = lu[sym.up_off[k]:sym.up_off[k] + u * s].reshape(u, s, order="F")
        ids = sym.idx[sym.idx_ptr[k] + s:sym.idx_ptr[k + 1]]
        r = xb[p0:p0 + s] - Ut.T @ xb[ids]
        for p in range(s - 1, -1, -1):
            r[p] /= Lp[p, p]
            r[:p] -= Lp[:p, p] * r[p]
        xb[p0:p0 + s] = r
    x = np.empty(n)
    x[sym.colp] = xb
    return x
