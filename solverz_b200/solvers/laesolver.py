"""The linear-solver seam of the reference for a batch of matrices that share one sparsity pattern.

Reference: Solverz/solvers/laesolver.py:11-45 -- ``solve(A, b)`` (SciPy ``spsolve``: a full symbolic +
numeric SuperLU factorisation per call) and ``lu_decomposition(A)`` -> object with ``.solve(b)`` and
``.perm_r / .perm_c / .nnz``.  Here the pattern is analysed once on the host (``lu/symbolic.py``) and every
``refactor`` / ``solve`` is one launch of the batched multifrontal kernels (``szb_lu_refactor`` /
``szb_lu_solve`` of the C-ABI) over ``B`` value sets -- no model library is involved, so user-written solvers
can use the seam on their own matrices the way they use the reference's.

    lu = BatchedLU(A0)                  # scipy sparse matrix: the pattern + representative values (static pivots)
    lu.refactor(values)                 # [B, nnz] in the CSC slot order of A0 (``A0.tocsc().data`` layout)
    X = lu.solve(rhs)                   # [B, n]  ->  [B, n]

Pivot safety.  The reference pivots inside every call (SuperLU, partial pivoting, ``diag_pivot_thresh = 1``); the
batched kernels use ONE pivot order fixed from ``A0``.  A matrix of the batch whose values differ enough from
``A0`` can make that order unstable, so nothing is returned unchecked: ``refactor`` reads the per-matrix status
(zero / non-finite pivot), ``solve`` measures ``max|b - A x| / max|b|`` of every matrix on the device
(``szb_lu_residual``), and the matrices that fail either test (``check_tol``, default 1e-9) are re-analysed with
their OWN values and solved again (still on the GPU, one at a time).  What still fails gets the reference's
behaviour for a singular matrix: a ``RuntimeWarning`` and ``nan`` in ``solve`` (SciPy's ``spsolve``), a
``RuntimeError`` from ``lu_decomposition`` for a single matrix (SciPy's ``splu``).  ``lu.unstable`` holds the indices
that needed the fallback.

``lu_decomposition(A)`` / ``solve(A, b)`` accept one matrix (B = 1, numpy in / numpy out, like the reference)
or a list of matrices with one pattern.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import scipy.sparse as sps

from ..lu.symbolic import analyze
from ..num_api import lib as L

__all__ = ["BatchedLU", "lu_decomposition", "solve"]


class BatchedLU:
    def __init__(self, A0, device: Optional[int] = None, **lu_options):
        import torch
        if not torch.cuda.is_available():
            raise L.SzbError("no CUDA device visible: the solverz_b200 backend has no CPU fallback")
        A = sps.csc_matrix(A0)
        A.sort_indices()
        if A.shape[0] != A.shape[1]:
            raise ValueError("BatchedLU needs a square matrix")
        self.n, self.nnz_a = A.shape[0], A.nnz
        self.indptr, self.indices = A.indptr.copy(), A.indices.copy()
        self.sym = analyze(A, **lu_options)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.lib = L.load()
        self._keep: list = []
        self.h = C.c_void_p()
        self.B = 0
        desc = L.DescC()
        desc.abi_version, desc.device, desc.n, desc.nnz_j = L.SZB_ABI_VERSION, self.device, self.n, self.nnz_a
        desc.model_lib = None
        desc.lu = L.lu_symbolic_c(self.sym, self._keep)
        rc = self.lib.szb_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise L.SzbError(f"szb_create failed ({rc}): {self.lib.szb_last_error(None).decode()}")
        self._A0_data = A.data.copy()
        self._lu_options = dict(lu_options)
        self.check_tol = 1e-9
        self.singular = np.zeros(0, dtype=bool)     # per matrix: the static pivots met a zero / non-finite pivot
        self.unstable = np.zeros(0, dtype=np.int64)  # matrices the last solve re-analysed with their own values
        cols = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.indptr))
        dev = torch.device("cuda", self.device)
        self._rows_d = torch.as_tensor(self.indices.astype(np.int32)).to(dev)
        self._cols_d = torch.as_tensor(cols).to(dev)

    # the reference object's attributes (laesolver.py:37-42), for the static pivot order used here
    @property
    def perm_r(self):
        return np.asarray(self.sym.rowp)

    @property
    def perm_c(self):
        return np.asarray(self.sym.colp)

    @property
    def nnz(self):
        return int(self.sym.nnz_lu)

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _set_batch(self, B: int):
        import torch
        if B == self.B:
            return
        L.check(self.lib.szb_set_batch(self.h, B), self.h, "szb_set_batch")
        dev = torch.device("cuda", self.device)
        self.JV = torch.zeros((B, self.nnz_a), dtype=torch.float64, device=dev)
        self.RHS = torch.zeros((B, self.n), dtype=torch.float64, device=dev)
        self.flags = None
        for slot, t in ((L.SLOT_JV, self.JV), (L.SLOT_RHS, self.RHS)):
            L.check(self.lib.szb_bind(self.h, slot, C.c_void_p(t.data_ptr()), t.numel(), L.DTYPE_F64), self.h, "szb_bind")
        self.B = B

    def refactor(self, values):
        """values: ``[B, nnz]`` (numpy or CUDA tensor) in the CSC slot order of the pattern, or ``[nnz]``."""
        import torch
        v = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
        if v.ndim == 1:
            v = v.reshape(1, -1)
        if v.shape[1] != self.nnz_a:
            raise ValueError(f"expected {self.nnz_a} values per matrix, got {v.shape[1]}")
        self._set_batch(int(v.shape[0]))
        self.JV.copy_(v)
        L.check(self.lib.szb_lu_refactor(self.h, 0.0, self._stream()), self.h, "szb_lu_refactor")
        fl = torch.zeros(self.B, dtype=torch.int32, device=self.JV.device)
        L.check(self.lib.szb_lu_status(self.h, C.c_void_p(fl.data_ptr()), self._stream()), self.h, "szb_lu_status")
        self.singular = (fl != 0).cpu().numpy()
        return self

    def _relres(self, rhs, x):
        import torch
        out = torch.empty(self.B, dtype=torch.float64, device=self.JV.device)
        L.check(self.lib.szb_lu_residual(self.h, 0.0, C.c_void_p(self._rows_d.data_ptr()),
                                         C.c_void_p(self._cols_d.data_ptr()), C.c_void_p(rhs.data_ptr()),
                                         C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), self._stream()),
                self.h, "szb_lu_residual")
        return out

    def factors(self, b: int = 0):
        """``(L, U)`` of matrix ``b`` as scipy CSC matrices in PIVOT order: ``A[perm_r][:, perm_c] = L @ U`` with
        ``L`` unit lower triangular (the ``.L`` / ``.U`` of the reference's object, laesolver.py:37-42, for the
        static pivot order used here)."""
        import torch
        sym = self.sym
        nq = int(sym.lu_size)
        buf = torch.empty(nq, dtype=torch.float64, device=self.JV.device)
        out = np.empty(nq)
        L.check(self.lib.szb_lu_factors(self.h, int(b), C.c_void_p(buf.data_ptr()), self._stream()), self.h,
                "szb_lu_factors")
        out[:] = buf.cpu().numpy()
        Lr, Lc, Lv, Ur, Uc, Uv = [], [], [], [], [], []
        for k in range(sym.ns):
            s_, f_ = int(sym.sn_s[k]), int(sym.sn_f[k])
            u_ = f_ - s_
            ids = np.asarray(sym.idx[sym.idx_ptr[k]:sym.idx_ptr[k + 1]], dtype=np.int64)
            Lp = out[sym.lp_off[k]:sym.lp_off[k] + f_ * s_].reshape(f_, s_, order="F")
            U12 = out[sym.up_off[k]:sym.up_off[k] + u_ * s_].reshape(s_, u_, order="F")
            for j in range(s_):
                Lr += ids[j:].tolist()
                Lc += [int(ids[j])] * (f_ - j)
                Lv += [1.0] + Lp[j + 1:, j].tolist()
                Ur += ids[:j + 1].tolist()
                Uc += [int(ids[j])] * (j + 1)
                Uv += Lp[:j, j].tolist() + [1.0 / Lp[j, j]]      # the factor stores 1 / u_jj
            for c in range(u_):
                Ur += ids[:s_].tolist()
                Uc += [int(ids[s_ + c])] * s_
                Uv += U12[:, c].tolist()
        n = self.n
        return (sps.csc_matrix((Lv, (Lr, Lc)), shape=(n, n)), sps.csc_matrix((Uv, (Ur, Uc)), shape=(n, n)))

    @property
    def L(self):
        return self.factors(0)[0]

    @property
    def U(self):
        return self.factors(0)[1]

    def solve(self, b):
        """``[B, n]`` (or ``[n]`` for B = 1) right-hand sides; returns the same kind of array it was given."""
        import torch
        on_device = isinstance(b, torch.Tensor)
        r = b if on_device else torch.as_tensor(np.asarray(b, dtype=np.float64))
        single = r.ndim == 1
        if single:
            r = r.reshape(1, -1)
        if tuple(r.shape) != (self.B, self.n):
            raise ValueError(f"expected right-hand sides of shape ({self.B}, {self.n}), got {tuple(r.shape)}")
        rhs_d = r.to(self.RHS.device)
        self.RHS.copy_(rhs_d)
        L.check(self.lib.szb_lu_solve(self.h, 1, self._stream()), self.h, "szb_lu_solve")
        x = self.RHS.clone()
        self.unstable = np.zeros(0, dtype=np.int64)
        if self.check_tol is not None:
            rr = self._relres(rhs_d.contiguous(), x)
            bad = (~torch.isfinite(rr)) | (rr > self.check_tol) | torch.as_tensor(self.singular, device=rr.device)
            if bool(bad.any()):
                x = self._fallback(torch.nonzero(bad).flatten().cpu().numpy(), rhs_d, x)
        if not on_device:
            x = x.cpu().numpy()
        return x[0] if single else x

    def _fallback(self, bad, rhs_d, x):
        """Matrices the static pivot order failed on: re-analysed with their own values, one at a time."""
        import warnings
        self.unstable = np.asarray(bad, dtype=np.int64)
        failed = []
        for b in bad.tolist():
            vals = self.JV[b].cpu().numpy()
            xb = None
            if np.all(np.isfinite(vals)):
                A = sps.csc_matrix((vals, self.indices, self.indptr), shape=(self.n, self.n))
                try:
                    own = BatchedLU(A, self.device, **self._lu_options)
                    own.check_tol = None
                    own.refactor(vals)
                    cand = own.solve(rhs_d[b])
                    ok = (not own.singular[0]) and float(own._relres(rhs_d[b:b + 1].contiguous(),
                                                                      cand.reshape(1, -1))[0]) <= self.check_tol
                    own.close()
                    if ok:
                        xb = cand
                except ValueError:                         # structurally singular for these values
                    xb = None
            if xb is None:
                failed.append(b)
                x[b] = float("nan")
            else:
                x[b] = xb
        if failed:
            warnings.warn(f"Matrix is exactly singular (or numerically so) for matrices {failed[:8]} of the batch",
                          RuntimeWarning, stacklevel=3)
        return x

    def close(self):
        if self.h:
            self.lib.szb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _values_of(mats, lu: BatchedLU) -> np.ndarray:
    out = np.empty((len(mats), lu.nnz_a))
    for b, M in enumerate(mats):
        Mc = sps.csc_matrix(M)
        Mc.sort_indices()
        if Mc.nnz != lu.nnz_a or not np.array_equal(Mc.indptr, lu.indptr) or not np.array_equal(Mc.indices, lu.indices):
            raise ValueError(f"matrix {b} does not have the pattern of matrix 0")
        out[b] = Mc.data
    return out


def lu_decomposition(A, **lu_options) -> BatchedLU:
    """``A``: one sparse matrix or a sequence of sparse matrices with one pattern (laesolver.py:18-45)."""
    mats = list(A) if isinstance(A, (list, tuple)) else [A]
    lu = BatchedLU(mats[0], **lu_options)
    lu.refactor(_values_of(mats, lu))
    if len(mats) == 1 and lu.singular[0]:
        lu.close()
        raise RuntimeError("Factor is exactly singular")          # what scipy.sparse.linalg.splu raises
    return lu


def solve(A, b, **lu_options):
    """``x = A^-1 b`` (laesolver.py:11-15); a sequence of matrices with ``b`` of shape ``[B, n]`` is a batch."""
    lu = lu_decomposition(A, **lu_options)
    try:
        return lu.solve(b)
    finally:
        lu.close()
