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
        return self

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
        self.RHS.copy_(r)
        L.check(self.lib.szb_lu_solve(self.h, 1, self._stream()), self.h, "szb_lu_solve")
        x = self.RHS.clone()
        if not on_device:
            x = x.cpu().numpy()
        return x[0] if single else x

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
    return lu.refactor(_values_of(mats, lu))


def solve(A, b, **lu_options):
    """``x = A^-1 b`` (laesolver.py:11-15); a sequence of matrices with ``b`` of shape ``[B, n]`` is a batch."""
    lu = lu_decomposition(A, **lu_options)
    try:
        return lu.solve(b)
    finally:
        lu.close()
