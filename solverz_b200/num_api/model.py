"""Device-resident numerical models: the ``nAE`` / ``nFDAE`` / ``nDAE`` duck-types of this backend.

Reference counterparts: ``nAE(F,J,p)``, ``nFDAE(F,J,p,nstep)``, ``nDAE(M,F,J,p)``
(Solverz/num_api/num_eqn.py:4-38) and the rendered package's ``mdl`` object
(code_printer/python/module/module_generator.py:277-314).  ``CudaModel`` keeps ``.F``, ``.J``, ``.p``
(and ``.M`` / ``.nstep``) so user code written against the reference keeps working for single
scenarios -- those calls run the same CUDA kernels with B = 1 and return numpy / scipy objects --
while ``session(B)`` exposes the batched device buffers the solvers drive through the C-ABI.
PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sps

from . import lib as L
from ..variables import Address, Vars

__all__ = ["CudaModel", "BatchSession", "load_rendered"]


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise L.SzbError("no CUDA device visible: the solverz_b200 backend has no CPU fallback")
    return torch


class BatchSession:
    """One szb handle + the torch tensors bound to it, for a fixed batch size on one device."""

    def __init__(self, model: "CudaModel", B: int, device: Optional[int] = None, with_lu: bool = True, sym=None):
        """``sym``: a symbolic LU analysis to use instead of the model's own (the pivot fallback of nr_method
        re-analyses on the Jacobian of a scenario the model's static pivots failed on)."""
        torch = _torch()
        self.model = model
        self.B = int(B)
        self.lib = L.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._keep: list = []
        self.h = C.c_void_p()
        desc = L.DescC()
        desc.abi_version = L.SZB_ABI_VERSION
        desc.device = self.device
        desc.n = model.n
        desc.nnz_j = model.nnz_j
        desc.model_lib = str(model.so_path).encode()
        if sym is None:
            sym = model.symbolic(self.device) if with_lu else None
        self.sym = sym
        desc.lu = L.lu_symbolic_c(sym, self._keep)
        if model.mass_vals is not None:
            mv = np.ascontiguousarray(model.mass_vals, dtype=np.float64)
            self._keep.append(mv)
            desc.mass_vals = mv.ctypes.data_as(C.POINTER(C.c_double))
        rc = self.lib.szb_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise L.SzbError(f"szb_create failed ({rc}): {self.lib.szb_last_error(None).decode()}")
        L.check(self.lib.szb_set_batch(self.h, self.B), self.h, "szb_set_batch")
        dev = torch.device("cuda", self.device)
        n, nz = model.n, model.nnz_j
        self.Y = torch.zeros((self.B, n), dtype=torch.float64, device=dev)
        self.F = torch.zeros((self.B, n), dtype=torch.float64, device=dev)
        self.JV = torch.zeros((self.B, nz), dtype=torch.float64, device=dev)
        self.RHS = torch.zeros((self.B, n), dtype=torch.float64, device=dev)
        self.iters = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self._bind(L.SLOT_Y, self.Y)
        self._bind(L.SLOT_F, self.F)
        self._bind(L.SLOT_JV, self.JV)
        self._bind(L.SLOT_RHS, self.RHS)
        self.P: Dict[str, "torch.Tensor"] = {}
        for i, pm in enumerate(model.params_meta):
            dt = torch.float64 if pm["dtype"] == "f64" else torch.int32
            shape = (self.B, pm["len"]) if pm["per_scenario"] else (pm["len"],)
            tns = torch.zeros(shape, dtype=dt, device=dev)
            self.P[pm["name"]] = tns
            self._bind(L.SLOT_PARAM0 + i, tns)
        self.set_params({pm["name"]: pm["value"] for pm in model.params_meta})

    # -- plumbing
    def _bind(self, slot: int, tns):
        dt = L.DTYPE_F64 if tns.dtype.is_floating_point else L.DTYPE_I32
        L.check(self.lib.szb_bind(self.h, slot, C.c_void_p(tns.data_ptr()), tns.numel(), dt), self.h,
                f"szb_bind(slot {slot})")

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if self.h:
            self.lib.szb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data
    def set_params(self, p: Dict[str, np.ndarray]):
        """Copy parameter values host -> device.  Per-scenario parameters accept ``[len]``
        (replicated) or ``[B, len]``; shared parameters ``[len]``."""
        torch = _torch()
        for pm in self.model.params_meta:
            name = pm["name"]
            if name not in p:
                continue
            v = p[name]
            dst = self.P[name]
            if isinstance(v, torch.Tensor):
                src = v.to(device=dst.device, dtype=dst.dtype)
            else:
                src = torch.as_tensor(np.ascontiguousarray(np.asarray(v)).astype(
                    np.float64 if pm["dtype"] == "f64" else np.int32, copy=False)).to(dst.device)
            if pm["per_scenario"]:
                if src.ndim == 1:
                    src = src.reshape(1, -1).expand(self.B, -1)
                if tuple(src.shape) != (self.B, pm["len"]):
                    raise ValueError(f"parameter {name}: expected shape ({self.B}, {pm['len']}) or "
                                     f"({pm['len']},), got {tuple(src.shape)}")
            else:
                src = src.reshape(-1)
                if src.numel() != pm["len"]:
                    raise ValueError(f"parameter {name}: expected {pm['len']} values, got {src.numel()}")
            dst.copy_(src)

    def set_y(self, y):
        torch = _torch()
        if isinstance(y, Vars):
            y = y.array
        src = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y, dtype=np.float64))
        src = src.to(self.Y.device)
        if src.ndim == 1:
            src = src.reshape(1, -1).expand(self.B, -1)
        self.Y.copy_(src)

    # -- operations (thin wrappers over the C-ABI)
    def eval_F(self, t: float = 0.0):
        L.check(self.lib.szb_eval_F(self.h, float(t), self._stream()), self.h, "szb_eval_F")
        return self.F

    def eval_J(self, t: float = 0.0):
        L.check(self.lib.szb_eval_J(self.h, float(t), self._stream()), self.h, "szb_eval_J")
        return self.JV

    def lu_refactor(self, h_gamma: float = 0.0):
        L.check(self.lib.szb_lu_refactor(self.h, float(h_gamma), self._stream()), self.h, "szb_lu_refactor")

    def lu_status(self):
        """Per-scenario status of the last refactorisation (device int32[B]; 0 ok, 2 singular / unstable)."""
        L.check(self.lib.szb_lu_status(self.h, C.c_void_p(self.flags.data_ptr()), self._stream()), self.h,
                "szb_lu_status")
        return self.flags.clone()      # self.flags is also where newton() reports: hand out a snapshot, not the buffer

    def slot_rows_cols(self):
        """Row and column of every Jacobian slot as device int32 tensors (CSC order of the frozen pattern)."""
        if getattr(self, "_slot_rc", None) is None:
            torch = _torch()
            m = self.model
            cols = np.repeat(np.arange(m.n, dtype=np.int32), np.diff(m.indptr))
            self._slot_rc = (torch.as_tensor(m.indices.astype(np.int32)).to(self.Y.device),
                             torch.as_tensor(cols).to(self.Y.device))
        return self._slot_rc

    def lu_residual(self, rhs, x, h_gamma: float = 0.0):
        """``max|rhs - A x| / max|rhs|`` per scenario for A = J (or M - h_gamma J) of the current SLOT_JV."""
        torch = _torch()
        rows, cols = self.slot_rows_cols()
        out = torch.empty(self.B, dtype=torch.float64, device=self.Y.device)
        L.check(self.lib.szb_lu_residual(self.h, float(h_gamma), C.c_void_p(rows.data_ptr()),
                                         C.c_void_p(cols.data_ptr()), C.c_void_p(rhs.data_ptr()),
                                         C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), self._stream()),
                self.h, "szb_lu_residual")
        return out

    def lu_solve(self, nrhs: int = 1):
        L.check(self.lib.szb_lu_solve(self.h, int(nrhs), self._stream()), self.h, "szb_lu_solve")
        return self.RHS

    def set_time(self, t: float):
        L.check(self.lib.szb_set_time(self.h, float(t)), self.h, "szb_set_time")

    def newton(self, tol: float, max_it: int):
        L.check(self.lib.szb_newton(self.h, float(tol), int(max_it), C.c_void_p(self.iters.data_ptr()),
                                    C.c_void_p(self.flags.data_ptr()), self._stream()), self.h, "szb_newton")
        return self.iters, self.flags

    def newton_host(self, tol: float, max_it: int, y_host, params_host: Dict[str, "object"], iters_host,
                    flags_host):
        """End-to-end call with HOST (pinned) buffers: H2D of y0 and the given parameters, the
        whole Newton solve, D2H of y / iters / flags -- one C-ABI call (szb_newton_host)."""
        np_ = len(self.model.params_meta)
        arr = (C.c_void_p * np_)()
        for i, pm in enumerate(self.model.params_meta):
            t = params_host.get(pm["name"])
            arr[i] = C.c_void_p(t.data_ptr()) if t is not None else None
        L.check(self.lib.szb_newton_host(self.h, float(tol), int(max_it), C.c_void_p(y_host.data_ptr()), arr,
                                         np_, C.c_void_p(iters_host.data_ptr()),
                                         C.c_void_p(flags_host.data_ptr()), self._stream()), self.h,
                "szb_newton_host")

    def set_profile(self, on: bool):
        L.check(self.lib.szb_set_profile(self.h, 1 if on else 0), self.h, "szb_set_profile")

    def get_profile(self):
        ms = (C.c_double * 5)()
        ln = (C.c_int64 * 5)()
        un = (C.c_int64 * 5)()
        L.check(self.lib.szb_get_profile(self.h, ms, ln, un), self.h, "szb_get_profile")
        names = ("F", "J", "lu_refactor", "lu_solve", "bookkeeping")
        return {k: dict(ms=ms[i], launches=ln[i], units=un[i]) for i, k in enumerate(names)}

    def query(self, what: int) -> int:
        out = C.c_int64()
        L.check(self.lib.szb_query(self.h, what, C.byref(out)), self.h, "szb_query")
        return out.value


class CudaModel:
    """Numerical model backed by a generated CUDA library."""

    def __init__(self, name: str, kind: str, so_path, n: int, nnz_j: int, indptr, indices, params_meta,
                 y0: np.ndarray, var_slices: Dict[str, slice], mass=None, mass_slots=None, nstep: int = 0,
                 symbolic=None, pkl_path=None, lu_options: Optional[dict] = None, hvp_dir=None):
        self.name = name
        self._hvp_dir = hvp_dir
        self._hvp_model = None
        self.kind = kind
        self.so_path = Path(so_path)
        self.n = int(n)
        self.nnz_j = int(nnz_j)
        self.indptr = np.asarray(indptr, dtype=np.int64)
        self.indices = np.asarray(indices, dtype=np.int64)
        self.params_meta: List[dict] = params_meta
        self.p: Dict[str, np.ndarray] = {pm["name"]: np.array(pm["value"]) for pm in params_meta}
        self.y0 = np.asarray(y0, dtype=float)
        self.a = Address()
        for k, s in var_slices.items():
            self.a.add(k, s.stop - s.start)
        self.mass = mass
        self.mass_vals = None
        if mass is not None:
            mv = np.zeros(self.nnz_j)
            np.add.at(mv, np.asarray(mass_slots, dtype=np.int64), 1.0)
            self.mass_vals = mv
        self.nstep = nstep
        from ..lu.symbolic import ANALYSIS_VERSION
        self._sym = symbolic if getattr(symbolic, "version", 0) == ANALYSIS_VERSION else None   # stale cached analyses are redone
        self._pkl_path = pkl_path
        self.lu_options = dict(lu_options or {})
        self._sessions: Dict[tuple, BatchSession] = {}
        if not self.so_path.exists():
            raise L.SzbError(f"model library {self.so_path} is missing; render the module again "
                             f"(the backend has no CPU fallback)")

    # -- reference-compatible single-scenario callables ------------------------------------
    @property
    def M(self):
        if self.mass is None:
            raise AttributeError("only a DAE has a mass matrix")
        r, c = self.mass
        return sps.csc_array((np.ones(len(r)), (r, c)), shape=(self.n, self.n))

    def _split(self, args):
        if self.kind == "AE":
            if len(args) != 2:
                raise TypeError("F/J of an AE take (y, p)")
            return 0.0, args[0], args[1]
        if len(args) < 3:
            raise TypeError("F/J of a DAE/FDAE take (t, y, p, ...)")
        p = args[2]
        if self.kind == "FDAE" and len(args) > 3:
            # F_(t, y_, p_, y_0, y_1, ...): history vectors become the alias parameters
            p = dict(p)
            for i, yh in enumerate(args[3:]):
                yh = np.asarray(yh.array if isinstance(yh, Vars) else yh, dtype=float)
                for pm in self.params_meta:
                    tag = f"_tag_{i}"
                    if pm["name"].endswith(tag) and pm["name"][:-len(tag)] in self.a:
                        p[pm["name"]] = yh[..., self.a[pm["name"][:-len(tag)]]]
        return float(args[0]), args[1], p

    def _one(self):
        return self.session(1, with_lu=False)

    def F(self, *args):
        t, y, p = self._split(args)
        s = self._one()
        s.set_params(p)
        s.set_y(y)
        return s.eval_F(t)[0].cpu().numpy()

    def J(self, *args):
        t, y, p = self._split(args)
        s = self._one()
        s.set_params(p)
        s.set_y(y)
        jv = s.eval_J(t)[0].cpu().numpy()
        return sps.csc_array((jv, self.indices, self.indptr), shape=(self.n, self.n))

    @property
    def hvp_model(self) -> "CudaModel":
        """The twin model whose Jacobian is ``d(J(y) v)/dy`` (rendered by ``make_hvp=True``); its per-scenario
        parameters ``v__<var>`` carry the vector."""
        if self._hvp_model is None:
            if self._hvp_dir is None:
                raise AttributeError("this module was rendered without make_hvp=True: no HVP (module_printer(..., "
                                     "make_hvp=True) / made_numerical(..., make_hvp=True))")
            self._hvp_model = load_rendered(str(Path(self._hvp_dir) / "__init__.py"))[0]
        return self._hvp_model

    def HVP(self, y, p, v):
        """``HVP(y_, p_, v_)`` of the reference's rendered module (module_printer.py, `make_hvp`): sparse
        ``d(J(y) v)/dy`` of one scenario."""
        hm = self.hvp_model
        s = hm.session(1, with_lu=False)
        v = np.asarray(v.array if isinstance(v, Vars) else v, dtype=float)
        pv = {k: val for k, val in p.items() if k in hm.p}
        for k in self.a.names:
            pv["v__" + k] = v[self.a[k]]
        s.set_params(pv)
        s.set_y(y)
        hv = s.eval_J(0.0)[0].cpu().numpy()
        return sps.csc_array((hv, hm.indices, hm.indptr), shape=(self.n, self.n))

    # -- batched access --------------------------------------------------------------------
    def session(self, B: int, device: Optional[int] = None, with_lu: bool = True) -> BatchSession:
        torch = _torch()
        dev = torch.cuda.current_device() if device is None else int(device)
        key = (int(B), dev, bool(with_lu))
        s = self._sessions.get(key)
        if s is None:
            # keep at most one big session alive per model to bound device memory
            for k in [k for k in self._sessions if k[0] > 1 and k != key]:
                self._sessions.pop(k).close()
            s = BatchSession(self, B, dev, with_lu)
            self._sessions[key] = s
        return s

    def symbolic(self, device: Optional[int] = None):
        """Host symbolic LU analysis, once per model; the representative Jacobian is evaluated on
        the GPU at (y0, default parameters)."""
        if self._sym is None:
            from ..lu.symbolic import analyze
            s = self.session(1, device, with_lu=False)
            s.set_params({pm["name"]: pm["value"] for pm in self.params_meta})
            s.set_y(self.y0)
            jv = s.eval_J(0.0)[0].cpu().numpy()
            if self.mass_vals is not None:
                jv = np.where(self.mass_vals != 0, jv + self.mass_vals, jv)
            if not np.all(np.isfinite(jv)):
                raise L.SzbError("the Jacobian at y0 is not finite; cannot fix the static pivot order")
            A = sps.csc_matrix((jv, self.indices, self.indptr), shape=(self.n, self.n))
            # a DAE pins every differential row to its mass entry (see static_row_matching)
            self._sym = analyze(A, forced_pairs=self.mass, **self.lu_options)
            if self._pkl_path is not None:
                try:
                    with open(self._pkl_path, "rb") as f:
                        d = pickle.load(f)
                    d["symbolic"] = self._sym
                    # atomic replace: other ranks of a multi-GPU job import the same rendered package
                    tmp = f"{self._pkl_path}.{os.getpid()}.tmp"
                    with open(tmp, "wb") as f:
                        pickle.dump(d, f)
                    os.replace(tmp, self._pkl_path)
                except (OSError, pickle.UnpicklingError, EOFError):
                    pass
        return self._sym

    def close(self):
        for s in self._sessions.values():
            s.close()
        self._sessions.clear()
        if self._hvp_model is not None:
            self._hvp_model.close()


def load_rendered(init_file: str):
    """Used by the ``__init__.py`` of a rendered package: returns ``(mdl, y)``."""
    d = Path(init_file).resolve().parent
    with open(d / "param_and_setting.pkl", "rb") as f:
        st = pickle.load(f)
    mdl = CudaModel(st["name"], st["kind"], d / st["so"], st["n"], st["nnz_j"], st["indptr"], st["indices"],
                    st["params_meta"], st["y0"], st["var_slices"], st["mass"], st["mass_slots"], st["nstep"],
                    st.get("symbolic"), d / "param_and_setting.pkl", st.get("lu_options"),
                    d / st["hvp"] if st.get("hvp") else None)
    a = Address()
    for k, s in st["var_slices"].items():
        a.add(k, s.stop - s.start)
    return mdl, Vars(a, st["y0"].copy())
