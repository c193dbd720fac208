"""ctypes binding of libszb.so (include/szb.h).  There is no fallback: if the library is missing
every entry point of the backend raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE.parent / "csrc" / "libszb.so"

SZB_ABI_VERSION = 4
SZB_FLAG_CONVERGED, SZB_FLAG_MAXIT, SZB_FLAG_SINGULAR, SZB_FLAG_DTMIN, SZB_FLAG_REJECTS = 0, 1, 2, 3, 4
SZB_FLAG_NAN, SZB_FLAG_RUNNING, SZB_FLAG_NOTCONV, SZB_FLAG_EVENT = 5, 6, 7, 8
SLOT_Y, SLOT_F, SLOT_JV, SLOT_RHS, SLOT_PARAM0 = 0, 1, 2, 3, 16
DTYPE_F64, DTYPE_I32 = 0, 1
FLAG_NAMES = {0: "converged", 1: "max_it", 2: "singular", 3: "dt_min", 4: "rejects", 5: "nan", 6: "running",
              7: "not_converged"}
Q_N, Q_NNZ_J, Q_LU_SIZE, Q_N_PARAMS, Q_BATCH, Q_LAUNCHES, Q_WORKSPACE_BYTES = range(7)

_i32p = C.POINTER(C.c_int32)


class LuSymbolicC(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "nnz_a", "ns", "lu_size", "cb_size", "cv_size", "n_warp_lists",
                                         "n_group_lists", "n_top", "f_warp_max", "f_group_max", "f_top_max",
                                         "s_warp_max", "s_group_max", "s_top_max", "threads", "wpg",
                                         "n_sched", "n_ga")] + \
               [(k, _i32p) for k in ("rowp", "colp", "sn_piv0", "sn_s", "sn_f", "idx_ptr", "idx", "rel_ptr",
                                     "rel", "child_ptr", "child", "a_ptr", "a_src", "a_dst", "lp_off",
                                     "up_off", "cb_off", "ea_ptr", "ea_dst", "wl_ptr", "wl", "gl_ptr", "gl", "top",
                                     "fh", "ga")]


class DescC(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n", C.c_int32), ("nnz_j", C.c_int32),
                ("model_lib", C.c_char_p), ("lu", LuSymbolicC), ("mass_vals", C.POINTER(C.c_double))]


_dp = C.POINTER(C.c_double)


class RodasOptC(C.Structure):
    _fields_ = [("t0", C.c_double), ("tend", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
                ("f_savety", C.c_double), ("fac1", C.c_double), ("fac2", C.c_double), ("facmax", C.c_double),
                ("hinit", C.c_double), ("hmax", C.c_double), ("fix_h", C.c_int32), ("dfdt_mode", C.c_int32),
                ("max_steps", C.c_int32), ("n_out", C.c_int32), ("t_out", _dp), ("record_cap", C.c_int32),
                ("s", C.c_int32), ("pord", C.c_int32), ("gamma", C.c_double), ("alpha", _dp),
                ("gammatilde", _dp), ("a", _dp), ("g", _dp), ("b", _dp), ("bd", _dp), ("c", _dp), ("d", _dp),
                ("e", _dp), ("n_events", C.c_int32), ("ev_idx", C.POINTER(C.c_int32)), ("ev_thr", _dp),
                ("ev_dir", C.POINTER(C.c_int32)), ("ev_term", C.POINTER(C.c_int32)), ("event_duration", C.c_double),
                ("ev_cap", C.c_int32), ("resume", C.c_int32), ("stop_after_accepted", C.c_int32)]


class RodasOutC(C.Structure):
    _fields_ = [("y_out", C.c_void_p), ("t_end", C.c_void_p), ("stats", C.c_void_p), ("y_steps", C.c_void_p),
                ("t_steps", C.c_void_p), ("ev_count", C.c_void_p), ("ev_t", C.c_void_p), ("ev_i", C.c_void_p),
                ("ev_y", C.c_void_p), ("n_rows", C.c_void_p)]


_lib = None


class SzbError(RuntimeError):
    pass


def load():
    """Load libszb.so; raise loudly when it is missing (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SzbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                       f"g.build()'` or `make -C solverz_b200/csrc`. The B200 backend has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    H = C.c_void_p
    lib.szb_abi_version.restype = C.c_int
    lib.szb_last_error.restype = C.c_char_p
    lib.szb_last_error.argtypes = [H]
    lib.szb_create.argtypes = [C.POINTER(DescC), C.POINTER(H)]
    lib.szb_destroy.argtypes = [H]
    lib.szb_set_batch.argtypes = [H, C.c_int32]
    lib.szb_bind.argtypes = [H, C.c_int32, C.c_void_p, C.c_int64, C.c_int32]
    lib.szb_eval_F.argtypes = [H, C.c_double, C.c_void_p]
    lib.szb_eval_J.argtypes = [H, C.c_double, C.c_void_p]
    lib.szb_lu_refactor.argtypes = [H, C.c_double, C.c_void_p]
    lib.szb_lu_solve.argtypes = [H, C.c_int32, C.c_void_p]
    lib.szb_lu_status.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.szb_lu_factors.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p]
    lib.szb_lu_residual.argtypes = [H, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
    lib.szb_newton.argtypes = [H, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.szb_newton_host.argtypes = [H, C.c_double, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    lib.szb_rodas.argtypes = [H, C.POINTER(RodasOptC), C.POINTER(RodasOutC), C.c_void_p]
    lib.szb_set_mass_rows.argtypes = [H, C.POINTER(C.c_int32), C.c_int32]
    lib.szb_rodas_state_get.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.szb_rodas_state_set.argtypes = [H, C.c_int32, C.c_double, C.c_int32, C.c_int32]
    lib.szb_rodas_stages.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.szb_lu_refactor_masked_rows.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.szb_query.argtypes = [H, C.c_int32, C.POINTER(C.c_int64)]
    lib.szb_set_time.argtypes = [H, C.c_double]
    lib.szb_set_implicit.argtypes = [H, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
    lib.szb_set_profile.argtypes = [H, C.c_int32]
    lib.szb_get_profile.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    if lib.szb_abi_version() != SZB_ABI_VERSION:
        raise SzbError(f"libszb.so ABI {lib.szb_abi_version()} != binding {SZB_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    if rc != 0:
        msg = load().szb_last_error(handle)
        raise SzbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def lu_symbolic_c(sym, keep: list) -> LuSymbolicC:
    """Convert solverz_b200.lu.symbolic.LUSymbolic into the C struct (int32 arrays kept alive in `keep`)."""
    s = LuSymbolicC()
    if sym is None:
        return s
    ns = sym.ns
    s.n, s.nnz_a, s.ns = sym.n, sym.nnzA, ns
    s.lu_size, s.cb_size, s.cv_size = sym.lu_size, sym.cb_size, int(sym.rel_ptr[-1])
    s.n_warp_lists, s.n_group_lists, s.n_top = sym.n_warp_lists, sym.n_group_lists, len(sym.top)
    s.f_warp_max, s.f_group_max, s.f_top_max = sym.f_warp_max, sym.f_group_max, sym.f_top_max

    def smax(ids):
        ids = np.asarray(ids, dtype=np.int64)
        return int(sym.sn_s[ids].max()) if len(ids) else 0
    s.s_warp_max, s.s_group_max, s.s_top_max = smax(sym.wl), smax(sym.gl), smax(sym.top)
    if max(sym.lu_size, sym.cb_size, int(sym.a_dst.max(initial=0))) >= 2 ** 31:
        raise SzbError("LU schedule exceeds int32 offsets")
    for name in ("rowp", "colp", "sn_piv0", "sn_s", "sn_f", "idx_ptr", "idx", "rel_ptr", "rel", "child_ptr",
                 "child", "a_ptr", "a_src", "a_dst", "lp_off", "up_off", "cb_off", "ea_ptr", "ea_dst", "wl_ptr",
                 "wl", "gl_ptr", "gl", "top"):
        arr = np.ascontiguousarray(np.asarray(getattr(sym, name)), dtype=np.int32)
        if arr.size == 0:
            arr = np.zeros(1, dtype=np.int32)
        keep.append(arr)
        setattr(s, name, arr.ctypes.data_as(_i32p))
    # device schedule: front descriptors + gather tables for exactly this thread layout
    from ..lu.schedule import build_tables, plan_threads
    plan = plan_threads(sym)
    tabs = build_tables(sym, plan["threads"], plan["wpg"])
    s.threads, s.wpg = plan["threads"], plan["wpg"]
    s.n_sched, s.n_ga = tabs["n_sched"], len(tabs["ga"])
    for name in ("fh", "ga"):
        arr = np.ascontiguousarray(tabs[name], dtype=np.int32)
        keep.append(arr)
        setattr(s, name, arr.ctypes.data_as(_i32p))
    return s
