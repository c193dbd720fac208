"""Host walk of the batch-interleaved LU -- TEST INFRASTRUCTURE ONLY.

Compiles ``tests/host_emu/lu_emu.cpp`` (which includes the product's ``csrc/szb_lu_front.cuh`` with
``-DSZB_HOST_EMU``: one lane, one warp) with g++ and drives it through ctypes with the tables of
``solverz_b200.lu.schedule.build_tables``, so the symbolic analysis, the schedule and the index arithmetic of
the device code are checked on a machine without a GPU.  Nothing in the product imports this file.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
_lib = None


def lib():
    global _lib
    if _lib is None:
        out = ROOT / "build" / "host_emu"
        out.mkdir(parents=True, exist_ok=True)
        so = out / "liblu_emu.so"
        src = ROOT / "tests" / "host_emu" / "lu_emu.cpp"
        hdr = ROOT / "solverz_b200" / "csrc" / "szb_lu_front.cuh"
        if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-ffp-contract=off",
                            f"-I{hdr.parent}", str(src), "-o", str(so)], check=True)
        _lib = C.CDLL(str(so))
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Tables:
    def __init__(self, sym):
        from solverz_b200.lu.schedule import WARP_LINES, build_tables
        t = build_tables(sym)
        self.sym = sym
        self.cta_lines = WARP_LINES
        self.fh, self.ct, self.n_sched = t["fh"], t["ct"], t["n_sched"]
        self.a_src, self.a_dst, self.atp, self.tcp, self.tce = t["a_src"], t["a_dst"], t["atp"], t["tcp"], t["tce"]
        self.rel, self.idx = _i32(sym.rel if len(sym.rel) else [0]), _i32(sym.idx)
        self.rowp, self.colp = _i32(sym.rowp), _i32(sym.colp)


def factor(sym, jv, mvals=None, scale=1.0, smask=None, rhs=None, tabs=None, buf_lines=None):
    """Returns (lu, y, monitors) with y = L^{-1} P rhs in pivot order (None without rhs)."""
    T = tabs or Tables(sym)
    buf_lines = T.cta_lines if buf_lines is None else buf_lines
    jv = np.ascontiguousarray(jv, dtype=float)
    mv = None if mvals is None else np.ascontiguousarray(mvals, dtype=float)
    sm = None if smask is None else np.ascontiguousarray(smask, dtype=float)
    r = None if rhs is None else np.ascontiguousarray(rhs, dtype=float)
    lu = np.full(sym.lu_size, np.nan)
    y = np.full(sym.n, np.nan)
    mon = np.zeros(3)
    lib().emu_factor(C.c_int(T.n_sched), _p(T.fh), _p(T.ct), _p(T.a_src), _p(T.a_dst), _p(T.rel), _p(T.idx),
                     _p(T.rowp), _p(T.colp), _p(T.atp), _p(T.tcp), _p(T.tce), C.c_int(int(sym.cb_size)), _p(jv), _p(mv), _p(sm), C.c_double(scale),
                     _p(r), _p(lu), _p(y), _p(mon), C.c_int(int(buf_lines)))
    return lu, (y if rhs is not None else None), mon


def solve(sym, lu, b=None, y=None, tabs=None):
    """x = A^{-1} b from stored factors, or the backward sweep only when y (pivot order) is given."""
    T = tabs or Tables(sym)
    x = np.empty(sym.n)
    if y is None:
        yy = np.full(sym.n, np.nan)
        r = np.ascontiguousarray(b, dtype=float)
        do_fwd = 1
    else:
        yy = np.array(y, dtype=float)
        r = None
        do_fwd = 0
    lib().emu_solve(C.c_int(sym.n), C.c_int(T.n_sched), _p(T.fh), _p(T.ct), _p(T.rel), _p(T.idx), _p(T.rowp),
                    _p(T.colp), C.c_int(int(sym.rel_ptr[-1])), _p(np.ascontiguousarray(lu)), _p(r), _p(yy), _p(x),
                    C.c_int(do_fwd))
    return x
