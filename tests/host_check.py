"""TEST INFRASTRUCTURE ONLY: a serial g++ build of the generated row functions (-DSZBM_HOST_CHECK), so that the
front-end lowering, the Jacobian derivation and the CUDA printer can be checked on a machine without a GPU.
Nothing in the product imports this file (the product has no CPU path)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

INCLUDE_DIR = Path(__file__).resolve().parents[1] / "include"


def compile_host_check(cu_path, so_path) -> str:
    cmd = ["g++", "-x", "c++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DSZBM_HOST_CHECK",
           f"-I{INCLUDE_DIR}", "-o", str(so_path), str(cu_path), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {cu_path}:\n{r.stdout}\n{r.stderr}")
    return str(so_path)


def host_lib(eqs, name, tmp_path):
    """(ir, pattern, ctypes library, source) of the host-check build of a lowered model."""
    from solverz_b200.code_printer.cuda import print_model
    from solverz_b200.jacobian import derive_jacobian
    ir = eqs.to_ir()
    pat = derive_jacobian(ir)
    src = print_model(ir, pat, name)
    cu = Path(tmp_path) / f"{name}.cu"
    cu.write_text(src)
    so = compile_host_check(cu, Path(tmp_path) / f"{name}_host.so")
    return ir, pat, ctypes.CDLL(so), src


def host_eval(lib, ir, pat, Y, params, t=0.0):
    """F [B, n] and the Jacobian values [B, nnz] of the host-check library at Y [B, n]."""
    Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=float)
    B = Y.shape[0]
    vals = []
    for p in ir.params.values():
        v = params.get(p.name, p.value)
        v = np.ascontiguousarray(v, dtype=np.float64 if p.dtype == "f64" else np.int32)
        if p.per_scenario and v.ndim == 1:
            v = np.ascontiguousarray(np.tile(v, (B, 1)))
        vals.append(v)
    ptrs = (ctypes.c_void_p * max(len(vals), 1))(*[v.ctypes.data for v in vals])
    F = np.zeros((B, ir.n))
    JV = np.zeros((B, pat.nnz))
    lib.szbm_host_F(B, ctypes.c_double(t), Y.ctypes.data_as(ctypes.c_void_p), ptrs, F.ctypes.data_as(ctypes.c_void_p))
    lib.szbm_host_J(B, ctypes.c_double(t), Y.ctypes.data_as(ctypes.c_void_p), ptrs, JV.ctypes.data_as(ctypes.c_void_p))
    return F, JV
