"""nvcc / g++ drivers for generated model sources."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[3]
INCLUDE_DIR = _ROOT / "include"

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA backend cannot build model kernels (there is no CPU "
                       "fallback)")


def compile_model(cu_path, so_path, extra=()) -> str:
    """Compile a generated model source for sm_100a into a shared library."""
    cmd = [_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE_DIR}", *extra, "-o", str(so_path), str(cu_path)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {cu_path}:\n{r.stdout}\n{r.stderr}")
    return str(so_path)
