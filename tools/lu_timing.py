"""Cycles per group of 32 scenarios spent in each phase of lu_factor_kernel (instrumented library, see
tools/lu_timing.sh).  Not a bench: the counters cost registers."""
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=32)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--lu-opt", default="")
a = ap.parse_args()
from bench import build_workload
from solverz_b200.num_api import lib as L
from solverz_b200.synthetic import pf_scenarios
eqs, y0, info, mdir, name = build_workload(a.nx)
sys.path.insert(0, str(mdir))
mdl = importlib.import_module(name).mdl
if a.lu_opt:
    mdl.lu_options = json.loads(a.lu_opt); mdl._sym = None; mdl._pkl_path = None
sess = mdl.session(a.batch, 0)
P, Q = pf_scenarios(info, a.batch)
sess.set_params({"Pinj": P, "Qinj": Q})
Y0 = torch.as_tensor(np.tile(y0.array, (a.batch, 1))).cuda()
lib = L.load()
out = (C.c_ulonglong * 24)()
sess.Y.copy_(Y0); sess.newton(1e-8, 100)
lib.szb_debug_lu_cycles(out, 1)
sess.Y.copy_(Y0); it, fl = sess.newton(1e-8, 100)
lib.szb_debug_lu_cycles(out, 1)
groups = (a.batch + 31) // 32
launches = int(it.max().item())
names = ["zero+rhs | tile: zero+entries", "entries | tile: contributors", "children | tile: k-loop", "diag", "panel",
         "trailing", "write_out | tile: finish+store", "TOTAL (thread 0)"]
v = np.array(list(out), dtype=float) / (groups * launches)
print(f"cycles per group per factorisation ({groups} groups, {launches} launches), {sess.sym.stats()}")
for lv, lab in enumerate(("level 0: warp lists, on-chip fronts", "level 1: warp lists, tiled fronts",
                          "level 2: CTA-level tiled fronts")):
    print(f"  {lab}")
    for i in range(8):
        print(f"      {names[i]:36s} {v[8 * lv + i]:12.0f}")
