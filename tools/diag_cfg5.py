#!/usr/bin/env python
"""Diagnostic for the DHS FDAE at moderate size: iterations and flags per step, residual of the linear solves."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from solverz_b200 import made_numerical
from solverz_b200.synthetic import dhs_fdae_model, dhs_load_profiles
nx, B, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m, info = dhs_fdae_model(nx, nx)
fdae, y0 = m.create_instance()
t_0 = time.perf_counter()
if nx >= 100:      # the rendered package with the precomputed analysis (tools/precompute_symbolic.py)
    import importlib
    from solverz_b200 import module_printer
    d = Path(__file__).resolve().parents[1] / "build" / "models"
    module_printer(fdae, y0, f"dhs_lattice_{nx}x{nx}", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module(f"dhs_lattice_{nx}x{nx}").mdl
else:
    mdl = made_numerical(fdae, y0, sparse=True)
print("model ready", time.perf_counter() - t_0, flush=True)
dt = info["dt"]
tst, tsv = dhs_load_profiles(B, dt)
mdl.p["sc__ts_t"], mdl.p["sc__ts_v"] = tst, tsv
t_0 = time.perf_counter()
s = mdl.session(B)
print("session", time.perf_counter() - t_0, flush=True)
print("n", mdl.n, "nnz_lu", s.sym.nnz_lu, "f_top_max", s.sym.f_top_max, "ordering", s.sym.ordering, flush=True)
s.set_params({k: v for k, v in mdl.p.items() if "_tag_" not in k})
s.set_y(np.tile(y0.array, (B, 1)))
o, ln = [(mdl.a[k].start, mdl.a[k].stop - mdl.a[k].start) for k in mdl.a.names if k == "Tout"][0]
t = 0.0
for k in range(steps):
    s.P["Tout_tag_0"].copy_(s.Y[:, o:o + ln])
    s.set_time(t + dt)
    # one linear solve check at the current point
    s.eval_J(t + dt); F = s.eval_F(t + dt).clone()
    s.RHS.copy_(F); s.lu_refactor(0.0); x = s.lu_solve(1).clone()
    rr = s.lu_residual(F, x)
    st = s.lu_status()
    t0 = time.perf_counter()
    it, fl = s.newton(1e-8, 30)
    torch.cuda.synchronize()
    print(f"step {k}: |F|max before {float(F.abs().max()):.3e} lin relres max {float(rr.max()):.2e} singular {int((st!=0).sum())} "
          f"iters {it.min().item()}..{it.max().item()} flags {sorted(set(fl.cpu().tolist()))} "
          f"|F| after {float(s.eval_F(t + dt).abs().max()):.3e}  {time.perf_counter()-t0:.2f}s", flush=True)
    t += dt
