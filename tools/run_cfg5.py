#!/usr/bin/env python
"""BASELINE.json configs[4] at size: district-heating hydraulic-thermal FDAE on an nx x nx lattice (nx = 158:
24 964 nodes, 49 612 pipes, n = 173 800 unknowns), a batch of time-stepped load profiles through
``fdae_solver(Opt(step_size=60))`` (one batched Newton solve per step, fdesolver.py:94-128).

2048 profiles is the whole-box batch (256 per GPU on 8 GPUs: the factors of one scenario are ~90 MB); this script runs
one GPU's share.  Prints one JSON line: Newton iterations/s and time steps/s over the batch, per-class device times of
the last step, and size-independent checks -- a scenario whose profile never moves stays at the steady state it
starts from, two scenarios with the same profile agree bit for bit, every scenario converges at every step.

Usage (GPU box):  python tools/run_cfg5.py [--nx 158] [--batch 256] [--steps 60]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=158)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--tol", type=float, default=1e-8, help="Opt.ite_tol (absolute, on max |F|)")
    ap.add_argument("--max-it", type=int, default=100, help="Newton iterations per time step at most")
    a = ap.parse_args()
    import torch
    from solverz_b200 import Opt, fdae_solver, module_printer
    from solverz_b200.synthetic import dhs_fdae_model, dhs_load_profiles
    t0 = time.perf_counter()
    m, info = dhs_fdae_model(a.nx, a.nx)
    fdae, y0 = m.create_instance()
    d = Path(__file__).resolve().parents[1] / "build" / "models"
    d.mkdir(parents=True, exist_ok=True)
    name = f"dhs_lattice_{a.nx}x{a.nx}"
    pkg = module_printer(fdae, y0, name, directory=str(d)).render()      # reuses a rendered package (and its analysis)
    sys.path.insert(0, str(d))
    import importlib
    mdl = importlib.import_module(name).mdl
    t_build = time.perf_counter() - t0
    B = a.batch
    dt = info["dt"]
    tst, tsv = dhs_load_profiles(B, dt)
    tsv[0] = 1.0                                         # scenario 0: the nominal load for ever
    tst[2], tsv[2] = tst[1], tsv[1]                      # scenario 2 repeats scenario 1
    mdl.p["sc__ts_t"], mdl.p["sc__ts_v"] = tst, tsv
    t0 = time.perf_counter()
    sess = mdl.session(B)                                # symbolic analysis (if not cached) + schedule upload
    t_sym = time.perf_counter() - t0
    dev = sess.Y.device
    Y0 = torch.as_tensor(np.tile(y0.array, (B, 1))).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opt = Opt(step_size=dt, ite_tol=a.tol, max_it=a.max_it)
    opt.trajectory = "last"                              # 61 nodes x B x n doubles would not fit next to the factors
    sol = fdae_solver(mdl, [0, a.steps * dt], Y0, opt)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = sol.stats
    Y = sol.Y                                            # [B, nt + 1, n] on the device
    y0d = torch.as_tensor(y0.array).to(dev)
    drift = float((Y[0] - y0d[None, :]).abs().max() / y0d.abs().max())
    same = bool(torch.equal(Y[1], Y[2]))
    moved = float((Y[1, -1] - y0d).abs().max())
    its = int(np.asarray(st.nsolve).sum())
    sym = sess.sym
    nnzj, nnzlu, n = int(mdl.nnz_j), int(sym.nnz_lu), int(mdl.n)
    it_bytes = 8 * (2 * nnzj + 2 * nnzlu + 9 * n + 2 * int(info["n_pipe"]))
    peak = 6539.2
    try:
        peak = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    line = {"metric": "batched_newton_iters_per_s", "value": its / wall, "unit": "iters/s",
            "workload": f"district-heating hydraulic-thermal FDAE, {info['n_pipe']} pipes, {info['n_node']} nodes, "
                        f"n={n}, {B} load profiles on one GPU (2048 = 8 x 256 on the box), fdae_solver step 60 s, "
                        f"{a.steps} steps",
            "time_steps_per_s": B * a.steps / wall, "seconds": wall, "newton_iterations": its,
            "iterations_per_step_min_max": [int(np.asarray(st.nsolve).min()), int(np.asarray(st.nsolve).max())],
            "all_converged": bool(np.asarray(st.succeed).all()),
            "converged_scenarios": int(np.asarray(st.succeed).sum()), "ite_tol": a.tol, "max_it": a.max_it,
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "alg_bytes_per_iteration": it_bytes,
                         "achieved": it_bytes * its / wall / 1e9, "frac": it_bytes * its / wall / 1e9 / peak,
                         "note": "whole Newton iteration (F + J + LU refactor + solve + update), algorithmic bytes x "
                                 "iterations / wall time of the whole call"},
            "nominal_scenario_drift_rel": drift, "repeat_scenario_bitwise_equal": same, "loaded_scenario_moved": moved,
            "nnz_J": nnzj, "nnz_LU": nnzlu, "lu_doubles_stored": int(sym.lu_size), "f_top_max": int(sym.f_top_max),
            "lu_flops_per_scenario": float(sym.flops), "build_s": t_build, "analysis_and_session_s": t_sym}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
