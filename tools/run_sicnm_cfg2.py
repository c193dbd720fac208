#!/usr/bin/env python
"""Batched sicnm on BASELINE.json configs[1] (1024-bus polar power flow) as a functional run (not a bench line):
the scenarios of the bench plus heavily loaded ones (injections scaled up until plain Newton from the flat start
stops converging for some of them).  Prints one JSON line: how many scenarios each solver converged, agreement
of the solutions where both did, sicnm steps and wall time.
Usage (GPU box):  python tools/run_sicnm_cfg2.py [--nx 32] [--batch 256] [--scale 1 6 9 12]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--scale", type=float, nargs="+", default=[1.0, 6.0, 9.0, 12.0])
    a = ap.parse_args()
    import torch
    from solverz_b200 import Opt, made_numerical, nr_method, sicnm
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    t0 = time.perf_counter()
    eqs, y0, info = polar_pf_model(a.nx, a.nx)
    mdl = made_numerical(eqs, y0, sparse=True, make_hvp=True)
    t_build = time.perf_counter() - t0
    B = a.batch
    P, Q = pf_scenarios(info, B)
    Y0 = np.tile(y0.array, (B, 1))
    out = dict(workload=f"polar power flow {a.nx}x{a.nx} (n={mdl.n}), {B} scenarios", build_s=round(t_build, 1),
               nnz_J=int(mdl.nnz_j), nnz_H=int(mdl.hvp_model.nnz_j), runs=[])
    for sc in a.scale:
        mdl.p["Pinj"], mdl.p["Qinj"] = sc * P, sc * Q
        nr = nr_method(mdl, Y0, Opt(ite_tol=1e-8))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        si = sicnm(mdl, Y0, Opt(ite_tol=1e-8))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ok_n, ok_s = np.asarray(nr.stats.succeed), np.asarray(si.stats.succeed)
        s = mdl.session(B)
        s.set_y(si.y)
        res = s.eval_F().abs().amax(dim=1).cpu().numpy()
        both = ok_n & ok_s & (res < 1e-8)
        same = np.abs(si.y[both] - nr.y[both]).max(axis=1) < 1e-6 if both.any() else np.zeros(0, bool)
        out["runs"].append(dict(
            scale=sc, newton_converged=int(ok_n.sum()), newton_iters_max=int(np.max(nr.stats.nstep)),
            sicnm_converged=int((ok_s & (res < 1e-8)).sum()), sicnm_steps_median=int(np.median(si.stats.nstep)),
            sicnm_steps_max=int(np.max(si.stats.nstep)), sicnm_rejects_max=int(np.max(si.stats.nreject)),
            same_solution=int(same.sum()), both=int(both.sum()), sicnm_wall_s=round(dt, 2),
            sicnm_stage_solves_per_s=round(float(np.sum(si.stats.nfeval)) / dt, 1)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
