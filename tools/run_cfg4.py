#!/usr/bin/env python
"""BASELINE.json configs[3] at full size as a functional run (not a bench line): 45 x 45 = 2025 machines
(swing, flux decay, exciter; n = 20 250 unknowns), Rodas (rodas4), a batch of bus-fault scenarios.

Prints one JSON line: accepted steps/s and stage solves/s over the batch, plus two size-independent checks --
an unfaulted scenario stays at the equilibrium it starts from, and two scenarios with the same fault agree bit
for bit.  Usage (GPU box):  python tools/run_cfg4.py [--nx 45] [--batch 64] [--tend 0.12]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=45)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tend", type=float, default=0.12)
    ap.add_argument("--rtol", type=float, default=1e-3)      # SURVEY.md 8d cfg#4: rodas4, rtol 1e-3, atol 1e-6, hinit 1e-5
    ap.add_argument("--atol", type=float, default=1e-6)
    ap.add_argument("--hinit", type=float, default=1e-5)
    a = ap.parse_args()
    import torch
    from solverz_b200 import Opt, Rodas, made_numerical
    from solverz_b200.synthetic import swing_dae_model, swing_fault_scenarios
    t0 = time.perf_counter()
    m, info = swing_dae_model(a.nx, a.nx)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    t_build = time.perf_counter() - t0
    B = a.batch
    fmask, tst, tsv = swing_fault_scenarios(info["nb"], B)
    fmask[0] = 0.0                      # scenario 0: no fault at all
    fmask[2], tst[2], tsv[2] = fmask[1], tst[1], tsv[1]     # scenario 2 repeats scenario 1
    nd.p["fmask"], nd.p["gf__ts_t"], nd.p["gf__ts_v"] = fmask, tst, tsv
    t0 = time.perf_counter()
    sess = nd.session(B)                # symbolic analysis + schedule upload happen here
    t_sym = time.perf_counter() - t0
    Y0 = np.tile(y0.array, (B, 1))
    tsp = np.linspace(0.0, a.tend, 7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = Rodas(nd, tsp, Y0, Opt(rtol=a.rtol, atol=a.atol, hinit=a.hinit))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = sol.stats
    drift = float(np.max(np.abs(sol.Y[0] - y0.array[None, :])))
    same = bool(np.array_equal(sol.Y[1], sol.Y[2]))
    moved = float(np.max(np.abs(sol.Y[1, -1] - y0.array)))
    peak = 6539.2
    try:
        peak = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    nnzj, nnzlu, n, s_ = int(nd.nnz_j), int(sess.sym.nnz_lu), int(nd.n), 6
    # algorithmic bytes of one Rodas step per scenario: 8 (nnzJ + nnzJ + nnzLU + s nnzLU + (6 s + 6) n)
    step_bytes = 8 * (2 * nnzj + (1 + s_) * nnzlu + (6 * s_ + 6) * n)
    attempts = int(st.nstep.sum() + st.nreject.sum())
    line = {"metric": "rodas_accepted_steps_per_s", "value": float(st.nstep.sum() / dt), "unit": "steps/s",
            "workload": f"swing/flux-decay/exciter DAE, {info['nb']} machines, n={info['n']}, Rodas rodas4 "
                        f"(rtol {a.rtol}, atol {a.atol}, hinit {a.hinit}), {B} bus-fault scenarios, tspan [0, {a.tend}]",
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "alg_bytes_per_step_per_scenario": step_bytes,
                         "achieved": step_bytes * attempts / dt / 1e9,
                         "frac": step_bytes * attempts / dt / 1e9 / peak,
                         "note": "whole step (1 J + 1 LU + s solves + s F), attempted steps (accepted + rejected) x "
                                 "algorithmic bytes / wall time of the whole call"},
            "nnz_J": nnzj, "attempted_steps": attempts,
            "accepted_steps_per_s": float(st.nstep.sum() / dt), "stage_solves_per_s": float(st.nsolve.sum() / dt),
            "seconds": dt, "steps_min_max": [int(st.nstep.min()), int(st.nstep.max())],
            "rejected_total": int(st.nreject.sum()), "all_finished": bool(st.succeed.all()),
            "equilibrium_drift_unfaulted": drift, "repeat_scenario_bitwise_equal": same,
            "faulted_scenario_moved": moved, "nnz_LU": int(sess.sym.nnz_lu), "f_top_max": int(sess.sym.f_top_max),
            "build_s": t_build, "analysis_s": t_sym}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
