#!/usr/bin/env python
"""Small driver for compute-sanitizer (racecheck / memcheck of the LU kernels): a 12 x 12 power-flow
lattice (warp-, group- and CTA-level fronts with reduced thresholds), 6 scenarios: Newton (fused factor + forward,
backward), then a stored-L refactorisation + forward / backward solve.

    compute-sanitizer --tool racecheck python tools/sanitize_lu.py
    compute-sanitizer --tool memcheck  python tools/sanitize_lu.py"""
import importlib
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    from solverz_b200 import Opt, module_printer, nr_method
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    d = tempfile.mkdtemp()
    eqs, y0, info = polar_pf_model(12, 12)
    module_printer(eqs, y0, "pf_san", directory=d,
                   lu_options=dict(warp_front_max=8, group_front_max=20, n_warps=4, n_groups=2)).render()
    sys.path.insert(0, d)
    mdl = importlib.import_module("pf_san").mdl
    B = 6
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    sol = nr_method(mdl, np.tile(y0.array, (B, 1)), Opt(ite_tol=1e-8))
    assert sol.stats.succeed.all()
    s = mdl.session(B)
    s.eval_J()
    s.lu_refactor(0.0)
    s.RHS.copy_(torch.ones_like(s.RHS))
    x = s.lu_solve(1)
    torch.cuda.synchronize()
    print("sanitize driver ok:", sol.stats.nstep.tolist(), float(x.abs().max()))


if __name__ == "__main__":
    main()
