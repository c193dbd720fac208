#!/usr/bin/env python
"""Build-container helper for the LARGE bench models: renders a model package and stores the host-side symbolic LU
analysis in its param_and_setting.pkl, so that the GPU box does not spend minutes of box time on host integer
work.  The product computes the representative Jacobian J(y0, default p) on the GPU at the first session; here there
is no GPU, so the values come from the serial host-check build of the generated row functions (tests/host_check.py --
test infrastructure, used as a one-off evaluator; the analysis code itself is the product's lu/symbolic.py::analyze
with the same arguments num_api/model.py::symbolic passes).

    python tools/precompute_symbolic.py dhs 158        # -> build/models/dhs_lattice_158x158
    python tools/precompute_symbolic.py swing 45       # -> build/models/swing_lattice_45x45
    python tools/precompute_symbolic.py dhs 100 --force   # redo a stored analysis (after a change of lu/symbolic.py)
    python tools/precompute_symbolic.py pf 32 --dry nd,mmd nd,amd   # compare ordering sets, store nothing
"""
import pickle
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sps

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    kind, nx = sys.argv[1], int(sys.argv[2])
    force = "--force" in sys.argv
    dry = sys.argv[sys.argv.index("--dry") + 1:] if "--dry" in sys.argv else None
    from host_check import host_eval, host_lib
    from solverz_b200 import module_printer
    from solverz_b200.lu.symbolic import analyze
    from solverz_b200.synthetic import dhs_fdae_model, polar_pf_model, swing_dae_model
    t0 = time.time()
    if kind == "dhs":
        m, info = dhs_fdae_model(nx, nx)
        eqs, y0 = m.create_instance()
        name = f"dhs_lattice_{nx}x{nx}"
    elif kind == "swing":
        m, info = swing_dae_model(nx, nx)
        eqs, y0 = m.create_instance()
        name = f"swing_lattice_{nx}x{nx}"
    elif kind == "pf":
        eqs, y0, info = polar_pf_model(nx, nx)
        name = f"pf_lattice_{nx}x{nx}"
    else:
        raise SystemExit("kind: dhs | swing | pf")
    d = ROOT / "build" / "models"
    d.mkdir(parents=True, exist_ok=True)
    pkg = module_printer(eqs, y0, name, directory=str(d)).render()
    print(f"rendered {pkg} in {time.time() - t0:.1f} s", flush=True)
    pkl = Path(pkg) / "param_and_setting.pkl"
    st = pickle.load(open(pkl, "rb"))
    from solverz_b200.lu.symbolic import ANALYSIS_VERSION
    if getattr(st.get("symbolic"), "version", 0) == ANALYSIS_VERSION and not force and dry is None:
        print("analysis already cached")
        return
    with tempfile.TemporaryDirectory() as tmp:
        ir, pat, lib, _ = host_lib(eqs, name + "_hc", tmp)
        hist = {}
        for pname, p in ir.params.items():                 # alias parameters start from the variable's own value
            if "_tag_" in pname and p.per_scenario:
                base = pname.rsplit("_tag_", 1)[0]
                o, ln = ir.var_off[base]
                hist[pname] = np.atleast_2d(y0.array[o:o + ln])
        _, JV = host_eval(lib, ir, pat, y0.array, hist, 0.0)
    jv = JV[0]
    mass = st.get("mass")
    if mass is not None:
        mv = np.zeros(len(jv))
        np.add.at(mv, np.asarray(st["mass_slots"], dtype=np.int64), 1.0)
        jv = np.where(mv != 0, jv + mv, jv)
    A = sps.csc_matrix((jv, st["indices"], st["indptr"]), shape=(st["n"], st["n"]))
    if dry is not None:
        for ords in dry:
            t0 = time.time()
            sym = analyze(A, forced_pairs=mass, **{**(st.get("lu_options") or {}), "orderings": tuple(ords.split(","))})
            print(f"{ords}: {time.time() - t0:.1f} s  ordering={sym.ordering} nnz_lu={sym.nnz_lu} lu_size={sym.lu_size} "
                  f"flops={sym.flops:.4g} f_top_max={sym.f_top_max}", flush=True)
        return
    t0 = time.time()
    sym = analyze(A, forced_pairs=mass, **(st.get("lu_options") or {}))
    print(f"analysis {time.time() - t0:.1f} s: n={sym.n} nnz_lu={sym.nnz_lu} lu_size={sym.lu_size} flops={sym.flops:.3g} "
          f"f_top_max={sym.f_top_max} cb_size={sym.cb_size} ordering={sym.ordering}", flush=True)
    st["symbolic"] = sym
    tmpf = str(pkl) + ".tmp"
    with open(tmpf, "wb") as f:
        pickle.dump(st, f)
    Path(tmpf).replace(pkl)


if __name__ == "__main__":
    main()
