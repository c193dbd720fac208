"""diagnostic: batched Rodas on the 3x3 swing ensemble vs the reference golden (run on the GPU box)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from solverz_b200 import Opt, Rodas, made_numerical
from solverz_b200.synthetic import swing_dae_model, swing_fault_scenarios
g = np.load(ROOT / "tests" / "golden" / "swing_ref_3x3.npz")
m, info = swing_dae_model(3, 3)
dae, y0 = m.create_instance()
nd = made_numerical(dae, y0, sparse=True)
B = g["Y"].shape[0]
fmask, tst, tsv = swing_fault_scenarios(info["nb"], B)
nd.p["fmask"] = fmask
nd.p["gf__ts_t"], nd.p["gf__ts_v"] = tst, tsv
opt = Opt(rtol=1e-4, atol=1e-7, hinit=1e-4)
opt.dfdt = "module"
sol = Rodas(nd, g["tspan"], np.tile(y0.array, (B, 1)), opt)
names = ["delta", "omega", "Eq", "Efd", "Ux", "Uy", "Ix", "Iy", "Fx", "Fy"]
for b in range(B):
    d = np.abs(sol.Y[b] - g["Y"][b])
    k = np.unravel_index(d.argmax(), d.shape)
    print(b, "steps", int(sol.stats.nstep[b]), int(g["stats"][b][0]), "rej", int(sol.stats.nreject[b]), int(g["stats"][b][4]),
          "max abs diff %.3e at t-index %d var %s[%d] (value %.4e)" % (d.max(), k[0], names[k[1] // 9], k[1] % 9, g["Y"][b][k]),
          "per-time max", np.array2string(d.max(axis=1), precision=1))
