#!/usr/bin/env python
"""bench.py -- batched Newton iterations/s (F + J + sparse LU) on the synthetic polar power flow.

Workload (BASELINE.json configs[1]): 1024-bus 32x32 lattice + 2 % chords, n = 2048 unknowns,
4096 load scenarios per GPU, flat start, Opt(ite_tol=1e-8).  A *step* is one pass of the hot path
over one batch: every scenario is solved from the flat start to convergence (``szb_newton``).
``value`` = Newton iterations of all scenarios of all ranks / device time, inputs resident in HBM;
``e2e`` = the same through ``szb_newton_host`` with pinned HOST buffers (H2D of y0 and the scenario
loads, D2H of the solutions, iteration counts and flags inside the timed region).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # CPU arm: oracle port on host cores

Multi-GPU: one process per GPU (torchrun), scenarios block-partitioned, no collective in the
iteration; per step one NCCL all_gather of iteration counts / flags and a gather of the
solutions to rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {"cfg2": dict(nx=32, batch=4096), "cfg3": dict(nx=100, batch=2048)}
METRIC = "batched_newton_iters_per_s"
UNIT = "iters/s"
TOL = 1e-8
MAX_IT = 100


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _ncu_traffic(cls: str, nx: int, B: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/ncu_traffic.json); None when there is none."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None, "no capture committed"
    d = json.loads(p.read_text())
    e = d.get(f"{cls}@nx{nx}B{B}")
    if not e:
        return None, "no capture of this kernel class at this workload"
    # a capture only describes the kernel it was taken from: the entry carries the hash of the kernel sources
    # (tools/ncu_traffic.py writes it); after any change to them the number is reported as stale, not as traffic
    if e.get("source_sha") != kernel_source_sha():
        return None, f"stale: captured on kernel sources {e.get('source_sha')}, now {kernel_source_sha()}"
    return float(e["dram_bytes_per_launch"]), "ncu capture of the same kernel sources"


def kernel_source_sha() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "solverz_b200" / "csrc").glob("*.cu*")):
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0: float = None, t1: float = None):
        """Samples that arrived inside [t0, t1] (the timed region); when the region was shorter than the sampling
        period, every sample since start() -- the sampler is started before the warm-up steps, so those are
        under the same load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [ln for ts, ln in self.lines if t0 is None or t0 <= ts <= t1 + 0.05]
        if len(lines) < 2:
            lines = [ln for _, ln in self.lines]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ workload
def build_workload(nx: int):
    from solverz_b200 import module_printer
    from solverz_b200.synthetic import polar_pf_model
    eqs, y0, info = polar_pf_model(nx, nx)
    d = ROOT / "build" / "models"
    d.mkdir(parents=True, exist_ok=True)
    name = f"pf_lattice_{nx}x{nx}"
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        module_printer(eqs, y0, name, directory=str(d)).render()
    return eqs, y0, info, d, name


def cpu_solve_block(args):
    """Worker of the CPU arm: oracle port (numpy + SciPy SuperLU) of nr_method on a block."""
    nx, first, count = args
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle.nr import nr_method as oracle_nr
    from oracle.pf import PFOracle
    from solverz_b200.synthetic import lattice_grid, pf_scenarios
    info = _info_cache(nx)
    orc = PFOracle(info["Ybus"], info["pv"], info["pq"])
    y0 = np.concatenate([np.zeros(info["nb"]), np.ones(info["nb"])])
    P, Q = pf_scenarios(info, count, first)
    its = 0
    for b in range(count):
        r = oracle_nr(lambda y: orc.F(y, P[b], Q[b]), orc.J, y0, TOL, MAX_IT)
        its += r["nstep"]
    return its


_INFO = {}


def _info_cache(nx):
    if nx not in _INFO:
        from solverz_b200.synthetic import lattice_grid
        nb, Y = lattice_grid(nx, nx)
        npv = max(1, int(0.1 * nb))
        rng = np.random.default_rng(1)
        _INFO[nx] = dict(nb=nb, Ybus=Y, pv=np.arange(1, 1 + npv), pq=np.arange(1 + npv, nb),
                         Pbase=rng.uniform(-.05, .05, nb), Qbase=rng.uniform(-.02, .02, nb))
    return _INFO[nx]


def cpu_arm(nx: int, sample: int, cores: int, steps: int, warmup: int, kind: str = "auto"):
    """Times the CPU path on `cores` host processes; returns (iters/s, iterations per step, s/step, scenarios
    per step, kind, description).  kind "reference": the reference's own numba module + nr_method from the
    private copy baseline/_ref (baseline/ref_pf.py); "port": the oracle restatement (numpy F/J + SciPy SuperLU),
    used only when the reference copy is not there."""
    import multiprocessing as mp
    if kind in ("auto", "reference"):
        from baseline import ref_pf
        why = ref_pf.available()
        if not why:
            v, its, sps_, used = ref_pf.timed_run(nx, sample, cores, steps, warmup, TOL)
            return v, its, sps_, used, "reference", ("reference numba module (module_printer(jit=True)) + nr_method "
                                                     "-> SciPy SuperLU spsolve, one process per core")
        if kind == "reference":
            raise SystemExit(f"bench.py: reference arm unavailable: {why}")
    per = max(1, sample // cores)
    blocks = [(nx, i * per, per) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(cpu_solve_block, [(nx, 0, 1)] * cores)
        t0 = time.perf_counter()
        total = 0
        for _ in range(steps):
            total += sum(pool.map(cpu_solve_block, blocks))
        dt = time.perf_counter() - t0
    return (total / dt, total // max(steps, 1), dt / max(steps, 1), per * cores, "port",
            "oracle port (numpy F/J + SciPy SuperLU spsolve per iteration), one process per core")


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: cfg2 = configs[1] (the default N=1 line), cfg3 = configs[2] "
                         "(10k-bus grid, 16384 scenarios over 8 GPUs = 2048 per GPU)")
    ap.add_argument("--nx", type=int, default=0, help="lattice side (overrides --config)")
    ap.add_argument("--batch", type=int, default=0, help="scenarios per GPU (overrides --config)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sessions", type=int, default=4,
                    help="end-to-end arm: concurrent sessions (host thread + stream each) the batch is split over, so that "
                         "one session's host<->device copies overlap the other's kernels")
    ap.add_argument("--lu-opt", default="", help="JSON dict of lu.symbolic.analyze options (experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nx = args.nx or CONFIGS[args.config]["nx"]
    B = args.batch or CONFIGS[args.config]["batch"]
    cpu_per_core = max(1, int(round(128 * (32.0 / nx) ** 3)))   # scenarios per core and step of the CPU arm
    workload = f"polar AC power flow, {nx * nx}-bus {nx}x{nx} lattice + 2% chords, n={2 * nx * nx}, " \
               f"{B} load scenarios per GPU, nr_method(ite_tol=1e-8), flat start"

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or min(B, cpu_per_core * cores)   # a few seconds of work per step
        val, its, sps_, used, kind, what = cpu_arm(nx, sample, cores, args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps_ * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample_scenarios_per_step": used},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                                 "sample": f"{used} scenarios per step ({its} Newton iterations), {what}, "
                                           f"{cores} processes"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # CPU baseline first: the worker pool is forked before this process creates a CUDA context
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or min(B, 2 * cpu_per_core * cores)   # ~10 s of CPU work
        val, its, sps_, used, kind, what = cpu_arm(nx, sample, cores, 1, 1)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{used} of the {B} scenarios ({its} Newton iterations, {sps_:.1f} s), {what}, "
                         f"{cores} processes"}
        if kind == "reference":          # the oracle port next to it (second number, same sample)
            v2, _, s2, _, _, w2 = cpu_arm(nx, sample, cores, 1, 1, kind="port")
            cpu["port"] = {"value": v2, "unit": UNIT, "seconds": s2, "what": w2}

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the backend has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eqs, y0, info, mdir, name = build_workload(nx)
    if world > 1:
        dist.barrier()
    sys.path.insert(0, str(mdir))
    mdl = importlib.import_module(name).mdl
    from solverz_b200.synthetic import pf_scenarios

    if args.lu_opt:
        mdl.lu_options = json.loads(args.lu_opt)
        mdl._sym = None
        mdl._pkl_path = None
    sess = mdl.session(B, local)
    sym = sess.sym
    P, Q = pf_scenarios(info, B, first=rank * B)
    dev = torch.device("cuda", local)
    Y0 = torch.as_tensor(np.tile(y0.array, (B, 1))).to(dev)
    sess.set_params({"Pinj": P, "Qinj": Q})
    sess.set_profile(True)
    from solverz_b200.parallel import gather_results_async
    pending = [None]

    def step():
        sess.Y.copy_(Y0)
        it, fl = sess.newton(TOL, MAX_IT)
        if world > 1:
            # the product's multi-GPU path (solverz_b200/parallel.py): the gather of solutions, iteration counts and
            # flags is ISSUED here from private send buffers and completes while the next step's kernels run; every
            # step's gather is waited for inside the timed region (the last one before the closing barrier)
            if pending[0] is not None:
                pending[0].wait()
            pending[0] = gather_results_async(sess.Y, it, fl, B * world)
        return it, fl

    def drain():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    def sync_all():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~0.2 s to deliver its first line: start before the warm-up
    for _ in range(max(args.warmup, 3)):
        it, fl = step()
    sync_all()
    launches0 = sess.query(5)
    prof = {k: dict(ms=0.0, units=0) for k in ("F", "J", "lu_refactor", "lu_solve", "bookkeeping")}
    pr0 = sess.get_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_region0 = time.perf_counter()
    e0.record()
    iters_total = 0
    its_steps = []
    for _ in range(args.steps):
        it, fl = step()
        its_steps.append(it)
        pr = sess.get_profile()
        for k in prof:
            prof[k]["ms"] += pr[k]["ms"]
            prof[k]["units"] += pr[k]["units"]
    drain()
    e1.record()
    sync_all()
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None
    pr1 = sess.get_profile()
    for k in prof:
        prof[k]["launches"] = pr1[k]["launches"] - pr0[k]["launches"]
    ms = e0.elapsed_time(e1)
    launches = sess.query(5) - launches0
    iters_total = int(sum(int(x.sum().item()) for x in its_steps))
    n_conv = int((fl == 0).sum().item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(iters_total), float(n_conv), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = tot[0].item() / (ms * 1e-3)

    # ---- end to end through the C-ABI host entry point (pinned host buffers).  The batch is split over
    # K independent sessions (handles), each driven by its own host thread on its own stream: every byte
    # still crosses PCIe inside the timed region, but one session's copies overlap the other's kernels.
    e2e = None
    if not args.no_e2e:
        import threading
        from solverz_b200.num_api.model import BatchSession
        K = max(1, args.e2e_sessions)
        while B % K:
            K -= 1
        Bk = B // K
        y0h = torch.as_tensor(np.tile(y0.array, (B, 1)))
        # one pinned in/out buffer per step (szb_newton_host returns the solution in place), filled before the
        # timed region: resetting a single buffer would put a 67 MB host-to-host copy inside every timed step
        n_buf = max(args.steps, 2)
        yhs = [torch.empty((B, mdl.n), dtype=torch.float64).pin_memory() for _ in range(n_buf)]
        yh = yhs[0]
        Ph = torch.as_tensor(P).pin_memory()
        Qh = torch.as_tensor(Q).pin_memory()
        ith = torch.zeros(B, dtype=torch.int32).pin_memory()
        flh = torch.zeros(B, dtype=torch.int32).pin_memory()
        h2d = yh.numel() * 8 + Ph.numel() * 8 + Qh.numel() * 8
        d2h = yh.numel() * 8 + 2 * B * 4
        if K == 1:
            parts = [sess]
        else:
            sess.close()                                  # free the full-batch factors first
            mdl._sessions.clear()
            parts = [BatchSession(mdl, Bk, local) for _ in range(K)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
        launches_e2e0 = sum(p_.query(5) for p_ in parts)

        def run_part(k, buf):
            sl = slice(k * Bk, (k + 1) * Bk)
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[k]):
                parts[k].newton_host(TOL, MAX_IT, buf[sl], {"Pinj": Ph[sl], "Qinj": Qh[sl]}, ith[sl], flh[sl])

        def e2e_step(buf):
            if K == 1:
                run_part(0, buf)
            else:
                th = [threading.Thread(target=run_part, args=(k, buf)) for k in range(K)]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()
            return int(ith.sum().item())

        for i in range(2):
            yhs[i].copy_(y0h)
            e2e_step(yhs[i])
        for b_ in yhs:
            b_.copy_(y0h)
        sync_all()
        e0.record()
        it_e2e = 0
        for i in range(args.steps):
            it_e2e += e2e_step(yhs[i])
        torch.cuda.synchronize()                          # all session streams have drained (newton_host syncs its own)
        e1.record()
        sync_all()
        ms_e2e = e0.elapsed_time(e1)
        te = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        ie = torch.tensor([float(it_e2e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ie, op=dist.ReduceOp.SUM)
        e2e = {"value": ie.item() / (te.item() * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": te.item() / args.steps,
               "api": f"szb_newton_host (C-ABI, pinned host buffers), {K} concurrent session(s) of {Bk} scenarios, "
                      f"one host thread + stream each",
               "converged": int((flh == 0).sum().item())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class (algorithmic bytes)
    peak, peak_kind = _peaks()
    n, nnzj, nnzlu = mdl.n, mdl.nnz_j, sym.nnz_lu
    npv = 2 * info["nb"]
    alg = {"F": 8 * (2 * n + npv), "J": 8 * (n + npv + nnzj), "lu_refactor": 8 * (nnzj + nnzlu),
           "lu_solve": 8 * (nnzlu + 4 * n)}  # solve+update: read LU, rhs, y; write y
    classes = {}
    for k, a in alg.items():
        if prof[k]["ms"] > 0:
            gbs = a * prof[k]["units"] / (prof[k]["ms"] * 1e-3) / 1e9
            classes[k] = {"ms_per_step": prof[k]["ms"] / args.steps, "alg_bytes_per_scenario": a,
                          "launches": prof[k]["launches"],
                          "ms_per_launch": prof[k]["ms"] / max(1, prof[k]["launches"]),
                          "achieved_gbs": gbs, "frac": gbs / peak}
    if "lu_refactor" in classes:     # the arithmetic side of the same launches (fp64, DMMA + FMA), for the large fronts
        classes["lu_refactor"]["fp64_tflops"] = (float(sym.flops) * prof["lu_refactor"]["units"]
                                                 / (prof["lu_refactor"]["ms"] * 1e-3) / 1e12)
    dom = max(classes, key=lambda k: classes[k]["ms_per_step"]) if classes else None
    roofline = None
    if dom:
        units_per_launch = prof[dom]["units"] / max(1, prof[dom]["launches"])
        roofline = {"bound": "hbm", "kernel": {"F": "szbm_F_kernel", "J": "szbm_J_kernel",
                                               "lu_refactor": "lu_factor_kernel", "lu_solve": "lu_backward_kernel"}[dom],
                    "achieved": classes[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": classes[dom]["frac"], "traffic": _ncu_traffic(dom, nx, B)[0],
                    "traffic_source": _ncu_traffic(dom, nx, B)[1], "peak_source": peak_kind,
                    "alg_bytes_per_launch": alg[dom] * units_per_launch, "classes": classes}
        if roofline["traffic"]:      # what the launch really moved through DRAM (ncu), as bandwidth and share of the peak
            roofline["traffic_gbs"] = roofline["traffic"] / (classes[dom]["ms_per_launch"] * 1e-3) / 1e9
            roofline["traffic_frac"] = roofline["traffic_gbs"] / peak

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "n": n, "nnz_J": nnzj, "nnz_LU": nnzlu,
                       "lu_doubles_stored": sym.lu_size, "lu_flops_per_scenario": sym.flops,
                       "ordering": sym.ordering, "iterations_per_step": int(tot[0].item() / args.steps),
                       "converged": int(tot[1].item()), "scenarios": B * world,
                       "l2": f"no flush needed: each step streams {(sym.lu_size + nnzj) * 8 * B / 1e9:.2f} GB "
                             f"of factors/Jacobians per GPU, far above the 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[2].item()), "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
