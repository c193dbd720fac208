#!/usr/bin/env python
"""Writes / refreshes an entry of profiles/ncu_traffic.json from an `ncu --set full` report of ONE launch of a kernel:
dram__bytes_read.sum + dram__bytes_write.sum per launch, tagged with the hash of the kernel sources so that bench.py
can tell a capture of the current kernels from a stale one.

    python tools/ncu_traffic.py gpurun_out/r2k_factor.ncu-rep lu_refactor 32 4096
    python tools/ncu_traffic.py <report> lu_refactor 100 2048 "ncu --metrics dram__bytes_read.sum,... one launch"

(optional fifth argument: how the report was taken, when it was not a --set full capture -- the large workloads are
captured with the three metrics alone, because every replay pass of --set full saves and restores 40 GB of factors)"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    rep, cls, nx, B = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    import bench
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        v, u = float(vals[col[name]]), units[col[name]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
        return v * scale
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    p = ROOT / "profiles" / "ncu_traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d[f"{cls}@nx{nx}B{B}"] = {"kernel": vals[col["Kernel Name"]], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
                              "dram_bytes_write": wr, "gpu_time_ms": get("gpu__time_duration.sum"),
                              "source": (sys.argv[5] if len(sys.argv) > 5 else "ncu --set full --clock-control none, one launch")
                              + f", report {Path(rep).name}",
                              "source_sha": bench.kernel_source_sha()}
    p.write_text(json.dumps(d, indent=1))
    print(d[f"{cls}@nx{nx}B{B}"])


if __name__ == "__main__":
    main()
