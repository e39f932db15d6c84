"""profiles/r02_ncu_traffic.json from `ncu --page raw --csv` dumps: per kernel the DRAM bytes per launch and per particle,
DRAM / L1 / issue percentages. bench.py reads it (roofline.traffic, roofline.dram_frac_ncu) as long as the kernel sources are
unchanged.   python tools/ncu_traffic_json.py N_PARTICLES label=raw.csv [label=raw.csv ...] > profiles/r02_ncu_traffic.json"""
import csv
import os
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402

out = {"kernel_source_sha": bench.kernel_source_hash(), "captures": {}}
for arg in sys.argv[1:]:
    label, _, path = arg.partition("=")
    n_str, _, label = label.partition(":")
    n = int(n_str)
    rows = list(csv.reader(open(path)))
    rows = [r for r in rows if len(r) > 10]
    hdr, data = rows[0], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]].replace(",", ""))
        except Exception:
            return None
    per = {}
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].replace("akua::", "").replace("rsort::", "").replace("void ", "")
        per.setdefault(name, []).append(r)
    cap = {"particles": n, "kernels": {}}
    for name, rs in per.items():
        r = rs[len(rs) // 2]
        rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
        units = rows[1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd_b = rd * scale.get(units[ix["dram__bytes_read.sum"]], 1.0) if rd is not None else None
        wr_b = wr * scale.get(units[ix["dram__bytes_write.sum"]], 1.0) if wr is not None else None
        dur = f(r, "gpu__time_duration.sum")
        dur_unit = units[ix["gpu__time_duration.sum"]]
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(dur_unit, 1.0) if dur is not None else None
        cap["kernels"][name] = {
            "launches_captured": len(rs), "duration_us_under_ncu": dur_us,
            "dram_bytes": (rd_b + wr_b) if rd_b is not None and wr_b is not None else None,
            "dram_bytes_per_particle": ((rd_b + wr_b) / n) if rd_b is not None and wr_b is not None else None,
            "dram_pct_of_peak": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_throughput_pct": f(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
            "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": f(r, "launch__registers_per_thread"),
        }
    out["captures"][label] = cap
os.write(bench._REAL_STDOUT, (json.dumps(out, indent=1) + "\n").encode())
