#!/usr/bin/env python
"""profiles/ncu_traffic.json (read by bench.py for roofline.traffic / roofline.ncu) from an `ncu --set full` raw page.

  ncu -i rep.ncu-rep --page raw --csv > raw.csv
  python profiles/tools/ncu_traffic.py raw.csv 20000000 "<source description>" > profiles/ncu_traffic.json

Keeps, per kernel name, the launch with the longest duration (the whole-stream launch of the `value` leg)."""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                 "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                 "sm__warps_active.avg.pct_of_peak_sustained_active",
                                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
out = {}
for r in rows[2:]:
    name = re.sub(r"[<(].*", "", r[col["Kernel Name"]].replace("<unnamed>::", "").replace("void ", ""))
    f = lambda k: float(r[col[k]].replace(",", ""))
    e = {"dram_bytes_read": f("dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]],
         "dram_bytes_write": f("dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]],
         "duration_ms_under_ncu": f("gpu__time_duration.sum") * tscale[units[col["gpu__time_duration.sum"]]],
         "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")}
    for key, metric in (("inst_executed", "smsp__inst_executed.sum"), ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                        ("dmma_pipe_active_pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")):
        if metric in hdr and r[hdr.index(metric)] not in ("", "n/a"):
            try:
                e[key] = float(r[hdr.index(metric)].replace(",", ""))
            except ValueError:
                pass
    if e.get("smem_wavefronts"):
        e["smem_bank_conflict_pct"] = 100.0 * e.get("smem_bank_conflicts", 0.0) / e["smem_wavefronts"]
    if name not in out or e["duration_ms_under_ncu"] > out[name]["duration_ms_under_ncu"]:
        out[name] = e
print(json.dumps({"events": int(sys.argv[2]), "config": sys.argv[4] if len(sys.argv) > 4 else "C2", "source": sys.argv[3], "kernels": out},
                 indent=1))
