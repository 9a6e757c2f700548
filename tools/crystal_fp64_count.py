#!/usr/bin/env python3
"""Executed FP64 thread-instructions per point of the crystal update (Newton + finish + fix-up kernels of one step-B call), from an ncu
metrics CSV:  ncu --csv --metrics <M> -k regex:crystal --launch-skip 3 --launch-count 3 python tools/profile_one.py crystal 2e6 > x.csv
usage: crystal_fp64_count.py x.csv points out.json"""
import csv
import io
import json
import sys

M = {"dfma": "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dmul": "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
     "dadd": "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "pipe": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "time": "gpu__time_duration.sum", "inst": "smsp__inst_executed.sum"}
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(txt)))
N = float(sys.argv[2])
tot = {"dfma": 0.0, "dmul": 0.0, "dadd": 0.0, "inst": 0.0}
kern = {}
for r in rows:
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    for k, m in M.items():
        if r["Metric Name"] == m:
            kern.setdefault(name, {})[k] = v * ({"msecond": 1.0, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}.get(r["Metric Unit"], 1.0) if k == "time" else 1.0)
            if k in tot:
                tot[k] += v
newton = [v for k, v in kern.items() if "newton" in k][0]
out = {"dfma_per_point": tot["dfma"] / N, "dmul_per_point": tot["dmul"] / N, "dadd_per_point": tot["dadd"] / N, "warp_inst_per_point": tot["inst"] / N,
       "pipe_fp64_active_pct": newton["pipe"], "kernels": kern, "points_in_capture": N, "source": sys.argv[1].split("/")[-1] + " (ncu --metrics, step B of configs[3])"}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out)[:600])
