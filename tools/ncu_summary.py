#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics -> text, and dram traffic per launch -> json.
usage: ncu_summary.py <report.ncu-rep> <out.txt> [--traffic-json out.json --points N --bytes-per-point B]"""
import argparse, csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("out")
ap.add_argument("--traffic-json"); ap.add_argument("--points", type=float); ap.add_argument("--bytes-per-point", type=float)
ap.add_argument("--note", default="")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = []
traffic = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    lines.append("kernel: %s   grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
    for k in KEYS:
        if k in d and d[k] != "":
            lines.append("  %-70s %s %s" % (k, d[k], u[k]))
    st = sorted(((float(d[k]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in d if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
    lines.append("  stall reasons (warps stalled per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
    def gb(x, unit):
        return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
    rd, wr = gb(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), gb(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    traffic = rd + wr
    lines.append("  dram traffic per launch: %.4g B (read %.4g + write %.4g)" % (traffic, rd, wr))
    if a.points and a.bytes_per_point:
        alg = a.points * a.bytes_per_point
        lines.append("  algorithmic bytes per launch: %.4g B (%g points x %g B) -> traffic / algorithmic = %.3f" % (alg, a.points, a.bytes_per_point, traffic / alg))
if a.note:
    lines.insert(0, a.note)
open(a.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if a.traffic_json and traffic is not None:
    json.dump({"dram_bytes_per_launch": traffic, "points_in_capture": a.points, "dram_bytes_per_point": traffic / a.points if a.points else None,
               "source": a.out}, open(a.traffic_json, "w"))
