#!/bin/bash
# usage: [MM_CRYSTAL_VARIANT=v] crystal_variant_ncu.sh <points> tag1 tag2 ...
# per-launch time / warp-instructions / active lanes / issue rate of every crystal kernel of ONE mm_crystal call (step B), for library variants
# built by tools/build_variants.py ("cur" = the library as it is)
n=$1; shift
cp materialmodels.jl_b200/libmmb200.so /tmp/libmmb200_orig.so
for t in "$@"; do
  [ "$t" != cur ] && cp variants/libmmb200_$t.so materialmodels.jl_b200/libmmb200.so
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:crystal --csv --log-file /tmp/v_$t.csv python tools/profile_one.py crystal $n > /dev/null 2>&1
  echo "== $t (MM_CRYSTAL_VARIANT=${MM_CRYSTAL_VARIANT:-default})"
  python - /tmp/v_$t.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
ids = sorted({int(r[0]) for r in rows})
calls = {}
for r in rows:
    calls.setdefault(int(r[0]), {"name": r[4].split("(")[0].replace("void ", "")[:60]})[r[-3]] = r[-1]
last = [i for i in ids][-6:]
# keep the launches of the last mm_crystal call: walk back until the second-to-last finish kernel
out = []
for i in reversed(ids):
    out.append(i)
    if "presort" in calls[i]["name"] or (len(out) > 1 and "fixup" in calls[i]["name"]):
        if "fixup" in calls[i]["name"]: out.pop()
        break
tot = 0.0
for i in reversed(out):
    c = calls[i]; ms = float(c["gpu__time_duration.sum"]) / 1e6; tot += ms
    print("  %-62s %8.3f ms  %6.2f Ginst  lanes %5s  issue %5s%%" % (c["name"], ms, float(c["smsp__inst_executed.sum"]) / 1e9, c["smsp__thread_inst_executed_per_inst_executed.ratio"], c["smsp__issue_active.avg.pct_of_peak_sustained_active"]))
print("  total %.3f ms" % tot)
PY
  cp /tmp/libmmb200_orig.so materialmodels.jl_b200/libmmb200.so
done
