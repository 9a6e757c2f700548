#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:90]
    t = float(r[-1]) / 1e6
    c = tot.setdefault(name, [0, 0.0, []])
    c[0] += 1; c[1] += t; c[2].append(t)
total = sum(v[1] for v in tot.values())
out = ["launch list: %s  (per-launch times under ncu are cold-cache and serialised: compare SHARES)" % sys.argv[1], "%-92s %6s %12s %8s %12s" % ("kernel", "count", "total ms", "share", "ms/launch")]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append("%-92s %6d %12.3f %7.1f%% %12.4f" % (k, v[0], v[1], 100 * v[1] / total, v[1] / v[0]))
if len(sys.argv) > 2:
    out.append(sys.argv[2])
print("\n".join(out))
