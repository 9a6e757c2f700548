#!/usr/bin/env python3
"""Static scan of the built objects (no GPU needed): per kernel, the number of CALL (IEEE division / sqrt slow paths and other
out-of-line subroutines), MUFU and local-memory (LDL/STL) instructions, plus the ptxas stack frame vs spill sizes when a
`-Xptxas -v` log is given.  A stack frame much larger than the spill stores means a dynamically indexed local array — the
signature of the UniaxialStress regression fixed in round 1.
usage: sass_scan.py [--json]            (objects: materialmodels.jl_b200/build/*.o)"""
import glob, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def scan(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); res[name] = {"instr": 0, "call": 0, "mufu": 0, "local": 0, "local_dynamic": 0}
            continue
        if name is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        r = res[name]; r["instr"] += 1
        if " CALL" in line: r["call"] += 1
        if "MUFU" in line: r["mufu"] += 1
        if re.search(r"\b(LDL|STL)\b", line.replace(".", " ")):
            r["local"] += 1
            if not re.search(r"\[R1(\+0x[0-9a-f]+)?\]", line): r["local_dynamic"] += 1      # not addressed off the stack pointer: a run-time index
    return res


def main():
    allres = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "materialmodels.jl_b200", "build", "*.o"))):
        for k, v in scan(obj).items():
            v["object"] = os.path.basename(obj); allres[k] = v
    names = demangle(list(allres))
    if "--json" in sys.argv:
        print(json.dumps({names[k]: v for k, v in allres.items()}, indent=1)); return
    print("%6s %5s %5s %6s %8s  %s" % ("instr", "call", "mufu", "local", "dynamic", "kernel"))
    for k, v in sorted(allres.items(), key=lambda kv: -kv[1]["local_dynamic"] * 100000 - kv[1]["local"]):
        print("%6d %5d %5d %6d %8d  %s" % (v["instr"], v["call"], v["mufu"], v["local"], v["local_dynamic"], names[k][:150]))


if __name__ == "__main__":
    main()
