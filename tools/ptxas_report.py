#!/usr/bin/env python3
"""Rebuild libmmb200.so with -Xptxas -v and print registers / spills / smem per kernel."""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "..", "materialmodels.jl_b200", "build.py"), "-v"],
                     capture_output=True, text=True).stdout
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void mm::", "")
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        stack = m.groups()
        continue
    m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
    if m and cur:
        rows.append((cur, m.group(1), stack[0], stack[1], stack[2]))
        cur = None
print("%-70s %5s %6s %7s %7s" % ("kernel", "regs", "stack", "spillst", "spillld"))
for r in rows:
    print("%-70s %5s %6s %7s %7s" % r)
