#!/usr/bin/env python
"""Registers / stack / shared memory per kernel of a built library (cuobjdump -res-usage), one line each.
    python tools/res_usage.py [q1physrl_b200/libq1phys.so] [substring filter]"""
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "q1physrl_b200/libq1phys.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        continue
    if name and "REG:" in line:
        r = dict(kv.split(":") for kv in line.split() if ":" in kv)
        if flt in name:
            print(f"{name:48s} REG {r.get('REG'):>4s}  STACK {r.get('STACK'):>4s}  SHARED {r.get('SHARED'):>6s}  LOCAL {r.get('LOCAL')}")
        name = None
