#!/usr/bin/env python
"""Trimmed SASS evidence for profiles/: per kernel the opcode histogram and every line carrying the
Blackwell-specific mnemonics (bulk copies / mbarriers for the step kernel; tcgen05 MMA, TMEM load /
store, commit for the policy kernel), each with two lines of context.  Runs on the CPU box:
    python tools/sass_excerpts.py            -> profiles/r2_sass_k_step_tma.txt, profiles/r2_sass_k_actor.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "q1physrl_b200", "libq1phys.so")
WANT = {
    "k_step_tma": ("k_step_tmaILb0ELb1ELb1ELi6E", ("UBLKCP", "SYNCS", "ACQBULK", "UTMA", "DADD", "DFMA", "DMUL", "MUFU.RCP64H")),
    "k_actor": ("k_actorILb1ELb1ELb1ELb0E", ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "MUFU.TANH", "ELECT", "FFMA2", "FMUL2", "USETMAXREG")),
}
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for short, (mangled, keys) in WANT.items():
    body = next((b for b in blocks if mangled in b.split("\n", 1)[0]), None)
    if body is None:
        raise SystemExit(f"{mangled} not found in {LIB}")
    lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    path = os.path.join(ROOT, "profiles", f"r2_sass_{short}.txt")
    with open(path, "w") as f:
        f.write(f"# cuobjdump -sass of {os.path.relpath(LIB, ROOT)}, function {body.split(chr(10), 1)[0].strip()}\n")
        f.write(f"# {len(lines)} instructions; opcode histogram (top 40):\n")
        for op, c in ops.most_common(40):
            f.write(f"#   {op:14s} {c}\n")
        f.write(f"# lines with {', '.join(keys)} (first 12 of each, two lines of context):\n")
        shown = collections.Counter()
        for i, l in enumerate(lines):
            k = next((k for k in keys if k in l), None)
            if k and shown[k] < 12:
                shown[k] += 1
                for j in range(max(0, i - 2), min(len(lines), i + 3)):
                    f.write(("  > " if j == i else "    ") + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", lines[j]).strip() + "\n")
                f.write("\n")
        f.write("# totals: " + ", ".join(f"{k} {sum(1 for l in lines if k in l)}" for k in keys) + "\n")
    print(path, {k: sum(1 for l in lines if k in l) for k in keys})
