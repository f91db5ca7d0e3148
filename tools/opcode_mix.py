#!/usr/bin/env python
"""Opcode mix per kernel phase from an ncu source-page dump of k_step_tiles.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; tools/opcode_mix.py src.csv"""
import collections
import csv
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_step = open(os.path.join(root, "mst-cfd_b200", "csrc", "step_tiles.cuh")).read().splitlines()
marks = {}
for i, l in enumerate(src_step):
    for k in ("phase 0", "phase 1", "phase 2", "phase 3"):
        if "---- " + k in l:
            marks[k] = i + 1
cur = hdr = line = None
ops = collections.defaultdict(collections.Counter)
stl = collections.defaultdict(collections.Counter)


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "":
        line = num(r[0]) or line
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3].strip())
    if not m:
        continue
    op = m.group(2).split(".")[0]
    n, s = num(r[hdr["Instructions Executed"]]), num(r[hdr["# Samples"]])
    if cur == "physics.cuh":
        ph = "2b math"
    elif cur == "step_tiles.cuh":
        ph = "0 stage" if line < marks["phase 1"] else "1" if line < marks["phase 2"] else "2a glue" if line < marks["phase 3"] else "3 update"
    else:
        ph = "x " + str(cur)
    ops[ph][op] += n
    stl[ph][op] += s
tot = sum(sum(c.values()) for c in ops.values())
ts = sum(sum(c.values()) for c in stl.values())
for ph in sorted(ops):
    t = sum(ops[ph].values())
    print(f"phase {ph}: {100 * t / tot:.1f}% of instructions, {100 * sum(stl[ph].values()) / max(ts, 1):.1f}% of stall samples")
    for op, n in ops[ph].most_common(12):
        print(f"    {op:10s} {100 * n / tot:5.2f}% inst  {100 * stl[ph][op] / max(ts, 1):5.2f}% stall")
