#!/usr/bin/env python
"""Aggregate an ncu source-page dump of k_step_tiles by kernel phase.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; tools/phase_breakdown.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
import os
_src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mst-cfd_b200", "csrc", "step_tiles.cuh")
_marks0 = {}
for _i, _l in enumerate(open(_src).read().splitlines()):
    for _k in ("phase 0", "phase 1", "phase 2", "phase 3"):
        if "---- " + _k in _l:
            _marks0[_k] = _i + 1
cur = None
hdr = None
phase, smp, wf, ex = (collections.Counter() for _ in range(4))
marks = dict(_marks0)  # comment lines carry no SASS: take the markers from the source file
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    src = r[1]
    if cur == "step_tiles.cuh":
        for k in ("phase 0", "phase 1", "phase 2", "phase 3"):
            if "---- " + k in src:
                marks[k] = ln
    try:
        inst = int(r[hdr["Instructions Executed"]] or 0)
        s = int(r[hdr["# Samples"]] or 0)
        w = int(r[hdr["L1 Wavefronts Shared"]] or 0)
        e = int(r[hdr["L1 Wavefronts Shared Excessive"]] or 0)
    except (ValueError, KeyError):
        continue
    if cur == "physics.cuh":
        ph = "2b flux math (physics.cuh)"
    elif cur == "step_tiles.cuh":
        p1, p2, p3 = marks.get("phase 1", 10**9), marks.get("phase 2", 10**9), marks.get("phase 3", 10**9)
        ph = "0  stage" if ln < p1 else "1  gradient + reconstruction" if ln < p2 else "2a flux glue" if ln < p3 else "3  update + residual"
    else:
        ph = "x  " + str(cur)
    phase[ph] += inst; smp[ph] += s; wf[ph] += w; ex[ph] += e
tot, ts = sum(phase.values()), sum(smp.values())
print(f"total warp instructions {tot:.3e}, stall samples {ts}")
for k in sorted(phase):
    print(f"{k:32s} inst {100 * phase[k] / tot:5.1f}%  stall samples {100 * smp[k] / max(ts, 1):5.1f}%  "
          f"smem wavefronts {wf[k] / 1e6:7.1f}M (excess {ex[k] / 1e6:6.1f}M)")
