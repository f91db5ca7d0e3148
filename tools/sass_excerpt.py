#!/usr/bin/env python
"""Trimmed SASS evidence of one kernel of libmstgpu.so for profiles/: resource usage, opcode histogram and
every TMA / mbarrier / warp-reduction instruction with its address.
usage: tools/sass_excerpt.py <mangled-name-substring> > profiles/rN_sass_<kernel>.txt"""
import collections
import re
import subprocess
import sys

so = "mst-cfd_b200/libmstgpu.so"
key = sys.argv[1]
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout.splitlines()
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
name = None
for i, l in enumerate(res):
    if "Function" in l and key in l:
        name = l.split("Function")[1].strip().rstrip(":")
        print(subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
        print(name)
        print(res[i + 1].strip())
        break
if name is None:
    sys.exit("kernel not found")
start = next(i for i, l in enumerate(sass) if "Function : " + name in l)
body = []
for l in sass[start + 1:]:
    if "Function : " in l:
        break
    body.append(l)
ops = collections.Counter()
keep = []
for l in body:
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if not m:
        continue
    op = m.group(3)
    ops[op.split(".")[0]] += 1
    if re.match(r"(UBLKCP|UBLKPF|SYNCS|REDUX|CREDUX|UTMA|LDGSTS|BAR|ATOMG|REDG|MUFU)", op):
        keep.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
print(f"\n{sum(ops.values())} SASS instructions; histogram:")
print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
print("\nTMA bulk copies (UBLKCP), L2 prefetch (UBLKPF), mbarrier (SYNCS), warp reductions (REDUX/CREDUX), barriers, atomics, MUFU:")
for l in keep:
    print("  " + l)
