"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum [+ dram bytes]) per kernel: python tools/launch_agg.py file.csv"""
import collections
import csv
import sys

for fn in sys.argv[1:]:
    rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in rows[1:]:
        a = agg[r[ik][:60]]
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            a[0] += 1; a[1] += v
        elif "read" in r[im]:
            a[2] += v
        elif "write" in r[im]:
            a[3] += v
    print(fn)
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:62s} n={a[0]:4d}  {a[1] / 1e6:9.3f} ms  ({a[1] / 1e6 / max(a[0], 1):7.4f} ms each)  rd {a[2] / 1e9:8.3f} GB  wr {a[3] / 1e9:8.3f} GB  {(a[2] + a[3]) / max(a[1], 1):7.1f} GB/s")
    print(f"  total {sum(a[1] for a in agg.values()) / 1e6:.3f} ms")
