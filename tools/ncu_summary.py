#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for
profiles/.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/x.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none; cold-cache, serialised replays)\n")
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "") for r in body]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in idx:
            continue
        vals = []
        for r in body:
            v = r[idx[key]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[idx[key]]}".strip())
        print(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")
    # top stall reasons
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("\nTop warp-stall reasons (warps per issue-active cycle):\n")
    for r, n in zip(body, names):
        s = sorted(((float(r[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:4]
        print(f"- `{n}`: " + ", ".join(f"{k} {v:.2f}" for v, k in s))


if __name__ == "__main__":
    main()
