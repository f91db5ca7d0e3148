#!/usr/bin/env python
"""Mesh-input and output path beside the reference's own, one JSON line.

  input   native reader (host/mshread.cpp) + flattener  vs  MshBlock::readMsh of the reference
  output  device node averaging (mstgpu_node_fields) + native writer (host/pltwrite.cpp)
          vs  Work::writedataRhoBasedMshNodePlt of the reference (host loops + iostream)

on the same synthetic 2-D mesh (forward-facing step, h = 1/N triangles; the reference is 2-D only),
written once with the native .msh writer.  The reference side is oracle/_ref/ref_io (the reference's
sources compiled by oracle/refbuild); without it only this repo's side is timed.  Without a GPU the
device part is skipped and the writer is timed on stand-in numbers (no CPU path computes node fields).

    python tools/bench_io.py --size 445 [--no-ref] [--out gpurun_out/io_bench.json]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mst-cfd_b200")]

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=445)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from mstgpu import host
    res = dict(workload=f"forward-facing step h=1/{a.size}", cpus=os.cpu_count(), threads=int(os.environ.get("OMP_NUM_THREADS", os.cpu_count())))
    tmp = tempfile.mkdtemp(prefix="mstio_")
    msh = os.path.join(tmp, "step.msh")
    raw0 = host.raw_zones_from_ftype(host.forward_step_raw(a.size))
    t = time.perf_counter(); host.write_msh(msh, raw0); res["msh_write_s"] = time.perf_counter() - t
    res["msh_bytes"] = os.path.getsize(msh)
    best = lambda fn: min(fn() for _ in range(a.reps))

    def t_read():
        t = time.perf_counter(); host.read_msh(msh); return time.perf_counter() - t
    res["read_s"] = best(t_read)
    raw = host.read_msh(msh)

    def t_flat():
        t = time.perf_counter(); host.flatten_raw(raw); return time.perf_counter() - t
    res["flatten_s"] = best(t_flat)
    f = host.flatten_raw(raw)
    nc, nn = f["ncells"], raw["nodes"].shape[0]
    res.update(cells=nc, faces=f["nfaces"], nodes=nn)
    t = time.perf_counter(); ptr, idx = host.node_faces(raw); cp, ci = host.cell_nodes(raw, f); res["connectivity_s"] = time.perf_counter() - t
    # state: Mach-3 free stream + a smooth perturbation, so that every printed digit is exercised
    x = f["cc"]
    u = 3.0 * np.sqrt(1.4)
    Q = np.empty((nc, 4))
    Q[:, 0] = 1.0 + 0.1 * np.sin(3 * x[:, 0]) * np.cos(5 * x[:, 1])
    Q[:, 1] = Q[:, 0] * u
    Q[:, 2] = 0.05 * np.sin(7 * x[:, 0] + x[:, 1])
    Q[:, 3] = 1.0 / 0.4 + 0.5 * (Q[:, 1] ** 2 + Q[:, 2] ** 2) / Q[:, 0]
    w = host.node_weights(f, nn)
    import torch
    gpu = torch.cuda.is_available()
    if gpu:
        import mstgpu
        ctx = mstgpu.Context(f, order=1, flux="ausm", inletQ=list(Q[0]) + [0.0])
        ctx.output_setup(f, ptr, idx, w)
        ctx.set_state(Q)
        fld = ctx.node_fields()
        ctx.enable_kernel_timing(True)

        def t_nf():
            t = time.perf_counter(); ctx.node_fields(fld); return time.perf_counter() - t

        def t_gs():
            out = np.empty((nc, 4))
            t = time.perf_counter(); ctx.get_state(out); return time.perf_counter() - t
        res["node_fields_s"] = best(t_nf)       # kernel + D2H of [nodes][6] (pageable host memory)
        res["get_state_s"] = best(t_gs)         # what the reference's writer needs first: D2H of [cells][4]
        ms, n = ctx.kernel_time("node_fields")
        res["node_fields_kernel_ms"] = ms / max(n, 1)
        # algorithmic bytes of the kernel: per node-face entry id 4 + (c0, c1, eta) 16 + two state rows 64; per node ptr 4 + w 8 + out 48
        if res["node_fields_kernel_ms"] > 0:
            res["node_fields_kernel_gbs"] = (idx.size * (4 + 16 + 64) + nn * 60) / (res["node_fields_kernel_ms"] * 1e-3) / 1e9
        ctx.close()
    else:
        # no device: the writer is still timed, on stand-in numbers of the same magnitude (the node
        # fields themselves are the device's job; nothing here computes them on the CPU)
        xn = raw["nodes"]
        fld = np.stack([1.0 + 0.1 * np.sin(3 * xn[:, 0]), 3.5 + 0 * xn[:, 0], 0.05 * np.cos(xn[:, 1]),
                        0.0035 + 0 * xn[:, 0], 1.0 + 0.1 * np.cos(xn[:, 0]), 3.0 + 0 * xn[:, 0]], axis=1)
    out = os.path.join(tmp, "ours.plt")

    def t_w():
        t = time.perf_counter(); host.plt_write(out, raw, fld, cp, ci, zone_t=10); return time.perf_counter() - t
    res["plt_write_s"] = best(t_w)
    res["plt_bytes"] = os.path.getsize(out)

    def t_wb():
        t = time.perf_counter(); host.plt_write(out + ".bin", raw, fld, cp, ci, zone_t=10, binary=True); return time.perf_counter() - t
    res["plt_write_binary_s"] = best(t_wb)
    ref_io = os.path.join(ROOT, "oracle", "_ref", "ref_io")
    if not a.no_ref and os.path.exists(ref_io):
        os.makedirs(os.path.join(tmp, "result"), exist_ok=True)
        Q.tofile(os.path.join(tmp, "q.bin"))
        r = subprocess.run([ref_io, msh, tmp, os.path.join(tmp, "q.bin"), "10"], check=True, capture_output=True, text=True)
        tok = r.stdout.split()
        res["ref_read_s"] = float(tok[tok.index("read_ms") + 1]) * 1e-3
        res["ref_write_s"] = float(tok[tok.index("write_ms") + 1]) * 1e-3
        ref = open(os.path.join(tmp, "result", "step.msh_TIME4000_u0_t10.plt"), "rb").read()
        res["plt_identical_to_reference"] = (hashlib.sha256(ref).hexdigest() == hashlib.sha256(open(out, "rb").read()).hexdigest()) if gpu else None
        res["input_speedup"] = res["ref_read_s"] / (res["read_s"] + res["flatten_s"])
        ours_out = res.get("node_fields_s", 0.0) + res["plt_write_s"]
        res["output_speedup"] = res["ref_write_s"] / ours_out if gpu else None
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
