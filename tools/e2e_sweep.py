"""Streamed step (mstgpu_step_host) on the box workload: ms per step for several chunk counts beside the
three-call sequence, with the PCIe copies alone for scale.  python tools/e2e_sweep.py --size 203 --chunks 1,8,16,24,48"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mst-cfd_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=203)
    ap.add_argument("--chunks", default="1,4,8,16,24,32,48,96")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    import mstgpu
    from mstgpu import host


    f = host.flatten_raw(host.box_tets_raw(a.size, a.size, a.size))
    nc, U = f["ncells"], f["dim"] + 2
    cc = f["cc"]
    s = np.sin(2 * np.pi * cc[:, 0]) * np.sin(2 * np.pi * cc[:, 1]) * np.sin(2 * np.pi * cc[:, 2])
    Q0 = np.zeros((nc, U)); Q0[:, 0] = 1 + 0.1 * s; Q0[:, -1] = (1 + 0.1 * s) / 0.4
    ctx = mstgpu.Context(f, order=2, flux="roe", device=0)
    hin = torch.empty((nc, U), dtype=torch.float64, pin_memory=True)
    hout = torch.empty((nc, U), dtype=torch.float64, pin_memory=True)
    dev = torch.empty((nc, U), dtype=torch.float64, device="cuda")
    res = {"cells": nc, "rows": []}

    def timed(fn, n=a.steps):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    hin.numpy()[:] = Q0
    res["h2d_ms"] = timed(lambda: dev.copy_(hin, non_blocking=True))
    res["d2h_ms"] = timed(lambda: hout.copy_(dev, non_blocking=True))
    s2 = torch.cuda.Stream()

    def both():
        dev.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dev, non_blocking=True)
    res["h2d_and_d2h_together_ms"] = timed(both)

    def three():
        ctx.set_state_ptr(hin.data_ptr()); ctx.step(1e-4, 1); ctx.get_state_ptr(hout.data_ptr()); ctx.residual()
    res["three_calls_ms"] = timed(three)
    ref = hout.numpy().copy() if nc < 3e7 else int(hout.numpy().view(np.uint64).sum(dtype=np.uint64))
    for g in [int(x) for x in a.chunks.split(",")]:
        hout.zero_()
        ms = timed(lambda: (ctx.step_host(hin.data_ptr(), hout.data_ptr(), 1e-4, g), ctx.residual()))
        same = np.array_equal(hout.numpy(), ref) if nc < 3e7 else int(hout.numpy().view(np.uint64).sum(dtype=np.uint64)) == ref
        row = dict(chunks=g, ms=ms, gcells_per_s=nc / ms / 1e6, identical=bool(same))
        print(json.dumps(row), flush=True)
        res["rows"].append(row)
    print(json.dumps({k: v for k, v in res.items() if k != "rows"}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
