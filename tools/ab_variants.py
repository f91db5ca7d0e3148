#!/usr/bin/env python
"""A/B of the launch variants of the fused kernel (csrc/step_tiles.cuh, mstgpu_set_tile_variant) and of tile
sizes: bit-identity of state and residual after 3 steps against the first run, then device time of K steps
(CUDA events per launch, and the same K steps issued as pairs from the CUDA graph).

    python tools/ab_variants.py --size 128 --steps 20 --variants 0,1,2,3,4,5,7,0           # 3-D, second order
    python tools/ab_variants.py --size 128 --variants 0 --tiles 512,440,544,560             # tile sizes
    python tools/ab_variants.py --workload step --size 445 --steps 200 --variants 32,16,8   # config 2: CTAs per SM
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mst-cfd_b200")]
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--variants", default="0,1,2,3,4,5,7,0")
    ap.add_argument("--workload", default="box", choices=["box", "step"], help="step: BASELINE config 2 (AUSM+, first order, Mach 3)")
    ap.add_argument("--order", type=int, default=0, help="0 = the workload's own (box 2, step 1)")
    ap.add_argument("--flux", default="", help="default: the workload's own (box roe, step ausm)")
    ap.add_argument("--block-threads", type=int, default=0, help="CTA size of the fused kernel (128|256), 0 = default")
    ap.add_argument("--tiles", default="0", help="comma list of tile sizes (cells per tile), 0 = default; one context each")
    ap.add_argument("--configs", default="", help="comma list of NT:T:variant[:tile_flags] tuples (one context each); overrides --tiles/--variants/--block-threads")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import mstgpu
    from mstgpu import host
    t = time.time()
    n = a.size
    kw, dt = dict(order=2, flux="roe"), 1e-4
    if a.workload == "step":
        f = host.flatten_raw(host.forward_step_raw(n))
        u = 3.0 * np.sqrt(1.4)
        Q0 = np.tile(np.array([1.0, u, 0.0, 1.0 / 0.4 + 0.5 * u * u]), (f["ncells"], 1))
        kw, dt = dict(order=1, flux="ausm", inletQ=list(Q0[0]) + [0.0]), 2e-5
    else:
        f = host.flatten_raw(host.box_tets_raw(n, n, n))
        x = f["cc"]
        pert = 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
        Q0 = np.zeros((f["ncells"], 5)); Q0[:, 0] = 1.0 + pert; Q0[:, 4] = (1.0 + pert) / 0.4
    if a.order:
        kw["order"] = a.order
    if a.flux:
        kw["flux"] = a.flux
    import hashlib
    ref, rows = None, []
    if a.configs:
        cfgs = [tuple(int(x) for x in (c.split(":") + ["0", "0"])[:5]) for c in a.configs.split(",")]
    else:
        cfgs = [(a.block_threads, int(T), int(v), 0, 0) for T in a.tiles.split(",") for v in a.variants.split(",")]
    ctx, key = None, None
    for NT, T, v, fl, fit in cfgs:
        if key != (NT, T, fl, fit):
            if ctx is not None:
                ctx.close()
            t = time.time()
            # 5th field: MSTGPU_TILE_FIT -- variable tile sizes, flux faces per tile <= fit (tile_cells is the cap)
            os.environ.pop("MSTGPU_TILE_FIT", None)
            if fit:
                os.environ["MSTGPU_TILE_FIT"] = str(fit)
            ctx = mstgpu.Context(f, tile_cells=T, block_threads=NT, tile_flags=fl, **kw)
            key = (NT, T, fl, fit)
            print(f"[ab] {f['ncells']} cells, NT={NT} T={T} fit={fit}: context in {time.time() - t:.1f}s", file=sys.stderr, flush=True)
        ctx.set_tile_variant(v)
        ctx.set_state(Q0)
        ctx.step(dt, 3)
        Q3 = ctx.get_state()
        r3 = ctx.residual()
        if ref is None:
            ref = (Q3.copy(), r3.copy())
        same = bool(np.array_equal(Q3, ref[0], equal_nan=True) and np.array_equal(r3, ref[1], equal_nan=True))
        ctx.step(dt, 5)
        ctx.sync()
        ctx.enable_kernel_timing(True)
        ms = ctx.step_timed(dt, a.steps)
        kms, kn = ctx.kernel_time("step_tiles")
        ctx.enable_kernel_timing(False)
        ctx.step(dt, 4)
        ctx.sync()
        gms = ctx.step_timed(dt, a.steps)   # no per-kernel events: pairs of steps from the CUDA graph
        row = dict(block_threads=NT, tile_cells=T, variant=v, tile_flags=fl, fit_faces=fit, identical_to_first=same, sha=hashlib.sha256(Q3.tobytes()).hexdigest()[:12],
                   ms_per_step=ms / a.steps, kernel_ms=kms / max(kn, 1),
                   graph_ms_per_step=gms / a.steps, gcells_per_s=f["ncells"] * a.steps / (ms * 1e-3) / 1e9)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if ctx is not None:
        ctx.close()
    if a.out:
        json.dump(dict(cells=int(f["ncells"]), steps=a.steps, rows=rows), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
