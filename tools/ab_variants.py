#!/usr/bin/env python
"""A/B of the experimental launch variants of the default fused kernel (csrc/step_tiles.cuh, MSTGPU_TILE_VAR)
on one context: bit-identity of the state against variant 0 after 3 steps, then device time of K steps.

    python tools/ab_variants.py --size 128 --steps 20 [--variants 0,1,2,3,4,5,7,0]
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
    ap.add_argument("--tiles", default="0", help="comma list of tile sizes (cells per tile), 0 = default; one context each")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import mstgpu
    from mstgpu import host
    t = time.time()
    n = a.size
    f = host.flatten_raw(host.box_tets_raw(n, n, n))
    x = f["cc"]
    pert = 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q0 = np.zeros((f["ncells"], 5)); Q0[:, 0] = 1.0 + pert; Q0[:, 4] = (1.0 + pert) / 0.4
    ref, rows = None, []
    for T in [int(s) for s in a.tiles.split(",")]:
      t = time.time()
      ctx = mstgpu.Context(f, order=2, flux="roe", tile_cells=T)
      print(f"[ab] {f['ncells']} cells, T={T}: context in {time.time() - t:.1f}s", file=sys.stderr, flush=True)
      for v in [int(s) for s in a.variants.split(",")]:
            ctx.set_tile_variant(v)
            ctx.set_state(Q0)
            ctx.step(1e-4, 3)
            Q3 = ctx.get_state()
            r3 = ctx.residual()
            if ref is None:
                ref = (Q3.copy(), r3.copy())
            same = bool(np.array_equal(Q3, ref[0]) and np.array_equal(r3, ref[1]))
            ctx.step(1e-4, 5)
            ctx.sync()
            ctx.enable_kernel_timing(True)
            ms = ctx.step_timed(1e-4, a.steps)
            kms, kn = ctx.kernel_time("step_tiles")
            ctx.enable_kernel_timing(False)
            ctx.step(1e-4, 4); ctx.sync()
            gms = ctx.step_timed(1e-4, a.steps)   # no per-kernel events: pairs of steps from the CUDA graph
            row = dict(tile_cells=T, variant=v, identical_to_v0=same, ms_per_step=ms / a.steps, kernel_ms=kms / max(kn, 1),
                       graph_ms_per_step=gms / a.steps, gcells_per_s=f["ncells"] * a.steps / (ms * 1e-3) / 1e9)
            rows.append(row)
            print(json.dumps(row), flush=True)
      ctx.close()
    if a.out:
        json.dump(dict(cells=int(f["ncells"]), steps=a.steps, rows=rows), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
