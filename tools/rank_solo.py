"""Diagnosis: one rank's partition of the box workload alone on one GPU, no neighbours (MSTGPU_DEBUG_NO_HALO=1: the
ghost rows keep the initial state, results are wrong at the cut, the timing of the tile launches is what it is).
    MSTGPU_DEBUG_NO_HALO=1 python tools/rank_solo.py --size 203 --parts 8 --ranks 0,5"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mst-cfd_b200")]
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=203)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--ranks", default="0,5")
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    os.environ["MSTGPU_DEBUG_NO_HALO"] = "1"
    os.environ["MSTGPU_NO_GRAPH"] = "1"
    import mstgpu
    from mstgpu import host
    f = host.flatten_raw(host.box_tets_raw(a.size, a.size, a.size))
    x = f["cc"]
    pert = 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q0 = np.zeros((f["ncells"], 5)); Q0[:, 0] = 1.0 + pert; Q0[:, 4] = (1.0 + pert) / 0.4
    for r in [int(v) for v in a.ranks.split(",")]:
        P = mstgpu.Partition(f, a.parts, r, order=2)
        ctx = mstgpu.Context(P, order=2, flux="roe", device=0)
        ctx.set_state(np.ascontiguousarray(Q0[P.cell_ids[:P.n_owned]]))
        ctx.step(1e-4, 3)
        ctx.sync()
        ctx.enable_kernel_timing(True)
        ms = ctx.step_timed(1e-4, a.steps)
        kt = {k: ctx.kernel_time(k) for k in ("halo_tiles", "interior_tiles", "step_tiles")}
        ctx.enable_kernel_timing(False)
        print(json.dumps(dict(rank=r, owned=int(P.n_owned), ghosts=int(P.n_local - P.n_owned), ms_per_step=ms / a.steps,
                              spans={k: v[0] / max(v[1], 1) for k, v in kt.items()})), flush=True)
        ctx.close()
        P.close()


if __name__ == "__main__":
    main()
