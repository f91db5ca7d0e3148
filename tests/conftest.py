import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mst-cfd_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

REF_MESHES = [
    "2d-shockwavepipe-2", "2d-stair-st-2", "2d-stair-un-3-loose-tri", "2d-stair-un-3-tri",
    "2d-stair-un-4-tri", "2d-stair-un-5-tri", "2d-stairW-1", "2d-stairW-2-st",
]
# 2d-stair-st-2 contains 20 degenerate sliver quads (width 1e-8) whose Heron
# volume is NaN in the reference's own formula (R/mesh/Cell.cpp:28-49); it is
# used for metric parity only, not for stepping.
STEP_MESHES = [m for m in REF_MESHES if m != "2d-stair-st-2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


_mesh_cache = {}


def load_raw(name: str) -> dict:
    from oracle import mshio
    return mshio.npz_to_raw(np.load(os.path.join(GOLDEN, f"mesh_{name}.npz")))


def load_flat(name: str, conv: str = "consistent") -> dict:
    """Flat reference-order tables of a reference mesh (oracle's numpy metrics)."""
    key = (name, conv)
    if key not in _mesh_cache:
        from oracle import mesh_np
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _mesh_cache[key] = mesh_np.flatten(load_raw(name), conv)
    return _mesh_cache[key]


def box_flat(nx, ny, nz, bc=(3, 3, 3, 3, 3, 3), l=(1.0, 1.0, 1.0)) -> dict:
    from mstgpu import host
    return host.flatten_raw(host.box_tets_raw(nx, ny, nz, *l, bc=bc))


def char_scales(Q, gamma=1.4):
    """Per-variable scales for a relative L-inf on the conserved variables.
    rho and E: their own max.  Momentum components: the characteristic momentum
    max(|m| + rho*a) -- a component that is physically zero (v in the Sod tube)
    would otherwise turn round-off into an O(1) 'relative' error."""
    D = Q.shape[1] - 2
    rho = Q[:, 0]
    m2 = (Q[:, 1:1 + D] ** 2).sum(1)
    # |.|: an unlimited 2nd-order step from a random state may leave p or rho
    # negative in a few cells; the scale only needs the order of magnitude
    p = np.abs((Q[:, -1] - 0.5 * m2 / rho) * (gamma - 1))
    a = np.sqrt(gamma * p / np.abs(rho))
    s = np.empty(D + 2)
    s[0] = np.abs(rho).max()
    s[1:1 + D] = (np.sqrt(m2) + np.abs(rho) * a).max()
    s[-1] = np.abs(Q[:, -1]).max()
    return s


def rel_linf(Qa, Qb, gamma=1.4):
    """max over cells and variables of |Qa - Qb| / scale, where scale is
      * for every variable the per-variable characteristic scale of `char_scales` (momentum components
        against max(|m| + rho a), so a physically zero component does not turn round-off into O(1)), and
      * for rho and E ALSO the cell's own value (pointwise relative, the north star's "relative L-inf"),
        floored at 1e-3 of the variable's scale so that a near-vacuum cell of an unlimited step from a
        random state does not divide by ~0."""
    fin = np.isfinite(Qb).all(axis=1)
    assert np.array_equal(np.isfinite(Qa).all(axis=1), fin), "finite masks differ"
    if not fin.any():
        return 0.0  # everything NaN on both sides (unlimited scheme on a random state)
    a, b = Qa[fin], Qb[fin]
    s = char_scales(b, gamma)
    d = np.abs(a - b)
    err = float((d / s).max())
    for k in (0, -1):
        err = max(err, float((d[:, k] / np.maximum(np.abs(b[:, k]), 1e-3 * s[k])).max()))
    return err


def rel_linf_pointwise(Qa, Qb, cols=(0, -1)):
    """max_c |Qa - Qb| / |Qb| on rho and E (no floor): for states that stay away from vacuum"""
    return max(float((np.abs(Qa[:, k] - Qb[:, k]) / np.abs(Qb[:, k])).max()) for k in cols)


def hex_box_flat(nx, ny, nz, bc=(3, 3, 3, 3, 3, 3)):
    """Structured hexahedral box (6 quad faces per cell) -> flat mesh.  Exercises
    the 7-point reconstruction stencil of the fused kernel (3-D extension; the
    reference's quad-face area vector has no 1/2, R/mesh/Face.cpp:30-35)."""
    from mstgpu import host
    nid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    cid = lambda i, j, k: (k * ny + j) * nx + i
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nodes = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    nodes[nid(I, J, K).ravel()] = np.stack([I.ravel() / nx, J.ravel() / ny * 0.8, K.ravel() / nz * 0.6], axis=1)
    inter, bnd = [], []
    for ax in range(3):
        n = [nx, ny, nz]
        rng_ = [range(n[0] + (ax == 0)), range(n[1] + (ax == 1)), range(n[2] + (ax == 2))]
        for i in rng_[0]:
            for j in rng_[1]:
                for k in rng_[2]:
                    p = [i, j, k]
                    a, b = (ax + 1) % 3, (ax + 2) % 3
                    q = []
                    for da, db in ((0, 0), (1, 0), (1, 1), (0, 1)):
                        v = list(p); v[a] += da; v[b] += db
                        q.append(nid(*v))
                    lo = list(p); lo[ax] -= 1
                    if p[ax] == 0:
                        bnd.append((q, cid(*p), -1, bc[2 * ax]))
                    elif p[ax] == n[ax]:
                        bnd.append((q, cid(*lo), -1, bc[2 * ax + 1]))
                    else:
                        inter.append((q, cid(*lo), cid(*p), 2))
    allf = inter + bnd
    raw = dict(dim=3, ncells=nx * ny * nz, nodes=nodes,
               face_nodes=np.array([f[0] for f in allf], dtype=np.int32),
               c0=np.array([f[1] for f in allf], dtype=np.int32), c1=np.array([f[2] for f in allf], dtype=np.int32),
               ftype=np.array([f[3] for f in allf], dtype=np.int32), nint=len(inter))
    return host.flatten_raw(raw)
