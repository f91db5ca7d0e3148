"""CPU suite, part 2: the C-ABI boundary without a GPU -- the shared library
loads, exports every symbol include/mstgpu.h declares, fails loudly when no
CUDA device exists, and its host-side plan (renumbering) is a valid, locality-
improving permutation."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, STEP_MESHES, have_gpu, load_flat, box_flat
import mstgpu


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mstgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mstgpu_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = mstgpu.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mstgpu.h but not exported"
    assert set(mstgpu.EXPORTS) == set(names)


def test_host_library_exports_every_declared_symbol():
    """include/msthost.h vs libmsthost.so (reader, flattener, writers, generators)"""
    from mstgpu import host
    hdr = open(os.path.join(ROOT, "include", "msthost.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(msthost_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 17
    L = host.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/msthost.h but not exported"


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_native_driver_fails_loudly_without_gpu(tmp_path):
    """mstrun (the reference's main program on the two C ABIs) reads and flattens the mesh on the host, then
    must stop with the CUDA error: there is no CPU path behind it"""
    import subprocess
    from conftest import load_raw
    from mstgpu import host
    exe = os.path.join(ROOT, "mst-cfd_b200", "mstrun")
    assert os.path.exists(exe), "mst-cfd_b200/mstrun is missing: make -C mst-cfd_b200"
    msh = str(tmp_path / "m.msh")
    host.write_msh(msh, load_raw("2d-stair-un-5-tri"))
    r = subprocess.run([exe, msh, "--steps", "2", "--out", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "mstgpu_create" in r.stderr and "cuda" in r.stderr.lower()
    r = subprocess.run([exe, str(tmp_path / "nope.msh")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr


def test_version_and_default_config():
    assert b"sm_100a" in mstgpu.lib().mstgpu_version()
    c = mstgpu.default_config(2)
    # R/include/CONST.h:38-48, SolverRoe.cpp:115
    assert (c.order, c.flux, c.viscous, c.renumber, c.kernel) == (2, 0, 0, 2, 1)
    assert (c.gamma, c.delta, c.eor, c.cv) == (1.4, 0.125, 1e-10, 715.8)
    assert c.inletQ[0] == 1.0 and c.inletQ[1] == 0.0 and abs(c.inletQ[3] - 2.5) < 1e-12
    c3 = mstgpu.default_config(3)
    assert c3.inletQ[3] == 0.0 and abs(c3.inletQ[4] - 2.5) < 1e-12


def test_struct_layout_matches_header():
    # 4 int32 + 12 pointers; 6 int32 + 6 doubles + 5 doubles
    assert C.sizeof(mstgpu.MstMesh) == 16 + 12 * 8
    assert C.sizeof(mstgpu.MstConfig) == 24 + 11 * 8 + 16 + 8 + 8 + 8  # + gradient, limiter, limiter_k, + tile_fit, reserved_


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    f = load_flat("2d-stair-un-5-tri")
    with pytest.raises(mstgpu.MstGpuError, match="CUDA"):
        mstgpu.Context(f)


def test_bad_arguments_are_rejected_before_touching_the_gpu():
    f = dict(load_flat("2d-stair-un-5-tri"))
    bad = dict(f); bad["dim"] = 4
    with pytest.raises(mstgpu.MstGpuError):
        mstgpu.plan_permutation(bad)
    bad = dict(f); c0 = f["c0"].copy(); c0[3] = f["ncells"] + 5; bad["c0"] = c0
    with pytest.raises(mstgpu.MstGpuError, match="c0 out of range"):
        mstgpu.plan_permutation(bad)
    with pytest.raises(mstgpu.MstGpuError):
        mstgpu.Context(f, order=3)
    # the connectivity is validated before anything indexes with it (plan, tile statistics, adjacency, partition)
    cases = []
    c1 = f["c1"].copy(); c1[5] = -7
    cases.append(("c1", c1, "c1 out of range"))
    ptr = f["cf_ptr"].copy(); ptr[4] = ptr[3] - 1
    cases.append(("cf_ptr", ptr, "non-decreasing"))
    ptr = f["cf_ptr"].copy(); ptr[0] = 1
    cases.append(("cf_ptr", ptr, r"cf_ptr\[0\]"))
    idx = f["cf_idx"].copy(); idx[7] = f["nfaces"] + 3
    cases.append(("cf_idx", idx, "cf_idx entry out of range"))
    idx = f["cf_idx"].copy(); idx[7] = -1
    cases.append(("cf_idx", idx, "cf_idx entry out of range"))
    idx = f["cf_idx"].copy()
    other = next(g for g in range(f["nfaces"]) if f["c0"][g] != 0 and f["c1"][g] != 0)
    idx[f["cf_ptr"][0]] = other
    cases.append(("cf_idx", idx, "does not touch the cell"))
    for key, arr, msg in cases:
        bad = dict(f); bad[key] = arr
        for call in (mstgpu.plan_permutation, mstgpu.tile_stats, mstgpu.mesh_adjacency, lambda m: mstgpu.Partition(m, 2, 0)):
            with pytest.raises(mstgpu.MstGpuError, match=msg):
                call(bad)


def _mean_neighbour_distance(f, perm):
    inv = np.empty_like(perm); inv[perm] = np.arange(perm.size, dtype=perm.dtype)
    i = f["c1"] >= 0
    return np.abs(inv[f["c0"][i]].astype(np.int64) - inv[f["c1"][i]]).mean()


@pytest.mark.parametrize("name", ["2d-shockwavepipe-2", "2d-stairW-1", "box"])
def test_plan_is_a_locality_improving_permutation(name):
    f = box_flat(12, 12, 12) if name == "box" else load_flat(name)
    cp, fp = mstgpu.plan_permutation(f, 2)
    cm, _ = mstgpu.plan_permutation(f, 1)
    assert np.array_equal(np.sort(cm), np.arange(f["ncells"]))
    assert np.array_equal(np.sort(cp), np.arange(f["ncells"]))
    assert np.array_equal(np.sort(fp), np.arange(f["nfaces"]))
    # interior faces stay in front of boundary faces
    nint = int((f["c1"] >= 0).sum())
    assert (f["c1"][fp[:nint]] >= 0).all() and (f["c1"][fp[nint:]] < 0).all()
    ident = np.arange(f["ncells"], dtype=np.int32)
    d_new, d_old = _mean_neighbour_distance(f, cp), _mean_neighbour_distance(f, ident)
    if name != "box":
        assert d_new < 0.5 * d_old, (d_new, d_old)
    # renumber = 0 keeps the reference order
    cp0, fp0 = mstgpu.plan_permutation(f, 0)
    assert np.array_equal(cp0, ident)


@pytest.mark.parametrize("order", [1, 2])
def test_tiling_statistics(order):
    """Host-side tiling of the fused kernel: every cell is owned by exactly one
    tile; Hilbert ranges are more compact than Morton ranges on a mesh that is
    not aligned with the octree; packets stay under the geometry budget."""
    f = box_flat(23, 23, 23)
    nc = f["ncells"]
    sh = mstgpu.tile_stats(f, order=order, tile_cells=96, renumber=2)
    sm = mstgpu.tile_stats(f, order=order, tile_cells=96, renumber=1)
    assert sh["tiles"] == -(-nc // 96)
    assert sh["sum_flux_faces"] >= f["nfaces"]            # every face is a flux face of some tile
    assert sh["sum_flux_faces"] < 1.5 * f["nfaces"]       # ... with bounded duplication
    assert sh["sum_ring1"] <= sm["sum_ring1"]
    assert sh["max_smem"] < 227 * 1024
    if order == 2:
        assert sh["packet_bytes"] / nc < 400
    # loop trips of a CTA (the tile-size predictor of DESIGN.md 8: padded phase-2 slots per cell)
    NT = sh["block_threads"]
    assert NT == 128   # tiles of <= 96 cells run 128-thread CTAs
    assert sh["face_trips"] * NT >= sh["sum_flux_faces"] and sh["face_trips"] <= sh["sum_flux_faces"] / NT + sh["tiles"]
    assert sh["cell_trips"] == sh["tiles"]                # 96 owned cells: one trip
    assert sh["ring_trips"] * NT >= sh["sum_ring1"] + sh["sum_ring2"]
