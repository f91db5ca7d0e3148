"""Host-side tiling of the fused kernel (csrc/tiles.cpp, csrc/plan.cpp, the curve-cube search in csrc/mstgpu.cu): no GPU.
What the kernel's launch shape relies on is checked here; that the results do not depend on any of it (tile shapes, cell
order, curve lattice) is the GPU suite's job (bit-identity across renumberings / tile sizes / partitions)."""
import os

import numpy as np
import pytest

import mstgpu
from conftest import box_flat


def _stats(f, **kw):
    st = mstgpu.tile_stats(f, order=2, **kw)
    return st, f["ncells"]


def test_default_tiles_on_tets_are_sized_by_their_flux_faces():
    """tets at second order: tiles grown until the next cell would push the flux faces past 4 x 128 -> every warp of the
    128-thread CTA makes at most 4 trips through the flux phase, the tiles are nearly full, and all of them fit the
    4-CTAs-per-SM launch class"""
    f = box_flat(24, 24, 24)
    st, nc = _stats(f)
    assert st["block_threads"] == 128
    assert st["face_trips"] <= 4 * st["tiles"]                      # no fifth, nearly empty trip
    assert st["sum_flux_faces"] >= 0.93 * 512 * (st["tiles"] - 1)   # ... and the four are full
    assert st["le56k"] == st["tiles"] and st["max_smem"] <= 57000
    # against fixed tiles of 240 cells: fewer padded flux-face slots per cell
    fx, _ = _stats(f, tile_cells=240)
    assert fx["face_trips"] * fx["block_threads"] / nc > st["face_trips"] * st["block_threads"] / nc


def test_fitted_tiles_cover_every_cell_once_with_even_starts(monkeypatch):
    """the bulk copies of the state move 16-byte units (two 40-byte rows): tile starts must be even; sizes vary"""
    f = box_flat(13, 11, 9)
    for fit in (256, 512):
        monkeypatch.setenv("MSTGPU_TILE_FIT", str(fit))
        st, nc = _stats(f, tile_cells=300)
        assert st["face_trips"] * st["block_threads"] >= st["sum_flux_faces"]
        assert st["sum_flux_faces"] <= fit * st["tiles"]
        assert st["cell_trips"] >= st["tiles"]
    monkeypatch.delenv("MSTGPU_TILE_FIT")


def test_curve_cube_search_is_deterministic_and_optional(monkeypatch):
    f = box_flat(20, 20, 20)
    a, _ = mstgpu.plan_permutation(f, 2)
    b, _ = mstgpu.plan_permutation(f, 2)
    assert np.array_equal(a, b)
    assert np.array_equal(np.sort(a), np.arange(f["ncells"]))
    monkeypatch.setenv("MSTGPU_CURVE_SEARCH", "0")
    c, _ = mstgpu.plan_permutation(f, 2)
    assert np.array_equal(np.sort(c), np.arange(f["ncells"]))
    monkeypatch.setenv("MSTGPU_CURVE_SCALE", "1.37")  # any cube gives a valid order
    d, _ = mstgpu.plan_permutation(f, 2)
    assert np.array_equal(np.sort(d), np.arange(f["ncells"])) and not np.array_equal(c, d)


def test_tile_locality_metrics_are_consistent():
    f = box_flat(16, 16, 16)
    L = mstgpu.tile_locality(f)
    st, nc = _stats(f)
    assert L["tiles"] == st["tiles"] and L["flux_faces"] == st["sum_flux_faces"]
    assert L["ring_rows"] == st["sum_ring1"] + st["sum_ring2"]
    assert 0 < L["runs"] <= L["ring_rows"]
    # a 40-byte row touches one or two 128-byte lines; contiguous rows share them
    assert L["ring_rows"] * 40 // 128 <= L["lines"] <= 2 * L["ring_rows"]
    assert L["gather_sectors"] >= L["lines"]


def test_partition_tiling_statistics_take_the_owned_cells_only():
    f = box_flat(14, 12, 10)
    P = mstgpu.Partition(f, 3, 1, order=2)
    lf = P.local_flat()
    st = mstgpu.tile_stats(lf, order=2, n_owned=P.n_owned)
    whole = mstgpu.tile_stats(lf, order=2)
    assert st["tiles"] < whole["tiles"]                               # the ghost cells are not tiled
    assert st["sum_flux_faces"] >= 2 * P.n_owned                       # every owned cell's faces are evaluated here
    with pytest.raises(mstgpu.MstGpuError):
        mstgpu.tile_stats(lf, order=2, n_owned=lf["ncells"] + 1)


def test_curve_cube_search_finds_the_lattice_of_a_kuhn_box(monkeypatch):
    """37^3 hexahedra x 6 tets: with the octree anchored at the domain corner and its finest boxes = hexahedra (lattice
    candidate k = 6 of choose_curve_frame) tiles are unions of whole hexahedra -- fewer cut faces, fewer tiles than on the
    bounding cube, whose boxes (0.57 of a hexahedron) drift against the mesh"""
    f = box_flat(37, 37, 37)
    nc = f["ncells"]
    st = mstgpu.tile_stats(f, order=2)
    monkeypatch.setenv("MSTGPU_CURVE_SEARCH", "0")
    st0 = mstgpu.tile_stats(f, order=2)
    assert st["sum_flux_faces"] / nc < 2.40 < 2.45 < st0["sum_flux_faces"] / nc
    assert st["tiles"] < st0["tiles"]
    assert st["sum_ring1"] + st["sum_ring2"] < st0["sum_ring1"] + st0["sum_ring2"]
