"""GPU suite: level-scheduled LU-SGS sweeps through the C ABI against the oracle
(which is pinned bit-exactly to the reference's SparseSolverNUM / SparseSolver).
Tolerance 1e-12 relative (FMA contraction and the 4x4 inverse algorithm differ)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle
import mstgpu
from golden.make_ref_lusgs_golden import system

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("block", [1, 4, 5])
@pytest.mark.parametrize("mesh", ["2d-stair-un-5-tri", "2d-stairW-1"])
def test_gpu_sweeps_match_reference_order(block, mesh):
    rowptr, col, val, b, x0 = system(mesh, block, 7)
    s = mstgpu.LuSgs(rowptr, col, block)
    xg, hist, it = s.solve(val, b, x0, 5, early_exit=False)
    xo, ho, _ = oracle.lusgs(rowptr, col, val, b, x0, block, 5, early_exit=False)
    assert _rel(xg, xo) <= 1e-12
    if block == 1:
        assert np.allclose(hist, ho, rtol=1e-9)
    f, bk = s.levels()
    assert 1 < f <= rowptr.size - 1 and bk > 1


def test_colour_ordered_system_has_few_levels_and_same_answer():
    rowptr, col, val, b, x0 = system("2d-stairW-1", 4, 9)
    n = rowptr.size - 1
    perm, ncol = mstgpu.lusgs_color_order(rowptr, col)
    A = sp.bsr_matrix((val, col, rowptr), shape=(4 * n, 4 * n))
    # permute the BLOCK system: P A P^T, P b, P x0
    Ab = sp.csr_matrix((np.arange(col.size) + 1, col, rowptr), shape=(n, n))[perm][:, perm].tocsr()
    Ab.sort_indices()
    valp = val[Ab.data - 1]
    sp_ = mstgpu.LuSgs(Ab.indptr, Ab.indices, 4)
    f, bk = sp_.levels()
    assert f <= ncol and bk <= ncol  # one level per colour
    xg, _, _ = sp_.solve(valp, b[perm], x0[perm], 5)
    xo, _, _ = oracle.lusgs(Ab.indptr, Ab.indices, valp, b[perm], x0[perm], 4, 5)
    assert _rel(xg, xo) <= 1e-12  # parity is defined on the permuted system
    # both orderings iterate towards the same solution of A x = b
    xs = sp.linalg.spsolve(A.tocsc(), b.ravel()).reshape(n, 4)
    x60, _, _ = sp_.solve(valp, b[perm], x0[perm], 60)
    assert np.abs(x60 - xs[perm]).max() < 1e-9 * np.abs(xs).max()


def test_scalar_early_exit_and_errors():
    rowptr, col, val, b, x0 = system("2d-stair-un-5-tri", 1, 3)
    s = mstgpu.LuSgs(rowptr, col, 1)
    xg, hist, it = s.solve(val, b, x0, 60, early_exit=True)
    xo, ho, ito = oracle.lusgs(rowptr, col, val, b, x0, 1, 60, early_exit=True)
    assert it == ito and it < 60
    assert _rel(xg, xo) <= 1e-12
    with pytest.raises(mstgpu.MstGpuError, match="block size"):
        mstgpu.LuSgs(rowptr, col, 3)
    bad = col.copy(); bad[0] = rowptr.size + 5
    with pytest.raises(mstgpu.MstGpuError, match="out of range"):
        mstgpu.LuSgs(rowptr, bad, 1)


@pytest.mark.parametrize("block", [1, 5])
def test_sweep_order_without_moving_the_data(block):
    """Colour-ordered sweeps on the matrix in STORAGE order == the reference's solver on the explicitly
    permuted system P A P^T (oracle), mapped back."""
    rowptr, col, val, b, x0 = system("2d-stairW-1", block, 11)
    n = rowptr.size - 1
    perm, ncol = mstgpu.lusgs_color_order(rowptr, col)
    s = mstgpu.LuSgs(rowptr, col, block, sweep_order=perm)
    f, bk = s.levels()
    assert f <= ncol and bk <= ncol
    xg, _, _ = s.solve(val, b, x0, 5)
    Ab = sp.csr_matrix((np.arange(col.size) + 1, col, rowptr), shape=(n, n))[perm][:, perm].tocsr()
    Ab.sort_indices()
    valp = val[Ab.data - 1]
    xo, _, _ = oracle.lusgs(Ab.indptr, Ab.indices, valp, b[perm], x0[perm], block, 5)
    assert _rel(xg[perm], xo) <= 1e-12
    # and bit-identical to the GPU solve of the explicitly permuted system
    xp, _, _ = mstgpu.LuSgs(Ab.indptr, Ab.indices, block).solve(valp, b[perm], x0[perm], 5)
    assert np.array_equal(xg[perm], xp)
    with pytest.raises(mstgpu.MstGpuError, match="permutation"):
        mstgpu.LuSgs(rowptr, col, block, sweep_order=np.zeros(n, np.int32))


def test_device_resident_solve_on_a_tet_mesh_pattern():
    """BASELINE config 5 in small: block-5 system on the tet adjacency (device cell order), colour
    sweeps, arrays resident on the device -- same bits as the host-array entry point, == oracle."""
    import torch
    from conftest import box_flat
    f = box_flat(7, 6, 5)
    rowptr, col = mstgpu.mesh_adjacency(f)
    n, nnz, B = rowptr.size - 1, col.size, 5
    assert n == f["ncells"] and nnz == n + 2 * f["nint"]
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    assert np.all(np.diff(col)[np.diff(rows) == 0] > 0)  # ascending within a row
    rng = np.random.default_rng(3)
    val = (rng.random((nnz, B, B)) - 0.5) * 0.2
    val[col == rows] += 3.0 * np.eye(B)
    b = rng.random((n, B)); x0 = np.ones((n, B))
    perm, ncol = mstgpu.lusgs_color_order(rowptr, col)
    assert ncol <= 6
    s = mstgpu.LuSgs(rowptr, col, B, sweep_order=perm)
    xh, _, _ = s.solve(val, b, x0, 5)
    dv, db, dx = torch.from_numpy(val).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(x0).cuda()
    ms = s.solve_device(dv.data_ptr(), db.data_ptr(), dx.data_ptr(), 5)
    assert ms > 0 and s.launch_count > 0
    assert np.array_equal(dx.cpu().numpy(), xh)
    Ab = sp.csr_matrix((np.arange(col.size) + 1, col, rowptr), shape=(n, n))[perm][:, perm].tocsr()
    Ab.sort_indices()
    xo, _, _ = oracle.lusgs(Ab.indptr, Ab.indices, val[Ab.data - 1], b[perm], x0[perm], B, 5)
    assert _rel(xh[perm], xo) <= 1e-12
    # 60 sweeps reach the solution of A x = b
    x60, _, _ = s.solve(val, b, x0, 60)
    A = sp.bsr_matrix((val, col, rowptr), shape=(n * B, n * B)).tocsc()
    xs = sp.linalg.spsolve(A, b.ravel()).reshape(n, B)
    assert np.abs(x60 - xs).max() < 1e-9 * np.abs(xs).max()


@pytest.mark.parametrize("block", [1, 4, 5])
def test_fused_mode_matches_the_reference_passes(block):
    """mstgpu_lusgs_set_mode(1): half the off-diagonal block traffic per iteration, same iterate (re-associated):
    against the oracle (reference order) and against mode 0 on the same handle, natural and colour sweep order"""
    rowptr, col, val, b, x0 = system("2d-stairW-1", block, 13)
    xo, ho, _ = oracle.lusgs(rowptr, col, val, b, x0, block, 5, early_exit=False)
    s = mstgpu.LuSgs(rowptr, col, block)
    x0_, h0, _ = s.solve(val, b, x0, 5)
    s.set_mode(1)
    x1, h1, _ = s.solve(val, b, x0, 5)
    assert _rel(x1, xo) <= 1e-12 and _rel(x1, x0_) <= 1e-13
    if block == 1:
        assert np.allclose(h1, ho, rtol=1e-9)
    xz, _, _ = s.solve(val, b, np.zeros_like(x0), 3)     # zero start vector: the implicit step's case
    xoz, _, _ = oracle.lusgs(rowptr, col, val, b, np.zeros_like(x0), block, 3, early_exit=False)
    assert _rel(xz, xoz) <= 1e-12
    perm, _ = mstgpu.lusgs_color_order(rowptr, col)
    so = mstgpu.LuSgs(rowptr, col, block, sweep_order=perm)
    xa, _, _ = so.solve(val, b, x0, 5)
    so.set_mode(1)
    xb, _, _ = so.solve(val, b, x0, 5)
    assert _rel(xb, xa) <= 1e-13
    # mode 2 (lean): mode 1 without the D (D^-1 v) round trip and with x = D^-1 w formed inside the backward sweep
    so.set_mode(2)
    xc, _, _ = so.solve(val, b, x0, 5)
    assert _rel(xc, xa) <= 1e-12
    s.set_mode(2)
    x2, h2, _ = s.solve(val, b, x0, 5)
    assert _rel(x2, xo) <= 1e-12
    xz2, _, _ = s.solve(val, b, np.zeros_like(x0), 3)
    assert _rel(xz2, xoz) <= 1e-12
    if block == 1:
        assert np.allclose(h2, ho, rtol=1e-9)  # residual history / early exit: the scalar solver keeps k_fin
    with pytest.raises(mstgpu.MstGpuError):
        so.set_mode(3)
