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
