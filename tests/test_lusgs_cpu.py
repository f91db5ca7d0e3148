"""CPU suite, part 5: LU-SGS.  The oracle restatement
against the reference's own SparseSolverNUM / SparseSolver<MT,VCT> (golden
vectors from oracle/_ref/ref_lusgs) and against the fixed point A x = b; the
host-side colour ordering of the GPU path."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg  # noqa: F401

from conftest import GOLDEN
from oracle import oracle
import mstgpu
from golden.make_ref_lusgs_golden import system

CASES = sorted(os.path.basename(p)[10:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_lusgs_*.npz")))


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_solver(case):
    g = np.load(os.path.join(GOLDEN, f"ref_lusgs_{case}.npz"))
    block = int(g["block"])
    rowptr, col, val, b, x0 = system(str(g["mesh"]), block, int(g["seed"]))
    # the reference runs LU_INTERVAL = 5 iterations (CONST.h:58); the scalar version may stop early
    x, hist, it = oracle.lusgs(rowptr, col, val, b, x0, block, 5, early_exit=(block == 1))
    assert np.array_equal(x.reshape(-1, block), g["x"]), np.abs(x.reshape(-1, block) - g["x"]).max()


@pytest.mark.parametrize("block", [1, 4, 5])
def test_fixed_point_is_the_linear_solve(block):
    rowptr, col, val, b, x0 = system("2d-stair-un-5-tri", block, 5)
    n = rowptr.size - 1
    x, _, _ = oracle.lusgs(rowptr, col, val, b, x0, block, 60)
    A = sp.bsr_matrix((val, col, rowptr), shape=(n * block, n * block)).tocsc()
    xs = sp.linalg.spsolve(A, b.ravel())
    assert np.abs(x.ravel() - xs).max() < 1e-9 * np.abs(xs).max()


def test_color_order_is_a_valid_colouring():
    rowptr, col, *_ = system("2d-stair-un-5-tri", 1, 1)
    perm, ncol = mstgpu.lusgs_color_order(rowptr, col)
    n = rowptr.size - 1
    assert np.array_equal(np.sort(perm), np.arange(n))
    assert 2 <= ncol <= 5  # triangle adjacency: at most 4 colours by greedy
    # rows of one colour are mutually uncoupled: permuted matrix is block-diagonal-free per colour
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    off = rows != col
    # recompute colours from the order: a row's colour = number of colour boundaries before it
    A = sp.csr_matrix((np.ones(col.size), col, rowptr), shape=(n, n))
    Ap = A[perm][:, perm].tocsr()
    # level count of the permuted lower triangle == number of colours at most
    lev = np.zeros(n, int)
    for r in range(n):
        c = Ap.indices[Ap.indptr[r]:Ap.indptr[r + 1]]
        c = c[c < r]
        lev[r] = (lev[c].max() + 1) if c.size else 0
    assert lev.max() + 1 <= ncol


@pytest.mark.parametrize("block, mesh", [(1, "2d-stair-un-3-loose-tri"), (4, "2d-stair-un-5-tri"), (5, "2d-stair-un-5-tri")])
def test_fused_iteration_is_the_reference_iteration_reassociated(block, mesh):
    """mode 1 of csrc/lusgs.cu (mstgpu_lusgs_set_mode), restated on the CPU in the kernels' order of operations
    (tests/lusgs_fused_np.py): the backward sweep's by-product is U x of the next iteration and the right-hand
    side folds into the forward sweep.  Against the oracle, which equals the reference build bit for bit."""
    import lusgs_fused_np
    rowptr, col, val, b, x0 = system(mesh, block, 11)
    xo, _, _ = oracle.lusgs(rowptr, col, val, b, x0, block, 5, early_exit=False)
    xf = lusgs_fused_np.solve(rowptr, col, val, b, x0, block, 5)
    assert np.abs(xf - xo.reshape(xf.shape)).max() <= 1e-13 * np.abs(xo).max()
    # and from a zero start vector (the implicit step's case: the first U x vanishes)
    z = np.zeros_like(x0)
    xo, _, _ = oracle.lusgs(rowptr, col, val, b, z, block, 3, early_exit=False)
    xf = lusgs_fused_np.solve(rowptr, col, val, b, z, block, 3)
    assert np.abs(xf - xo.reshape(xf.shape)).max() <= 1e-13 * np.abs(xo).max()
    # mode 2 (lean): additionally drops the reference's w0 = D (D^-1 v) round trip (identity up to cond(D) eps)
    xo, _, _ = oracle.lusgs(rowptr, col, val, b, x0, block, 5, early_exit=False)
    xl = lusgs_fused_np.solve(rowptr, col, val, b, x0, block, 5, lean=True)
    assert np.abs(xl - xo.reshape(xl.shape)).max() <= 1e-13 * np.abs(xo).max()
