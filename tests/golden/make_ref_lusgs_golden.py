"""Regenerate tests/golden/ref_lusgs_*.npz by running the reference's own
SparseSolverNUM::solveILUSGS and SparseSolver<MT,VCT>::solveILU
(oracle/_ref/ref_lusgs, built by oracle/refbuild/Makefile) on seeded systems
with the sparsity of the reference meshes' cell adjacency.

    make -C oracle ref && python tests/golden/make_ref_lusgs_golden.py
"""
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_flat  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.dirname(os.path.abspath(__file__))


def system(mesh, block, seed):
    """Diagonally dominant block system on the cell adjacency graph of a mesh."""
    f = load_flat(mesh)
    n = f["ncells"]
    i = f["c1"] >= 0
    a, b = f["c0"][i], f["c1"][i]
    pat = sp.coo_matrix((np.ones(2 * a.size), (np.r_[a, b], np.r_[b, a])), shape=(n, n)).tocsr()
    pat = (pat + sp.eye(n)).tocsr()
    pat.sort_indices()
    rowptr, col = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
    rng = np.random.default_rng(seed)
    nnz = col.size
    val = rng.uniform(-1.0, 1.0, (nnz, block, block)) * 0.3
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    diag = rows == col
    val[diag] += np.eye(block) * (2.0 + 1.2 * block * np.diff(rowptr)[rows[diag]])[:, None, None] * 0.3
    bvec = rng.normal(size=(n, block))
    x0 = rng.uniform(0.5, 1.5, (n, block))
    return rowptr, col, val, bvec, x0


CASES = [("stair5_scalar", "2d-stair-un-5-tri", 1, 1), ("stair5_block4", "2d-stair-un-5-tri", 4, 2),
         ("stair3loose_scalar", "2d-stair-un-3-loose-tri", 1, 3)]

if __name__ == "__main__":
    tmp = os.path.join(REF, "tmp")
    os.makedirs(tmp, exist_ok=True)
    for name, mesh, block, seed in CASES:
        rowptr, col, val, b, x0 = system(mesh, block, seed)
        fin, fout = os.path.join(tmp, name + ".in"), os.path.join(tmp, name + ".out")
        with open(fin, "wb") as fh:
            np.array([rowptr.size - 1, block, col.size], dtype=np.int32).tofile(fh)
            rowptr.tofile(fh); col.tofile(fh); val.tofile(fh); b.tofile(fh); x0.tofile(fh)
        subprocess.run([os.path.join(REF, "ref_lusgs"), fin, fout], check=True, stdout=subprocess.DEVNULL)
        x = np.fromfile(fout).reshape(-1, block)
        np.savez_compressed(os.path.join(OUT, f"ref_lusgs_{name}.npz"), mesh=mesh, block=block, seed=seed, x=x)
        print(name, x.shape, float(np.abs(x).max()))
