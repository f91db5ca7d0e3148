"""Test helper: numpy emulation of the FUSED LU-SGS iteration of csrc/lusgs.cu (mode 1), in the order of
operations of the kernels, to show on the CPU that the re-association it makes is harmless (<= 1e-12 against
the oracle, which is pinned bit for bit to the reference's SparseSolver<MT,VCT>::solveILU).

The reference iteration (R/lusolver/SparseSolver.cpp:54-104), with XD[r,c] = X[r,c] D_c^-1:
    s   = U x                       (RHSUx before the D^-1)
    rhs = b + L D^-1 s              (RHSLDUx)
    v   : v[r] = rhs[r] - sum_{c<r} LD[r,c] v[c]          forward sweep in RHS space
    w0  = D (D^-1 v)
    w   : w[r] = w0[r] - sum_{c>r} UD[r,c] w[c]            backward sweep
    x   = D^-1 w
Fused: (1) the backward sweep already forms sum_{c>r} UD[r,c] w[c] = sum U[r,c] x[c] = s[r] of the NEXT
iteration; (2) rhs and the forward sweep collapse into v[r] = b[r] + sum_{c<r} LD[r,c] (s[c] - v[c]).
The unscaled blocks are then read once per solve (for the first s) instead of twice per iteration."""
import numpy as np


def solve(rowptr, col, val, b, x0, B, iters, lean=False):
    """lean = mode 2: no D (D^-1 v) round trip before the backward sweep (identity up to cond(D) eps)"""
    n = rowptr.size - 1
    val = val.reshape(-1, B, B)
    x = x0.reshape(n, B).copy()
    b = b.reshape(n, B)
    D = np.zeros((n, B, B))
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    for k in np.flatnonzero(col == rows):
        D[rows[k]] += val[k]
    Dinv = np.linalg.inv(D)
    XD = np.einsum("kij,kjl->kil", val, Dinv[col])            # k_scale: X[e] * Dinv[col[e]]
    # first s from the start vector and the unscaled blocks (k_ux without the D^-1)
    s = np.zeros((n, B))
    for r in range(n):
        for k in range(rowptr[r], rowptr[r + 1]):
            if col[k] > r:
                s[r] += val[k] @ x[col[k]]
    for _ in range(iters):
        v = np.zeros((n, B)); t = np.zeros((n, B))
        for r in range(n):                                     # forward, natural order
            acc = b[r].copy()
            for k in range(rowptr[r], rowptr[r + 1]):
                if col[k] < r:
                    acc += XD[k] @ t[col[k]]
            v[r] = acc
            t[r] = s[r] - acc
        w = v.copy() if lean else np.einsum("nij,nj->ni", D, np.einsum("nij,nj->ni", Dinv, v))   # k_mid
        s = np.zeros((n, B))
        for r in range(n - 1, -1, -1):                         # backward
            acc = w[r].copy(); ss = np.zeros(B)
            for k in range(rowptr[r + 1] - 1, rowptr[r] - 1, -1):
                if col[k] > r:
                    d = XD[k] @ w[col[k]]
                    acc -= d
                    ss += d
            w[r] = acc
            s[r] = ss
        x = np.einsum("nij,nj->ni", Dinv, w)                   # k_fin
    return x
