"""TEST INFRASTRUCTURE: the partitioned implicit step, restated with the oracle.

Definition being checked (SURVEY.md 8e "LU-SGS"): every rank assembles the rows of
its OWNED cells (its ghost cells supply states for the residual and the Jacobians); the couplings
to ghost cells are LAGGED -- block Jacobi across partitions:

    for every LU-SGS iteration:
        b_eff = b - G x_ghost          (x_ghost: the neighbours' dQ of the previous iteration, 0 at first)
        x_owned <- one iteration of the reference's solveILU on the local system (A_owned, b_eff)
        exchange x (ghost rows <- owners)
    Q_owned += x_owned

The local sweep order is whatever the rank uses (colour order of its owned pattern on the GPU);
the oracle runs the reference solver on the explicitly permuted local system."""
import numpy as np
import scipy.sparse as sp

from oracle import oracle


def exchange(parts, arrs):
    """ghost rows <- owners' rows, for every rank (in-process stand-in for the NCCL send / recv)"""
    for r, P in enumerate(parts):
        for nb in P.neighbors:
            src = parts[nb["rank"]]
            back = [x for x in src.neighbors if x["rank"] == r][0]
            arrs[r][nb["recv_first"]:nb["recv_first"] + nb["recv_count"]] = arrs[nb["rank"]][back["send_local"]]


def local_systems(ors, parts, Qs, dt):
    """per rank: (A_owned CSR pattern+blocks, ghost blocks as (row, col, block) lists, b_owned)"""
    out = []
    for o, P, Q in zip(ors, parts, Qs):
        rowptr, col, val, b = o.implicit_system(dt, Q)
        no = P.n_owned
        rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
        keep = rows < no
        own = keep & (col < no)
        gho = keep & (col >= no)
        rp = np.zeros(no + 1, dtype=np.int32)
        np.add.at(rp, rows[own] + 1, 1)
        rp = np.cumsum(rp).astype(np.int32)
        out.append(dict(rowptr=rp, col=col[own].astype(np.int32), val=val[own], b=b[:no],
                        grow=rows[gho], gcol=col[gho], gval=val[gho]))
    return out


def one_iteration(sysr, x, order, U):
    """one sweep pair of the reference's block solver on the local system, ghost couplings lagged"""
    no = sysr["rowptr"].size - 1
    beff = sysr["b"].copy()
    if sysr["grow"].size:
        np.subtract.at(beff, sysr["grow"], np.einsum("eij,ej->ei", sysr["gval"], x[sysr["gcol"]]))
    rp, col, val = sysr["rowptr"], sysr["col"], sysr["val"]
    if order is None:
        xn, _, _ = oracle.lusgs(rp, col, val, beff, x[:no], U, 1)
        return xn.reshape(no, U)
    perm = np.asarray(order)
    Ab = sp.csr_matrix((np.arange(col.size) + 1, col, rp), shape=(no, no))[perm][:, perm].tocsr()
    Ab.sort_indices()
    xp, _, _ = oracle.lusgs(Ab.indptr, Ab.indices, val[Ab.data - 1], beff[perm], x[:no][perm], U, 1)
    xn = np.empty((no, U))
    xn[perm] = xp.reshape(no, U)
    return xn


def step(ors, parts, Qs, dt, iters, orders=None):
    """One partitioned implicit step, in place on Qs (list of [n_local, U] arrays, owned rows first)."""
    U = Qs[0].shape[1]
    exchange(parts, Qs)
    systems = local_systems(ors, parts, Qs, dt)
    xs = [np.zeros_like(Q) for Q in Qs]
    for it in range(iters):
        new = [one_iteration(s, x, None if orders is None else orders[r], U)
               for r, (s, x) in enumerate(zip(systems, xs))]
        for r, P in enumerate(parts):
            xs[r][:P.n_owned] = new[r]
        exchange(parts, xs)
    for r, P in enumerate(parts):
        Qs[r][:P.n_owned] += xs[r][:P.n_owned]
    return Qs
