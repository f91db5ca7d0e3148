"""ORACLE / TEST INFRASTRUCTURE -- not product code.

numpy restatement of the arithmetic of the reference's Tecplot writer
(Work::writedataRhoBasedMshNodePlt, R/work/Work.cpp:243-304, R = /root/reference/MST-CFD),
in the reference's order of operations (numpy rounds every elementwise operation separately,
like the reference build with -ffp-contract=off):

  Work.cpp:248-285  face state: eta*Q[c0] + (1-eta)*Q[c1] on type-2 faces, Q[c0] on inlet (10),
                    wall (3; FLAG_WALLSMOOTH = 1, CONST.h:16) and outlet (5) faces.  Zone types the
                    switch does not list (symmetry 7) leave the reference's array uninitialised;
                    Q[c0] is used here (build-defined, stated in DESIGN.md).
  Work.cpp:288-295  node state: sum_f w*Qf / sum_f w over the node's faces in Node::addNbFace order
                    with w = 1/area(face[NODE id]) -- the reference indexes the face list with the
                    node index, so the weight is the same for every face of a node.
  Work.cpp:299-303  rho, u_i = Q_i/rho, getT, getP, getMa (R/work/FUNCTION.cpp:8-20; the 3-D form
                    adds the third momentum component: extension).

Pinned to the reference itself: tests/test_output_cpu.py compares the file written from these numbers
with the one the reference's own writer produces (oracle/_ref/ref_io), byte for byte."""
import numpy as np

GAMMA, CV = 1.4, 715.8  # CONST.h:39,41


def node_weights(f, raw):
    """w[i] = 1 / Face::getArea() of face i (Face.cpp:17,27: |S|), for node i (Work.cpp:292)."""
    nn = raw["nodes"].shape[0]
    S = f["S"]
    if S.shape[0] < nn:
        raise ValueError("the reference reads face[node id]: needs nfaces >= nnodes")
    area = np.sqrt((S[:nn] ** 2).sum(axis=1)) if S.shape[1] == 3 else np.sqrt(S[:nn, 0] * S[:nn, 0] + S[:nn, 1] * S[:nn, 1])
    return 1.0 / area


def node_fields(f, raw, Q, nf_ptr, nf_idx, w=None):
    """[nnodes, dim+4]: rho, u_i, T, p, Ma."""
    D = int(f["dim"])
    U = D + 2
    c0, c1, eta = f["c0"], f["c1"], f["eta"]
    interior = f["ftype"] == 2
    Qf = Q[c0].copy()
    e = eta[interior][:, None]
    Qf[interior] = e * Q[c0[interior]] + (1.0 - e) * Q[c1[interior]]
    nn = raw["nodes"].shape[0]
    w = node_weights(f, raw) if w is None else w
    acc = np.zeros((nn, U))
    d = np.zeros(nn)
    deg = np.diff(nf_ptr)
    for j in range(int(deg.max())):          # j-th face of every node that has one: the reference's order
        m = deg > j
        fj = nf_idx[nf_ptr[:-1][m] + j]
        acc[m] = acc[m] + w[m][:, None] * Qf[fj]
        d[m] = d[m] + w[m]
    with np.errstate(all="ignore"):
        nq = acc / d[:, None]
        rho = nq[:, 0]
        m2 = nq[:, 1] * nq[:, 1] + nq[:, 2] * nq[:, 2]
        if D == 3:
            m2 = m2 + nq[:, 3] * nq[:, 3]
        ek = 0.5 * m2 / rho
        T = (nq[:, U - 1] - ek) / rho / CV
        p = (nq[:, U - 1] - ek) * (GAMMA - 1)
        Ma = np.sqrt(m2 / (GAMMA * p * rho))
        out = np.empty((nn, D + 4))
        out[:, 0] = rho
        for i in range(D):
            out[:, 1 + i] = nq[:, 1 + i] / rho
        out[:, D + 1], out[:, D + 2], out[:, D + 3] = T, p, Ma
    return out
