"""ORACLE / TEST INFRASTRUCTURE -- not product code.

numpy restatement of the mesh metrics the reference computes after reading a
mesh (R = /root/reference/MST-CFD), SURVEY.md Appendix B:

  face centre / area vector   R/mesh/Face.cpp:8-44
  eta0                        R/mesh/Face.cpp:62-69
  cell centre                 R/mesh/Cell.cpp:6-14   (mean of FACE centres)
  cell volume 2-D             R/mesh/Cell.cpp:15-51  (Heron; quad = two Herons)
  directAndCells, flags       R/mesh/MshBlock.cpp:281-305
  outward Sout per cell slot  R/mesh/MshBlock.cpp:307-318
  per-cell face order         R/mesh/MshBlock.cpp:238-239,254-255 (file order)

3-D ("extension", parity unpinned): the reference's 3-D volume
(R/mesh/Cell.cpp:52-59) lacks the 1/3 and uses un-oriented normals, and the
3-D reader leaves flags uninitialised in one branch (MshBlock.cpp:635-647).
Here 3-D uses V = 1/3 sum Sout.(fc - cc) and the `consistent` flag rule.

The result ("flat mesh") is a dict of plain arrays in REFERENCE ORDER; it is
what both the CPU oracle and the C-ABI (`mstgpu_mesh`) consume.
"""
from __future__ import annotations

import numpy as np

EOR2 = 1e-7  # R/include/CONST.h:42


def _norm(v):
    # Eigen .norm() = sqrt(sum of squares), accumulated left to right
    s = v[..., 0] * v[..., 0]
    for k in range(1, v.shape[-1]):
        s = s + v[..., k] * v[..., k]
    return np.sqrt(s)


def cell_face_lists(c0, c1, ncells):
    """CSR cell->face lists in file order (faces appended as they are read:
    c0 first, then c1, R/mesh/MshBlock.cpp:238-239,254-255)."""
    nf = c0.shape[0]
    interior = c1 >= 0
    fid = np.arange(nf, dtype=np.int64)
    # key = 2*f for the c0 append, 2*f+1 for the c1 append -> global read order
    cells = np.concatenate([c0.astype(np.int64), c1[interior].astype(np.int64)])
    keys = np.concatenate([2 * fid, 2 * fid[interior] + 1])
    order = np.lexsort((keys, cells))
    cells_s = cells[order]
    faces_s = (keys[order] // 2).astype(np.int32)
    ptr = np.zeros(ncells + 1, dtype=np.int64)
    np.add.at(ptr, cells_s + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr.astype(np.int32), faces_s


def flatten(raw: dict, flag_convention: str = "consistent") -> dict:
    dim = int(raw["dim"])
    nodes = raw["nodes"]
    fn = raw["face_nodes"]
    c0 = raw["c0"].astype(np.int32)
    c1 = raw["c1"].astype(np.int32)
    ncells = int(raw["ncells"])
    nf = c0.shape[0]

    # ---- zones -> per-face type, nint (MshBlock.cpp:223-226) ----------------
    ftype = np.zeros(nf, dtype=np.int32)
    nint = 0
    for z in raw["zones"]:
        ftype[z["start"] : z["end"]] = z["type"]
        if z["type"] == 2:
            nint = z["end"]

    # ---- face centre and area vector (Face.cpp:8-44) ------------------------
    npf = (fn >= 0).sum(axis=1)
    fc = np.zeros((nf, dim))
    for k in range(fn.shape[1]):
        m = fn[:, k] >= 0
        fc[m] += nodes[fn[m, k]]
    fc = fc / npf[:, None].astype(np.float64)
    S = np.zeros((nf, dim))
    if dim == 2:
        f = nodes[fn[:, 0]] - nodes[fn[:, 1]]
        S[:, 0] = -f[:, 1]
        S[:, 1] = f[:, 0]
        area = _norm(f)
    else:
        f1 = nodes[fn[:, 1]] - nodes[fn[:, 0]]
        f2 = nodes[fn[:, 2]] - nodes[fn[:, 0]]
        cr = np.cross(f1, f2)
        tri = npf == 3
        S[tri] = 0.5 * cr[tri]
        S[~tri] = cr[~tri]  # quad: no 1/2, first three nodes only (Face.cpp:30-35)
        area = _norm(S)

    # ---- per-cell face lists -----------------------------------------------
    cf_ptr, cf_idx = cell_face_lists(c0, c1, ncells)
    nslot_c = (cf_ptr[1:] - cf_ptr[:-1]).astype(np.int32)
    nslot = int(nslot_c.max())

    # ---- cell centre = mean of face centres, summed in list order ----------
    cc = np.zeros((ncells, dim))
    for j in range(nslot):
        m = nslot_c > j
        cc[m] += fc[cf_idx[cf_ptr[:-1][m] + j]]
    cc = cc / nslot_c[:, None].astype(np.float64)

    # ---- volume --------------------------------------------------------------
    vol = np.zeros(ncells)
    ln = _norm(S)  # getDirect().norm()
    if dim == 2:
        t3 = np.nonzero(nslot_c == 3)[0]
        if t3.size:
            b = cf_ptr[:-1][t3]
            l0, l1, l2 = ln[cf_idx[b]], ln[cf_idx[b + 1]], ln[cf_idx[b + 2]]
            s = 0.5 * (l0 + l1 + l2)
            vol[t3] = np.sqrt(s * (s - l0) * (s - l1) * (s - l2))
        t4 = np.nonzero(nslot_c == 4)[0]
        if t4.size:
            b = cf_ptr[:-1][t4]
            f0, f1_, f2_, f3 = (cf_idx[b + k] for k in range(4))
            v1 = fc[f0] - fc[f1_]
            v2 = fc[f2_] - fc[f3]
            par = np.abs(v1[:, 0] * v2[:, 1] - v2[:, 0] * v1[:, 1]) < EOR2

            def heron2(fa, fb, fc_, fd, mid_other):
                mid = 2.0 * _norm(fc[fa] - fc[mid_other])
                s1 = 0.5 * (ln[fa] + ln[fb] + mid)
                v = np.sqrt(s1 * (s1 - ln[fa]) * (s1 - ln[fb]) * (s1 - mid))
                s2 = 0.5 * (ln[fc_] + ln[fd] + mid)
                v = v + np.sqrt(s2 * (s2 - ln[fc_]) * (s2 - ln[fd]) * (s2 - mid))
                return v

            va = heron2(f0, f1_, f2_, f3, f1_)  # Cell.cpp:36-40
            vb = heron2(f0, f2_, f1_, f3, f2_)  # Cell.cpp:43-47
            vol[t4] = np.where(par, va, vb)

    # ---- orientation (MshBlock.cpp:282-283) ---------------------------------
    dot = np.einsum("ij,ij->i", S, fc - cc[c0])
    dac = np.where(dot < 0, -1, 1).astype(np.int8)

    if dim == 3:
        # extension: V = 1/3 sum_j Sout_j . (fc_j - cc)
        for j in range(nslot):
            m = nslot_c > j
            f = cf_idx[cf_ptr[:-1][m] + j]
            cid = np.nonzero(m)[0]
            sgn = np.where(c0[f] == cid, 1.0, -1.0) * dac[f]
            vol[m] += sgn * np.einsum("ij,ij->i", S[f], fc[f] - cc[cid])
        vol = vol / 3.0

    # ---- eta0 (Face.cpp:62-69) ------------------------------------------------
    eta = np.ones(nf)
    ii = np.nonzero(c1 >= 0)[0]
    d0 = _norm(cc[c0[ii]] - fc[ii])
    d1 = _norm(cc[c1[ii]] - fc[ii])
    eta[ii] = d1 / (d0 + d1)

    # ---- left/right flags (MshBlock.cpp:284-303) --------------
    outward_c0 = dac[:, None] * S
    if flag_convention == "consistent":
        flag = outward_c0 >= 0
    elif flag_convention == "as_shipped":
        if dim == 2:
            flag = np.where((dac == -1)[:, None], outward_c0 >= 0, ~(outward_c0 >= 0))
        else:
            # 3-D reader sets flags only in the `< 0` branch; the rest is
            # uninitialised memory in the reference -> defined as false here.
            flag = np.where((dac == -1)[:, None], outward_c0 >= 0, False)
    else:
        raise ValueError(flag_convention)

    return dict(
        dim=dim,
        ncells=ncells,
        nfaces=nf,
        nint=int(nint),
        c0=np.ascontiguousarray(c0),
        c1=np.ascontiguousarray(c1),
        S=np.ascontiguousarray(S),
        dac=np.ascontiguousarray(dac),
        fc=np.ascontiguousarray(fc),
        area=np.ascontiguousarray(area),
        eta=np.ascontiguousarray(eta),
        flag=np.ascontiguousarray(flag.astype(np.uint8)),
        ftype=np.ascontiguousarray(ftype),
        cc=np.ascontiguousarray(cc),
        vol=np.ascontiguousarray(vol),
        cf_ptr=np.ascontiguousarray(cf_ptr),
        cf_idx=np.ascontiguousarray(cf_idx),
    )


def sod_initial_state(flat: dict, gamma: float = 1.4) -> np.ndarray:
    """R/time/Time.cpp:13-38 with the CONST.h:70-75 base state
    (rho=1, u=v=0, E = rho*(T*CV) = 715.8/286.32 -> p = 1)."""
    dim = flat["dim"]
    U = dim + 2
    iniT = 1 / 286.32
    iniE = 1 * (iniT * 715.8 + 0.5 * (0 * 0 + 0 * 0))
    Q = np.zeros((flat["ncells"], U))
    Q[:, 0] = 1.0
    Q[:, U - 1] = iniE
    right = flat["cc"][:, 0] > 0.5
    Q[right, 0] *= 0.125
    Q[right, U - 1] *= 0.1
    return Q


def random_state(flat: dict, seed: int = 20231017, gamma: float = 1.4) -> np.ndarray:
    """Parity stress input of SURVEY.md 8d: rho, p in [0.5,1.5], velocity
    components in [-1.5,1.5] (|M|>1 both signs)."""
    rng = np.random.default_rng(seed)
    dim = flat["dim"]
    n = flat["ncells"]
    U = dim + 2
    rho = rng.uniform(0.5, 1.5, n)
    p = rng.uniform(0.5, 1.5, n)
    vel = rng.uniform(-1.5, 1.5, (n, dim))
    Q = np.zeros((n, U))
    Q[:, 0] = rho
    Q[:, 1 : 1 + dim] = rho[:, None] * vel
    Q[:, U - 1] = p / (gamma - 1) + 0.5 * rho * (vel**2).sum(axis=1)
    return Q
