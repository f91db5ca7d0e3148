"""ctypes binding of libmsthost.so (mst-cfd_b200/host/msthost.cpp): synthetic
mesh generators for the BASELINE configs and the multi-threaded flattener that
turns raw mesh tables into the flat tables of include/mstgpu.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "libmsthost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C mst-cfd_b200`")
        L = C.CDLL(LIB_PATH)
        vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
        L.msthost_box_tets_sizes.argtypes = [C.c_int] * 3 + [i64p] * 4
        L.msthost_box_tets_sizes.restype = None
        L.msthost_box_tets.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [vp] * 6
        L.msthost_grid_tris_sizes.argtypes = [C.c_int, C.c_int, vp] + [i64p] * 4
        L.msthost_grid_tris_sizes.restype = None
        L.msthost_grid_tris.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double] + [vp] * 7
        L.msthost_flatten.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int] + [vp] * 4 + [C.c_int] + [vp] * 9
        L.msthost_last_error.restype = C.c_char_p
        L.msthost_msh_read.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.msthost_msh_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(vp)]
        L.msthost_msh_free.argtypes = [vp]
        L.msthost_msh_free.restype = None
        L.msthost_msh_sizes.argtypes = [vp, i64p]
        L.msthost_msh_tables.argtypes = [vp] * 7
        L.msthost_msh_zone_name.argtypes = [vp, C.c_int32]
        L.msthost_msh_zone_name.restype = C.c_char_p
        L.msthost_node_faces.argtypes = [C.c_int64, C.c_int64, C.c_int32, vp, vp, vp]
        L.msthost_cell_nodes.argtypes = [C.c_int32, C.c_int64, C.c_int32] + [vp] * 6
        L.msthost_plt_write.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64] + [vp] * 4 + [C.c_int32, C.c_int32]
        L.msthost_plt_write_binary.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64] + [vp] * 4 + [C.c_int32]
        L.msthost_msh_write.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32] + [vp] * 4 + [C.c_int32, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data


def flatten_raw(raw: dict, flag_convention: str = "consistent") -> dict:
    """raw tables -> flat mesh dict (same keys as oracle/mesh_np.flatten)."""
    dim = int(raw["dim"])
    nodes = np.ascontiguousarray(raw["nodes"], dtype=np.float64)
    fn = np.ascontiguousarray(raw["face_nodes"], dtype=np.int32)
    c0 = np.ascontiguousarray(raw["c0"], dtype=np.int32)
    c1 = np.ascontiguousarray(raw["c1"], dtype=np.int32)
    nc, nf = int(raw["ncells"]), c0.shape[0]
    if "ftype" in raw:
        ftype = np.ascontiguousarray(raw["ftype"], dtype=np.int32)
        nint = int(raw["nint"])
    else:
        ftype = np.zeros(nf, dtype=np.int32)
        nint = 0
        for z in raw["zones"]:
            ftype[z["start"]:z["end"]] = z["type"]
            if z["type"] == 2:
                nint = z["end"]
    S = np.empty((nf, dim)); fc = np.empty((nf, dim)); dac = np.empty(nf, dtype=np.int8)
    eta = np.empty(nf); flag = np.empty((nf, dim), dtype=np.uint8)
    cc = np.empty((nc, dim)); vol = np.zeros(nc)
    cf_ptr = np.empty(nc + 1, dtype=np.int32)
    cf_idx = np.empty(int(nf + (c1 >= 0).sum()), dtype=np.int32)
    conv = {"consistent": 0, "as_shipped": 1}[flag_convention]
    rc = lib().msthost_flatten(dim, nodes.shape[0], nc, nf, fn.shape[1], _p(nodes), _p(fn), _p(c0),
                               _p(c1), conv, _p(S), _p(fc), _p(dac), _p(eta), _p(flag), _p(cc),
                               _p(vol), _p(cf_ptr), _p(cf_idx))
    if rc != 0:
        raise RuntimeError("msthost_flatten failed")
    return dict(dim=dim, ncells=nc, nfaces=nf, nint=nint, c0=c0, c1=c1, S=S, dac=dac, fc=fc,
                eta=eta, flag=flag, ftype=ftype, cc=cc, vol=vol, cf_ptr=cf_ptr, cf_idx=cf_idx)


def _msh_handle_to_raw(L, h) -> dict:
    try:
        sz = (C.c_int64 * 8)()
        L.msthost_msh_sizes(h, sz)
        dim, nn, nc, nf, nint, nz, npf = (int(v) for v in sz[:7])
        nodes = np.empty((nn, dim)); fn = np.empty((nf, npf), dtype=np.int32)
        c0 = np.empty(nf, dtype=np.int32); c1 = np.empty(nf, dtype=np.int32); ft = np.empty(nf, dtype=np.int32)
        zt = np.empty((nz, 5), dtype=np.int32)
        L.msthost_msh_tables(h, _p(nodes), _p(fn), _p(c0), _p(c1), _p(ft), _p(zt))
        zones = [dict(id=int(z[0]), start=int(z[1]), end=int(z[2]), type=int(z[3]), npf=int(z[4]),
                      name=L.msthost_msh_zone_name(h, k).decode("ascii", "replace")) for k, z in enumerate(zt)]
    finally:
        L.msthost_msh_free(h)
    return dict(dim=dim, ncells=nc, nodes=nodes, face_nodes=fn, c0=c0, c1=c1, ftype=ft, nint=nint, zones=zones)


def read_msh(path: str) -> dict:
    """Native reader for the Fluent ASCII .msh subset of the reference's MshBlock
    (R/mesh/MshBlock.cpp:75-271, host/mshread.cpp): raw tables, same keys as the generators."""
    L = lib()
    h = C.c_void_p()
    rc = L.msthost_msh_read(os.fsencode(path), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"msthost_msh_read({path}): {L.msthost_last_error().decode()}")
    return _msh_handle_to_raw(L, h)


def parse_msh(text: bytes) -> dict:
    """read_msh on a memory image of the file."""
    L = lib()
    h = C.c_void_p()
    rc = L.msthost_msh_parse(text, len(text), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"msthost_msh_parse: {L.msthost_last_error().decode()}")
    return _msh_handle_to_raw(L, h)


def write_msh(path: str, raw: dict):
    """Raw tables -> .msh file the reference's own reader accepts (multi-threaded writer)."""
    L = lib()
    raw = raw if "zones" in raw else raw_zones_from_ftype(raw)
    nodes = np.ascontiguousarray(raw["nodes"], dtype=np.float64)
    fn = np.ascontiguousarray(raw["face_nodes"], dtype=np.int32)
    c0 = np.ascontiguousarray(raw["c0"], dtype=np.int32); c1 = np.ascontiguousarray(raw["c1"], dtype=np.int32)
    zt = np.array([[z.get("id", k + 7), z["start"], z["end"], z["type"],
                    z.get("npf", int((fn[z["start"]] >= 0).sum()) if z["end"] > z["start"] else fn.shape[1])]
                   for k, z in enumerate(raw["zones"])], dtype=np.int32).reshape(-1, 5)
    rc = L.msthost_msh_write(os.fsencode(path), int(raw["dim"]), nodes.shape[0], int(raw["ncells"]), c0.shape[0],
                             fn.shape[1], _p(nodes), _p(fn), _p(c0), _p(c1), zt.shape[0], _p(zt))
    if rc != 0:
        raise RuntimeError(f"msthost_msh_write({path}): {L.msthost_last_error().decode()}")


def node_faces(raw: dict):
    """(nf_ptr, nf_idx): a node's faces in the order Node::addNbFace builds it (R/mesh/Node.cpp:13-15)."""
    fn = np.ascontiguousarray(raw["face_nodes"], dtype=np.int32)
    nn = int(np.asarray(raw["nodes"]).shape[0])
    ptr = np.empty(nn + 1, dtype=np.int32)
    idx = np.empty(int((fn >= 0).sum()), dtype=np.int32)
    if lib().msthost_node_faces(nn, fn.shape[0], fn.shape[1], _p(fn), _p(ptr), _p(idx)) != 0:
        raise RuntimeError("msthost_node_faces failed")
    return ptr, idx


def node_weights(flat: dict, nnodes: int) -> np.ndarray:
    """The per-node weight of the reference's node averaging: 1 / Face::getArea() of face[NODE id]
    (the reference indexes its face list with the node index, R/work/Work.cpp:292-293)."""
    S = np.asarray(flat["S"])[:nnodes]
    if S.shape[0] < nnodes:
        raise ValueError("the reference reads face[node id]: needs nfaces >= nnodes")
    a2 = S[:, 0] * S[:, 0] + S[:, 1] * S[:, 1]
    if S.shape[1] == 3:
        a2 = a2 + S[:, 2] * S[:, 2]
    return 1.0 / np.sqrt(a2)


def cell_nodes(raw: dict, flat: dict):
    """(cn_ptr, cn_idx): Cell::getBeginItPNbNodes as the reference builds it (MshBlock.cpp:335-368)."""
    fn = np.ascontiguousarray(raw["face_nodes"], dtype=np.int32)
    nodes = np.ascontiguousarray(raw["nodes"], dtype=np.float64)
    nc = int(flat["ncells"])
    ptr = np.empty(nc + 1, dtype=np.int32)
    L = lib()
    args = (int(raw["dim"]), nc, fn.shape[1], _p(fn), _p(flat["cf_ptr"]), _p(flat["cf_idx"]), _p(nodes))
    if L.msthost_cell_nodes(*args, _p(ptr), None) != 0:
        raise RuntimeError("msthost_cell_nodes failed")
    idx = np.empty(int(ptr[-1]), dtype=np.int32)
    if L.msthost_cell_nodes(*args, _p(ptr), _p(idx)) != 0:
        raise RuntimeError("msthost_cell_nodes failed")
    return ptr, idx


def plt_write(path: str, raw: dict, fields: np.ndarray, cn_ptr, cn_idx, zone_t: int = 0, felnum: int = 4, binary: bool = False):
    """The reference's node-averaged Tecplot file (Work.cpp:204-319) from node fields [nnodes, dim+4]
    (mstgpu.Context.node_fields); byte-identical to the reference's own writer.  binary=True: raw dump."""
    nodes = np.ascontiguousarray(raw["nodes"], dtype=np.float64)
    fields = np.ascontiguousarray(fields, dtype=np.float64)
    dim = int(raw["dim"])
    assert fields.shape == (nodes.shape[0], dim + 4)
    L = lib()
    if binary:
        rc = L.msthost_plt_write_binary(os.fsencode(path), dim, nodes.shape[0], cn_ptr.size - 1, _p(nodes), _p(fields),
                                        _p(cn_ptr), _p(cn_idx), int(zone_t))
    else:
        rc = L.msthost_plt_write(os.fsencode(path), dim, nodes.shape[0], cn_ptr.size - 1, _p(nodes), _p(fields),
                                 _p(cn_ptr), _p(cn_idx), int(zone_t), int(felnum))
    if rc != 0:
        raise RuntimeError(f"msthost_plt_write({path}) failed ({rc})")


def box_tets_raw(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, bc=(3, 3, 3, 3, 3, 3)) -> dict:
    """Kuhn 6-tet split of an nx*ny*nz hex box (BASELINE configs 4-5)."""
    L = lib()
    nn, nc, nf, ni = (C.c_int64() for _ in range(4))
    L.msthost_box_tets_sizes(nx, ny, nz, C.byref(nn), C.byref(nc), C.byref(nf), C.byref(ni))
    if nf.value >= 2**30:
        raise ValueError("mesh too large for 32-bit face ids with the side bit")
    nodes = np.empty((nn.value, 3)); fn = np.empty((nf.value, 3), dtype=np.int32)
    c0 = np.empty(nf.value, dtype=np.int32); c1 = np.empty(nf.value, dtype=np.int32)
    ft = np.empty(nf.value, dtype=np.int32)
    bca = np.asarray(bc, dtype=np.int32)
    rc = L.msthost_box_tets(nx, ny, nz, lx, ly, lz, _p(bca), _p(nodes), _p(fn), _p(c0), _p(c1), _p(ft))
    if rc != 0:
        raise RuntimeError("msthost_box_tets failed")
    return dict(dim=3, ncells=nc.value, nodes=nodes, face_nodes=fn, c0=c0, c1=c1, ftype=ft, nint=ni.value)


def grid_tris_raw(nx, ny, lx, ly, mask=None, bc=(10, 5, 3)) -> dict:
    """Structured 2-D grid, active quads (mask) split into two triangles."""
    L = lib()
    mask = np.ones((ny, nx), dtype=np.uint8) if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    nn, nc, nf, ni = (C.c_int64() for _ in range(4))
    L.msthost_grid_tris_sizes(nx, ny, _p(mask), C.byref(nn), C.byref(nc), C.byref(nf), C.byref(ni))
    nodes = np.empty((nn.value, 2)); fn = np.empty((nf.value, 2), dtype=np.int32)
    c0 = np.empty(nf.value, dtype=np.int32); c1 = np.empty(nf.value, dtype=np.int32)
    ft = np.empty(nf.value, dtype=np.int32)
    bca = np.asarray(bc, dtype=np.int32)
    rc = L.msthost_grid_tris(nx, ny, lx, ly, _p(mask), _p(bca), _p(nodes), _p(fn), _p(c0), _p(c1), _p(ft))
    if rc != 0:
        raise RuntimeError("msthost_grid_tris failed")
    return dict(dim=2, ncells=nc.value, nodes=nodes, face_nodes=fn, c0=c0, c1=c1, ftype=ft, nint=ni.value)


def forward_step_raw(h_inv=445, bc=(10, 5, 3)) -> dict:
    """BASELINE config 2: 3 x 1 channel, step of height 0.2 from x = 0.6,
    uniform h = 1/h_inv, quads split into triangles (h_inv = 445 -> 998 046)."""
    nx, ny = 3 * h_inv, h_inv
    mask = np.ones((ny, nx), dtype=np.uint8)
    i0 = int(round(0.6 * h_inv)); j1 = int(round(0.2 * h_inv))
    mask[:j1, i0:] = 0
    return grid_tris_raw(nx, ny, 3.0, 1.0, mask, bc)


def raw_zones_from_ftype(raw: dict) -> dict:
    """Add a `zones` list (contiguous runs of ftype) so oracle/mesh_np.flatten
    accepts a generated mesh."""
    ft = raw["ftype"]
    cuts = np.flatnonzero(np.diff(ft)) + 1
    starts = np.concatenate([[0], cuts]); ends = np.concatenate([cuts, [ft.shape[0]]])
    out = dict(raw)
    out["zones"] = [dict(id=k, start=int(s), end=int(e), type=int(ft[s]), name="")
                    for k, (s, e) in enumerate(zip(starts, ends))]
    return out


def tets_to_raw(nodes: np.ndarray, tets: np.ndarray, boundary_type) -> dict:
    """Raw mesh tables of a conforming tetrahedral mesh given as (nodes, cell -> 4 nodes): faces are
    found by matching sorted node triples (twice = interior, once = boundary).  boundary_type(fc) maps
    the boundary face centres [nb,3] to zone types; boundary faces are grouped by type, interior first."""
    nc = tets.shape[0]
    combos = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]])
    tri = np.sort(tets[:, combos].reshape(-1, 3), axis=1)           # [4 nc, 3]
    cell = np.repeat(np.arange(nc, dtype=np.int64), 4)
    nn = int(nodes.shape[0])
    key = (tri[:, 0].astype(np.int64) * nn + tri[:, 1]) * nn + tri[:, 2]
    order = np.argsort(key, kind="stable")
    key, tri, cell = key[order], tri[order], cell[order]
    first = np.ones(key.size, bool); first[1:] = key[1:] != key[:-1]
    idx = np.flatnonzero(first)
    cnt = np.diff(np.append(idx, key.size))
    if cnt.max() > 2:
        raise ValueError("non-manifold face")
    inter = idx[cnt == 2]
    bnd = idx[cnt == 1]
    fc_b = nodes[tri[bnd]].mean(axis=1)
    bt = np.asarray(boundary_type(fc_b), dtype=np.int32)
    bo = np.argsort(bt, kind="stable")
    bnd, bt = bnd[bo], bt[bo]
    face_nodes = np.concatenate([tri[inter], tri[bnd]]).astype(np.int32)
    c0 = np.concatenate([np.minimum(cell[inter], cell[inter + 1]), cell[bnd]]).astype(np.int32)
    c1 = np.concatenate([np.maximum(cell[inter], cell[inter + 1]), np.full(bnd.size, -1)]).astype(np.int32)
    ftype = np.concatenate([np.full(inter.size, 2, np.int32), bt])
    return dict(dim=3, ncells=nc, nodes=np.ascontiguousarray(nodes, dtype=np.float64), face_nodes=face_nodes,
                c0=c0, c1=c1, ftype=ftype, nint=int(inter.size))


def sphere_shell_raw(n=42, m=40, r0=0.5, r1=10.0) -> dict:
    """BASELINE config 3: cubed-sphere shell r in [r0, r1] around a sphere, 6 patches x n^2 x m
    hexahedra (geometric radial stretching), each split into 24 tetrahedra about its centroid and its
    face centres (always conforming): n = 42, m = 40 -> 10 160 640 tets.  Sphere = wall (3), outer
    boundary = inlet (10) upstream (x < 0) / outlet (5) downstream."""
    # surface lattice of the cube [0, n]^3: nodes with a coordinate on 0 or n, numbered by coordinates
    g = np.arange(n + 1)
    I, J, K = np.meshgrid(g, g, g, indexing="ij")
    on = (I == 0) | (I == n) | (J == 0) | (J == n) | (K == 0) | (K == n)
    sid = -np.ones((n + 1,) * 3, dtype=np.int64)
    sid[on] = np.arange(int(on.sum()))
    ns = int(on.sum())
    # equi-angular direction of every surface node
    t = np.tan(np.pi / 4 * (2.0 * np.stack([I[on], J[on], K[on]], axis=1) / n - 1.0))
    dirs = t / np.linalg.norm(t, axis=1, keepdims=True)
    radii = r0 * (r1 / r0) ** (np.arange(m + 1) / m)
    shell_nodes = (radii[:, None, None] * dirs[None]).reshape(-1, 3)   # node (layer l, surface s) = l * ns + s
    # surface quads of the 6 cube faces
    quads = []
    a, b = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a, b = a.ravel(), b.ravel()
    for ax in range(3):
        for side in (0, n):
            def nid(da, db):
                c = [None, None, None]
                c[ax] = np.full(a.size, side)
                c[(ax + 1) % 3] = a + da
                c[(ax + 2) % 3] = b + db
                return sid[c[0], c[1], c[2]]
            quads.append(np.stack([nid(0, 0), nid(1, 0), nid(1, 1), nid(0, 1)], axis=1))
    quads = np.concatenate(quads)                                        # [6 n^2, 4]
    nq = quads.shape[0]
    lay = np.arange(m)
    lo = (lay[:, None, None] * ns + quads[None]).reshape(-1, 4)          # inner quad of each hex
    hi = lo + ns
    nh = lo.shape[0]
    hexes = np.concatenate([lo, hi], axis=1)                             # 0-3 inner, 4-7 outer
    # hex faces as node quads (local ids), each split about its centre
    hf = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]])
    fq = hexes[:, hf]                                                    # [nh, 6, 4] global node ids
    # face-centre nodes: unique per geometric face (shared between the two hexes of an interior quad)
    fkey = np.sort(fq.reshape(-1, 4), axis=1)
    nn0 = shell_nodes.shape[0]
    uniq, inv = np.unique(fkey, axis=0, return_inverse=True)
    fcen = shell_nodes[uniq].mean(axis=1)
    fc_id = nn0 + inv.reshape(nh, 6)
    hc_id = nn0 + uniq.shape[0] + np.arange(nh)
    hcen = shell_nodes[hexes].mean(axis=1)
    nodes = np.concatenate([shell_nodes, fcen, hcen])
    # 24 tets: (corner e, corner e+1, face centre, hex centre) for every edge of every face
    e0 = fq
    e1 = np.roll(fq, -1, axis=2)
    tets = np.stack([e0, e1, np.broadcast_to(fc_id[:, :, None], e0.shape),
                     np.broadcast_to(hc_id[:, None, None], e0.shape)], axis=3).reshape(-1, 4)

    def btype(fc):
        r = np.linalg.norm(fc, axis=1)
        return np.where(r < np.sqrt(r0 * r1), 3, np.where(fc[:, 0] < 0.0, 10, 5))

    return tets_to_raw(nodes, tets, btype)
