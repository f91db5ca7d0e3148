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
