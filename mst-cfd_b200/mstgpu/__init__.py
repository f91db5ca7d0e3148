"""mstgpu -- ctypes binding of the C ABI in include/mstgpu.h (libmstgpu.so).

This is the thinnest possible host layer: the same calls a cgo/JNI/C++ host
would make.  `GpuRhoSolver` mirrors the reference's seven-method solver
interface (R/rhoSolver/RhoSolver.h:17-24) so tests read like the reference's
call sequence in R/time/Time.cpp:54-81.

There is no CPU fallback: importing works without a GPU (the library loads and
its symbols resolve), but every compute entry point returns an error when no
CUDA device is usable, and a missing libmstgpu.so raises at import.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("MSTGPU_LIB", os.path.join(_ROOT, "libmstgpu.so"))  # override: A/B builds

EXPORTS = [
    "mstgpu_default_config", "mstgpu_create", "mstgpu_destroy", "mstgpu_set_state",
    "mstgpu_get_state", "mstgpu_get_prev_state", "mstgpu_step", "mstgpu_step_timed", "mstgpu_step_host", "mstgpu_host_register", "mstgpu_host_unregister",
    "mstgpu_residual_linf", "mstgpu_sync", "mstgpu_cfl_dt", "mstgpu_step_cfl", "mstgpu_step_cfl_timed", "mstgpu_implicit_setup",
    "mstgpu_implicit_sweep_order", "mstgpu_step_implicit", "mstgpu_debug_gradient", "mstgpu_debug_face_flux",
    "mstgpu_launch_count", "mstgpu_enable_kernel_timing", "mstgpu_kernel_time",
    "mstgpu_device_bytes", "mstgpu_plan_permutation", "mstgpu_tile_stats", "mstgpu_tile_stats_owned", "mstgpu_tile_locality",
    "mstgpu_partition_create", "mstgpu_partition_destroy", "mstgpu_partition_mesh", "mstgpu_partition_sizes",
    "mstgpu_partition_cell_ids", "mstgpu_partition_neighbor", "mstgpu_create_partitioned",
    "mstgpu_comm_unique_id", "mstgpu_comm_init", "mstgpu_peer_blob_bytes", "mstgpu_peer_export", "mstgpu_peer_connect", "mstgpu_peer_disable", "mstgpu_lusgs_create", "mstgpu_lusgs_destroy",
    "mstgpu_lusgs_solve", "mstgpu_lusgs_levels", "mstgpu_lusgs_create_ordered", "mstgpu_lusgs_solve_device",
    "mstgpu_lusgs_create_partitioned", "mstgpu_lusgs_color_order_partitioned",
    "mstgpu_lusgs_launch_count", "mstgpu_lusgs_device_bytes", "mstgpu_mesh_adjacency", "mstgpu_lusgs_color_order", "mstgpu_lusgs_last_error",
    "mstgpu_output_setup", "mstgpu_output_setup_partitioned", "mstgpu_output_node_count", "mstgpu_output_node_ids", "mstgpu_node_fields", "mstgpu_set_tile_variant", "mstgpu_lusgs_set_mode",
    "mstgpu_last_error", "mstgpu_version",
]


class MstMesh(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ncells", C.c_int32), ("nfaces", C.c_int32), ("nint", C.c_int32),
        ("c0", C.c_void_p), ("c1", C.c_void_p), ("S", C.c_void_p), ("dac", C.c_void_p),
        ("fc", C.c_void_p), ("eta", C.c_void_p), ("flag", C.c_void_p), ("ftype", C.c_void_p),
        ("cc", C.c_void_p), ("vol", C.c_void_p), ("cf_ptr", C.c_void_p), ("cf_idx", C.c_void_p),
    ]


class MstConfig(C.Structure):
    _fields_ = [
        ("order", C.c_int32), ("flux", C.c_int32), ("viscous", C.c_int32),
        ("qf_copy_from", C.c_int32), ("renumber", C.c_int32), ("device", C.c_int32),
        ("gamma", C.c_double), ("delta", C.c_double), ("eor", C.c_double),
        ("mu", C.c_double), ("kappa", C.c_double), ("cv", C.c_double),
        ("inletQ", C.c_double * 5),
        ("kernel", C.c_int32), ("tile_cells", C.c_int32), ("block_threads", C.c_int32),
        ("tile_flags", C.c_int32),
        # extension (absent from the reference): gradient / limiter choice, include/mstgpu.h
        ("gradient", C.c_int32), ("limiter", C.c_int32), ("limiter_k", C.c_double),
        ("tile_fit", C.c_int32), ("reserved_", C.c_int32),
    ]

GRADIENTS = {"gg": 0, "lsq": 1}
LIMITERS = {"none": 0, "bj": 1, "venkat": 2}


class MstGpuError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libmstgpu.so.  Raises if the CUDA extension has not been built --
    the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MstGpuError(
                f"{LIB_PATH} is missing: build it with `make -C mst-cfd_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback"
            )
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.mstgpu_default_config.argtypes = [C.POINTER(MstConfig), i32]
        L.mstgpu_default_config.restype = None
        L.mstgpu_create.argtypes = [C.POINTER(vp), C.POINTER(MstMesh), C.POINTER(MstConfig)]
        L.mstgpu_destroy.argtypes = [vp]
        L.mstgpu_destroy.restype = None
        L.mstgpu_set_state.argtypes = [vp, vp, i64]
        L.mstgpu_get_state.argtypes = [vp, vp]
        L.mstgpu_get_prev_state.argtypes = [vp, vp]
        L.mstgpu_step.argtypes = [vp, dbl, i32]
        L.mstgpu_step_timed.argtypes = [vp, dbl, i32, C.POINTER(C.c_float)]
        L.mstgpu_step_host.argtypes = [vp, vp, vp, dbl, i32]
        L.mstgpu_host_register.argtypes = [vp, C.c_size_t]
        L.mstgpu_host_unregister.argtypes = [vp]
        L.mstgpu_cfl_dt.argtypes = [vp, dbl, C.POINTER(dbl)]
        L.mstgpu_step_cfl.argtypes = [vp, dbl, i32, C.POINTER(dbl)]
        L.mstgpu_step_cfl_timed.argtypes = [vp, dbl, i32, C.POINTER(dbl), C.POINTER(C.c_float)]
        L.mstgpu_implicit_setup.argtypes = [vp, i32]
        L.mstgpu_implicit_sweep_order.argtypes = [vp, vp]
        L.mstgpu_step_implicit.argtypes = [vp, dbl, i32, i32, C.POINTER(C.c_float)]
        L.mstgpu_residual_linf.argtypes = [vp, vp]
        L.mstgpu_sync.argtypes = [vp]
        L.mstgpu_output_setup.argtypes = [vp, C.POINTER(MstMesh), i32, vp, vp, vp]
        L.mstgpu_node_fields.argtypes = [vp, vp]
        L.mstgpu_output_setup_partitioned.argtypes = [vp, vp, C.POINTER(MstMesh), i32, vp, vp, vp]
        L.mstgpu_output_node_count.argtypes = [vp]
        L.mstgpu_output_node_count.restype = i32
        L.mstgpu_output_node_ids.argtypes = [vp, vp]
        L.mstgpu_set_tile_variant.argtypes = [vp, i32]
        L.mstgpu_debug_gradient.argtypes = [vp, vp]
        L.mstgpu_debug_face_flux.argtypes = [vp, vp]
        L.mstgpu_launch_count.argtypes = [vp]
        L.mstgpu_launch_count.restype = i64
        L.mstgpu_enable_kernel_timing.argtypes = [vp, i32]
        L.mstgpu_kernel_time.argtypes = [vp, C.c_char_p, C.POINTER(dbl), C.POINTER(i64)]
        L.mstgpu_device_bytes.argtypes = [vp]
        L.mstgpu_device_bytes.restype = i64
        L.mstgpu_plan_permutation.argtypes = [C.POINTER(MstMesh), C.POINTER(MstConfig), vp, vp]
        L.mstgpu_tile_stats.argtypes = [C.POINTER(MstMesh), C.POINTER(MstConfig), vp]
        L.mstgpu_partition_create.argtypes = [C.POINTER(vp), C.POINTER(MstMesh), C.POINTER(MstConfig), i32, i32, vp]
        L.mstgpu_partition_destroy.argtypes = [vp]
        L.mstgpu_partition_destroy.restype = None
        L.mstgpu_partition_mesh.argtypes = [vp]
        L.mstgpu_partition_mesh.restype = C.POINTER(MstMesh)
        L.mstgpu_partition_sizes.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.mstgpu_partition_cell_ids.argtypes = [vp]
        L.mstgpu_partition_cell_ids.restype = C.POINTER(i32)
        L.mstgpu_partition_neighbor.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.POINTER(i32)),
                                                C.POINTER(i32), C.POINTER(i32)]
        L.mstgpu_create_partitioned.argtypes = [C.POINTER(vp), vp, C.POINTER(MstConfig)]
        L.mstgpu_comm_unique_id.argtypes = [vp]
        L.mstgpu_comm_init.argtypes = [vp, i32, i32, vp]
        L.mstgpu_peer_blob_bytes.argtypes = []
        L.mstgpu_peer_blob_bytes.restype = i64
        L.mstgpu_peer_export.argtypes = [vp, i32, vp]
        L.mstgpu_peer_connect.argtypes = [vp, i32, i32, vp]
        L.mstgpu_peer_disable.argtypes = [vp]
        L.mstgpu_lusgs_create.argtypes = [C.POINTER(vp), i32, i32, vp, vp, i32]
        L.mstgpu_lusgs_create_ordered.argtypes = [C.POINTER(vp), i32, i32, vp, vp, vp, i32]
        L.mstgpu_lusgs_solve_device.argtypes = [vp, vp, vp, vp, i32, C.POINTER(C.c_float)]
        L.mstgpu_lusgs_set_mode.argtypes = [vp, i32]
        L.mstgpu_lusgs_launch_count.argtypes = [vp]
        L.mstgpu_lusgs_launch_count.restype = i64
        L.mstgpu_lusgs_device_bytes.argtypes = [vp]
        L.mstgpu_lusgs_device_bytes.restype = i64
        L.mstgpu_mesh_adjacency.argtypes = [C.POINTER(MstMesh), C.POINTER(MstConfig), vp, vp, i64]
        L.mstgpu_lusgs_destroy.argtypes = [vp]
        L.mstgpu_lusgs_destroy.restype = None
        L.mstgpu_lusgs_solve.argtypes = [vp, vp, vp, vp, i32, i32, vp, C.POINTER(i32)]
        L.mstgpu_lusgs_levels.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.mstgpu_lusgs_color_order.argtypes = [i32, vp, vp, vp, C.POINTER(i32)]
        L.mstgpu_lusgs_last_error.restype = C.c_char_p
        L.mstgpu_last_error.argtypes = [vp]
        L.mstgpu_last_error.restype = C.c_char_p
        L.mstgpu_version.restype = C.c_char_p
        _lib = L
    return _lib


_MESH_SPEC = dict(c0=np.int32, c1=np.int32, S=np.float64, dac=np.int8, fc=np.float64,
                  eta=np.float64, flag=np.uint8, ftype=np.int32, cc=np.float64,
                  vol=np.float64, cf_ptr=np.int32, cf_idx=np.int32)


def default_config(dim: int) -> MstConfig:
    cfg = MstConfig()
    lib().mstgpu_default_config(C.byref(cfg), dim)
    return cfg


def _mesh_struct(flat: dict):
    m = MstMesh()
    m.dim, m.ncells, m.nfaces, m.nint = int(flat["dim"]), int(flat["ncells"]), int(flat["nfaces"]), int(flat["nint"])
    keep = {}
    for k, dt in _MESH_SPEC.items():
        a = np.ascontiguousarray(flat[k], dtype=dt)
        keep[k] = a
        setattr(m, k, a.ctypes.data)
    return m, keep


def plan_permutation(flat: dict, renumber: int = 2):
    """(cell_new2old, face_new2old) that mstgpu_create would use; host only."""
    m, keep = _mesh_struct(flat)
    cfg = default_config(int(flat["dim"]))
    cfg.renumber = renumber
    c = np.empty(int(flat["ncells"]), dtype=np.int32)
    f = np.empty(int(flat["nfaces"]), dtype=np.int32)
    rc = lib().mstgpu_plan_permutation(C.byref(m), C.byref(cfg), c.ctypes.data, f.ctypes.data)
    if rc != 0:
        raise MstGpuError(f"plan_permutation failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    return c, f


def mesh_adjacency(flat: dict, renumber: int = 2):
    """(rowptr, col): CSR cell adjacency + diagonal in the device cell order (host only)."""
    m, keep = _mesh_struct(flat)
    cfg = default_config(int(flat["dim"]))
    cfg.renumber = renumber
    rowptr = np.empty(int(flat["ncells"]) + 1, dtype=np.int32)
    rc = lib().mstgpu_mesh_adjacency(C.byref(m), C.byref(cfg), rowptr.ctypes.data, None, 0)
    if rc != 0:
        raise MstGpuError(f"mesh_adjacency failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    col = np.empty(int(rowptr[-1]), dtype=np.int32)
    rc = lib().mstgpu_mesh_adjacency(C.byref(m), C.byref(cfg), rowptr.ctypes.data, col.ctypes.data, col.size)
    if rc != 0:
        raise MstGpuError(f"mesh_adjacency failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    return rowptr, col


def tile_stats(flat: dict, order: int = 2, tile_cells: int = 0, renumber: int = 2, n_owned: int = -1) -> dict:
    """Statistics of the fused kernel's tiling (host only).  n_owned >= 0: `flat` is a partition's local mesh
    (Partition.local_flat()), only its first n_owned cells are renumbered and tiled."""
    m, keep = _mesh_struct(flat)
    cfg = default_config(int(flat["dim"]))
    cfg.order, cfg.tile_cells, cfg.renumber = order, tile_cells, renumber
    out = np.zeros(16, dtype=np.int64)
    lib().mstgpu_tile_stats_owned.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    rc = lib().mstgpu_tile_stats_owned(C.byref(m), C.byref(cfg), int(n_owned), out.ctypes.data)
    if rc != 0:
        raise MstGpuError(f"tile_stats failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    keys = ("tiles", "max_smem", "mean_smem", "sum_ring1", "sum_ring2", "sum_flux_faces", "sum_local_faces",
            "packet_bytes", "le56k", "le75k", "le113k", "more", "face_trips", "cell_trips", "ring_trips", "block_threads")
    return dict(zip(keys, (int(x) for x in out)))


def tile_locality(flat: dict, order: int = 2, renumber: int = 2, n_owned: int = -1) -> dict:
    """How scattered the ring rows of the tiles are in memory (host only; include/mstgpu.h)."""
    m, keep = _mesh_struct(flat)
    cfg = default_config(int(flat["dim"]))
    cfg.order, cfg.renumber = order, renumber
    out = np.zeros(6, dtype=np.int64)
    lib().mstgpu_tile_locality.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    rc = lib().mstgpu_tile_locality(C.byref(m), C.byref(cfg), int(n_owned), out.ctypes.data)
    if rc != 0:
        raise MstGpuError(f"tile_locality failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    return dict(tiles=int(out[0]), ring_rows=int(out[1]), lines=int(out[2]), runs=int(out[3]), gather_sectors=int(out[4]), flux_faces=int(out[5]))


def _apply_consts(cfg, consts):
    for k, v in consts.items():
        if k == "gradient" and isinstance(v, str):
            v = GRADIENTS[v]
        if k == "limiter" and isinstance(v, str):
            v = LIMITERS[v]
        if not hasattr(cfg, k):
            raise MstGpuError(f"unknown config field {k!r}")
        setattr(cfg, k, v)


def make_config(dim, order=2, flux="roe", viscous=0, qf_copy_from=None, renumber=2, device=-1, inletQ=None,
                kernel=None, tile_cells=0, block_threads=0, **consts) -> MstConfig:
    cfg = default_config(dim)
    cfg.order = order
    cfg.flux = {"roe": 0, "ausm": 1}[flux] if isinstance(flux, str) else int(flux)
    cfg.viscous = viscous
    cfg.qf_copy_from = -1 if qf_copy_from is None else qf_copy_from
    cfg.renumber = renumber
    cfg.device = device
    if kernel is not None:
        cfg.kernel = {"tiles": 1, "split": 0}[kernel] if isinstance(kernel, str) else int(kernel)
    cfg.tile_cells = tile_cells
    cfg.block_threads = block_threads
    _apply_consts(cfg, consts)
    if inletQ is not None:
        for k in range(5):
            cfg.inletQ[k] = float(inletQ[k]) if k < len(inletQ) else 0.0
    return cfg


class Partition:
    """One rank's piece of a global flat mesh (host only): local mesh tables,
    local->global cell ids, neighbour send / receive lists."""

    def __init__(self, flat: dict, nparts: int, rank: int, cell_part=None, **cfg_kw):
        L = lib()
        m, keep = _mesh_struct(flat)
        self.cfg = make_config(int(flat["dim"]), **cfg_kw)
        cp = None if cell_part is None else np.ascontiguousarray(cell_part, dtype=np.int32)
        h = C.c_void_p()
        rc = L.mstgpu_partition_create(C.byref(h), C.byref(m), C.byref(self.cfg), nparts, rank,
                                       None if cp is None else cp.ctypes.data)
        if rc != 0:
            raise MstGpuError(f"partition_create failed ({rc}): {L.mstgpu_last_error(None).decode()}")
        self.h = h
        self.dim = int(flat["dim"])
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        L.mstgpu_partition_sizes(h, C.byref(a), C.byref(b), C.byref(c))
        self.n_owned, self.n_local, self.n_neighbors = a.value, b.value, c.value
        self.cell_ids = np.ctypeslib.as_array(L.mstgpu_partition_cell_ids(h), shape=(self.n_local,)).copy()
        self.neighbors = []
        for i in range(self.n_neighbors):
            r, sc, rf, rcnt = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            sp = C.POINTER(C.c_int32)()
            L.mstgpu_partition_neighbor(h, i, C.byref(r), C.byref(sc), C.byref(sp), C.byref(rf), C.byref(rcnt))
            send = np.ctypeslib.as_array(sp, shape=(sc.value,)).copy() if sc.value else np.zeros(0, np.int32)
            self.neighbors.append(dict(rank=r.value, send_local=send, recv_first=rf.value, recv_count=rcnt.value))

    def local_flat(self) -> dict:
        """The local mesh as a flat-mesh dict (copies), e.g. to run the oracle on it."""
        m = lib().mstgpu_partition_mesh(self.h).contents
        D, nc, nf = m.dim, m.ncells, m.nfaces

        def arr(ptr, dt, shape):
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).reshape(shape).copy()

        cf_ptr = arr(m.cf_ptr, np.int32, (nc + 1,))
        return dict(dim=D, ncells=nc, nfaces=nf, nint=m.nint, c0=arr(m.c0, np.int32, (nf,)), c1=arr(m.c1, np.int32, (nf,)),
                    S=arr(m.S, np.float64, (nf, D)), dac=arr(m.dac, np.int8, (nf,)), fc=arr(m.fc, np.float64, (nf, D)),
                    eta=arr(m.eta, np.float64, (nf,)), flag=arr(m.flag, np.uint8, (nf, D)),
                    ftype=arr(m.ftype, np.int32, (nf,)), cc=arr(m.cc, np.float64, (nc, D)),
                    vol=arr(m.vol, np.float64, (nc,)), cf_ptr=cf_ptr, cf_idx=arr(m.cf_idx, np.int32, (int(cf_ptr[-1]),)))

    def close(self):
        if getattr(self, "h", None):
            lib().mstgpu_partition_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().mstgpu_comm_unique_id(buf)
    if rc != 0:
        raise MstGpuError(f"comm_unique_id failed ({rc}): {lib().mstgpu_last_error(None).decode()}")
    return buf.raw


class Context:
    """Owns one mstgpu_ctx.  `flat` is the flattened reference-order mesh (the
    dict produced by the host's mesh flattener)."""

    def __init__(self, flat, order=2, flux="roe", viscous=0, qf_copy_from=None,
                 renumber=2, device=-1, inletQ=None, kernel=None, tile_cells=0, block_threads=0,
                 **consts):
        L = lib()
        if isinstance(flat, Partition):
            # one rank of a multi-GPU run: context on the partition's local mesh
            part = flat
            self.dim, self.U = part.dim, part.dim + 2
            self.ncells = part.n_owned
            self.nfaces = int(L.mstgpu_partition_mesh(part.h).contents.nfaces)
            cfg = make_config(self.dim, order=order, flux=flux, viscous=viscous, qf_copy_from=qf_copy_from,
                              renumber=renumber, device=device, inletQ=inletQ, kernel=kernel,
                              tile_cells=tile_cells, block_threads=block_threads, **consts)
            self.cfg = cfg
            h = C.c_void_p()
            rc = L.mstgpu_create_partitioned(C.byref(h), part.h, C.byref(cfg))
            if rc != 0:
                raise MstGpuError(f"mstgpu_create_partitioned failed ({rc}): {L.mstgpu_last_error(None).decode()}")
            self.h = h
            return
        self.dim = int(flat["dim"])
        self.U = self.dim + 2
        self.ncells = int(flat["ncells"])
        self.nfaces = int(flat["nfaces"])
        m, keep = _mesh_struct(flat)
        cfg = default_config(self.dim)
        cfg.order = order
        cfg.flux = {"roe": 0, "ausm": 1}[flux] if isinstance(flux, str) else int(flux)
        cfg.viscous = viscous
        cfg.qf_copy_from = -1 if qf_copy_from is None else qf_copy_from
        cfg.renumber = renumber
        cfg.device = device
        if kernel is not None:
            cfg.kernel = {"tiles": 1, "split": 0}[kernel] if isinstance(kernel, str) else int(kernel)
        cfg.tile_cells = tile_cells
        cfg.block_threads = block_threads
        _apply_consts(cfg, consts)
        if inletQ is not None:
            for k in range(5):
                cfg.inletQ[k] = float(inletQ[k]) if k < len(inletQ) else 0.0
        self.cfg = cfg
        h = C.c_void_p()
        rc = L.mstgpu_create(C.byref(h), C.byref(m), C.byref(cfg))
        if rc != 0:
            raise MstGpuError(f"mstgpu_create failed ({rc}): {L.mstgpu_last_error(None).decode()}")
        self.h = h

    def _check(self, rc, what):
        if rc != 0:
            raise MstGpuError(f"{what} failed ({rc}): {lib().mstgpu_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            lib().mstgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        self._check(lib().mstgpu_comm_init(self.h, nranks, rank, unique_id), "comm_init")

    def peer_export(self, rank: int) -> bytes:
        """this rank's blob for the peer-memory halo (CUDA IPC handles + row offsets)"""
        n = int(lib().mstgpu_peer_blob_bytes())
        buf = C.create_string_buffer(n)
        self._check(lib().mstgpu_peer_export(self.h, rank, buf), "peer_export")
        return buf.raw

    def peer_connect(self, nranks: int, rank: int, blobs: bytes):
        """blobs = the nranks blobs of peer_export, concatenated by rank (collective)"""
        assert len(blobs) == nranks * int(lib().mstgpu_peer_blob_bytes())
        self._check(lib().mstgpu_peer_connect(self.h, nranks, rank, blobs), "peer_connect")

    def peer_connect_torch(self, dist, nranks: int, rank: int, device=None) -> bool:
        """all-gather the blobs with torch.distributed and connect; False (and the NCCL exchange stays) if any
        rank cannot open a neighbour's memory.  The decision is collective: all ranks switch or none does."""
        import torch
        mine = torch.frombuffer(bytearray(self.peer_export(rank)), dtype=torch.uint8)
        if device is not None:
            mine = mine.to(device)
        parts = [torch.empty_like(mine) for _ in range(nranks)]
        dist.all_gather(parts, mine)
        blobs = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        ok = torch.ones(1, dtype=torch.int32, device=mine.device)
        try:
            self.peer_connect(nranks, rank, blobs)
        except MstGpuError:
            ok[0] = 0
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self.peer_disable()
            return False
        return True

    def peer_disable(self):
        self._check(lib().mstgpu_peer_disable(self.h), "peer_disable")

    def set_state(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        assert Q.shape == (self.ncells, self.U)
        self._check(lib().mstgpu_set_state(self.h, Q.ctypes.data, self.ncells), "set_state")

    def set_state_ptr(self, ptr: int):
        self._check(lib().mstgpu_set_state(self.h, ptr, self.ncells), "set_state")

    def get_state(self, out=None):
        out = np.empty((self.ncells, self.U)) if out is None else out
        self._check(lib().mstgpu_get_state(self.h, out.ctypes.data), "get_state")
        return out

    def get_state_ptr(self, ptr: int):
        self._check(lib().mstgpu_get_state(self.h, ptr), "get_state")

    def set_tile_variant(self, variant: int):
        """experimental launch variants of the default fused kernel (include/mstgpu.h); 0 = default"""
        self._check(lib().mstgpu_set_tile_variant(self.h, int(variant)), "set_tile_variant")

    def output_setup(self, flat, nf_ptr, nf_idx, node_weight=None):
        """Output path (Work.cpp:243-304 on the device): `flat` is the mesh the context was created
        from, (nf_ptr, nf_idx) the node -> faces lists in Node::addNbFace order (host.node_faces),
        node_weight the reference's 1 / area(face[node id]) (None = 1)."""
        m, keep = _mesh_struct(flat)
        nf_ptr = np.ascontiguousarray(nf_ptr, dtype=np.int32); nf_idx = np.ascontiguousarray(nf_idx, dtype=np.int32)
        w = None if node_weight is None else np.ascontiguousarray(node_weight, dtype=np.float64)
        self.nnodes = nf_ptr.size - 1
        self._check(lib().mstgpu_output_setup(self.h, C.byref(m), self.nnodes, nf_ptr.ctypes.data, nf_idx.ctypes.data,
                                              None if w is None else w.ctypes.data), "output_setup")

    def output_setup_partitioned(self, part, flat_global, nf_ptr, nf_idx, node_weight=None):
        """Output path on a partitioned context: `part` is the Partition the context was created from, the other
        arguments are GLOBAL (mesh, node -> faces lists, per-node weight).  Afterwards node_ids holds the global ids
        of the nodes this rank computes and node_fields() (collective) returns their rows."""
        m, keep = _mesh_struct(flat_global)
        nf_ptr = np.ascontiguousarray(nf_ptr, dtype=np.int32); nf_idx = np.ascontiguousarray(nf_idx, dtype=np.int32)
        w = None if node_weight is None else np.ascontiguousarray(node_weight, dtype=np.float64)
        self._check(lib().mstgpu_output_setup_partitioned(self.h, part.h, C.byref(m), nf_ptr.size - 1, nf_ptr.ctypes.data, nf_idx.ctypes.data,
                                                          None if w is None else w.ctypes.data), "output_setup_partitioned")
        self.nnodes = int(lib().mstgpu_output_node_count(self.h))
        self.node_ids = np.empty(self.nnodes, dtype=np.int32)
        self._check(lib().mstgpu_output_node_ids(self.h, self.node_ids.ctypes.data), "output_node_ids")

    def node_fields(self, out=None):
        """[nnodes, dim+4] = rho, u_i, T, p, Ma per node of the current state (bit-identical to the
        numbers the reference's writer prints)."""
        out = np.empty((self.nnodes, self.dim + 4)) if out is None else out
        self._check(lib().mstgpu_node_fields(self.h, out.ctypes.data), "node_fields")
        return out

    def get_prev_state(self):
        out = np.empty((self.ncells, self.U))
        self._check(lib().mstgpu_get_prev_state(self.h, out.ctypes.data), "get_prev_state")
        return out

    def step(self, dt: float, nsteps: int = 1):
        self._check(lib().mstgpu_step(self.h, dt, nsteps), "step")

    def step_timed(self, dt: float, nsteps: int = 1) -> float:
        ms = C.c_float()
        self._check(lib().mstgpu_step_timed(self.h, dt, nsteps, C.byref(ms)), "step_timed")
        return float(ms.value)

    def step_host(self, q_in, q_out, dt: float, nchunks: int = 0):
        """One step with HOST arrays on both sides (set_state + step + get_state, pipelined over chunks of host
        rows; include/mstgpu.h).  q_in / q_out: numpy arrays [ncells][U] or raw pointers (page-locked for overlap)."""
        pi = q_in if isinstance(q_in, int) else q_in.ctypes.data
        po = q_out if isinstance(q_out, int) else q_out.ctypes.data
        self._check(lib().mstgpu_step_host(self.h, pi, po, dt, nchunks), "step_host")

    def cfl_dt(self, cfl: float) -> float:
        """Global CFL time step of the current state (extension; collective with a communicator)."""
        dt = C.c_double()
        self._check(lib().mstgpu_cfl_dt(self.h, cfl, C.byref(dt)), "cfl_dt")
        return float(dt.value)

    def step_cfl(self, cfl: float, nsteps: int = 1) -> float:
        """nsteps steps, each at its own CFL step (dt stays on the device); returns the time advanced."""
        t = C.c_double()
        self._check(lib().mstgpu_step_cfl(self.h, cfl, nsteps, C.byref(t)), "step_cfl")
        return float(t.value)

    def step_cfl_timed(self, cfl: float, nsteps: int = 1) -> float:
        """step_cfl bracketed by CUDA events on the solver's stream; returns milliseconds."""
        t, ms = C.c_double(), C.c_float()
        self._check(lib().mstgpu_step_cfl_timed(self.h, cfl, nsteps, C.byref(t), C.byref(ms)), "step_cfl_timed")
        return float(ms.value)

    def implicit_setup(self, colour_sweeps: bool = True):
        self._check(lib().mstgpu_implicit_setup(self.h, 1 if colour_sweeps else 0), "implicit_setup")

    def implicit_sweep_order(self) -> np.ndarray:
        out = np.empty(self.ncells, dtype=np.int32)
        self._check(lib().mstgpu_implicit_sweep_order(self.h, out.ctypes.data), "implicit_sweep_order")
        return out

    def step_implicit(self, dt: float, nsteps: int = 1, lusgs_iters: int = 5) -> float:
        """Implicit steps (LU-SGS sweeps of the reference's lusolver); returns the CUDA-event time in ms."""
        ms = C.c_float()
        self._check(lib().mstgpu_step_implicit(self.h, dt, nsteps, lusgs_iters, C.byref(ms)), "step_implicit")
        return float(ms.value)

    def residual(self):
        out = np.empty(self.U)
        self._check(lib().mstgpu_residual_linf(self.h, out.ctypes.data), "residual_linf")
        return out

    def sync(self):
        self._check(lib().mstgpu_sync(self.h), "sync")

    def debug_gradient(self):
        out = np.empty((self.ncells, self.U, self.dim))
        self._check(lib().mstgpu_debug_gradient(self.h, out.ctypes.data), "debug_gradient")
        return out

    def debug_face_flux(self):
        out = np.empty((self.nfaces, self.U))
        self._check(lib().mstgpu_debug_face_flux(self.h, out.ctypes.data), "debug_face_flux")
        return out

    @property
    def launch_count(self) -> int:
        return int(lib().mstgpu_launch_count(self.h))

    @property
    def device_bytes(self) -> int:
        return int(lib().mstgpu_device_bytes(self.h))

    def enable_kernel_timing(self, on=True):
        self._check(lib().mstgpu_enable_kernel_timing(self.h, 1 if on else 0), "enable_kernel_timing")

    def kernel_time(self, name: str):
        ms, n = C.c_double(), C.c_int64()
        self._check(lib().mstgpu_kernel_time(self.h, name.encode(), C.byref(ms), C.byref(n)), "kernel_time")
        return float(ms.value), int(n.value)


class GpuRhoSolver:
    """Python mirror of the reference's solver duck type
    (R/rhoSolver/RhoSolver.h:17-24; PSolver has the same shape).  The reference
    constructs the solver every step (R/time/Time.cpp:58); here the heavy
    context is built once and handed in, the object itself is cheap."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.DT = None

    def setDT(self, dt: float):  # RhoSolver.cpp:33-35
        self.DT = float(dt)

    def solve(self):  # RhoSolver.cpp:37-89 (+ pointer swap; see updateNewToOld)
        if self.DT is None:
            raise MstGpuError("solve() before setDT()")
        self.ctx.step(self.DT, 1)

    def getNewValue(self):  # RhoSolver.cpp:507-509
        return self.ctx.get_state()

    def getOldValue(self):  # RhoSolver.cpp:504-506 (as seen before updateNewToOld)
        return self.ctx.get_prev_state()

    def getOldNTimeValue(self):  # RhoSolver.cpp:510-512 (pseudo-time off: same as old)
        return self.ctx.get_prev_state()

    def updateNewToOld(self):  # RhoSolver.cpp:513-517: already a pointer swap on the device
        return None


def lusgs_color_order(rowptr, col):
    """(perm_new2old, ncolors): greedy colour ordering of a CSR pattern (host only)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    n = rowptr.shape[0] - 1
    perm = np.empty(n, dtype=np.int32)
    nc = C.c_int32()
    rc = lib().mstgpu_lusgs_color_order(n, rowptr.ctypes.data, col.ctypes.data, perm.ctypes.data, C.byref(nc))
    if rc != 0:
        raise MstGpuError(f"lusgs_color_order failed ({rc}): {lib().mstgpu_lusgs_last_error().decode()}")
    return perm, nc.value


class LuSgs:
    """GPU LU-SGS solver for one sparsity pattern (mirror of the reference's
    SparseSolverNUM for block = 1 and SparseSolver<MT,VCT> for block = DIMU)."""

    def __init__(self, rowptr, col, block=1, device=-1, sweep_order=None):
        """sweep_order: new2old permutation of the rows (e.g. lusgs_color_order); the data stay in
        storage order, the sweeps are those of the reference on the permuted system."""
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.n, self.block = self.rowptr.shape[0] - 1, block
        h = C.c_void_p()
        so = None if sweep_order is None else np.ascontiguousarray(sweep_order, dtype=np.int32)
        rc = lib().mstgpu_lusgs_create_ordered(C.byref(h), self.n, block, self.rowptr.ctypes.data, self.col.ctypes.data,
                                               None if so is None else so.ctypes.data, device)
        if rc != 0:
            raise MstGpuError(f"lusgs_create failed ({rc}): {lib().mstgpu_lusgs_last_error().decode()}")
        self.h = h

    def set_mode(self, mode: int):
        """0 = the reference's passes one by one (default), 1 = fused iteration (experimental, include/mstgpu.h)"""
        rc = lib().mstgpu_lusgs_set_mode(self.h, int(mode))
        if rc != 0:
            raise MstGpuError(f"lusgs_set_mode failed ({rc}): {lib().mstgpu_lusgs_last_error().decode()}")

    def levels(self):
        a, b = C.c_int32(), C.c_int32()
        lib().mstgpu_lusgs_levels(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def solve(self, val, b, x0, max_iter=5, early_exit=False):
        B = self.block
        val = np.ascontiguousarray(val, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, order="C", copy=True)
        assert val.size == self.col.shape[0] * B * B and b.size == self.n * B and x.size == self.n * B
        hist = np.zeros(max_iter)
        it = C.c_int32()
        rc = lib().mstgpu_lusgs_solve(self.h, val.ctypes.data, b.ctypes.data, x.ctypes.data, max_iter,
                                      1 if early_exit else 0, hist.ctypes.data, C.byref(it))
        if rc != 0:
            raise MstGpuError(f"lusgs_solve failed ({rc}): {lib().mstgpu_lusgs_last_error().decode()}")
        return x, hist[:it.value], it.value

    def solve_device(self, d_val: int, d_b: int, d_x: int, max_iter=5) -> float:
        """Device pointers in, x updated in place on the device; returns the CUDA-event time in ms."""
        ms = C.c_float()
        rc = lib().mstgpu_lusgs_solve_device(self.h, d_val, d_b, d_x, max_iter, C.byref(ms))
        if rc != 0:
            raise MstGpuError(f"lusgs_solve_device failed ({rc}): {lib().mstgpu_lusgs_last_error().decode()}")
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(lib().mstgpu_lusgs_launch_count(self.h))

    @property
    def device_bytes(self) -> int:
        return int(lib().mstgpu_lusgs_device_bytes(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().mstgpu_lusgs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
