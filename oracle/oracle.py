"""ORACLE / TEST INFRASTRUCTURE -- not product code.

ctypes front end of liboracle.so (rho_oracle.cpp, lusgs_oracle.cpp): the CPU
restatement of the reference hot path.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("rho_oracle.cpp", "lusgs_oracle.cpp")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


class OmMesh(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ncells", C.c_int32), ("nfaces", C.c_int32), ("nint", C.c_int32),
        ("c0", C.c_void_p), ("c1", C.c_void_p), ("S", C.c_void_p), ("dac", C.c_void_p),
        ("fc", C.c_void_p), ("eta", C.c_void_p), ("flag", C.c_void_p), ("ftype", C.c_void_p),
        ("cc", C.c_void_p), ("vol", C.c_void_p), ("cf_ptr", C.c_void_p), ("cf_idx", C.c_void_p),
    ]


class OmCfg(C.Structure):
    _fields_ = [
        ("order", C.c_int32), ("flux", C.c_int32), ("viscous", C.c_int32),
        ("qf_copy_from", C.c_int32), ("nthreads", C.c_int32), ("pad_", C.c_int32),
        ("gamma", C.c_double), ("delta", C.c_double), ("eor", C.c_double),
        ("mu", C.c_double), ("kappa", C.c_double), ("cv", C.c_double),
        ("inletQ", C.c_double * 5),
        ("gradient", C.c_int32), ("limiter", C.c_int32), ("limiter_k", C.c_double),
    ]


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OmMesh), C.POINTER(OmCfg)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_nthreads.argtypes = [C.c_void_p]
        L.oracle_solve.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_run.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_cfl_dt.restype = C.c_double
        L.oracle_cfl_dt.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.oracle_run_cfl.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_implicit_system.restype = C.c_int64
        L.oracle_implicit_system.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_riemann.argtypes = [C.c_int, C.c_int, C.POINTER(OmCfg), C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_lusgs.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32)]
        _LIB = L
    return _LIB


# reference constants, R/include/CONST.h:38-48, 70-83
GAMMA = 1.4
CV = 715.8
INI_T = 1 / 286.32


# build-defined extension (absent from the reference): gradient / limiter choices
GRADIENTS = {"gg": 0, "lsq": 1}
LIMITERS = {"none": 0, "bj": 1, "venkat": 2}


def default_inlet(dim: int) -> np.ndarray:
    """inletQ as the shipped macros build it (CONST.h:78-83): rho=1, u=0,
    E = rho*(T*CV)."""
    q = np.zeros(5)
    q[0] = 1.0
    q[dim + 1] = 1.0 * (INI_T * CV + 0.0)
    return q


def make_cfg(flat, order=2, flux="roe", viscous=0, qf_copy_from=None, nthreads=0,
             inletQ=None, gamma=GAMMA, delta=0.125, eor=1e-10,
             mu=1.7894e-05, kappa=0.0242, cv=CV, gradient="gg", limiter="none", limiter_k=5.0) -> OmCfg:
    c = OmCfg()
    c.order = order
    c.flux = {"roe": 0, "ausm": 1}[flux] if isinstance(flux, str) else int(flux)
    c.viscous = viscous
    c.qf_copy_from = flat["nint"] - 1 if qf_copy_from is None else qf_copy_from
    c.nthreads = nthreads
    c.gamma, c.delta, c.eor = gamma, delta, eor
    c.mu, c.kappa, c.cv = mu, kappa, cv
    c.gradient = GRADIENTS[gradient] if isinstance(gradient, str) else int(gradient)
    c.limiter = LIMITERS[limiter] if isinstance(limiter, str) else int(limiter)
    c.limiter_k = limiter_k
    iq = default_inlet(flat["dim"]) if inletQ is None else np.asarray(inletQ, dtype=np.float64)
    for k in range(5):
        c.inletQ[k] = float(iq[k]) if k < len(iq) else 0.0
    return c


class Oracle:
    """One reference solver instance on a flat mesh (oracle/mesh_np.flatten)."""

    def __init__(self, flat: dict, **cfg_kw):
        self.flat = flat
        self.dim = flat["dim"]
        self.U = self.dim + 2
        self._keep = {}
        m = OmMesh()
        m.dim, m.ncells, m.nfaces, m.nint = flat["dim"], flat["ncells"], flat["nfaces"], flat["nint"]
        spec = dict(c0=np.int32, c1=np.int32, S=np.float64, dac=np.int8, fc=np.float64,
                    eta=np.float64, flag=np.uint8, ftype=np.int32, cc=np.float64,
                    vol=np.float64, cf_ptr=np.int32, cf_idx=np.int32)
        for k, dt in spec.items():
            a = np.ascontiguousarray(flat[k], dtype=dt)
            self._keep[k] = a
            setattr(m, k, a.ctypes.data)
        self.mesh = m
        self.cfg = make_cfg(flat, **cfg_kw)
        self.h = lib().oracle_create(C.byref(self.mesh), C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("oracle_create failed")

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nthreads(self) -> int:
        return lib().oracle_nthreads(self.h)

    def solve(self, dt: float, Qold: np.ndarray) -> np.ndarray:
        Qold = np.ascontiguousarray(Qold, dtype=np.float64)
        Qnew = np.empty_like(Qold)
        lib().oracle_solve(self.h, dt, Qold.ctypes.data, Qnew.ctypes.data)
        return Qnew

    def run(self, dt: float, nsteps: int, Q: np.ndarray, residuals: bool = False):
        Q = np.array(Q, dtype=np.float64, order="C", copy=True)
        r = np.zeros((nsteps, self.U)) if residuals else None
        lib().oracle_run(self.h, dt, nsteps, Q.ctypes.data, r.ctypes.data if residuals else None)
        return (Q, r) if residuals else Q

    def cfl_dt(self, cfl: float, Q: np.ndarray) -> float:
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        return float(lib().oracle_cfl_dt(self.h, cfl, Q.ctypes.data))

    def run_cfl(self, cfl: float, nsteps: int, Q: np.ndarray):
        """nsteps steps at dt = CFL step of each start state; returns (Q, dts)."""
        Q = np.array(Q, dtype=np.float64, order="C", copy=True)
        dts = np.zeros(nsteps)
        lib().oracle_run_cfl(self.h, cfl, nsteps, Q.ctypes.data, dts.ctypes.data)
        return Q, dts

    def implicit_system(self, dt: float, Q: np.ndarray):
        """(rowptr, col, val [nnz,U,U], b [nc,U]) of one implicit step (extension; reference cell order)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        nc, U = self.flat["ncells"], self.U
        rowptr = np.zeros(nc + 1, dtype=np.int32)
        nnz = lib().oracle_implicit_system(self.h, dt, Q.ctypes.data, rowptr.ctypes.data, None, None, None)
        col = np.zeros(nnz, dtype=np.int32); val = np.zeros((nnz, U, U)); b = np.zeros((nc, U))
        lib().oracle_implicit_system(self.h, dt, Q.ctypes.data, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data)
        return rowptr, col, val, b

    def step_implicit(self, dt: float, Q: np.ndarray, iters: int = 5, sweep_order=None) -> np.ndarray:
        """One implicit step: assemble, `iters` LU-SGS sweeps of the REFERENCE block solver
        (lusgs_oracle.cpp == SparseSolver<MT,VCT>::solveILU) from dQ = 0, Q + dQ.  sweep_order
        (new2old, reference cell ids): the solver runs on the explicitly permuted system P A P^T."""
        import scipy.sparse as sp
        rowptr, col, val, b = self.implicit_system(dt, Q)
        nc, U = self.flat["ncells"], self.U
        x0 = np.zeros((nc, U))
        if sweep_order is None:
            x, _, _ = lusgs(rowptr, col, val, b, x0, U, iters)
            return Q + x.reshape(nc, U)
        perm = np.asarray(sweep_order)
        Ab = sp.csr_matrix((np.arange(col.size) + 1, col, rowptr), shape=(nc, nc))[perm][:, perm].tocsr()
        Ab.sort_indices()
        xp, _, _ = lusgs(Ab.indptr, Ab.indices, val[Ab.data - 1], b[perm], x0[perm], U, iters)
        x = np.empty((nc, U)); x[perm] = xp.reshape(nc, U)
        return Q + x

    def probe(self):
        """(Qf [nf,U], G [nc,U,D], F [nf,D,U]) of the last solve()."""
        f, n, U, D = self.flat["nfaces"], self.flat["ncells"], self.U, self.dim
        Qf = np.empty((f, U)); G = np.empty((n, U, D)); F = np.empty((f, D, U))
        lib().oracle_probe(self.h, Qf.ctypes.data, G.ctypes.data, F.ctypes.data)
        return Qf, G, F


def riemann(dim, flux, L, R, d, **cfg_kw) -> np.ndarray:
    flat = dict(dim=dim, nint=1)
    cfg = make_cfg(flat, flux=flux, **cfg_kw)
    L = np.ascontiguousarray(L, dtype=np.float64); R = np.ascontiguousarray(R, dtype=np.float64)
    out = np.empty(dim + 2)
    lib().oracle_riemann(dim, cfg.flux, C.byref(cfg), L.ctypes.data, R.ctypes.data, d, out.ctypes.data)
    return out


def lusgs(rowptr, col, val, b, x0, block=1, max_iter=5, early_exit=False):
    """Reference LU-SGS sweeps (lusgs_oracle.cpp).  Returns (x, residual history, iterations)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, order="C", copy=True)
    n = rowptr.shape[0] - 1
    hist = np.zeros(max_iter)
    it = C.c_int32()
    rc = lib().oracle_lusgs(n, block, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data,
                            x.ctypes.data, max_iter, 1 if early_exit else 0, hist.ctypes.data, C.byref(it))
    if rc != 0:
        raise RuntimeError(f"oracle_lusgs failed ({rc})")
    return x, hist[:it.value], it.value
