"""CPU suite, part 6: the build-defined extension named by the north star and
absent from the reference: least-squares gradient,
Barth-Jespersen / Venkatakrishnan limiter, CFL time step.

There is no reference code for any of this ("parity unpinned"); the oracle's
restatement of the formulas (oracle/rho_oracle.cpp: lsqGradient, limitGradient,
cflDt) is checked here against first principles, so that it can in turn check
the CUDA path (tests/test_extension_gpu.py)."""
import numpy as np
import pytest

from conftest import load_flat, box_flat
from oracle import mesh_np, oracle
import mstgpu


def _interior_cells(f):
    b = np.zeros(f["ncells"], bool)
    b[f["c0"][f["c1"] < 0]] = True
    return ~b


def test_defaults_are_the_reference_scheme():
    f = load_flat("2d-stair-un-5-tri")
    Q0 = mesh_np.random_state(f, seed=3)
    a = oracle.Oracle(f, order=2, flux="roe").solve(1e-4, Q0)
    b = oracle.Oracle(f, order=2, flux="roe", gradient="gg", limiter="none").solve(1e-4, Q0)
    assert np.array_equal(a, b, equal_nan=True)
    cfg = mstgpu.default_config(2)
    assert cfg.gradient == 0 and cfg.limiter == 0


@pytest.mark.parametrize("dim", [2, 3])
def test_lsq_gradient_is_exact_for_linear_fields(dim):
    f = load_flat("2d-stair-un-4-tri") if dim == 2 else box_flat(6, 5, 4)
    D, U = dim, dim + 2
    cc = f["cc"].reshape(-1, D)
    rng = np.random.default_rng(5)
    slope = rng.uniform(-1, 1, size=(U, D))
    Q = 3.0 + cc @ slope.T
    Q[:, 0] += 5.0
    Q[:, -1] += 50.0
    o = oracle.Oracle(f, order=2, gradient="lsq")
    o.solve(1e-5, Q)
    _, G, _ = o.probe()
    inner = _interior_cells(f)
    assert inner.sum() > 20
    assert np.abs(G[inner] - slope[None]).max() < 1e-10
    # Green-Gauss with eta interpolation is not exact on a skewed mesh: the two gradients differ
    o2 = oracle.Oracle(f, order=2, gradient="gg")
    o2.solve(1e-5, Q)
    _, G2, _ = o2.probe()
    assert np.isfinite(G2).all()


@pytest.mark.parametrize("limiter", ["bj", "venkat"])
@pytest.mark.parametrize("gradient", ["gg", "lsq"])
def test_limited_reconstruction_is_bounded_by_the_neighbours(limiter, gradient):
    f = load_flat("2d-stairW-1")
    D, U = 2, 4
    Q0 = mesh_np.random_state(f, seed=9)
    o = oracle.Oracle(f, order=2, gradient=gradient, limiter=limiter, limiter_k=0.0)
    o.solve(1e-5, Q0)
    _, G, _ = o.probe()
    ou = oracle.Oracle(f, order=2, gradient=gradient)
    ou.solve(1e-5, Q0)
    _, Gu, _ = ou.probe()
    # phi in [0, 1], per variable: G = phi * Gu
    with np.errstate(all="ignore"):
        ratio = np.where(np.abs(Gu) > 1e-12, G / Gu, np.nan)
    assert np.nanmin(ratio) >= -1e-12 and np.nanmax(ratio) <= 1 + 1e-12
    # reconstructed values at all face centres of a cell stay within the min / max of the cell and its neighbours
    qmin, qmax = Q0.copy(), Q0.copy()
    c0, c1 = f["c0"], f["c1"]
    it = c1 >= 0
    np.minimum.at(qmin, c0[it], Q0[c1[it]]); np.minimum.at(qmin, c1[it], Q0[c0[it]])
    np.maximum.at(qmax, c0[it], Q0[c1[it]]); np.maximum.at(qmax, c1[it], Q0[c0[it]])
    fc, cc = f["fc"].reshape(-1, D), f["cc"].reshape(-1, D)
    worst = 0.0
    for cells, faces in ((c0, np.arange(f["nfaces"])), (c1[it], np.nonzero(it)[0])):
        r = fc[faces] - cc[cells]
        rec = Q0[cells] + np.einsum("fkd,fd->fk", G[cells], r)
        worst = max(worst, (rec - qmax[cells]).max(), (qmin[cells] - rec).max())
    assert worst < 1e-12  # Barth-Jespersen exactly; Venkatakrishnan with K = 0 as well


def test_limiter_removes_the_overshoots_of_the_unlimited_scheme_on_sod():
    f = load_flat("2d-shockwavepipe-2")
    Q0 = mesh_np.sod_initial_state(f)
    un = oracle.Oracle(f, order=2, flux="roe").run(2.5e-4, 400, Q0)
    assert un[:, 0].max() > 1.02 and un[:, 0].min() < 0.11  # the reference scheme over/undershoots
    for lim in ("bj", "venkat"):
        Q = oracle.Oracle(f, order=2, flux="roe", limiter=lim, limiter_k=1.0).run(2.5e-4, 400, Q0)
        assert Q[:, 0].max() < 1.002 and Q[:, 0].min() > 0.1245, lim


def test_cfl_time_step_formula():
    f = box_flat(5, 4, 3)
    Q = mesh_np.random_state(f, seed=2)
    o = oracle.Oracle(f, order=2)
    D = 3
    S = f["S"].reshape(-1, D)
    rho = Q[:, 0]
    u = Q[:, 1:4] / rho[:, None]
    p = (Q[:, 4] - 0.5 * (Q[:, 1:4] ** 2).sum(1) / rho) * 0.4
    a = np.sqrt(1.4 * p / rho)
    lam = np.zeros(f["ncells"])
    Sn = np.linalg.norm(S, axis=1)
    for c in range(f["ncells"]):
        for j in range(f["cf_ptr"][c], f["cf_ptr"][c + 1]):
            fa = f["cf_idx"][j]
            lam[c] += abs(u[c] @ S[fa]) + a[c] * Sn[fa]
    want = 0.7 * (f["vol"] / lam).min()
    assert abs(o.cfl_dt(0.7, Q) - want) <= 1e-14 * want
    # a NaN cell never wins the minimum
    Q2 = Q.copy(); Q2[3, 4] = -1.0
    assert np.isfinite(o.cfl_dt(0.7, Q2))
    Qe, dts = o.run_cfl(0.5, 3, Q)
    assert dts.shape == (3,) and (dts > 0).all()
    assert np.array_equal(Qe, o.solve(dts[2], o.solve(dts[1], o.solve(dts[0], Q))), equal_nan=True)


@pytest.mark.parametrize("nparts", [2, 3])
@pytest.mark.parametrize("limiter,gradient", [("bj", "gg"), ("venkat", "lsq")])
def test_partitioned_limited_scheme_is_bit_identical(nparts, limiter, gradient):
    """Two ghost layers are enough for the limiter too: layer 1 needs its gradient AND its
    limiter value, both built from layer-2 states.  Global CFL step = min over the partitions."""
    f = box_flat(7, 6, 5, bc=(10, 5, 3, 7, 3, 3))
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    kw = dict(order=2, flux="roe", inletQ=inlet, limiter=limiter, gradient=gradient, limiter_k=2.0)
    Q0 = mesh_np.random_state(f, seed=4)
    full = oracle.Oracle(f, **kw)
    ref, dts = full.run_cfl(0.4, 3, Q0)
    parts = [mstgpu.Partition(f, nparts, r, order=2) for r in range(nparts)]
    locs = [P.local_flat() for P in parts]
    ors = [oracle.Oracle(lf, qf_copy_from=lf["nint"], **kw) for lf in locs]
    Qs = [np.zeros((P.n_local, 5)) for P in parts]
    for P, Q in zip(parts, Qs):
        Q[:P.n_owned] = Q0[P.cell_ids[:P.n_owned]]
    from test_partition_cpu import _exchange_inprocess
    for step in range(3):
        _exchange_inprocess(parts, Qs)
        # CFL step over OWNED cells only, then the minimum over partitions (ncclAllReduce(min) on the GPUs)
        dt = min(oracle.Oracle({**lf, **_owned_only(lf, P.n_owned)}, **kw).cfl_dt(0.4, Qs[r][:P.n_owned])
                 for r, (lf, P) in enumerate(zip(locs, parts)))
        assert dt == dts[step]
        for r, P in enumerate(parts):
            Qn = ors[r].solve(dt, Qs[r])
            Qs[r][:P.n_owned] = Qn[:P.n_owned]
    out = np.empty_like(ref)
    for P, Q in zip(parts, Qs):
        out[P.cell_ids[:P.n_owned]] = Q[:P.n_owned]
    assert np.array_equal(out, ref, equal_nan=True)


def _owned_only(lf, n_owned):
    """cell tables of a local mesh cut down to its owned cells (the CFL step looks at cells, their
    faces' area vectors and volumes only)"""
    return dict(ncells=n_owned, cf_ptr=lf["cf_ptr"][:n_owned + 1].copy(), vol=lf["vol"][:n_owned].copy(),
                cc=lf["cc"].reshape(-1, lf["dim"])[:n_owned].copy())


def test_config_is_validated_without_a_gpu():
    f = box_flat(2, 2, 2)
    for bad in (dict(limiter=7), dict(gradient=3), dict(limiter=2, limiter_k=-1.0)):
        with pytest.raises(mstgpu.MstGpuError):
            mstgpu.Context(f, order=2, **bad)
    with pytest.raises(mstgpu.MstGpuError, match="unknown config field"):
        mstgpu.make_config(3, limitr=1)


# ---- implicit operator (oracle side) -----------------------------------------------------------------

def test_implicit_system_blocks_follow_the_stated_operator():
    """off-diagonal O_ij = 1/2 (A(Q_j,S) - lam I), and the Euler flux is homogeneous of degree one:
    A(Q,S) Q = F(Q).S, so (2 O_ij + lam I) Q_j must be the physical flux of Q_j through S."""
    f = box_flat(4, 3, 3)
    D, U = 3, 5
    Q = mesh_np.random_state(f, seed=6)
    o = oracle.Oracle(f, order=2, flux="roe")
    dt = 1e-3
    rowptr, col, val, b = o.implicit_system(dt, Q)
    assert rowptr[-1] == f["ncells"] + 2 * f["nint"] and val.shape == (rowptr[-1], U, U)
    rows = np.repeat(np.arange(f["ncells"]), np.diff(rowptr))
    assert np.all(np.diff(col)[np.diff(rows) == 0] > 0)
    S_all = f["S"].reshape(-1, D) * f["dac"][:, None]  # outward from c0

    def lam(q, S):
        u = q[1:4] / q[0]
        p = (q[4] - 0.5 * q[0] * (u @ u)) * 0.4
        return abs(u @ S) + np.sqrt(1.4 * p / q[0]) * np.linalg.norm(S)

    def flux(q, S):
        u = q[1:4] / q[0]
        p = (q[4] - 0.5 * q[0] * (u @ u)) * 0.4
        un = u @ S
        return np.concatenate([[q[0] * un], q[1:4] * un + p * S, [(q[4] + p) * un]])

    checked = 0
    for fa in range(0, f["nint"], 7):
        i, j = f["c0"][fa], f["c1"][fa]
        S = S_all[fa]
        l = max(lam(Q[i], S), lam(Q[j], S))
        k = rowptr[i] + np.searchsorted(col[rowptr[i]:rowptr[i + 1]], j)
        assert col[k] == j
        assert np.allclose((2 * val[k] + l * np.eye(U)) @ Q[j], flux(Q[j], S), rtol=1e-11, atol=1e-13)
        k2 = rowptr[j] + np.searchsorted(col[rowptr[j]:rowptr[j + 1]], i)
        assert np.allclose((2 * val[k2] + l * np.eye(U)) @ Q[i], flux(Q[i], -S), rtol=1e-11, atol=1e-13)
        checked += 1
    assert checked > 20
    # b = -R: the explicit step is Q - dt/V R
    Qe = o.solve(dt, Q)
    assert np.allclose(Q + dt / f["vol"][:, None] * b, Qe, rtol=1e-12, atol=1e-14)


def test_implicit_step_limits():
    f = load_flat("2d-stair-un-5-tri")
    Q0 = mesh_np.random_state(f, seed=8)
    o = oracle.Oracle(f, order=2, flux="roe", limiter="bj")
    dt = 1e-10
    Qi, Qe = o.step_implicit(dt, Q0, 5), o.solve(dt, Q0)
    assert np.abs(Qi - Qe).max() <= 1e-4 * np.abs(Qe - Q0).max()
    # a gas at rest in a closed box stays exactly at rest
    fb = box_flat(3, 3, 3)
    Qr = np.tile(np.array([1.0, 0, 0, 0, 2.5]), (fb["ncells"], 1))
    ob = oracle.Oracle(fb, order=2, flux="roe")
    assert np.abs(ob.step_implicit(1e-2, Qr, 5) - Qr).max() < 1e-13
    # a sweep order changes the 5-sweep iterate, not the converged solution
    perm = np.random.default_rng(0).permutation(f["ncells"])
    a = o.step_implicit(1e-3, Q0, 60)
    bq = o.step_implicit(1e-3, Q0, 60, sweep_order=perm)
    assert np.abs(a - bq).max() < 1e-9 * np.abs(a).max()
