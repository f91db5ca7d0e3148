"""GPU parity of the build-defined extension (least-squares gradient, slope
limiter, CFL time step) against the oracle's restatement of the same formulas,
through the C ABI.  Tolerances are the hot path's: relative L-inf on the
conserved variables <= 1e-12 after one step, <= 1e-9 after many.

The reference has none of this: "parity unpinned" -- the
oracle side is validated from first principles in tests/test_extension_cpu.py."""
import numpy as np
import pytest

from conftest import have_gpu, load_flat, box_flat, hex_box_flat, rel_linf
from oracle import mesh_np, oracle

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]

KERNELS = ["tiles", "split"]
BOX_BC = (10, 5, 3, 7, 3, 3)
BOX_INLET = np.array([1.0, 0.4, 0.0, 0.0, 2.58])


def _case(name):
    if name == "box":
        return box_flat(9, 8, 7, bc=BOX_BC), BOX_INLET
    return load_flat(name), None


@pytest.mark.parametrize("name", ["2d-stairW-1", "2d-stair-un-5-tri", "2d-stairW-2-st", "box"])
@pytest.mark.parametrize("limiter", ["none", "bj", "venkat"])
@pytest.mark.parametrize("gradient", ["gg", "lsq"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_one_step_limited_random_state(name, limiter, gradient, kernel):
    import mstgpu
    if limiter == "none" and gradient == "gg":
        pytest.skip("the reference scheme: covered by test_parity_gpu.py")
    f, inlet = _case(name)
    Q0 = mesh_np.random_state(f, seed=11)
    kw = dict(order=2, flux="roe", inletQ=inlet, limiter=limiter, gradient=gradient, limiter_k=2.0)
    want = oracle.Oracle(f, **kw).solve(1e-4, Q0)
    ctx = mstgpu.Context(f, kernel=kernel, **kw)
    ctx.set_state(Q0)
    ctx.step(1e-4, 1)
    got = ctx.get_state()
    assert rel_linf(got, want) <= 1e-12
    if kernel == "split":
        # stage probe: the limited gradient itself
        ow = oracle.Oracle(f, **kw)
        ow.solve(1e-4, Q0)
        _, G, _ = ow.probe()
        Gg = ctx.debug_gradient()
        scale = np.abs(G).max()
        assert np.abs(Gg - G).max() <= 1e-11 * scale
    ctx.close()


@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_one_step_limited_ausm_and_hexes(flux, kernel):
    import mstgpu
    f = hex_box_flat(6, 5, 4, bc=(3, 3, 3, 3, 7, 7))
    Q0 = mesh_np.random_state(f, seed=12)
    kw = dict(order=2, flux=flux, limiter="venkat", gradient="lsq", limiter_k=1.0)
    want = oracle.Oracle(f, **kw).solve(1e-4, Q0)
    ctx = mstgpu.Context(f, kernel=kernel, **kw)
    ctx.set_state(Q0)
    ctx.step(1e-4, 1)
    assert rel_linf(ctx.get_state(), want) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("limiter,tol", [("venkat", 1e-9), ("bj", 1e-9)])
def test_sod_limited_400_steps(limiter, tol):
    """Second-order *limited* Roe on the reference's SOD tube: no over/undershoot, shock in the exact
    cell, GPU == oracle.  (Barth-Jespersen is not differentiable; both sides evaluate the same
    expressions, so the difference stays at round-off here as well.)"""
    import mstgpu
    f = load_flat("2d-shockwavepipe-2")
    Q0 = mesh_np.sod_initial_state(f)
    kw = dict(order=2, flux="roe", limiter=limiter, limiter_k=1.0)
    want = oracle.Oracle(f, **kw).run(2.5e-4, 400, Q0)
    ctx = mstgpu.Context(f, **kw)
    ctx.set_state(Q0)
    ctx.step(2.5e-4, 400)
    got = ctx.get_state()
    assert rel_linf(got, want) <= tol
    assert got[:, 0].max() < 1.002 and got[:, 0].min() > 0.1245
    # shock location: the cell column where rho crosses the mid value between the two plateaus
    x = f["cc"].reshape(-1, 2)[:, 0]
    mid = 0.5 * (0.2656 + 0.125)
    band = np.abs(f["cc"].reshape(-1, 2)[:, 1] - 0.1) < 0.02  # the tube is 1 x 0.2
    xs_g = x[band][np.argmin(np.abs(got[band, 0] - mid))]
    xs_o = x[band][np.argmin(np.abs(want[band, 0] - mid))]
    assert xs_g == xs_o
    assert abs(xs_g - 0.675) < 0.02  # exact Sod shock position at t = 0.1
    ctx.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["2d-stair-un-4-tri", "box"])
def test_cfl_time_step_and_cfl_stepping(kernel, name):
    import mstgpu
    f, inlet = _case(name)
    Q0 = mesh_np.random_state(f, seed=13)
    kw = dict(order=2, flux="roe", inletQ=inlet, limiter="venkat", limiter_k=2.0)
    o = oracle.Oracle(f, **kw)
    ctx = mstgpu.Context(f, kernel=kernel, **kw)
    ctx.set_state(Q0)
    dt_g, dt_o = ctx.cfl_dt(0.5), o.cfl_dt(0.5, Q0)
    assert abs(dt_g - dt_o) <= 1e-14 * dt_o
    want, dts = o.run_cfl(0.3, 10, Q0)
    t = ctx.step_cfl(0.3, 10)
    assert abs(t - dts.sum()) <= 1e-13 * dts.sum()
    assert rel_linf(ctx.get_state(), want) <= 1e-11
    assert ctx.residual().shape == (f["dim"] + 2,)
    with pytest.raises(mstgpu.MstGpuError):
        ctx.cfl_dt(0.0)
    ctx.close()


def test_limited_shock_box_at_cfl_steps():
    """BASELINE config 4's input as SURVEY 8d proposes it (SOD split + smooth perturbation on Kuhn
    tets), advanced with the limited scheme at CFL steps: bounded, conservative, GPU == oracle."""
    import mstgpu
    f = box_flat(24, 24, 24)
    cc = f["cc"].reshape(-1, 3)
    s = 1e-2 * np.sin(2 * np.pi * cc[:, 0]) * np.sin(2 * np.pi * cc[:, 1]) * np.sin(2 * np.pi * cc[:, 2])
    rho = np.where(cc[:, 0] > 0.5, 0.125, 1.0) + s
    p = np.where(cc[:, 0] > 0.5, 0.1, 1.0) + s
    Q0 = np.zeros((f["ncells"], 5)); Q0[:, 0] = rho; Q0[:, 4] = p / 0.4
    ctx = mstgpu.Context(f, order=2, flux="roe", limiter="bj", gradient="lsq")
    ctx.set_state(Q0)
    t = ctx.step_cfl(0.4, 60)
    Q = ctx.get_state()
    assert np.isfinite(Q).all() and t > 0
    assert Q[:, 0].min() > 0.11 and Q[:, 0].max() < 1.02
    vol = f["vol"]
    assert abs((vol * Q[:, 0]).sum() - (vol * Q0[:, 0]).sum()) < 1e-12  # all walls: mass is conserved
    want, _ = oracle.Oracle(f, order=2, flux="roe", limiter="bj", gradient="lsq").run_cfl(0.4, 60, Q0)
    assert rel_linf(Q, want) <= 1e-9
    ctx.close()


# ---- implicit step: the reference's LU-SGS sweeps applied to rhoSolver (build-defined operator) ----------

@pytest.mark.parametrize("name", ["2d-stairW-1", "box"])
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("colour", [True, False])
def test_implicit_step_matches_the_reference_solver_on_the_oracle_system(name, kernel, colour):
    """GPU: residual (fused or split kernels) + block assembly + level-scheduled sweeps in device order.
    Oracle: its own assembly + the restatement of SparseSolver<MT,VCT>::solveILU (pinned bit-exactly
    to the reference build) on the system permuted into the GPU's sweep order."""
    import mstgpu
    f, inlet = _case(name)
    Q0 = mesh_np.random_state(f, seed=21)
    kw = dict(order=2, flux="roe", inletQ=inlet, limiter="venkat", limiter_k=2.0)
    o = oracle.Oracle(f, **kw)
    dt = 20 * o.cfl_dt(1.0, Q0)
    ctx = mstgpu.Context(f, kernel=kernel, **kw)
    ctx.set_state(Q0)
    ctx.implicit_setup(colour)
    order = ctx.implicit_sweep_order()
    assert np.array_equal(np.sort(order), np.arange(f["ncells"]))
    ms = ctx.step_implicit(dt, 1, 5)
    assert ms > 0
    got = ctx.get_state()
    want = o.step_implicit(dt, Q0, 5, sweep_order=order)
    assert rel_linf(got, want) <= 1e-11
    assert np.array_equal(ctx.get_prev_state(), Q0)
    r = ctx.residual()
    with np.errstate(all="ignore"):
        x = np.abs(got - Q0) / Q0
    assert np.allclose(r, np.nanmax(np.where(x > 0, x, 0), axis=0), rtol=1e-12)
    ctx.close()


def test_implicit_step_reduces_to_the_explicit_step_for_small_dt_and_survives_large_ones():
    import mstgpu
    f = load_flat("2d-shockwavepipe-2")
    Q0 = mesh_np.sod_initial_state(f)
    kw = dict(order=2, flux="roe", limiter="venkat", limiter_k=1.0)
    ci = mstgpu.Context(f, **kw); ce = mstgpu.Context(f, **kw)
    ci.set_state(Q0); ce.set_state(Q0)
    dt = 1e-9
    ci.step_implicit(dt, 1, 5); ce.step(dt, 1)
    Qi, Qe = ci.get_state(), ce.get_state()
    assert np.abs(Qi - Qe).max() <= 1e-4 * np.abs(Qe - Q0).max()
    # CFL 10, 40 steps: bounded (the explicit scheme at this dt blows up), GPU == oracle
    o = oracle.Oracle(f, **kw)
    dt = 10 * o.cfl_dt(1.0, Q0)
    ci.set_state(Q0)
    order = ci.implicit_sweep_order()
    ci.step_implicit(dt, 40, 5)
    Q = ci.get_state()
    assert np.isfinite(Q).all() and Q[:, 0].min() > 0.12 and Q[:, 0].max() < 1.01
    want = Q0
    for _ in range(40):
        want = o.step_implicit(dt, want, 5, sweep_order=order)
    assert rel_linf(Q, want) <= 1e-9
    ce.set_state(Q0)
    ce.step(dt, 40)
    assert not np.isfinite(ce.get_state()).all()
    ci.close(); ce.close()
