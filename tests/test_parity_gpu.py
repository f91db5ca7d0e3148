"""GPU suite: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.

Tolerances (BASELINE.json north_star): relative L-inf on the conserved
variables <= 1e-12 after one step and <= 1e-9 after 1000 steps in FP64, plus
identical shock location on SOD.  "Relative" = per variable, against the
characteristic scale defined in conftest.char_scales.
"""
import numpy as np
import pytest

from conftest import STEP_MESHES, load_flat, box_flat, rel_linf, char_scales
from oracle import mesh_np, oracle
import mstgpu

pytestmark = pytest.mark.gpu

TOL_1STEP = 1e-12
TOL_1000 = 1e-9
DT_SOD = 1.0 / 4e3


KERNELS = ["tiles", "split"]  # fused tile kernel (default) and the three-kernel path


def _run_pair(f, Q0, dt, nsteps, kernel="tiles", **kw):
    o = oracle.Oracle(f, **kw)
    Qo = o.run(dt, nsteps, Q0)
    g = mstgpu.Context(f, kernel=kernel, **kw)
    g.set_state(Q0)
    g.step(dt, nsteps)
    Qg = g.get_state()
    return o, g, Qo, Qg


@pytest.mark.parametrize("name", STEP_MESHES)
@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
def test_one_step_random_state(name, flux, order, kernel):
    """Stress input of SURVEY.md 8d: |M| > 1 both signs -> entropy fix, AUSM
    quirk branches, every boundary type of the mesh."""
    f = load_flat(name)
    Q0 = mesh_np.random_state(f)
    o, g, Qo, Qg = _run_pair(f, Q0, 1e-4, 1, kernel=kernel, flux=flux, order=order)
    assert rel_linf(Qg, Qo) <= TOL_1STEP
    # stage probes: gradient and contracted face flux
    Qf, G, F = o.probe()
    phi_ref = np.einsum("fd,fdk->fk", f["dac"][:, None] * f["S"], F)
    phi = g.debug_face_flux()
    m = np.isfinite(phi_ref)  # unlimited 2nd order from a random state can produce NaN (sqrt of p < 0): same faces
    assert np.array_equal(np.isfinite(phi), m)
    assert np.abs(phi[m] - phi_ref[m]).max() <= TOL_1STEP * np.abs(phi_ref[m]).max()
    if order == 2:
        Gg = g.debug_gradient()
        assert np.abs(Gg - G).max() <= TOL_1STEP * np.abs(G).max()
    # residual (Time.cpp:69-76)
    with np.errstate(invalid="ignore", divide="ignore"):
        x = np.abs(Qo - Q0) / Q0
    ro = np.nanmax(np.where(x > 0, x, 0), axis=0)
    if np.isfinite(Qo).all():
        rg = g.residual()
        assert np.allclose(rg, ro, rtol=1e-9, atol=0)
    else:  # failure detection: a NaN state is reported through the C ABI
        with pytest.raises(mstgpu.MstGpuError, match="NaN"):
            g.residual()


@pytest.mark.parametrize("flux", ["roe", "ausm"])
def test_one_step_as_shipped_flags(flux):
    """The kernels take the flag table verbatim: parity holds under the
    reference's as-shipped (inverted) convention too."""
    f = load_flat("2d-stairW-1", "as_shipped")
    Q0 = mesh_np.random_state(f, seed=7)
    _, _, Qo, Qg = _run_pair(f, Q0, 1e-4, 1, flux=flux, order=2)
    assert rel_linf(Qg, Qo) <= TOL_1STEP


def test_off_by_one_and_its_correction():
    f = load_flat("2d-stair-un-5-tri")
    Q0 = mesh_np.random_state(f, seed=11)
    for qf in (None, f["nint"]):
        _, _, Qo, Qg = _run_pair(f, Q0, 1e-4, 1, order=2, qf_copy_from=qf)
        assert rel_linf(Qg, Qo) <= TOL_1STEP


@pytest.mark.parametrize("order", [1, 2])
def test_sod_1000_steps(order):
    """BASELINE config 1: SOD tube, Roe, explicit, DT = 2.5e-4."""
    f = load_flat("2d-shockwavepipe-2")
    Q0 = mesh_np.sod_initial_state(f)
    o = oracle.Oracle(f, order=order, flux="roe")
    g = mstgpu.Context(f, order=order, flux="roe")
    g.set_state(Q0)
    Qo = Q0
    for n in (1, 9, 390, 600):  # checkpoints at 1, 10, 400, 1000 steps
        Qo = o.run(DT_SOD, n, Qo)
        g.step(DT_SOD, n)
        Qg = g.get_state()
        err = rel_linf(Qg, Qo)
        assert err <= (TOL_1STEP if n == 1 else TOL_1000), (n, err)
    # identical shock location: same cell carries the steepest density drop
    x = f["cc"][:, 0]
    band = np.abs(f["cc"][:, 1] - f["cc"][:, 1].mean()) < 0.02
    idx = np.nonzero(band)[0][np.argsort(x[band])]
    jo = np.argmax(-np.diff(Qo[idx, 0]) * (x[idx][1:] > 0.6))
    jg = np.argmax(-np.diff(Qg[idx, 0]) * (x[idx][1:] > 0.6))
    assert jo == jg


@pytest.mark.parametrize("flux", ["roe", "ausm"])
def test_forward_step_ausm_and_roe_50_steps(flux):
    """Inlet 10 / outlet 5 / wall 3 mesh with a Mach-3 inlet (BASELINE config 2
    physics on the reference's own stair mesh)."""
    f = load_flat("2d-stair-un-3-tri")
    u = 3.0 * np.sqrt(1.4)
    inlet = np.array([1.0, u, 0.0, 1.0 / 0.4 + 0.5 * u * u, 0.0])
    Q0 = np.tile(inlet[:4], (f["ncells"], 1))
    o = oracle.Oracle(f, order=1, flux=flux, inletQ=inlet)
    g = mstgpu.Context(f, order=1, flux=flux, inletQ=inlet)
    g.set_state(Q0)
    Qo = o.run(2e-5, 50, Q0)
    g.step(2e-5, 50)
    assert np.isfinite(Qo).all()
    assert rel_linf(g.get_state(), Qo) <= 1e-10


@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
def test_3d_tets_one_step_and_20_steps(flux, order, kernel):
    """3-D extension (parity unpinned against the reference): GPU
    vs oracle on a Kuhn-split box with all boundary types."""
    f = box_flat(6, 5, 4, bc=(10, 5, 3, 7, 3, 3), l=(1.0, 0.8, 0.6))
    Q0 = mesh_np.random_state(f, seed=3)
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    o = oracle.Oracle(f, order=order, flux=flux, inletQ=inlet)
    g = mstgpu.Context(f, order=order, flux=flux, inletQ=inlet, kernel=kernel)
    g.set_state(Q0)
    Q1 = o.run(1e-4, 1, Q0)
    g.step(1e-4, 1)
    assert rel_linf(g.get_state(), Q1) <= TOL_1STEP
    Q20 = o.run(1e-4, 19, Q1)
    g.step(1e-4, 19)
    assert rel_linf(g.get_state(), Q20) <= 1e-10


def test_oversized_tile_is_rejected():
    f = box_flat(14, 14, 14)
    with pytest.raises(mstgpu.MstGpuError, match="shared memory"):
        mstgpu.Context(f, order=2, kernel="tiles", tile_cells=2048)


@pytest.mark.parametrize("tile_cells,block_threads", [(32, 128), (64, 256), (126, 128), (256, 256), (384, 256)])
@pytest.mark.parametrize("dim", [2, 3])
def test_tile_shapes(tile_cells, block_threads, dim):
    """Any tiling gives the oracle's answer: tiles smaller / larger than the
    default, odd-sized last tile, a handful of tiles for the whole mesh."""
    if dim == 2:
        f = load_flat("2d-stair-un-5-tri")
        inlet = None
    else:
        f = box_flat(6, 5, 4, bc=(10, 5, 3, 7, 3, 3), l=(1.0, 0.8, 0.6))
        inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    Q0 = mesh_np.random_state(f, seed=2)
    for order in (1, 2):
        o = oracle.Oracle(f, order=order, flux="roe", inletQ=inlet)
        Qo = o.run(1e-4, 3, Q0)
        g = mstgpu.Context(f, order=order, flux="roe", inletQ=inlet, kernel="tiles", tile_cells=tile_cells,
                           block_threads=block_threads)
        g.set_state(Q0)
        g.step(1e-4, 3)
        assert rel_linf(g.get_state(), Qo) <= 1e-11
        ro = o.run(1e-4, 1, Qo)
        g.step(1e-4, 1)
        assert rel_linf(g.get_state(), ro) <= 1e-11


def test_fused_and_split_kernels_agree():
    f = load_flat("2d-stairW-1")
    Q0 = mesh_np.random_state(f, seed=5)
    outs = {}
    for k in KERNELS:
        g = mstgpu.Context(f, order=2, flux="roe", kernel=k)
        g.set_state(Q0)
        g.step(1e-4, 10)
        outs[k] = (g.get_state(), g.residual())
    assert rel_linf(outs["tiles"][0], outs["split"][0]) <= 1e-13
    assert np.allclose(outs["tiles"][1], outs["split"][1], rtol=1e-10)


def test_launch_variants_do_not_change_the_bits():
    """mstgpu_set_tile_variant: L2 prefetch ahead, persistent CTAs, cp.async ring rows (3-D second order) and
    the register allocations for 3 / 4 resident CTAs (2-D first order on triangles; the default there) run
    the same arithmetic in the same order: state and residual are bit-identical to the plain launch, also
    when the steps come from the CUDA graph (12 steps per call)."""
    f3 = box_flat(12, 10, 9, bc=(10, 5, 3, 3, 7, 3))
    f2 = load_flat("2d-stair-un-3-tri")
    for f, kw, plain, variants in ((f3, dict(order=2, flux="roe"), 0, (1, 2, 3, 4, 5, 7)),
                                   (f2, dict(order=1, flux="ausm"), 32, (0, 8, 16)),
                                   (f2, dict(order=1, flux="roe"), 32, (0, 8, 16))):
        Q0 = mesh_np.random_state(f, seed=8)
        g = mstgpu.Context(f, tile_cells=96, block_threads=256, **kw)
        ref = None
        for v in (plain,) + tuple(variants):
            g.set_tile_variant(v)
            g.set_state(Q0)
            g.step(1e-4, 12)
            out = (g.get_state(), g.residual())
            if ref is None:
                ref = out
                assert rel_linf(out[0], oracle.Oracle(f, **kw).run(1e-4, 12, Q0)) <= 1e-11
            assert np.array_equal(out[0], ref[0], equal_nan=True) and np.array_equal(out[1], ref[1], equal_nan=True), (kw, v)
        with pytest.raises(mstgpu.MstGpuError, match="tile variant"):
            g.set_tile_variant(6)
        g.close()


def test_renumbering_does_not_change_the_bits():
    """Per-cell arithmetic is independent of the memory order: Morton-renumbered
    and reference-ordered runs agree bit for bit; two runs are bit-identical
    (no atomics on floating-point data)."""
    f = load_flat("2d-stairW-1")
    Q0 = mesh_np.random_state(f, seed=5)
    outs = []
    for ren in (1, 0, 1):
        g = mstgpu.Context(f, order=2, flux="roe", renumber=ren)
        g.set_state(Q0)
        g.step(1e-4, 25)
        outs.append(g.get_state())
    assert np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[0], outs[1])


def test_state_roundtrip_and_solver_interface():
    """set_state/get_state are exact; GpuRhoSolver mirrors the 7-method duck
    type of R/rhoSolver/RhoSolver.h:17-24 in the order Time.cpp:58-80 uses it."""
    f = load_flat("2d-stair-un-4-tri")
    Q0 = mesh_np.random_state(f, seed=9)
    ctx = mstgpu.Context(f, order=2, flux="roe")
    ctx.set_state(Q0)
    assert np.array_equal(ctx.get_state(), Q0)
    s = mstgpu.GpuRhoSolver(ctx)
    s.setDT(1e-4)
    s.solve()
    old, new = s.getOldValue(), s.getNewValue()
    s.updateNewToOld()
    assert np.array_equal(old, Q0)
    Qo = oracle.Oracle(f, order=2, flux="roe").solve(1e-4, Q0)
    assert rel_linf(new, Qo) <= TOL_1STEP
    assert ctx.launch_count >= 3


def test_api_errors():
    f = load_flat("2d-stair-un-5-tri")
    ctx = mstgpu.Context(f)
    with pytest.raises(mstgpu.MstGpuError, match="before set_state"):
        ctx.step(1e-4, 1)
    rc = mstgpu.lib().mstgpu_set_state(ctx.h, np.zeros(8).ctypes.data, 2)
    assert rc == -1
    # a NaN state is reported, not hidden (failure detection)
    Q = mesh_np.random_state(f); Q[17, 0] = np.nan
    ctx.set_state(Q)
    ctx.step(1e-4, 1)
    with pytest.raises(mstgpu.MstGpuError, match="NaN"):
        ctx.residual()


def test_large_box_properties():
    """Size-independent properties at a size the oracle does not need to run:
    fluid at rest stays at rest, mass/energy are conserved in a closed box,
    free-stream is preserved in the interior (1.3 M tets)."""
    f = box_flat(60, 60, 60)
    n = f["ncells"]
    ctx = mstgpu.Context(f, order=2, flux="roe")
    Q = np.zeros((n, 5)); Q[:, 0] = 1.0; Q[:, 4] = 2.5
    ctx.set_state(Q); ctx.step(1e-4, 3)
    assert np.abs(ctx.get_state() - Q).max() < 1e-12
    x = f["cc"]
    Q[:, 0] += 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q[x[:, 0] > 0.5, 4] *= 0.5
    ctx.set_state(Q); ctx.step(1e-4, 10)
    Qn = ctx.get_state()
    V = f["vol"]
    for k in (0, 4):
        assert abs((V * Qn[:, k]).sum() - (V * Q[:, k]).sum()) < 1e-12 * (V * Q[:, k]).sum()
    assert np.isfinite(Qn).all()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("limiter", ["none", "venkat"])
def test_viscous_extension(dim, order, kernel, limiter):
    """Laminar viscous term (documented extension): GPU (fused tile kernel at second
    order, split kernels otherwise) vs the oracle's corrected formulation; viscous wall = negated
    momentum ghost (RhoSolver.cpp:301-305)."""
    if limiter != "none" and order == 1:
        pytest.skip("the limiter acts on the second-order reconstruction")
    if dim == 2:
        f = load_flat("2d-stairW-1"); inlet = None
    else:
        f = box_flat(6, 5, 4, bc=(10, 5, 3, 7, 3, 3), l=(1.0, 0.8, 0.6)); inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    Q0 = mesh_np.random_state(f, seed=6)
    kw = dict(order=order, flux="roe", viscous=1, mu=0.01, kappa=0.5, inletQ=inlet, limiter=limiter, limiter_k=2.0)
    o = oracle.Oracle(f, **kw)
    g = mstgpu.Context(f, kernel=kernel, **kw)
    g.set_state(Q0)
    Q1 = o.run(1e-5, 1, Q0)
    g.step(1e-5, 1)
    assert rel_linf(g.get_state(), Q1) <= TOL_1STEP
    # the term is active: the inviscid result differs
    Qi = oracle.Oracle(f, **dict(kw, viscous=0)).run(1e-5, 1, Q0)
    assert rel_linf(Q1, Qi) > 1e-8
    Q5 = o.run(1e-5, 4, Q1)
    g.step(1e-5, 4)
    assert rel_linf(g.get_state(), Q5) <= 1e-10


def test_mid_size_direct_parity_against_oracle():
    """6 M tets, one second-order Roe step, GPU (fused kernel, ~12 k tiles, every
    tile shape the Hilbert order produces) vs the oracle directly."""
    f = box_flat(100, 100, 100)
    x = f["cc"]
    Q0 = np.zeros((f["ncells"], 5))
    s = np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q0[:, 0] = 1.0 + 0.1 * s
    Q0[:, 1] = 0.3 * Q0[:, 0] * np.cos(2 * np.pi * x[:, 1])
    Q0[:, 4] = (1.0 + 0.1 * s) / 0.4 + 0.5 * Q0[:, 1] ** 2 / Q0[:, 0]
    o = oracle.Oracle(f, order=2, flux="roe")
    Qo = o.run(1e-4, 2, Q0)
    g = mstgpu.Context(f, order=2, flux="roe")
    g.set_state(Q0)
    g.step(1e-4, 2)
    assert rel_linf(g.get_state(), Qo) <= TOL_1STEP


@pytest.fixture(scope="module")
def full_size():
    """BASELINE config 4 at its full size: mesh + context built once for the two tests below"""
    f = box_flat(203, 203, 203)
    assert f["ncells"] == 50192562
    ctx = mstgpu.Context(f, order=2, flux="roe")
    yield f, ctx
    ctx.close()


def test_full_size_direct_parity_against_oracle(full_size):
    """50 192 562 tets (the size the metric is quoted on), one second-order Roe step from the bench's initial
    state plus a velocity field: GPU vs the oracle DIRECTLY, <= 1e-12 on every conserved variable against
    its characteristic scale and pointwise-relative on rho and E (about 5 s of CPU per oracle step)."""
    from conftest import rel_linf_pointwise
    f, ctx = full_size
    x = f["cc"]
    s = np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q0 = np.zeros((f["ncells"], 5))
    Q0[:, 0] = 1.0 + 0.1 * s
    Q0[:, 1] = 0.3 * Q0[:, 0] * np.cos(2 * np.pi * x[:, 1])
    Q0[:, 3] = -0.2 * Q0[:, 0] * np.sin(2 * np.pi * x[:, 0])
    Q0[:, 4] = (1.0 + 0.1 * s) / 0.4 + 0.5 * (Q0[:, 1] ** 2 + Q0[:, 3] ** 2) / Q0[:, 0]
    o = oracle.Oracle(f, order=2, flux="roe")
    Qo = o.run(1e-4, 1, Q0)
    o.close()
    ctx.set_state(Q0)
    ctx.step(1e-4, 1)
    Qg = ctx.get_state()
    assert rel_linf(Qg, Qo) <= TOL_1STEP
    assert rel_linf_pointwise(Qg, Qo) <= TOL_1STEP
    # and the step did something: the comparison is not of two copies of the input
    assert np.abs(Qo - Q0).max() > 1e-6


def test_full_size_properties(full_size):
    """BASELINE config 4 at its full size (50 192 562 tets): size-independent
    properties -- fluid at rest stays at rest, mass and energy are conserved in
    the closed box to round-off, the state stays finite, runs are bit-identical."""
    f, ctx = full_size
    n = f["ncells"]
    Q = np.zeros((n, 5)); Q[:, 0] = 1.0; Q[:, 4] = 2.5
    ctx.set_state(Q); ctx.step(1e-4, 2)
    out = ctx.get_state()
    assert np.abs(out - Q).max() < 1e-12
    x = f["cc"]
    pert = 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
    Q[:, 0] = 1.0 + pert; Q[:, 4] = (1.0 + pert) / 0.4
    V = f["vol"]
    m0, e0 = (V * Q[:, 0]).sum(), (V * Q[:, 4]).sum()
    ctx.set_state(Q); ctx.step(1e-4, 5)
    a = ctx.get_state()
    assert np.isfinite(a).all()
    assert abs((V * a[:, 0]).sum() - m0) < 1e-12 * m0 and abs((V * a[:, 4]).sum() - e0) < 1e-12 * e0
    ctx.set_state(Q); ctx.step(1e-4, 5)
    assert np.array_equal(ctx.get_state(), a)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
def test_hexahedra_seven_point_stencil(order, kernel):
    """Cells with 6 faces: NS = 7 instantiation of the fused kernel (extension)."""
    from conftest import hex_box_flat
    f = hex_box_flat(7, 6, 5, bc=(10, 5, 3, 7, 3, 3))
    assert (np.diff(f["cf_ptr"]) == 6).all()
    Q0 = mesh_np.random_state(f, seed=12)
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    o = oracle.Oracle(f, order=order, flux="roe", inletQ=inlet)
    g = mstgpu.Context(f, order=order, flux="roe", inletQ=inlet, kernel=kernel, tile_cells=64)
    g.set_state(Q0)
    Q1 = o.run(1e-4, 3, Q0)
    g.step(1e-4, 3)
    assert rel_linf(g.get_state(), Q1) <= 1e-11


@pytest.mark.parametrize("viscous,kernel", [(1, "split"), (1, "tiles"), (0, "tiles"), (0, "split")])
def test_config3_sphere_shell_roe_viscous(viscous, kernel):
    """BASELINE config 3 in small: flow over a sphere on the cubed-sphere shell (24 tets per hex), Roe,
    second order, laminar viscous term (extension), wall / inlet / outlet zones; 1 and 20 steps."""
    import mstgpu
    from mstgpu import host
    f = host.flatten_raw(host.sphere_shell_raw(6, 5))
    u = 0.5 * np.sqrt(1.4)
    q = np.array([1.0, u, 0.0, 0.0, 2.5 + 0.5 * u * u])
    rng = np.random.default_rng(7)
    Q0 = np.tile(q, (f["ncells"], 1)) * (1.0 + 0.05 * rng.standard_normal((f["ncells"], 5)))
    kw = dict(order=2, flux="roe", viscous=viscous, inletQ=q, mu=1e-2, kappa=10.0)  # viscous term large enough to matter
    o = oracle.Oracle(f, **kw)
    dt = o.cfl_dt(0.2, Q0)
    ctx = mstgpu.Context(f, kernel=kernel, **kw)
    ctx.set_state(Q0)
    ctx.step(dt, 1)
    assert rel_linf(ctx.get_state(), o.solve(dt, Q0)) <= 1e-12
    ctx.step(dt, 19)
    assert rel_linf(ctx.get_state(), o.run(dt, 20, Q0)) <= 1e-10
    if viscous:
        # the viscous term is not a no-op in this test
        oi = oracle.Oracle(f, **dict(kw, viscous=0))
        assert rel_linf(o.solve(dt, Q0), oi.solve(dt, Q0)) > 1e-6
    ctx.close()


# ---- streamed step: host rows in, host rows out (mstgpu_step_host) ----------------------------------
def _three_calls(f, Q0, dt, **kw):
    g = mstgpu.Context(f, **kw)
    g.set_state(Q0)
    g.step(dt, 1)
    Q1 = g.get_state()
    r = g.residual()
    g.close()
    return Q1, r


@pytest.mark.parametrize("case", ["sod-roe2", "stair-ausm1", "box-roe2", "box-roe1"])
@pytest.mark.parametrize("nchunks", [1, 5, 24])
def test_step_host_equals_set_step_get(case, nchunks):
    """mstgpu_step_host = mstgpu_set_state + mstgpu_step(dt, 1) + mstgpu_get_state, pipelined over chunks of host rows:
    same bits whatever the chunk count, same residual, the context left with out = current and in = previous."""
    if case == "sod-roe2":
        f = load_flat("2d-shockwavepipe-2"); Q0 = mesh_np.sod_initial_state(f); kw = dict(order=2, flux="roe"); dt = DT_SOD
    elif case == "stair-ausm1":
        f = load_flat("2d-stair-un-3-tri"); Q0 = mesh_np.random_state(f, seed=3); kw = dict(order=1, flux="ausm"); dt = 1e-4
    else:
        f = box_flat(14, 11, 9); Q0 = mesh_np.random_state(f, seed=5); kw = dict(order=2 if case == "box-roe2" else 1, flux="roe"); dt = 1e-4
    Qref, rref = _three_calls(f, Q0, dt, **kw)
    g = mstgpu.Context(f, **kw)
    out = np.full_like(Q0, -7.0)
    g.step_host(Q0, out, dt, nchunks)
    assert np.array_equal(out, Qref, equal_nan=True)
    if np.isfinite(Qref).all():
        assert np.array_equal(g.residual(), rref)
    assert np.array_equal(g.get_state(), Qref, equal_nan=True)
    assert np.array_equal(g.get_prev_state(), Q0)
    # a second streamed step from the first one's output, in place (q_in == q_out), then plain steps on top
    Qref2, _ = _three_calls(f, Qref, dt, **kw)
    g.step_host(out, out, dt, nchunks)
    assert np.array_equal(out, Qref2, equal_nan=True)
    g.step(dt, 1)
    Qref3, _ = _three_calls(f, Qref2, dt, **kw)
    assert np.array_equal(g.get_state(), Qref3, equal_nan=True)
    g.close()


def test_step_host_with_a_shuffled_host_numbering():
    """The schedule is derived from the numbering and must be valid for ANY numbering: cells renumbered at random
    on the host side (every tile waits for the last chunk, every chunk leaves after the last group)."""
    f = box_flat(10, 9, 8)
    rng = np.random.default_rng(11)
    perm = rng.permutation(f["ncells"])          # new host id -> old host id
    inv = np.empty_like(perm); inv[perm] = np.arange(perm.size)
    fp = dict(f)
    for k in ("cc", "vol"):
        fp[k] = np.ascontiguousarray(f[k][perm])
    fp["c0"] = inv[f["c0"]].astype(np.int32)
    fp["c1"] = np.where(f["c1"] >= 0, inv[np.maximum(f["c1"], 0)], -1).astype(np.int32)
    # per-cell face lists follow their cells
    ptr, idx = f["cf_ptr"], f["cf_idx"]
    cnt = np.diff(ptr)[perm]
    nptr = np.zeros(perm.size + 1, dtype=ptr.dtype); nptr[1:] = np.cumsum(cnt)
    nidx = np.concatenate([idx[ptr[c]:ptr[c + 1]] for c in perm]).astype(idx.dtype)
    fp["cf_ptr"], fp["cf_idx"] = nptr, nidx
    if "Sout" in f:
        fp["Sout"] = np.concatenate([f["Sout"][ptr[c]:ptr[c + 1]] for c in perm])
    Q0 = mesh_np.random_state(f, seed=2)
    Qref, _ = _three_calls(f, Q0, 1e-4, order=2, flux="roe")
    g = mstgpu.Context(fp, order=2, flux="roe")
    out = np.empty_like(Q0)
    g.step_host(np.ascontiguousarray(Q0[perm]), out, 1e-4, 7)
    g.close()
    assert rel_linf(out[inv], Qref) <= TOL_1STEP
