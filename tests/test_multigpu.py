"""GPU suite, multi-GPU part: two ranks, one partition per GPU, NCCL halo
exchange + residual allreduce through the C ABI -- against the single-GPU run
(bit-identical) and the oracle.  Skipped on a box with fewer than two GPUs;
the host logic is covered on CPU by test_partition_cpu.py."""
import os

import numpy as np
import pytest

from conftest import load_flat, box_flat, rel_linf
from oracle import mesh_np, oracle
import mstgpu

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


EXT = dict(limiter="bj", gradient="lsq")  # extension scheme for the CFL variant


def _worker(rank, world, port, case, kernel, q, cfl=0.0, halo="nccl", nsteps=5, host_chunks=0):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    f = box_flat(12, 10, 8, bc=(10, 5, 3, 7, 3, 3)) if case == "box" else load_flat(case)
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58]) if case == "box" else None
    Q0 = mesh_np.random_state(f, seed=4)
    P = mstgpu.Partition(f, world, rank, order=2)
    ctx = mstgpu.Context(P, order=2, flux="roe", inletQ=inlet, device=rank, kernel=kernel, **(EXT if cfl > 0 else {}))
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(mstgpu.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    ctx.comm_init(world, rank, bytes(idt.numpy().tobytes()))
    if halo == "peer":
        # boundary rows stored straight into the neighbour's ghost block (CUDA IPC + NVLink), epoch flags
        assert ctx.peer_connect_torch(dist, world, rank), "peer memory not available between the two GPUs"
    ctx.set_state(Q0[P.cell_ids[:P.n_owned]])
    if host_chunks:
        # streamed steps: the rank's rows live on the host, every step is host rows in -> host rows out
        a, b = np.ascontiguousarray(Q0[P.cell_ids[:P.n_owned]]), np.empty((P.n_owned, f["dim"] + 2))
        for _ in range(nsteps):
            ctx.step_host(a, b, 1e-4, host_chunks)
            a, b = b, a
        assert np.array_equal(ctx.get_state(), a, equal_nan=True)
    elif cfl > 0:
        t = ctx.step_cfl(cfl, nsteps)  # global time step: min over ranks on the device (ncclAllReduce(min))
        assert t > 0
    else:
        ctx.step(1e-4, nsteps)
    res = ctx.residual()  # collective
    full = torch.zeros((f["ncells"], f["dim"] + 2), dtype=torch.float64)
    full[torch.from_numpy(P.cell_ids[:P.n_owned].astype(np.int64))] = torch.from_numpy(ctx.get_state())
    dist.all_reduce(full)
    if rank == 0:
        q.put((full.numpy(), res))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["2d-stairW-1", "box"])
@pytest.mark.parametrize("kernel", ["tiles", "split"])
def test_two_gpus_match_one(case, kernel):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, case, kernel, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = box_flat(12, 10, 8, bc=(10, 5, 3, 7, 3, 3)) if case == "box" else load_flat(case)
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58]) if case == "box" else None
    Q0 = mesh_np.random_state(f, seed=4)
    one = mstgpu.Context(f, order=2, flux="roe", inletQ=inlet, kernel=kernel)
    one.set_state(Q0)
    one.step(1e-4, 5)
    assert np.array_equal(got, one.get_state(), equal_nan=True)  # same arithmetic per cell
    assert np.array_equal(res, one.residual())
    ref = oracle.Oracle(f, order=2, flux="roe", inletQ=inlet).run(1e-4, 5, Q0)
    assert rel_linf(got, ref) <= 1e-11


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kernel", ["tiles", "split"])
def test_two_gpus_limited_scheme_at_the_global_cfl_step(kernel):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29650 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, "box", kernel, q, 0.4)) for r in range(2)]
    for p in procs:
        p.start()
    got, res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = box_flat(12, 10, 8, bc=(10, 5, 3, 7, 3, 3))
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    Q0 = mesh_np.random_state(f, seed=4)
    one = mstgpu.Context(f, order=2, flux="roe", inletQ=inlet, kernel=kernel, **EXT)
    one.set_state(Q0)
    one.step_cfl(0.4, 5)
    assert np.array_equal(got, one.get_state(), equal_nan=True)  # same dt (a minimum is exact), same arithmetic per cell
    ref, _ = oracle.Oracle(f, order=2, flux="roe", inletQ=inlet, **EXT).run_cfl(0.4, 5, Q0)
    assert rel_linf(got, ref) <= 1e-11


def _implicit_worker(rank, world, port, kernel, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    f, kw, Q0, dt = _implicit_case()
    P = mstgpu.Partition(f, world, rank, order=2)
    ctx = mstgpu.Context(P, device=rank, kernel=kernel, **kw)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(mstgpu.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    ctx.comm_init(world, rank, bytes(idt.numpy().tobytes()))
    ctx.set_state(Q0[P.cell_ids[:P.n_owned]])
    ctx.implicit_setup(True)
    order = ctx.implicit_sweep_order()  # partition-local ids of the owned cells, in sweep order
    ctx.step_implicit(dt, 2, 5)
    res = ctx.residual()  # collective
    orders = [None] * world
    dist.all_gather_object(orders, order)
    full = torch.zeros((f["ncells"], f["dim"] + 2), dtype=torch.float64)
    full[torch.from_numpy(P.cell_ids[:P.n_owned].astype(np.int64))] = torch.from_numpy(ctx.get_state())
    dist.all_reduce(full)
    if rank == 0:
        q.put((full.numpy(), res, orders))
    ctx.close()
    dist.destroy_process_group()


def _implicit_case():
    f = box_flat(10, 8, 6, bc=(10, 5, 3, 7, 3, 3))
    kw = dict(order=2, flux="roe", inletQ=np.array([1.0, 0.4, 0.0, 0.0, 2.58]), limiter="venkat", limiter_k=2.0)
    rng = np.random.default_rng(4)
    Q0 = np.array([1.0, 0.4, 0.1, -0.2, 2.7]) * (1.0 + 0.1 * rng.standard_normal((f["ncells"], 5)))
    dt = 10 * oracle.Oracle(f, **kw).cfl_dt(1.0, Q0)
    return f, kw, Q0, dt


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kernel", ["tiles", "split"])
def test_two_gpus_implicit_step_with_lagged_ghosts(kernel):
    """BASELINE config 5 across GPUs: every rank sweeps its own rows (colour order), the couplings to
    the neighbour's rows lag one sweep (halo exchange of dQ between sweeps).  Checked against the same
    algorithm restated with the oracle + the reference's block solver (tests/implicit_partitioned.py)."""
    import torch.multiprocessing as mp
    import implicit_partitioned as ip
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_implicit_worker, args=(r, 2, port, kernel, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, res, orders = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f, kw, Q0, dt = _implicit_case()
    parts = [mstgpu.Partition(f, 2, r, order=2) for r in range(2)]
    locs = [P.local_flat() for P in parts]
    ors = [oracle.Oracle(lf, qf_copy_from=lf["nint"], **kw) for lf in locs]
    Qs = [np.zeros((P.n_local, 5)) for P in parts]
    for P, Q in zip(parts, Qs):
        Q[:P.n_owned] = Q0[P.cell_ids[:P.n_owned]]
    for _ in range(2):
        ip.step(ors, parts, Qs, dt, 5, orders=orders)
    want = np.empty_like(Q0)
    for P, Q in zip(parts, Qs):
        want[P.cell_ids[:P.n_owned]] = Q[:P.n_owned]
    assert rel_linf(got, want) <= 1e-10
    # close to (not equal to) the single-domain 5-sweep iterate
    one = oracle.Oracle(f, **kw)
    Qs1 = one.step_implicit(dt, one.step_implicit(dt, Q0, 5), 5)
    assert rel_linf(got, Qs1) < 0.05


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_one_host_thread_drives_contexts_on_two_devices():
    """The opt-in to > 48 KB of dynamic shared memory is a per-device attribute of a kernel: a host thread that
    creates contexts on two devices must configure it on both (it used to be cached per thread, so the second
    device's launches failed with cudaErrorInvalidValue).  Same mesh, same state, one thread: equal bits."""
    f = box_flat(20, 20, 20)
    Q0 = mesh_np.random_state(f, seed=3)
    a = mstgpu.Context(f, order=2, flux="roe", device=0)
    b = mstgpu.Context(f, order=2, flux="roe", device=1)
    a.set_state(Q0); b.set_state(Q0)
    for ctx in (a, b, a, b):  # interleaved: every entry point switches the device itself
        ctx.step(1e-4, 3)
    Qa, Qb = a.get_state(), b.get_state()
    assert np.array_equal(Qa, Qb)
    assert rel_linf(Qa, oracle.Oracle(f, order=2, flux="roe").run(1e-4, 6, Q0)) <= 1e-11
    # a smaller context created afterwards must not shrink what the first ones configured
    c = mstgpu.Context(box_flat(4, 4, 4), order=2, flux="roe", device=1)
    c.set_state(mesh_np.random_state(box_flat(4, 4, 4), seed=1)); c.step(1e-4, 1)
    b.step(1e-4, 1); a.step(1e-4, 1)
    assert np.array_equal(a.get_state(), b.get_state())
    for ctx in (a, b, c):
        ctx.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("kernel,nsteps", [("tiles", 5), ("tiles", 12), ("split", 5)])
def test_two_gpus_peer_memory_halo_matches_one(kernel, nsteps):
    """The halo through peer memory (mstgpu_peer_connect: one kernel stores the boundary rows into the
    neighbour's ghost block and publishes an epoch flag; no pack buffer, no ncclSend/ncclRecv) gives the bits of
    the single-GPU run.  12 steps go through the CUDA graph of the partitioned step (pairs of steps, epochs on
    the device), 5 are issued directly."""
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29750 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, "box", kernel, q, 0.0, "peer", nsteps)) for r in range(2)]
    for p in procs:
        p.start()
    got, res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = box_flat(12, 10, 8, bc=(10, 5, 3, 7, 3, 3))
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    Q0 = mesh_np.random_state(f, seed=4)
    one = mstgpu.Context(f, order=2, flux="roe", inletQ=inlet, kernel=kernel)
    one.set_state(Q0)
    one.step(1e-4, nsteps)
    assert np.array_equal(got, one.get_state(), equal_nan=True)
    assert np.array_equal(res, one.residual(), equal_nan=True)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("halo", ["nccl", "peer"])
def test_two_gpus_streamed_step_matches_one(halo):
    """mstgpu_step_host on a partitioned context (collective): the tiles away from the cut stream with the host
    chunks, the tiles next to ghost cells run after the halo exchange that follows the last chunk.  Same bits
    as the single-GPU run with the state resident on the device."""
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29800 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, "box", "tiles", q, 0.0, halo, 3, 6)) for r in range(2)]
    for p in procs:
        p.start()
    got, res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = box_flat(12, 10, 8, bc=(10, 5, 3, 7, 3, 3))
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    Q0 = mesh_np.random_state(f, seed=4)
    one = mstgpu.Context(f, order=2, flux="roe", inletQ=inlet, kernel="tiles")
    one.set_state(Q0)
    one.step(1e-4, 3)
    assert np.array_equal(got, one.get_state(), equal_nan=True)
    assert np.array_equal(res, one.residual(), equal_nan=True)


def _output_worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist
    from mstgpu import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    raw, f, inlet = _output_case(case)
    Q0 = mesh_np.random_state(f, seed=9)
    ptr, idx = host.node_faces(raw)
    w = host.node_weights(f, raw["nodes"].shape[0])
    P = mstgpu.Partition(f, world, rank, order=2)
    ctx = mstgpu.Context(P, order=2, flux="roe", inletQ=inlet, device=rank)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(mstgpu.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    ctx.comm_init(world, rank, bytes(idt.numpy().tobytes()))
    ctx.output_setup_partitioned(P, f, ptr, idx, w)
    ctx.set_state(Q0[P.cell_ids[:P.n_owned]])
    ctx.step(1e-4, 3)
    fld = ctx.node_fields()  # collective: fresh rows of the other rank's cells around this rank's nodes
    got = [None] * world
    dist.all_gather_object(got, (ctx.node_ids.copy(), fld))
    if rank == 0:
        q.put(got)
    ctx.close()
    dist.destroy_process_group()


def _output_case(case):
    from mstgpu import host
    if case == "box":
        raw = host.box_tets_raw(9, 8, 7, bc=(10, 5, 3, 7, 3, 3))
        return raw, host.flatten_raw(raw), np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    from conftest import load_raw
    raw = load_raw(case)
    return raw, host.flatten_raw(raw, "consistent"), None


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", ["box", "2d-stairW-1"])
def test_two_gpus_node_fields_match_one(case):
    """Output path on partitioned contexts (mstgpu_output_setup_partitioned): every node is computed by one rank from
    fresh rows of all cells around it -- also the other rank's, also beyond the ghost layers of the step.  The union of
    the two ranks' rows is every node exactly once and equals the single-GPU node fields bit for bit."""
    import torch.multiprocessing as mp
    from mstgpu import host
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29850 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_output_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    raw, f, inlet = _output_case(case)
    Q0 = mesh_np.random_state(f, seed=9)
    ptr, idx = host.node_faces(raw)
    one = mstgpu.Context(f, order=2, flux="roe", inletQ=inlet)
    one.output_setup(f, ptr, idx, host.node_weights(f, raw["nodes"].shape[0]))
    one.set_state(Q0)
    one.step(1e-4, 3)
    want = one.node_fields()
    nn = raw["nodes"].shape[0]
    seen = np.zeros(nn, dtype=np.int32)
    full = np.full_like(want, np.nan)
    for ids, fld in got:
        seen[ids] += 1
        full[ids] = fld
    has_faces = np.diff(ptr) > 0
    assert np.array_equal(seen[has_faces], np.ones(int(has_faces.sum()), dtype=np.int32))  # every node exactly once
    assert all(len(ids) > 0 for ids, _ in got)                                           # both ranks own nodes
    assert np.array_equal(full[has_faces], want[has_faces], equal_nan=True)
