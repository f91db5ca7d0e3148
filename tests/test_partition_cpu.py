"""CPU suite, part 4: the multi-GPU host logic without GPUs.

The partitioner (owned + 2 ghost layers, send / receive lists) is validated by
running the ORACLE on every partition's local mesh with the ghost states
exchanged (a) in-process and (b) between two real processes over
torch.distributed / gloo, and comparing the owned cells with the
single-domain oracle: the results must be bit-identical, which is what the
NCCL path on the GPUs relies on."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_flat, box_flat
from oracle import mesh_np, oracle
import mstgpu


def _exchange_inprocess(parts, Qs):
    for r, P in enumerate(parts):
        for nb in P.neighbors:
            src = parts[nb["rank"]]
            back = [x for x in src.neighbors if x["rank"] == r][0]
            assert len(back["send_local"]) == nb["recv_count"]
            # what the neighbour sends is exactly, and in order, what we hold as its ghosts
            assert np.array_equal(src.cell_ids[back["send_local"]],
                                  P.cell_ids[nb["recv_first"]:nb["recv_first"] + nb["recv_count"]])
            Qs[r][nb["recv_first"]:nb["recv_first"] + nb["recv_count"]] = Qs[nb["rank"]][back["send_local"]]


@pytest.mark.parametrize("case", ["2d-stairW-1", "box"])
@pytest.mark.parametrize("nparts", [2, 3, 5])
@pytest.mark.parametrize("order", [1, 2])
def test_partitioned_oracle_is_bit_identical(case, nparts, order):
    f = box_flat(7, 6, 5, bc=(10, 5, 3, 7, 3, 3)) if case == "box" else load_flat(case)
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58]) if case == "box" else None
    Q0 = mesh_np.random_state(f, seed=4)
    ref = oracle.Oracle(f, order=order, flux="roe", inletQ=inlet).run(1e-4, 3, Q0)
    parts = [mstgpu.Partition(f, nparts, r, order=order) for r in range(nparts)]
    assert sum(P.n_owned for P in parts) == f["ncells"]
    owned_all = np.concatenate([P.cell_ids[:P.n_owned] for P in parts])
    assert np.array_equal(np.sort(owned_all), np.arange(f["ncells"]))
    locs = [P.local_flat() for P in parts]
    ors = [oracle.Oracle(lf, order=order, flux="roe", inletQ=inlet, qf_copy_from=lf["nint"]) for lf in locs]
    Qs = [np.zeros((P.n_local, f["dim"] + 2)) for P in parts]
    for P, Q in zip(parts, Qs):
        Q[:P.n_owned] = Q0[P.cell_ids[:P.n_owned]]
    for step in range(3):
        _exchange_inprocess(parts, Qs)
        for r, P in enumerate(parts):
            Qn = ors[r].solve(1e-4, Qs[r])
            Qs[r][:P.n_owned] = Qn[:P.n_owned]
    out = np.empty_like(ref)
    for P, Q in zip(parts, Qs):
        out[P.cell_ids[:P.n_owned]] = Q[:P.n_owned]
    assert np.array_equal(out, ref, equal_nan=True)


def test_explicit_cell_part_and_errors():
    f = load_flat("2d-stair-un-5-tri")
    part = (f["cc"][:, 0] > np.median(f["cc"][:, 0])).astype(np.int32)  # slab split
    P0 = mstgpu.Partition(f, 2, 0, cell_part=part)
    assert P0.n_owned == int((part == 0).sum())
    assert np.array_equal(P0.cell_ids[:P0.n_owned], np.nonzero(part == 0)[0])
    with pytest.raises(mstgpu.MstGpuError):
        mstgpu.Partition(f, 2, 5)
    with pytest.raises(mstgpu.MstGpuError, match="owns no cells"):
        mstgpu.Partition(f, 2, 1, cell_part=np.zeros(f["ncells"], np.int32))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = load_flat("2d-stair-un-4-tri")
    Q0 = mesh_np.random_state(f, seed=8)
    P = mstgpu.Partition(f, world, rank, order=2)
    lf = P.local_flat()
    o = oracle.Oracle(lf, order=2, flux="ausm", qf_copy_from=lf["nint"], nthreads=2)
    Q = np.zeros((P.n_local, 4))
    Q[:P.n_owned] = Q0[P.cell_ids[:P.n_owned]]
    res = np.zeros(4)
    for step in range(4):
        reqs, bufs = [], []
        for nb in P.neighbors:  # the NCCL path groups the same sends / receives
            s = torch.from_numpy(np.ascontiguousarray(Q[nb["send_local"]]))
            rbuf = torch.empty((nb["recv_count"], 4), dtype=torch.float64)
            reqs += [dist.isend(s, nb["rank"]), dist.irecv(rbuf, nb["rank"])]
            bufs.append((nb, rbuf, s))
        for rq in reqs:
            rq.wait()
        for nb, rbuf, _ in bufs:
            Q[nb["recv_first"]:nb["recv_first"] + nb["recv_count"]] = rbuf.numpy()
        Qn = o.solve(1e-4, Q)
        with np.errstate(all="ignore"):
            x = np.abs(Qn[:P.n_owned] - Q[:P.n_owned]) / Q[:P.n_owned]
        res = np.nanmax(np.where(x > 0, x, 0), axis=0)
        Q[:P.n_owned] = Qn[:P.n_owned]
    t = torch.from_numpy(res.copy())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # residual: max over ranks (ncclAllReduce on the GPUs)
    # gather the owned rows on rank 0
    full = torch.zeros((f["ncells"], 4), dtype=torch.float64)
    full[torch.from_numpy(P.cell_ids[:P.n_owned].astype(np.int64))] = torch.from_numpy(Q[:P.n_owned].copy())
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((full.numpy(), t.numpy()))
    dist.destroy_process_group()


def test_two_process_gloo_halo_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = load_flat("2d-stair-un-4-tri")
    Q0 = mesh_np.random_state(f, seed=8)
    o = oracle.Oracle(f, order=2, flux="ausm")
    ref, rr = o.run(1e-4, 4, Q0, residuals=True)
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.allclose(res, rr[-1], rtol=1e-12)


@pytest.mark.parametrize("nparts", [1, 2, 3])
def test_partitioned_implicit_step_definition(nparts):
    """The lagged-ghost (block-Jacobi across partitions) implicit step of tests/implicit_partitioned.py:
    with one partition it IS the single-domain step; with several it converges to the same
    solution of the linear system as the sweeps are iterated (the 5-sweep iterates differ)."""
    import implicit_partitioned as ip
    f = box_flat(6, 5, 4, bc=(10, 5, 3, 7, 3, 3))
    inlet = np.array([1.0, 0.4, 0.0, 0.0, 2.58])
    kw = dict(order=2, flux="roe", inletQ=inlet, limiter="venkat", limiter_k=2.0)
    Q0 = mesh_np.random_state(f, seed=4)
    single = oracle.Oracle(f, **kw)
    dt = 10 * single.cfl_dt(1.0, Q0)
    parts = [mstgpu.Partition(f, nparts, r, order=2) for r in range(nparts)]
    locs = [P.local_flat() for P in parts]
    ors = [oracle.Oracle(lf, qf_copy_from=lf["nint"], **kw) for lf in locs]

    def run(iters):
        Qs = [np.zeros((P.n_local, 5)) for P in parts]
        for P, Q in zip(parts, Qs):
            Q[:P.n_owned] = Q0[P.cell_ids[:P.n_owned]]
        ip.step(ors, parts, Qs, dt, iters)
        out = np.empty_like(Q0)
        for P, Q in zip(parts, Qs):
            out[P.cell_ids[:P.n_owned]] = Q[:P.n_owned]
        return out

    if nparts == 1:
        # same rows in the partition's own numbering: the single-domain step with that sweep order
        order = parts[0].cell_ids[:parts[0].n_owned]
        assert np.allclose(run(5), single.step_implicit(dt, Q0, 5, sweep_order=order), rtol=1e-13, atol=1e-13)
    conv = single.step_implicit(dt, Q0, 80)
    assert np.abs(run(80) - conv).max() < 1e-9 * np.abs(conv).max()
    if nparts > 1:
        d5 = np.abs(run(5) - single.step_implicit(dt, Q0, 5)).max()
        assert 0 < d5 < 0.05 * np.abs(conv - Q0).max()  # a different iterate, close to the same answer
