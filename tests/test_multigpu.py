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


def _worker(rank, world, port, case, kernel, q):
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
    ctx = mstgpu.Context(P, order=2, flux="roe", inletQ=inlet, device=rank, kernel=kernel)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(mstgpu.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    ctx.comm_init(world, rank, bytes(idt.numpy().tobytes()))
    ctx.set_state(Q0[P.cell_ids[:P.n_owned]])
    ctx.step(1e-4, 5)
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
