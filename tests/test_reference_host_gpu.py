"""GPU suite: the drop-in claim end to end.  The reference's OWN host code
(MshBlock .msh reader, AllData, Time::initialization -- compiled from
/root/reference by oracle/refbuild) drives the GPU solver through
mst-cfd_b200/host/GpuRhoSolver.h in the call sequence of Time::goNextTimeStep,
and must reproduce what the same host produced with the reference's CPU
RhoSolver (tests/golden/ref_*.npz)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_raw, rel_linf
from oracle import mesh_np, refdump
from msh_writer import write_msh

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
CASES = ["sod_roe2_consistent", "sod_roe1_consistent", "stair5_roe1_random", "stair5_ausm1_random",
         "stair5_roe2_random_5to7", "stairW1_ausm1_random"]


@pytest.mark.parametrize("host_mirror", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_reference_time_loop_on_the_gpu(case, host_mirror, tmp_path):
    """host_mirror = 1: GpuRhoSolver::options().host_mirror -- the fields stay in AllData and solve() is one streamed
    step (mstgpu_step_host), the data flow of the reference's own RhoSolver::solve."""
    g = np.load(os.path.join(GOLDEN, f"ref_{case}.npz"))
    variant = str(g["variant"])
    exe = os.path.join(REF, f"ref_gpu_{variant}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gpu_* not built (needs /root/reference at build time)")
    raw = load_raw(str(g["mesh"]))
    msh = str(tmp_path / "mesh.msh")
    write_msh(msh, raw)
    init = "-"
    if int(g["seed"]) >= 0:
        flat = mesh_np.flatten(raw)
        init = str(tmp_path / "init.bin")
        mesh_np.random_state(flat, seed=int(g["seed"])).tofile(init)
    out = str(tmp_path / "out.bin")
    steps = [int(s) for s in g["steps"] if int(s) <= 400]
    cmd = [exe, msh, out, str(int(g["flagmode"])), str(g["retag"]), init] + [str(s) for s in steps]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, timeout=600, env=dict(os.environ, MST_HOST_MIRROR=str(host_mirror)))
    d = refdump.read_dump(out)
    nc = int(d["hdr"][1])
    for s in steps:
        Qg, Qr = d[f"Q{s}"].reshape(nc, -1), g[f"Q{s}"]
        tol = 1e-12 if s == 1 else 1e-9
        assert rel_linf(Qg, Qr) <= tol, (case, s)
    # output path through the reference's own Node / Face lists: the device's node fields of the final state
    # equal the restatement of Work.cpp:243-304 (itself pinned to the reference's writer, test_output_cpu.py)
    # bit for bit
    if "node_fields" in d and np.isfinite(d[f"Q{steps[-1]}"]).all():
        from mstgpu import host
        from oracle import output_np
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = mesh_np.flatten(raw, "consistent" if int(g["flagmode"]) == 1 else "as_shipped")
        retag = str(g["retag"])
        if retag != "-":
            a, b = (int(x) for x in retag.split(":"))
            f["ftype"] = np.where(f["ftype"] == a, b, f["ftype"]).astype(f["ftype"].dtype)
        ptr, idx = host.node_faces(raw)
        exp = output_np.node_fields(f, raw, d[f"Q{steps[-1]}"].reshape(nc, -1), ptr, idx)
        assert np.array_equal(d["node_fields"].reshape(exp.shape), exp, equal_nan=True)
    # the residual the reference's host loop computes from the downloaded arrays equals the device reduction
    r = d["resid_host_dev"].reshape(-1, 2)
    fin = np.isfinite(r).all(axis=1)
    assert np.allclose(r[fin, 0], r[fin, 1], rtol=1e-12, atol=0)
