"""GPU suite: the whole application, end to end.  `mstrun` (mst-cfd_b200/host/mstrun.cpp) is the reference's
main program for the density-based solver -- main.cpp + Work::work + Time -- rebuilt on the two C ABIs:
native .msh reader, flattener, SOD initial state, GPU steps with the residual log, device node averaging and
the Tecplot writer.  Compared with what the REFERENCE ITSELF produced for the same case (its reader, its CPU
RhoSolver with Roe / ACCURACY 2, its writer): golden states recorded from the reference build
(tests/golden/ref_sod_roe2_consistent.npz), the golden digest of its t = 0 file, and -- where
oracle/_ref/ref_io travelled to the box -- a fresh run of the reference program beside it."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_raw
from msh_writer import write_msh
from mstgpu import host
from oracle import mesh_np, output_np
from test_output_cpu import GOLD, REF_IO

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "mst-cfd_b200", "mstrun")
MESH = "2d-shockwavepipe-2"


def read_plt(path, nn):
    lines = open(path).read().split("\n")
    return np.array([[float(x) for x in ln.split()] for ln in lines[3:3 + nn]]), lines[3 + nn:]


def read_log(path):
    return np.array([[float(x) for x in ln.split()] for ln in open(path).read().split("\n") if ln.strip()])


@pytest.fixture(scope="module")
def run(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("mstrun")
    msh = str(tmp / (MESH + ".msh"))
    write_msh(msh, load_raw(MESH))
    r = subprocess.run([EXE, msh, "--steps", "10", "--out", str(tmp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return tmp, msh, r.stdout


def test_initial_file_is_the_reference_writers(run):
    tmp, _, _ = run
    data = open(str(tmp / f"{MESH}.msh_TIME4000_u0_t0.plt"), "rb").read()
    g = GOLD["sod_init"]
    assert len(data) == g["bytes"] and hashlib.sha256(data).hexdigest() == g["sha256"]


def test_ten_steps_match_the_reference_run(run):
    """file at t = 10 == the writer's arithmetic applied to the state the reference's own solver reached"""
    tmp, msh, _ = run
    g = np.load(os.path.join(GOLDEN, "ref_sod_roe2_consistent.npz"))
    raw = host.read_msh(msh)
    f = host.flatten_raw(raw, "consistent")
    ptr, idx = host.node_faces(raw)
    nn = raw["nodes"].shape[0]
    exp = output_np.node_fields(f, raw, g["Q10"], ptr, idx)
    got, elems = read_plt(str(tmp / f"{MESH}.msh_TIME4000_u0_t10.plt"), nn)
    assert got.shape == (nn, 8)
    assert np.array_equal(got[:, :2], np.round(raw["nodes"], 15)) or np.abs(got[:, :2] - raw["nodes"]).max() < 1e-15
    # 15 printed decimals: absolute 5e-16 from the rounding of the text + 1e-12 relative from the solver
    assert np.abs(got[:, 2:] - exp).max() <= 1e-12 * max(1.0, np.abs(exp).max()) + 1e-15
    cp, ci = host.cell_nodes(raw, f)
    assert elems[0].split() == [str(v + 1) for v in ci[cp[0]:cp[1]]]
    # first residual line of the log == Time.cpp:69-76 on the reference's own states
    log = read_log(str(tmp / f"{MESH}.msh_TIME4000_u0-log.lhblog"))
    assert log.shape == (10, 4)
    Q0 = mesh_np.sod_initial_state(f)
    with np.errstate(all="ignore"):
        x = np.abs(g["Q1"] - Q0) / Q0
    x = np.where(np.isnan(x) | (x < 0), 0.0, x).max(axis=0)
    fin = np.isfinite(x)
    assert np.allclose(log[0][fin], x[fin], rtol=1e-9, atol=1e-15) and np.array_equal(np.isinf(log[0]), np.isinf(x))


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="oracle/_ref/ref_io did not travel to this box")
def test_beside_a_fresh_run_of_the_reference_program(run, tmp_path):
    tmp, msh, _ = run
    (tmp_path / "result").mkdir()
    subprocess.run([REF_IO, msh, str(tmp_path), "-", "10", "10", "1"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    nn = load_raw(MESH)["nodes"].shape[0]
    ref, ref_el = read_plt(str(tmp_path / "result" / f"{MESH}.msh_TIME4000_u0_t10.plt"), nn)
    got, got_el = read_plt(str(tmp / f"{MESH}.msh_TIME4000_u0_t10.plt"), nn)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) + 1e-15
    assert got_el == ref_el
    a, b = read_log(str(tmp / f"{MESH}.msh_TIME4000_u0-log.lhblog")), read_log(str(tmp_path / "result" / "ref-log.lhblog"))
    fin = np.isfinite(b)
    assert a.shape == b.shape and np.array_equal(np.isfinite(a), fin)
    assert np.allclose(a[fin], b[fin], rtol=1e-9, atol=1e-15)
