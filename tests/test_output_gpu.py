"""GPU suite: the device half of the output path (csrc/output.cuh) through the C ABI.

mstgpu_node_fields must return, BIT FOR BIT, the numbers the reference's Tecplot writer prints
(Work::writedataRhoBasedMshNodePlt, R/work/Work.cpp:243-304): compared with oracle/output_np.py (pinned
to the reference's own writer byte for byte in tests/test_output_cpu.py), with the golden digests of
files the reference wrote (tests/golden/ref_plt.json), and -- where oracle/_ref/ref_io travelled to the
box -- with a fresh run of the reference's reader + writer on a state no golden holds."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import mstgpu
from conftest import GOLDEN, ROOT, box_flat
from mstgpu import host
from oracle import mesh_np, output_np
from test_output_cpu import GOLD, REF_IO, case_inputs

pytestmark = pytest.mark.gpu


def device_fields(f, raw, Q, ptr, idx, steps=0, dt=1e-4, **kw):
    ctx = mstgpu.Context(f, **kw)
    ctx.output_setup(f, ptr, idx, host.node_weights(f, raw["nodes"].shape[0]))
    ctx.set_state(Q)
    if steps:
        ctx.step(dt, steps)
    out = ctx.node_fields()
    Qd = ctx.get_state()
    n = ctx.launch_count
    ctx.close()
    return out, Qd, n


@pytest.mark.parametrize("name", sorted(GOLD))
def test_device_node_fields_reproduce_the_reference_writers_file(name, tmp_path):
    g = GOLD[name]
    _, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    fld, _, launches = device_fields(f, raw, Q, ptr, idx, order=1)
    assert launches > 0
    assert np.array_equal(fld, output_np.node_fields(f, raw, Q, ptr, idx), equal_nan=True)
    cp, ci = host.cell_nodes(raw, f)
    out = str(tmp_path / "o.plt")
    host.plt_write(out, raw, fld, cp, ci, zone_t=g["t"])
    data = open(out, "rb").read()
    assert len(data) == g["bytes"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.parametrize("renumber", [0, 1, 2])
def test_node_fields_after_stepping_any_cell_order(renumber, tmp_path):
    """the state the kernel reads is the device's own (renumbered) one: 5 second-order steps, then the node
    fields equal the restatement applied to the downloaded state, bit for bit"""
    g = GOLD["sod_init"]
    _, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    f = host.flatten_raw(raw, "consistent")
    fld, Qd, _ = device_fields(f, raw, Q, ptr, idx, steps=5, dt=2.5e-4, order=2, flux="roe", renumber=renumber)
    assert np.array_equal(fld, output_np.node_fields(f, raw, Qd, ptr, idx), equal_nan=True)


def test_node_fields_3d_tets():
    """3-D form (extension: q^2 includes w): tet box with inlet / outlet / wall / symmetry zones"""
    raw = host.raw_zones_from_ftype(host.box_tets_raw(6, 5, 4, 1.0, 0.8, 0.6, bc=(10, 5, 3, 3, 7, 7)))
    f = host.flatten_raw(raw)
    Q = mesh_np.random_state(f, seed=5)
    ptr, idx = host.node_faces(raw)
    fld, Qd, _ = device_fields(f, raw, Q, ptr, idx, steps=2, order=2)
    assert fld.shape == (raw["nodes"].shape[0], 7)
    assert np.array_equal(fld, output_np.node_fields(f, raw, Qd, ptr, idx), equal_nan=True)


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="oracle/_ref/ref_io did not travel to this box")
def test_against_a_fresh_run_of_the_reference_reader_and_writer(tmp_path):
    g = dict(mesh="2d-stair-un-3-tri", seed=99, t=12)
    p, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    (tmp_path / "result").mkdir()
    Q.tofile(str(tmp_path / "q.bin"))
    subprocess.run([REF_IO, p, str(tmp_path), str(tmp_path / "q.bin"), "12"], check=True, stdout=subprocess.DEVNULL)
    ref = open(str(tmp_path / "result" / (g["mesh"] + ".msh_TIME4000_u0_t12.plt")), "rb").read()
    fld, _, _ = device_fields(f, raw, Q, ptr, idx, order=1)
    cp, ci = host.cell_nodes(raw, f)
    out = str(tmp_path / "o.plt")
    host.plt_write(out, raw, fld, cp, ci, zone_t=12)
    assert open(out, "rb").read() == ref


def test_output_errors():
    f = box_flat(3, 3, 3)
    ctx = mstgpu.Context(f)
    with pytest.raises(mstgpu.MstGpuError, match="before output_setup"):
        ctx.nnodes = 4
        ctx.node_fields()
    ctx.close()
