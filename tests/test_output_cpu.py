"""CPU suite: the host half of the output path against the REFERENCE'S OWN
Tecplot writer (Work::writedataRhoBasedMshNodePlt, R/work/Work.cpp:204-319).

tests/golden/ref_plt.json holds SHA-256 digests of files the reference's writer produced
(tests/golden/make_ref_plt_golden.py).  Here the native .msh reader, the flattener, the oracle's
numpy restatement of the node averaging (oracle/output_np.py) and the native writer
(host/pltwrite.cpp) must reproduce those files byte for byte.  The device half
(mstgpu_node_fields) is compared with the same numbers in tests/test_output_gpu.py."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_raw
from msh_writer import write_msh
from mstgpu import host
from oracle import mesh_np, output_np

_ALL = json.load(open(os.path.join(GOLDEN, "ref_plt.json")))
GOLD = {k: v for k, v in _ALL.items() if not k.startswith("_")}     # .plt cases
REF_LOG = _ALL["_log_sod_roe2_consistent_10steps"]                  # residual log of the reference program
REF_IO = os.path.join(ROOT, "oracle", "_ref", "ref_io")


def case_inputs(g, tmp_path):
    """(raw, flat, Q, node->face CSR) of a golden case, through the native reader"""
    p = str(tmp_path / (g["mesh"] + ".msh"))
    src = load_raw(g["mesh"])
    for z in src["zones"]:   # symmetry zones: undefined in the reference's writer, see make_ref_plt_golden.py
        if g.get("retag") and z["type"] == g["retag"][0]:
            z["type"] = g["retag"][1]
    write_msh(p, src)
    raw = host.read_msh(p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # 2d-stair*-st: sliver quads with NaN Heron volumes (Cell.cpp:28-49)
        f = host.flatten_raw(raw, "as_shipped")
        fo = mesh_np.flatten(load_raw(g["mesh"]), "as_shipped")
    Q = mesh_np.sod_initial_state(fo) if g["seed"] is None else mesh_np.random_state(fo, seed=g["seed"])
    return p, raw, f, Q, host.node_faces(raw)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_plt_file_equals_the_reference_writers(name, tmp_path):
    g = GOLD[name]
    _, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    fld = output_np.node_fields(f, raw, Q, ptr, idx)
    cp, ci = host.cell_nodes(raw, f)
    out = str(tmp_path / "o.plt")
    host.plt_write(out, raw, fld, cp, ci, zone_t=g["t"])
    data = open(out, "rb").read()
    lines = data.split(b"\n")
    assert [ln.decode() for ln in lines[:6]] == g["head"]
    assert [ln.decode() for ln in lines[-4:]] == g["tail"]
    assert len(data) == g["bytes"]
    assert hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="oracle/_ref/ref_io not built (make -C oracle ref)")
def test_plt_file_against_a_fresh_run_of_the_reference_writer(tmp_path):
    """the reference's reader + writer run here on a state no golden holds"""
    g = dict(mesh="2d-stair-un-4-tri", seed=424242, t=3)
    p, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    (tmp_path / "result").mkdir()
    Q.tofile(str(tmp_path / "q.bin"))
    subprocess.run([REF_IO, p, str(tmp_path), str(tmp_path / "q.bin"), "3"], check=True, stdout=subprocess.DEVNULL)
    ref = open(str(tmp_path / "result" / (g["mesh"] + ".msh_TIME4000_u0_t3.plt")), "rb").read()
    out = str(tmp_path / "o.plt")
    cp, ci = host.cell_nodes(raw, f)
    host.plt_write(out, raw, output_np.node_fields(f, raw, Q, ptr, idx), cp, ci, zone_t=3)
    assert open(out, "rb").read() == ref


def test_writer_spells_non_finite_and_negative_zero_like_iostream(tmp_path):
    raw = dict(dim=2, nodes=np.array([[0.0, -0.0], [1.5, 2.25], [1e6, -3.0], [0.1, 0.2]]))
    fld = np.array([[np.nan, -np.nan, np.inf, -np.inf, -np.nan, np.inf],
                    [-0.0, 1.0 / 3.0, -np.inf, 1e-20, np.nan, -0.0],
                    [1.0, 2.0, 3.0, 4.0, 5.0, 6.0],
                    [123456789.125, -1e-300, 5e-16, 4.9999e-16, 0.5, 2.5]])
    cp = np.array([0, 3], dtype=np.int32); ci = np.array([0, 1, 2], dtype=np.int32)
    n = fld.shape[0]
    out = str(tmp_path / "x.plt")
    host.plt_write(out, raw, fld, cp, ci, zone_t=1, felnum=3)
    lines = open(out).read().split("\n")
    assert lines[2].endswith("ZONETYPE=FETRIANGLE")
    # glibc printf / libstdc++ num_put spell a NaN with the sign bit set "-nan" (Python drops the sign)
    # setw(15) precedes the coordinates, rho, u, v, T (not p, Ma): it only shows on nan / inf
    fmt = lambda v, k: ((("-nan" if np.signbit(v) else "nan") if np.isnan(v) else "%.15f" % v).rjust(15 if k < 6 else 0))
    exp = lambda row: " ".join(fmt(v, k) for k, v in enumerate(row)) + " "
    for i in range(n):
        assert lines[3 + i] == exp(list(raw["nodes"][i]) + list(fld[i]))
    assert lines[3 + n] == "1 2 3 "


def test_binary_twin_round_trips(tmp_path):
    g = GOLD["stair5_random"]
    _, raw, f, Q, (ptr, idx) = case_inputs(g, tmp_path)
    fld = output_np.node_fields(f, raw, Q, ptr, idx)
    cp, ci = host.cell_nodes(raw, f)
    out = str(tmp_path / "o.bin")
    host.plt_write(out, raw, fld, cp, ci, zone_t=5, binary=True)
    b = open(out, "rb").read()
    assert b[:8] == b"MSTPLT1\0"
    dim, t = np.frombuffer(b, np.int32, 2, 8)
    nn, nc, nconn = np.frombuffer(b, np.int64, 3, 16)
    assert (dim, t, nn, nc, nconn) == (2, 5, raw["nodes"].shape[0], f["ncells"], ci.size)
    o = 40
    assert np.array_equal(np.frombuffer(b, np.float64, nn * 2, o).reshape(nn, 2), raw["nodes"]); o += nn * 16
    assert np.array_equal(np.frombuffer(b, np.float64, nn * 6, o).reshape(nn, 6), fld, equal_nan=True); o += nn * 48
    assert np.array_equal(np.frombuffer(b, np.int32, nc + 1, o), cp); o += (nc + 1) * 4
    assert np.array_equal(np.frombuffer(b, np.int32, nconn, o), ci)
    assert o + nconn * 4 == len(b)


def test_cell_nodes_of_a_tet_box():
    raw = host.box_tets_raw(3, 2, 2)
    f = host.flatten_raw(raw)
    cp, ci = host.cell_nodes(raw, f)
    assert (np.diff(cp) == 4).all()
    # every cell's 4 nodes are exactly the union of its faces' nodes
    for c in (0, 5, f["ncells"] - 1):
        faces = f["cf_idx"][f["cf_ptr"][c]:f["cf_ptr"][c + 1]]
        assert set(ci[cp[c]:cp[c + 1]]) == set(raw["face_nodes"][faces].ravel())


def test_oracle_residuals_equal_the_reference_programs_log():
    """Row I of SURVEY 8a (Time.cpp:69-76): the residual lines the reference PROGRAM logged for 10 SOD steps
    (signed denominator, inf where the old momentum is zero) against the oracle's restatement"""
    from conftest import load_flat
    from oracle import oracle
    f = load_flat("2d-shockwavepipe-2", "consistent")
    _, r = oracle.Oracle(f, order=2, flux="roe").run(2.5e-4, 10, mesh_np.sod_initial_state(f), residuals=True)
    ref = np.array([[float(x) for x in ln.split()] for ln in REF_LOG["lines"]])
    assert ref.shape == (10, 4)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(r[:, :4]), fin) and np.array_equal(np.isinf(r[:, :4]), np.isinf(ref))
    # 15 printed decimals
    assert np.abs(r[:, :4][fin] - ref[fin]).max() <= 1e-12 * np.abs(ref[fin]).max() + 5e-16
