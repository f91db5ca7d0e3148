"""CPU suite: the native .msh reader (mst-cfd_b200/host/mshread.cpp) against
(a) the raw tables of the 8 reference meshes (tests/golden/mesh_*.npz, made from the shipped files),
(b) the digests of the REFERENCE BUILD's own mesh getters (tests/golden/ref_*.npz) after msthost_flatten,
(c) the shipped CRLF files themselves where /root/reference exists (this container only)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_raw
from msh_writer import write_msh
from mstgpu import host

MESHES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "mesh_*.npz")))
REF_MSH = "/root/reference/MST-CFD/msh"


def _same_raw(a, b):
    assert a["dim"] == b["dim"] and a["ncells"] == b["ncells"]
    for k in ("nodes", "c0", "c1"):
        assert np.array_equal(a[k], b[k]), k
    w = min(a["face_nodes"].shape[1], b["face_nodes"].shape[1])
    assert np.array_equal(a["face_nodes"][:, :w], b["face_nodes"][:, :w])
    assert (a["face_nodes"][:, w:] < 0).all() and (b["face_nodes"][:, w:] < 0).all()
    za = [(z["start"], z["end"], z["type"]) for z in a["zones"]]
    zb = [(z["start"], z["end"], z["type"]) for z in b["zones"]]
    assert za == zb


@pytest.mark.parametrize("name", MESHES)
def test_reader_round_trip_of_reference_meshes(name, tmp_path):
    """npz -> .msh (python test writer, repr() decimals) -> native reader: identical tables, bit for bit"""
    raw = load_raw(name)
    p = str(tmp_path / "m.msh")
    write_msh(p, raw)
    got = host.read_msh(p)
    _same_raw(got, raw)
    assert got["nint"] == max([z["end"] for z in raw["zones"] if z["type"] == 2], default=0)
    # CRLF image of the same file (what the reference ships): '\r' must not leak into the last id
    crlf = open(p, "rb").read().replace(b"\n", b"\r\n")
    _same_raw(host.parse_msh(crlf), raw)


@pytest.mark.parametrize("name", MESHES[:3])
def test_native_writer_round_trip(name, tmp_path):
    raw = load_raw(name)
    p = str(tmp_path / "w.msh")
    host.write_msh(p, raw)
    _same_raw(host.read_msh(p), raw)
    from oracle import mshio   # the oracle's reader accepts the native writer's file too
    _same_raw(mshio.read_msh(p), raw)


@pytest.mark.skipif(not os.path.isdir(REF_MSH), reason="shipped meshes live in /root/reference (this container only)")
@pytest.mark.parametrize("name", MESHES)
def test_reader_on_the_shipped_crlf_files(name):
    _same_raw(host.read_msh(os.path.join(REF_MSH, name + ".msh")), load_raw(name))


def test_reader_plus_flattener_equal_the_reference_build_tables(tmp_path):
    """native reader -> msthost_flatten == what the reference's MshBlock getters return (SHA-256 digests
    recorded from the reference build, tests/golden/make_ref_golden.py)"""
    done = set()
    for path in sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))):
        if os.path.basename(path).startswith("ref_lusgs_"):
            continue
        g = np.load(path)
        key = (str(g["mesh"]), int(g["flagmode"]), str(g["retag"]))
        if key in done:
            continue
        done.add(key)
        p = str(tmp_path / "m.msh")
        write_msh(p, load_raw(key[0]))
        raw = host.read_msh(p)
        if key[2] != "-":
            a, b = (int(x) for x in key[2].split(":"))
            for z in raw["zones"]:
                if z["type"] == a:
                    z["type"] = b
            del raw["ftype"]
        f = host.flatten_raw(raw, "consistent" if key[1] == 1 else "as_shipped")
        cid = np.repeat(np.arange(f["ncells"]), np.diff(f["cf_ptr"]))
        fid = f["cf_idx"]
        sgn = np.where(f["c0"][fid] == cid, 1.0, -1.0) * f["dac"][fid]
        f = dict(f, sout=sgn[:, None] * f["S"][fid])
        for k in ("c0", "c1", "S", "dac", "fc", "eta", "flag", "ftype", "cc", "vol", "cf_ptr", "cf_idx", "sout"):
            h = hashlib.sha256(np.ascontiguousarray(f[k]).tobytes()).hexdigest()
            assert h == str(g["sha_" + k]), (key, k)
    assert len(done) >= 5


def test_node_faces_follow_file_order():
    raw = load_raw(MESHES[0])
    ptr, idx = host.node_faces(raw)
    fn = raw["face_nodes"]
    exp = [[] for _ in range(raw["nodes"].shape[0])]
    for f in range(fn.shape[0]):
        for v in fn[f]:
            if v >= 0:
                exp[v].append(f)
    assert ptr[-1] == idx.size
    assert all(list(idx[ptr[v]:ptr[v + 1]]) == exp[v] for v in range(len(exp)))


def test_3d_file_with_triangle_faces(tmp_path):
    raw = host.raw_zones_from_ftype(host.box_tets_raw(3, 2, 4, 1.0, 0.5, 2.0, bc=(10, 5, 3, 3, 7, 7)))
    p = str(tmp_path / "b.msh")
    host.write_msh(p, raw)
    got = host.read_msh(p)
    _same_raw(got, raw)
    assert got["dim"] == 3 and got["face_nodes"].shape[1] == 3
    assert np.array_equal(got["ftype"], raw["ftype"])


_HEAD = b"(2 2)\n(10 (0 1 2 0 2))\n(10 (1 1 2 1 2)\n(\n"


@pytest.mark.parametrize("text, msg", [
    (_HEAD + b"0 0\n))\n(12 (0 1 1 0 0))\n(13 (0 1 1 0 0))\n", "file ends inside the node block"),
    (_HEAD + b"0 0\n1 x\n))\n(12 (0 1 1 0 0))\n(13 (0 1 1 0 0))\n", "unreadable coordinate"),
    (_HEAD + b"0 0\n1 0\n))\n(12 (0 1 1 0 0))\n(13 (0 1 1 0 0))\n(13 (3 1 1 3 2)(\n1 9 1 0\n))\n", "out-of-range id"),
    (_HEAD + b"0 0\n1 0\n))\n(12 (0 1 1 0 0))\n(13 (0 1 2 0 0))\n(13 (3 1 1 3 2)(\n1 2 1 0\n))\n", "face zones cover 1 of 2"),
    (b"(10 (0 1 2 0 2))\n", "before the"),
    # zone header fields must fit 32-bit ids (100000005 hex used to wrap to 5)
    (_HEAD + b"0 0\n1 0\n))\n(12 (0 1 1 0 0))\n(13 (0 1 1 0 0))\n(13 (3 1 100000001 3 2)(\n1 2 1 0\n))\n", "32-bit"),
])
def test_malformed_files_fail_loudly(text, msg):
    with pytest.raises(RuntimeError, match=msg):
        host.parse_msh(text)


@pytest.mark.parametrize("tail", [b"(2 ", b"(13", b"(1", b"("])
def test_truncated_header_at_the_end_of_the_input_is_not_read_past(tail):
    """a header line shorter than 4 bytes at the very end of the buffer (exact-size copy: no slack behind it)"""
    text = bytes(_HEAD + b"0 0\n1 0\n))\n(12 (0 1 1 0 0))\n(13 (0 1 1 0 0))\n(13 (3 1 1 3 2)(\n1 2 1 0\n))\n" + tail)
    try:
        host.parse_msh(text)
    except RuntimeError:
        pass  # an error is fine; reading b[3] of a 3-byte line is not


def test_missing_file_is_an_error():
    with pytest.raises(RuntimeError, match="cannot open"):
        host.read_msh("/nonexistent/x.msh")
