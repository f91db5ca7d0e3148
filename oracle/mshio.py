"""ORACLE / TEST INFRASTRUCTURE -- not product code.

Reader for the subset of the Fluent ASCII ``.msh`` format that the reference's
``MshBlock`` accepts (R = /root/reference/MST-CFD):

  * R/mesh/MshBlock.cpp:75-110   splitInformationData2D  (header / data split)
  * R/mesh/MshBlock.cpp:111-271  dataToMesh2D            (sections 10, 12, 13)
  * R/mesh/MshBlock.cpp:436-627  3-D twin
  * R/work/FUNCTION.cpp:41-55    hexStringToInt (lower-case hex only)

Only the raw tables are produced here (nodes, face->nodes, c0, c1, zones);
the metrics of R/mesh/{Face,Cell}.cpp are restated in ``mesh_np.py``.

Portability trap handled on purpose: the shipped files are
CRLF.  ``hexStringToInt`` weights digits by *string length*, so a trailing
``\\r`` would multiply the last cell id of each line by 16 on Linux.  The
reference was written on Windows where the C runtime strips it; we strip it.
"""
from __future__ import annotations

import re
import numpy as np

_HDR = re.compile(r"^\((\d+)\s*\(([^)]*)\)")


def read_msh(path: str) -> dict:
    """Return the raw mesh tables of a Fluent ASCII .msh file.

    keys: dim, nodes (nn,dim) f64, face_nodes (nf,k) i32 0-based (k = nodes per
    face; -1 padded), c0, c1 (nf,) i32 0-based with c1 = -1 on boundary faces,
    zones = list of dict(id, start, end, type, name) with 0-based half-open
    [start, end) face ranges in file order, ncells.
    """
    with open(path, "rb") as fh:
        text = fh.read().decode("ascii", errors="replace").replace("\r", "")
    lines = text.split("\n")
    dim = None
    nodes = None
    nn = ncells = nfaces = None
    face_nodes = None
    c0 = c1 = None
    zones = []
    last_comment = ""
    i = 0
    nl = len(lines)
    while i < nl:
        ln = lines[i]
        m = _HDR.match(ln)
        if not ln.startswith("("):
            i += 1
            continue
        if ln.startswith("(0 "):
            q = ln.split('"')
            last_comment = q[1] if len(q) > 1 else ""
            i += 1
            continue
        if ln.startswith("(2 "):
            dim = int(ln[3])
            i += 1
            continue
        if m is None:
            i += 1
            continue
        sec = int(m.group(1))
        tok = m.group(2).split()
        if sec == 10:
            zid = int(tok[0], 16)
            first, last = int(tok[1], 16), int(tok[2], 16)
            if zid == 0:
                nn = last
                nodes = np.zeros((nn, dim), dtype=np.float64)
                i += 1
                continue
            # data follows: optional "(" line then one node per line
            i += 1
            if lines[i].strip() == "(":
                i += 1
            cnt = last - first + 1
            blk = np.array(
                [[float(t) for t in lines[i + k].split()[:dim]] for k in range(cnt)],
                dtype=np.float64,
            )
            nodes[first - 1 : last] = blk
            i += cnt
            continue
        if sec == 12:
            zid = int(tok[0], 16)
            if zid == 0:
                ncells = int(tok[2], 16)
            i += 1
            continue
        if sec == 13:
            zid = int(tok[0], 16)
            first, last = int(tok[1], 16), int(tok[2], 16)
            if zid == 0:
                nfaces = last
                c0 = np.full(nfaces, -1, dtype=np.int32)
                c1 = np.full(nfaces, -1, dtype=np.int32)
                i += 1
                continue
            btype = int(tok[3], 16)
            npf = int(tok[4], 16)  # nodes per face (2 = line, 3 = tri, 4 = quad)
            if face_nodes is None:
                face_nodes = np.full((nfaces, max(npf, dim)), -1, dtype=np.int32)
            if npf > face_nodes.shape[1]:
                ext = np.full((nfaces, npf), -1, dtype=np.int32)
                ext[:, : face_nodes.shape[1]] = face_nodes
                face_nodes = ext
            i += 1
            cnt = last - first + 1
            rows = np.array(
                [[int(t, 16) for t in lines[i + k].split()] for k in range(cnt)],
                dtype=np.int64,
            )
            face_nodes[first - 1 : last, :npf] = rows[:, :npf] - 1
            c0[first - 1 : last] = rows[:, npf] - 1
            # R/mesh/MshBlock.cpp:242-259: only type-2 zones read a second cell
            if btype == 2:
                c1[first - 1 : last] = rows[:, npf + 1] - 1
            zones.append(
                dict(id=zid, start=first - 1, end=last, type=btype, name=last_comment)
            )
            i += cnt
            continue
        i += 1
    return dict(
        dim=dim,
        nodes=nodes,
        face_nodes=face_nodes,
        c0=c0,
        c1=c1,
        zones=zones,
        ncells=ncells,
    )


def raw_to_npz_dict(raw: dict) -> dict:
    """Flatten the zone list so the raw mesh can go through np.savez."""
    z = raw["zones"]
    return dict(
        dim=np.int32(raw["dim"]),
        ncells=np.int32(raw["ncells"]),
        nodes=raw["nodes"],
        face_nodes=raw["face_nodes"],
        c0=raw["c0"],
        c1=raw["c1"],
        zone_start=np.array([q["start"] for q in z], dtype=np.int32),
        zone_end=np.array([q["end"] for q in z], dtype=np.int32),
        zone_type=np.array([q["type"] for q in z], dtype=np.int32),
    )


def npz_to_raw(d) -> dict:
    zones = [
        dict(id=k, start=int(s), end=int(e), type=int(t), name="")
        for k, (s, e, t) in enumerate(zip(d["zone_start"], d["zone_end"], d["zone_type"]))
    ]
    return dict(
        dim=int(d["dim"]),
        ncells=int(d["ncells"]),
        nodes=np.asarray(d["nodes"], dtype=np.float64),
        face_nodes=np.asarray(d["face_nodes"], dtype=np.int32),
        c0=np.asarray(d["c0"], dtype=np.int32),
        c1=np.asarray(d["c1"], dtype=np.int32),
        zones=zones,
    )
