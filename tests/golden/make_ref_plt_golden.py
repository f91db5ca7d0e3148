"""Regenerate tests/golden/ref_plt.json by RUNNING THE REFERENCE'S OWN reader and Tecplot writer
(oracle/_ref/ref_io = MshBlock::readMsh + Work::writedataRhoBasedMshNodePlt compiled from
/root/reference, oracle/refbuild/ref_io_driver.cpp).  Stored per case: SHA-256 and size of the
.plt file, plus the first node lines verbatim.  Runs in the build container only.

    make -C oracle ref && python tests/golden/make_ref_plt_golden.py
"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mst-cfd_b200")]
from oracle import mesh_np, mshio  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
MSH = "/root/reference/MST-CFD/msh"
OUT = os.path.dirname(os.path.abspath(__file__))

# name, mesh, state seed (None = the reference's SOD initial state), step counter printed in ZONE T, retag.
# The writer's switch has no case for symmetry zones (type 7): the reference reads its face array
# uninitialised there (Work.cpp:243, 250-284).  Meshes with such zones are retagged 7 -> 3 in the
# zone headers of the file both sides read, so that every face state is defined.
CASES = [
    ("sod_init", "2d-shockwavepipe-2", None, 0, None),
    ("sod_random", "2d-shockwavepipe-2", 3, 10, None),
    ("stair5_random", "2d-stair-un-5-tri", 20231017, 20, None),       # inlet 10, outlet 5, wall 3
    ("stair3loose_random", "2d-stair-un-3-loose-tri", 5, 30, None),
    ("stairW1_random_7to3", "2d-stairW-1", 7, 40, (7, 3)),            # quadrilateral cells
    ("stairW2st_random_7to3", "2d-stairW-2-st", 9, 50, (7, 3)),
]


def run_case(mesh, seed, t, tmp, retag=None):
    """-> path of the .plt the reference wrote"""
    msh = os.path.join(tmp, mesh + ".msh")
    with open(os.path.join(MSH, mesh + ".msh"), "rb") as fi, open(msh, "wb") as fo:
        text = fi.read().replace(b"\r", b"")  # CRLF trap, SURVEY.md 8c
        if retag:
            a, b = (format(v, "x").encode() for v in retag)
            text = re.sub(rb"^(\(13 \([0-9a-f]+ [0-9a-f]+ [0-9a-f]+ )" + a + rb"( [0-9a-f]+\))", rb"\g<1>" + b + rb"\2", text, flags=re.M)
        fo.write(text)
    init = "-"
    if seed is not None:
        flat = mesh_np.flatten(mshio.read_msh(msh))
        init = os.path.join(tmp, f"{mesh}_{seed}_init.bin")
        mesh_np.random_state(flat, seed=seed).tofile(init)
    os.makedirs(os.path.join(tmp, "result"), exist_ok=True)
    subprocess.run([os.path.join(REF, "ref_io"), msh, tmp, init, str(t)], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(tmp, "result", f"{mesh}.msh_TIME4000_u0_t{t}.plt")


if __name__ == "__main__":
    tmp = os.path.join(REF, "tmp")
    os.makedirs(tmp, exist_ok=True)
    gold = {}
    for name, mesh, seed, t, retag in CASES:
        data = open(run_case(mesh, seed, t, tmp, retag), "rb").read()
        lines = data.split(b"\n")
        gold[name] = dict(mesh=mesh, seed=seed, t=t, retag=retag, bytes=len(data), sha256=hashlib.sha256(data).hexdigest(),
                          head=[ln.decode() for ln in lines[:6]], tail=[ln.decode() for ln in lines[-4:]])
        print(name, len(data), gold[name]["sha256"][:16])
    # the residual log of the reference PROGRAM (Time.cpp:69-78): SOD, Roe / ACCURACY 2, consistent flags, 10 steps
    msh = os.path.join(tmp, "2d-shockwavepipe-2.msh")
    subprocess.run([os.path.join(REF, "ref_io"), msh, tmp, "-", "10", "10", "1"], check=True, stdout=subprocess.DEVNULL)
    lines = [ln for ln in open(os.path.join(tmp, "result", "ref-log.lhblog")).read().split("\n") if ln.strip()]
    gold["_log_sod_roe2_consistent_10steps"] = dict(
        mesh="2d-shockwavepipe-2", lines=lines,
        note="residual lines of Time.cpp:78 written by the reference program (ref_io <msh> <dir> - 10 10 1)")
    json.dump(gold, open(os.path.join(OUT, "ref_plt.json"), "w"), indent=1)
