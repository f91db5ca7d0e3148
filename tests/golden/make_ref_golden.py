"""Regenerate tests/golden/ref_*.npz by RUNNING THE REFERENCE ITSELF.

The binaries oracle/_ref/ref_{roe,ausm}{1,2} are the reference's own sources
(MshBlock reader, AllData, Time::goNextTimeStep, RhoSolver, solverRoe /
SolverAusm) compiled by oracle/refbuild/Makefile against the Eigen stand-in;
see that Makefile for exactly what is and is not the reference.  Runs in the
build container only (needs /root/reference); the outputs are committed so the
GPU box and the CPU suite can pin the oracle without the reference.

    make -C oracle ref && python tests/golden/make_ref_golden.py
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mesh_np, mshio, refdump  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
MSH = "/root/reference/MST-CFD/msh"
OUT = os.path.dirname(os.path.abspath(__file__))

# name, mesh, variant, flagmode, retag, init seed (None = the reference's SOD init), steps, dt is fixed 1/STEP_TIME
CASES = [
    ("sod_roe2_consistent", "2d-shockwavepipe-2", "roe2", 1, "-", None, (1, 10, 400)),
    ("sod_roe1_consistent", "2d-shockwavepipe-2", "roe1", 1, "-", None, (1, 400)),
    ("sod_roe2_as_shipped", "2d-shockwavepipe-2", "roe2", 0, "-", None, (1, 10)),
    ("stair5_roe1_random", "2d-stair-un-5-tri", "roe1", 1, "-", 20231017, (1, 5)),
    ("stair5_ausm1_random", "2d-stair-un-5-tri", "ausm1", 1, "-", 20231017, (1, 5)),
    # ACCURACY 2 writes through uninitialised pointers in the outlet case
    # (RhoSolver.cpp:360-361): outlet zones are retagged to symmetry (7) through
    # FacesInf::setType so the 2nd-order inlet / wall / symmetry code runs.
    ("stair5_roe2_random_5to7", "2d-stair-un-5-tri", "roe2", 1, "5:7", 20231017, (1, 5)),
    ("stair5_ausm2_random_5to7", "2d-stair-un-5-tri", "ausm2", 1, "5:7", 20231017, (1, 5)),
    ("stairW1_ausm1_random", "2d-stairW-1", "ausm1", 0, "-", 7, (1,)),
    ("stair4_roe2_random_5to3_as_shipped", "2d-stair-un-4-tri", "roe2", 0, "5:3", 11, (1,)),
]

if __name__ == "__main__":
    tmp = os.path.join(REF, "tmp")
    os.makedirs(tmp, exist_ok=True)
    for name, mesh, variant, flagmode, retag, seed, steps in CASES:
        msh = os.path.join(tmp, mesh + ".msh")
        with open(os.path.join(MSH, mesh + ".msh"), "rb") as fi, open(msh, "wb") as fo:
            fo.write(fi.read().replace(b"\r", b""))  # CRLF trap, SURVEY.md 8c
        init = "-"
        if seed is not None:
            flat = mesh_np.flatten(mshio.read_msh(msh))
            init = os.path.join(tmp, name + "_init.bin")
            mesh_np.random_state(flat, seed=seed).tofile(init)
        out = os.path.join(tmp, name + ".bin")
        cmd = [os.path.join(REF, "ref_" + variant), msh, out, str(flagmode), retag, init] + [str(s) for s in steps]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        d = refdump.read_dump(out)
        nc = int(d["hdr"][1])
        keep = dict(mesh=mesh, variant=variant, flagmode=flagmode, retag=retag,
                    seed=-1 if seed is None else seed, steps=np.array(steps))
        for s in steps:
            keep[f"Q{s}"] = d[f"Q{s}"].reshape(nc, -1)
        # per-face probes of the FIRST checkpoint's last solve, thinned to every 7th face
        nf = int(d["hdr"][2])
        keep["F_every7"] = d["F"].reshape(nf, 2, 4)[::7]
        # mesh tables as the reference's getters return them: SHA-256 of the raw
        # bytes (bit-exact pin of oracle/mesh_np.py without storing the arrays)
        for k in ("c0", "c1", "S", "dac", "fc", "eta", "flag", "ftype", "cc", "vol", "cf_ptr", "cf_idx", "sout"):
            keep["sha_" + k] = hashlib.sha256(np.ascontiguousarray(d[k]).tobytes()).hexdigest()
        np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **keep)
        print(name, os.path.getsize(os.path.join(OUT, f"ref_{name}.npz")) // 1024, "KiB")
