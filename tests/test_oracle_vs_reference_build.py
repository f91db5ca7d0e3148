"""CPU suite, part 3: the oracle against OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/ref_*.npz were produced by the reference's own sources (MshBlock
reader, Time::goNextTimeStep, RhoSolver, solverRoe / SolverAusm) compiled with
an Eigen stand-in -- oracle/refbuild/Makefile, tests/golden/make_ref_golden.py.
This pins: the .msh reader and every mesh metric (bit-exact digests), the
initial state, both flux schemes, both orders, inlet / wall / symmetry / outlet
handling, the off-by-one face, both flag conventions and 400 steps of SOD.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_raw, rel_linf
from oracle import mesh_np, oracle

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz"))
               if not os.path.basename(p).startswith("ref_lusgs_"))
DT = 1.0 / 4e3  # Time.cpp:62, CONST.h:51


def _flat_for(g):
    import warnings
    raw = load_raw(str(g["mesh"]))
    retag = str(g["retag"])
    if retag != "-":
        a, b = (int(x) for x in retag.split(":"))
        for z in raw["zones"]:
            if z["type"] == a:
                z["type"] = b
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mesh_np.flatten(raw, "consistent" if int(g["flagmode"]) == 1 else "as_shipped")


def test_golden_cases_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("case", CASES)
def test_mesh_tables_bit_exact(case):
    g = np.load(os.path.join(GOLDEN, f"ref_{case}.npz"))
    f = _flat_for(g)
    # Sout per cell slot as Cell::getBeginItDirectOfNbFaces returns it (MshBlock.cpp:307-318)
    cid = np.repeat(np.arange(f["ncells"]), np.diff(f["cf_ptr"]))
    fid = f["cf_idx"]
    sgn = np.where(f["c0"][fid] == cid, 1.0, -1.0) * f["dac"][fid]
    f = dict(f, sout=sgn[:, None] * f["S"][fid])
    for k in ("c0", "c1", "S", "dac", "fc", "eta", "flag", "ftype", "cc", "vol", "cf_ptr", "cf_idx", "sout"):
        h = hashlib.sha256(np.ascontiguousarray(f[k]).tobytes()).hexdigest()
        assert h == str(g["sha_" + k]), f"{k} differs from the reference's table"


@pytest.mark.parametrize("case", CASES)
def test_states_match_reference_run(case):
    g = np.load(os.path.join(GOLDEN, f"ref_{case}.npz"))
    f = _flat_for(g)
    variant = str(g["variant"])
    flux, order = variant[:-1], int(variant[-1])
    seed = int(g["seed"])
    Q = mesh_np.sod_initial_state(f) if seed < 0 else mesh_np.random_state(f, seed=seed)
    o = oracle.Oracle(f, order=order, flux=flux, nthreads=8)
    done = 0
    first = True
    for s in g["steps"]:
        Q = o.run(DT, int(s) - done, Q)
        done = int(s)
        Qr = g[f"Q{int(s)}"]
        assert np.array_equal(np.isfinite(Q), np.isfinite(Qr))
        if np.isfinite(Qr).all():
            # same arithmetic order + same inverse algorithm -> expected bit-exact;
            # 1e-13 leaves room for a different compiler's sqrt/div scheduling
            assert rel_linf(Q, Qr) <= 1e-13, (case, s)
        if first:
            F = o.probe()[2]
            Fr = g["F_every7"]
            m = np.isfinite(Fr)
            assert np.array_equal(np.isfinite(F[::7]), m)
            assert np.abs(F[::7][m] - Fr[m]).max() <= 1e-13 * max(1.0, np.abs(Fr[m]).max())
            first = False


def test_reference_as_shipped_flags_fail_like_the_oracle_says():
    """The reference itself, flags as its reader builds them, is NaN-bound on
    its own SOD case; 10 steps are still finite."""
    g = np.load(os.path.join(GOLDEN, "ref_sod_roe2_as_shipped.npz"))
    assert np.isfinite(g["Q10"]).all()
    f = _flat_for(g)
    Q = oracle.Oracle(f, order=2, flux="roe").run(DT, 60, mesh_np.sod_initial_state(f))
    assert not np.isfinite(Q).all()
