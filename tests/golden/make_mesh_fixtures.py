"""Regenerate tests/golden/mesh_*.npz from the reference's own .msh files.

Run in the build container only (needs /root/reference); the GPU box never
reads /root/reference, it uses the committed .npz files.  The fixtures hold the
RAW tables of each mesh (nodes, face->nodes, c0, c1, zone table) -- data, not
source -- so that both the oracle's metric restatement (oracle/mesh_np.py) and
the product's flattener are exercised on the reference's real inputs.

    python tests/golden/make_mesh_fixtures.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mshio  # noqa: E402

SRC = "/root/reference/MST-CFD/msh"
OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for p in sorted(glob.glob(os.path.join(SRC, "*.msh"))):
        raw = mshio.read_msh(p)
        name = os.path.basename(p)[:-4]
        d = mshio.raw_to_npz_dict(raw)
        out = os.path.join(OUT, f"mesh_{name}.npz")
        np.savez_compressed(out, **d)
        print(name, raw["ncells"], raw["c0"].shape[0], os.path.getsize(out) // 1024, "KiB")
