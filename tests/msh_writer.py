"""Test helper: write raw mesh tables (the tests/golden/mesh_*.npz fixtures) back
to the Fluent ASCII subset the reference's MshBlock reader accepts (SURVEY.md
Appendix C), so the reference's own reader can run on the GPU box where
/root/reference does not exist."""
import numpy as np


def write_msh(path: str, raw: dict):
    nodes, fn, c0, c1 = raw["nodes"], raw["face_nodes"], raw["c0"], raw["c1"]
    dim, nc, nn, nf = int(raw["dim"]), int(raw["ncells"]), nodes.shape[0], c0.shape[0]
    with open(path, "w") as f:
        f.write('(0 " written by tests/msh_writer.py")\n')
        f.write(f"(2 {dim})\n")
        f.write('(0 "Node Section")\n')
        f.write(f"(10 (0 1 {nn:x} 0 {dim}))\n")
        f.write(f"(10 (5 1 {nn:x} 1 {dim})\n(\n")
        for p in nodes:
            f.write(" ".join(repr(float(x)) for x in p) + "\n")
        f.write("))\n")
        f.write(f"(12 (0 1 {nc:x} 0 0))\n")
        f.write(f"(12 (6 1 {nc:x} 1 1))\n")
        f.write(f"(13 (0 1 {nf:x} 0 0))\n")
        for zi, z in enumerate(raw["zones"]):
            npf = int((fn[z["start"]] >= 0).sum())
            f.write(f'(0 "Faces of zone Z{zi}")\n')
            f.write(f"(13 ({zi + 7:x} {z['start'] + 1:x} {z['end']:x} {z['type']:x} {npf:x})(\n")
            for i in range(z["start"], z["end"]):
                ids = " ".join(f"{int(v) + 1:x}" for v in fn[i, :npf])
                f.write(f"{ids} {int(c0[i]) + 1:x} {int(c1[i]) + 1 if c1[i] >= 0 else 0:x}\n")
            f.write(")\n)\n")
        f.write('(0 "Zone Sections")\n')
