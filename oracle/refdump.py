"""ORACLE / TEST INFRASTRUCTURE -- reader for the dumps of oracle/refbuild/ref_driver.cpp."""
import numpy as np

_DT = {"i4": np.int32, "f8": np.float64, "i1": np.int8, "u1": np.uint8}


def read_dump(path: str) -> dict:
    out = {}
    with open(path, "rb") as fh:
        while True:
            line = fh.readline()
            if not line:
                break
            name, dt, cnt = line.decode().split()
            dt = np.dtype(_DT[dt])
            out[name] = np.frombuffer(fh.read(int(cnt) * dt.itemsize), dtype=dt).copy()
    return out


def dump_to_flat(d: dict) -> dict:
    D, nc, nf, nint = (int(x) for x in d["hdr"])
    return dict(dim=D, ncells=nc, nfaces=nf, nint=nint, c0=d["c0"], c1=d["c1"], S=d["S"].reshape(nf, D),
                dac=d["dac"], fc=d["fc"].reshape(nf, D), eta=d["eta"], flag=d["flag"].reshape(nf, D),
                ftype=d["ftype"], cc=d["cc"].reshape(nc, D), vol=d["vol"], cf_ptr=d["cf_ptr"], cf_idx=d["cf_idx"])
