"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own
form code) into oracle Problems."""
import glob
import json
import os

import numpy as np

from oracle import tp_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    prm = orc.Params(**meta["params"])
    g = orc.Grid(meta["nx"], meta["ny"], meta["nz"], meta["dx"], meta["dy"], meta["dz"], meta["dim"])
    srcs = [orc.Source(int(r[0]), int(r[1]), float(r[2]), float(r[3]), float(r[4]), bool(r[5]))
            for r in z["sources"]]
    pb = orc.Problem(grid=g, nphase=meta["nphase"], prm=prm, phi=z["phi"], Kx=z["Kx"], Ky=z["Ky"],
                     Kz=z["Kz"] if g.dim == 3 else None, kT=z["kT"], sources=srcs)
    return meta, pb, z


def rel_err_rows(a, b):
    """max over rows of |a-b|_inf / |b|_inf  (rows = leading axes, cells = last axis)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.abs(b).max(axis=-1, keepdims=True)
    scale = np.where(scale == 0, 1.0, scale)
    return float((np.abs(a - b) / scale).max())
