"""SPE10 on-disk formats - mirror of data/create_SPE10_slice.py, create_SPE10_slice2D.py of the reference.

The public SPE10 model-2 files (not distributed with the reference, .MISSING_LARGE_BLOBS:1-2):
  spe_phi.dat   60*220*85 porosities, x fastest, then y, then layers TOP-DOWN
  spe_perm.dat  3 blocks (Kx, Ky, Kz) of 60*220*85 permeabilities in mD, same order
`create_slice` / `create_slice2D` cut a window and return / save the four `slice_*.npy` arrays that
SPE10model3D.py:26-68 / SPE10model.py:23-60 load: indexed [i, j(, k)], K converted to mm^2
(x 9.869233e-10, create_SPE10_slice.py:42) and layers flipped so that z is up (:31).
`write_dat` writes fields back in the .dat layout (used to round-trip the synthetic generator in the tests).
"""
from __future__ import annotations

import os

import numpy as np

NX, NY, NZ = 60, 220, 85
MD_TO_MM2 = 9.869233e-10


def read_dat(data_dir):
    """-> (phi, kx_md, ky_md, kz_md), each shaped (NZ, NY, NX) with layer 0 = TOP (file order)."""
    phi = np.loadtxt(os.path.join(data_dir, "spe_phi.dat")).reshape(-1)
    perm = np.loadtxt(os.path.join(data_dir, "spe_perm.dat")).reshape(-1)
    n = NX * NY * NZ
    if phi.size != n or perm.size != 3 * n:
        raise ValueError("spe_phi.dat / spe_perm.dat do not hold 60x220x85 (x3) values")
    shp = (NZ, NY, NX)
    return phi.reshape(shp), perm[:n].reshape(shp), perm[n:2 * n].reshape(shp), perm[2 * n:].reshape(shp)


def write_dat(data_dir, phi, kx_md, ky_md, kz_md, per_line=6):
    """inverse of read_dat (arrays (NZ, NY, NX), top-down layers)."""
    os.makedirs(data_dir, exist_ok=True)

    def dump(path, flat):
        pad = (-flat.size) % per_line
        rows = np.concatenate([flat, np.zeros(pad)]).reshape(-1, per_line)
        with open(path, "w") as f:
            for r, row in enumerate(rows):
                vals = row if (r + 1) * per_line <= flat.size else row[:per_line - pad]
                f.write(" ".join("%.10g" % v for v in vals) + "\n")
    dump(os.path.join(data_dir, "spe_phi.dat"), np.asarray(phi).reshape(-1))
    dump(os.path.join(data_dir, "spe_perm.dat"), np.concatenate([np.asarray(a).reshape(-1) for a in (kx_md, ky_md, kz_md)]))


def create_slice(Nx, Ny, Nz, x_shift=0, y_shift=0, z_shift=0, data_dir=None, fields=None, save_dir=None, perm_factor=1.0):
    """create_SPE10_slice.py:10-71.  field[i][j][Nz-1-kk] = line[(i+x_shift) + (j+y_shift)*60 + (kk+z_shift)*220*60]."""
    phi, kx, ky, kz = fields if fields is not None else read_dat(data_dir)

    def cut(a, scale):
        w = a[z_shift:z_shift + Nz, y_shift:y_shift + Ny, x_shift:x_shift + Nx]       # (kk, j, i), top-down
        return np.ascontiguousarray(w[::-1].transpose(2, 1, 0)) * scale               # [i, j, Nz-1-kk]
    out = (cut(phi, 1.0), cut(kx, MD_TO_MM2 * perm_factor), cut(ky, MD_TO_MM2 * perm_factor), cut(kz, MD_TO_MM2 * perm_factor))
    if save_dir is not None:
        _save(save_dir, out, True)
    return out


def create_slice2D(Nx, Ny, x_shift=0, y_shift=0, z_shift=0, data_dir=None, fields=None, save_dir=None, perm_factor=1.0):
    """create_SPE10_slice2D.py:11-60: layer `z_shift` (top-down index), arrays [i, j]."""
    phi, kx, ky, _ = fields if fields is not None else read_dat(data_dir)

    def cut(a, scale):
        return np.ascontiguousarray(a[z_shift, y_shift:y_shift + Ny, x_shift:x_shift + Nx].T) * scale
    out = (cut(phi, 1.0), cut(kx, MD_TO_MM2 * perm_factor), cut(ky, MD_TO_MM2 * perm_factor))
    if save_dir is not None:
        _save(save_dir, out, False)
    return out


def _save(save_dir, arrays, three_d):
    os.makedirs(save_dir, exist_ok=True)
    names = ["slice_phi.npy", "slice_perm_x.npy", "slice_perm_y.npy"] + (["slice_perm_z.npy"] if three_d else [])
    for n, a in zip(names, arrays):
        np.save(os.path.join(save_dir, n), a)
