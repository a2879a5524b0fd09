"""Field output for `save=True` - stands in for the reference's `File("results/pressure.pvd").write(pvec)` calls
(thermalmodel.py:113-133, 304-320): one ParaView collection (.pvd) per field, one VTK ImageData file (.vti, cell
data, base64 inline binary) per saved step.  Under torchrun every rank writes its own slab (files carry the rank
suffix; the extents place the slabs side by side in the global grid).  Host-side I/O, outside the hot path.

Also the dump format of the reference's debugging hooks `utils.ExportJacobian` / `utils.ExportResidual`
(utils.py:27-43: PETSc ASCII_MATLAB viewers writing `matrix.txt` / `rhs.txt`): `export_jacobian`, `export_residual`
write it from the block-stencil Jacobian, `load_matlab_matrix` / `load_matlab_vector` read it back (or read a dump a
Firedrake install of the reference made) as SciPy/NumPy objects in the reference's field-major dof order.
"""
from __future__ import annotations

import base64
import os
import re

import numpy as np


class PvdWriter:
    def __init__(self, directory, names, geo, slab, suffix=""):
        self.dir, self.names, self.geo, self.slab, self.suffix = directory, list(names), geo, slab, suffix
        os.makedirs(directory, exist_ok=True)
        self.steps = []          # (time, index)

    def _extent(self):
        g, s = self.geo, self.slab
        if g.dim == 3:
            return (0, g.Nx, 0, g.Ny, s.k0, s.k1), (g.Dx, g.Dy, g.Dz)
        return (0, g.Nx, s.k0, s.k1, 0, 1), (g.Dx, g.Dy, 1.0)

    def write(self, t, fields):
        k = len(self.steps)
        ext, sp = self._extent()
        es = " ".join(str(v) for v in ext)
        for name, arr in zip(self.names, fields):
            raw = np.ascontiguousarray(arr, dtype="<f8").tobytes()
            payload = base64.b64encode(np.array([len(raw)], dtype="<u4").tobytes() + raw).decode()
            with open(os.path.join(self.dir, "%s%s_%d.vti" % (name, self.suffix, k)), "w") as f:
                f.write('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" '
                        'header_type="UInt32">\n<ImageData WholeExtent="%s" Origin="0 0 0" Spacing="%r %r %r">\n'
                        '<Piece Extent="%s">\n<CellData Scalars="%s">\n<DataArray type="Float64" Name="%s" '
                        'format="binary">\n%s\n</DataArray>\n</CellData>\n</Piece>\n</ImageData>\n</VTKFile>\n'
                        % (es, sp[0], sp[1], sp[2], es, name, name, payload))
        self.steps.append((float(t), k))
        self._write_pvd()

    def _write_pvd(self):
        for name in self.names:
            with open(os.path.join(self.dir, "%s%s.pvd" % (name, self.suffix)), "w") as f:
                f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
                for t, k in self.steps:
                    f.write('<DataSet timestep="%r" part="0" file="%s%s_%d.vti"/>\n' % (t, name, self.suffix, k))
                f.write('</Collection>\n</VTKFile>\n')

    def close(self):
        self._write_pvd()


def read_vti(path):
    """(extent, spacing, {name: array}) of a file PvdWriter wrote (tests; quick looks without ParaView)."""
    txt = open(path).read()
    ext = tuple(int(v) for v in re.search(r'Piece Extent="([^"]+)"', txt).group(1).split())
    sp = tuple(float(v) for v in re.search(r'Spacing="([^"]+)"', txt).group(1).split())
    out = {}
    for m in re.finditer(r'<DataArray type="Float64" Name="([^"]+)" format="binary">\s*([^<]+?)\s*</DataArray>', txt):
        blob = base64.b64decode(m.group(2))
        nbytes = int(np.frombuffer(blob[:4], dtype="<u4")[0])
        out[m.group(1)] = np.frombuffer(blob[4:4 + nbytes], dtype="<f8").copy()
    return ext, sp, out


# ---- utils.ExportJacobian / ExportResidual (utils.py:27-43): PETSc ASCII_MATLAB dumps ------------------------------
_SLOT_OFF = ((0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))


def jacobian_coo(J, nx, ny, nz):
    """(rows, cols, vals) of a block-stencil Jacobian J[s, r, c, cell] in the reference's field-major dof order
    (dof = field * ncell + cell); entries that leave the grid are dropped."""
    J = np.asarray(J)
    ns, nf, _, n = J.shape
    cell = np.arange(n)
    i, j, k = cell % nx, (cell // nx) % ny, cell // (nx * ny)
    rows, cols, vals = [], [], []
    for s in range(ns):
        di, dj, dk = _SLOT_OFF[s]
        ok = (i + di >= 0) & (i + di < nx) & (j + dj >= 0) & (j + dj < ny) & (k + dk >= 0) & (k + dk < nz)
        nb = cell + di + nx * (dj + ny * dk)
        for r in range(nf):
            for c in range(nf):
                rows.append(r * n + cell[ok])
                cols.append(c * n + nb[ok])
                vals.append(J[s, r, c][ok])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def export_jacobian(J, nx, ny, nz, path="matrix.txt", name="Mat_0"):
    """matrix.txt as PETSc's MatView(ASCII_MATLAB) of an aij matrix writes it: 1-based (row, col, value) triplets
    in `zzz = [...]` and a spconvert line."""
    r, c, v = jacobian_coo(J, nx, ny, nz)
    order = np.lexsort((c, r))
    nd = np.asarray(J).shape[1] * np.asarray(J).shape[3]
    with open(path, "w") as f:
        f.write("%% Size = %d %d \n%% Nonzeros = %d \nzzz = zeros(%d,3);\nzzz = [\n" % (nd, nd, len(v), len(v)))
        for q in order:
            f.write("%d %d  %.16e\n" % (r[q] + 1, c[q] + 1, v[q]))
        f.write("];\n %s = spconvert(zzz);\n" % name)


def export_residual(F, path="rhs.txt", name="Vec_0"):
    """rhs.txt as PETSc's VecView(ASCII_MATLAB) writes it."""
    with open(path, "w") as f:
        f.write("%s = [\n" % name)
        for v in np.asarray(F).reshape(-1):
            f.write("%.16e\n" % v)
        f.write("];\n")


def _numeric_rows(path):
    rows = []
    for line in open(path):
        t = line.strip()
        if not t or t[0] in "%]" or "=" in t:
            continue
        rows.append([float(x) for x in t.split()])
    return rows


def load_matlab_matrix(path, shape=None):
    """scipy.sparse.csr_matrix of a PETSc ASCII_MATLAB matrix dump (the reference's matrix.txt)."""
    import scipy.sparse as sp
    a = np.array(_numeric_rows(path))
    if shape is None:
        m = re.search(r"Size = (\d+) (\d+)", open(path).read())
        shape = (int(m.group(1)), int(m.group(2))) if m else (int(a[:, 0].max()), int(a[:, 1].max()))
    return sp.csr_matrix((a[:, 2], (a[:, 0].astype(int) - 1, a[:, 1].astype(int) - 1)), shape=shape)


def load_matlab_vector(path):
    return np.array([r[0] for r in _numeric_rows(path)])
