"""Structured-grid geometry and coefficient fields - the host-side mirror of the reference's
geo classes (rectanglegeo.py, boxgeo.py, homogeneousgeo.py, homogeneousboxgeo.py, SPE10model.py,
SPE10model3D.py).  Same constructor arguments and attribute names (Nx, Ny, Nz, dim, Length*, Dx, Dy,
Dz, phi, K_x, K_y, K_z, kT, name); fields are NumPy arrays of length Ncell in the C-ABI cell order
c = i + Nx*(j + Ny*k) (x fastest, z up) instead of Firedrake Functions.
"""
from __future__ import annotations

import os

import numpy as np

MD_TO_MM2 = 9.869233e-10          # data/create_SPE10_slice.py:42
SPE10_D = (6.096, 3.048, 0.6096)  # SPE10model3D.py:11-13


class _Geo:
    dim = 2

    @property
    def shape(self):
        """(Nz, Ny, Nx): array shape of a field reshaped from the flat cell order."""
        return (getattr(self, "Nz", 1), self.Ny, self.Nx)

    @property
    def ncell(self):
        return self.Nx * self.Ny * getattr(self, "Nz", 1)

    @property
    def cell_volume(self):
        return self.Dx * self.Dy * (self.Dz if self.dim == 3 else 1.0)

    def _full(self, v):
        return np.full(self.ncell, float(v))

    def cell_centres(self):
        """(ncell, dim) coordinates of the cell centres in cell order."""
        k, j, i = np.meshgrid(np.arange(getattr(self, "Nz", 1)), np.arange(self.Ny), np.arange(self.Nx),
                              indexing="ij")
        cols = [(i.ravel() + 0.5) * self.Dx, (j.ravel() + 0.5) * self.Dy]
        if self.dim == 3:
            cols.append((k.ravel() + 0.5) * self.Dz)
        return np.stack(cols, axis=1)

    def _finish(self):
        # rectanglegeo.py:21-27 / boxgeo.py:20-29: isotropic fall-back
        if not hasattr(self, "K_x"):
            self.K_x = self.K
            self.K_y = self.K
            if self.dim == 3:
                self.K_z = self.K


class RectangleGeo(_Geo):
    """rectanglegeo.py:4-34 (quadrilateral RectangleMesh, DQ0)."""

    def __init__(self, Nx, Ny, params, Length=365.76, Length_y=365.76, mg=None):
        if mg:
            raise NotImplementedError("mesh hierarchies exist only for the FAS/PatchPC research branch (out of scope)")
        self.Nx, self.Ny, self.dim, self.params = int(Nx), int(Ny), 2, params
        self.Length, self.Length_y = Length, Length_y
        self.Dx, self.Dy = Length / Nx, Length_y / Ny
        self.gravity2D = False
        self.generate_geo_fields()
        self._finish()


class BoxGeo(_Geo):
    """boxgeo.py:3-44 (extruded quadrilateral mesh, z up)."""

    def __init__(self, Nx, Ny, Nz, params, Length=365.76, Length_y=365.76, Length_z=1.8288, mg=False):
        if mg:
            raise NotImplementedError("mesh hierarchies exist only for the FAS/PatchPC research branch (out of scope)")
        self.Nx, self.Ny, self.Nz, self.dim, self.params = int(Nx), int(Ny), int(Nz), 3, params
        self.Length, self.Length_y, self.Length_z = Length, Length_y, Length_z
        self.Dx, self.Dy, self.Dz = Length / Nx, Length_y / Ny, Length_z / Nz
        self.generate_geo_fields()
        self._finish()


class HomogeneousGeo(RectangleGeo):
    """homogeneousgeo.py:4-20."""

    def __init__(self, Nx, Ny, params, Length, Length_y, mg=None):
        self.geotype = "Homogeneous"
        RectangleGeo.__init__(self, Nx, Ny, params, Length, Length_y, mg)
        self.name = self.geotype + " " + str(self.Nx) + "X" + str(self.Ny) + " grid"

    def generate_geo_fields(self):
        self.phi = self._full(0.2)
        self.K = self._full(3e-7)  # mm^2
        self.kT = self.phi * self.params.ko + (1 - self.phi) * self.params.kr


class HomogeneousBoxGeo(BoxGeo):
    """homogeneousboxgeo.py:4-19."""

    def __init__(self, Nx, Ny, Nz, params, Length, Length_y, Length_z, mg=False):
        self.geotype = "Homogeneous"
        BoxGeo.__init__(self, Nx, Ny, Nz, params, Length, Length_y, Length_z, mg)
        self.name = self.geotype + " " + str(self.Nx) + "X" + str(self.Ny) + "X" + str(self.Nz) + " grid"

    def generate_geo_fields(self):
        self.phi = self._full(0.2)
        self.K = self._full(3e-7)
        self.kT = self.phi * self.params.ko + (1 - self.phi) * self.params.kr


def _from_ijk(arr, geo):
    """SPE10model3D.py:30-68: the slice arrays are indexed [i, j(, k)]; cell order is x fastest."""
    a = np.asarray(arr, dtype=np.float64)
    if geo.dim == 2:
        assert a.shape[:2] == (geo.Nx, geo.Ny), (a.shape, geo.Nx, geo.Ny)
        return np.ascontiguousarray(a[:, :].T).reshape(-1)
    assert a.shape == (geo.Nx, geo.Ny, geo.Nz), (a.shape, geo.Nx, geo.Ny, geo.Nz)
    return np.ascontiguousarray(a.transpose(2, 1, 0)).reshape(-1)


class _SPE10Fields:
    def _load(self, fields, data_dir):
        if fields is None:
            data_dir = data_dir or os.path.join(os.path.dirname(__file__), "..", "data")
            names = ["slice_phi.npy", "slice_perm_x.npy", "slice_perm_y.npy"] + \
                    (["slice_perm_z.npy"] if self.dim == 3 else [])
            try:
                fields = [np.load(os.path.join(data_dir, n)) for n in names]
            except FileNotFoundError as e:
                raise FileNotFoundError(
                    "SPE10 slice files are not distributed with the reference (.MISSING_LARGE_BLOBS); pass "
                    "fields=spe10_synthetic(...) or a data_dir holding slice_*.npy") from e
        return fields

    def generate_geo_fields(self):
        f = self._fields
        self.phi = _from_ijk(f[0], self) + 1e-10          # SPE10model.py:34, SPE10model3D.py:28
        self.K_x = _from_ijk(f[1], self)
        self.K_y = _from_ijk(f[2], self)
        if self.dim == 3:
            self.K_z = _from_ijk(f[3], self)
        p = self.params
        self.kT = self.phi * p.ko + (1 - self.phi) * p.kr  # SPE10model.py:64, SPE10model3D.py:72


class SPE10Model(_SPE10Fields, RectangleGeo):
    """SPE10model.py:6-64.  `fields` = (phi, Kx, Ky) arrays indexed [i, j] (the slice_*.npy layout)."""

    def __init__(self, Nx, Ny, params, save=False, plane="xy", fields=None, data_dir=None):
        self.geotype = "SPE10"
        self.name = self.geotype
        Dx, Dy = {"xy": (6.096, 3.048), "xz": (6.096, 0.6096), "yz": (3.048, 0.6096)}[plane]
        self.dim = 2
        self._fields = self._load(fields, data_dir)
        RectangleGeo.__init__(self, Nx, Ny, params, Length=Nx * Dx, Length_y=Ny * Dy)


class SPE10Model3D(_SPE10Fields, BoxGeo):
    """SPE10model3D.py:6-72.  `fields` = (phi, Kx, Ky, Kz) arrays indexed [i, j, k], z up.
    `refine_z` repeats every layer that many times with Dz/refine_z (BASELINE config 5, weak scaling)."""

    def __init__(self, Nx, Ny, Nz, params, save=False, fields=None, data_dir=None, refine_z=1):
        self.geotype = "SPE10 " + str(Nx) + "X" + str(Ny) + "X" + str(Nz)
        self.name = self.geotype
        self.dim = 3
        Dx, Dy, Dz = SPE10_D
        fields = self._load(fields, data_dir)
        if refine_z > 1:
            fields = [np.repeat(np.asarray(a), refine_z, axis=2) for a in fields]
            Nz = Nz * refine_z
            Dz = Dz / refine_z
        self._fields = fields
        BoxGeo.__init__(self, Nx, Ny, Nz, params, Length=Nx * Dx, Length_y=Ny * Dy, Length_z=Nz * Dz)


def spe10_synthetic(Nx=60, Ny=220, Nz=85, seed=10):
    """Seeded SPE10-shaped fields in the slice_*.npy layout: arrays [i, j, k], z up,
    K in mm^2, phi without the +1e-10.  The real spe_perm.dat / spe_phi.dat are not distributed with the
    reference; this reproduces their statistics: log-normal permeability over ~8 decades, a smooth
    'Tarbert' top (35/85 of the layers) over a channelised 'Upper Ness' bottom, Kz/Kx = 0.3 | 1e-3,
    porosity correlated with log K and 2.5 % zero-porosity cells.  Returns (phi, Kx, Ky, Kz)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    shape = (Nx, Ny, Nz)
    ntop = int(round(Nz * 35.0 / 85.0))

    def grf(sig):
        f = gaussian_filter(rng.standard_normal(shape), sigma=sig, mode="wrap")
        return (f - f.mean()) / f.std()

    g = grf((8.0 / 2.355 * 2, 16.0 / 2.355 * 2, 2.0 / 2.355 * 2))
    logk = 1.0 + 1.3 * g                                    # log10 Kx [mD], top-down layer index here
    # bottom layers: sinuous channels along y, +2.5 decades inside
    chan = grf((3.0, 30.0, 1.0)) > 0.9
    layer = np.arange(Nz)[None, None, :]
    bottom = layer >= ntop                                   # top-down: first ntop layers are Tarbert
    logk = np.where(bottom & chan, logk + 2.5, np.where(bottom, logk - 0.5, logk))
    kx_md = np.clip(10.0 ** logk, 6.65e-4, 2e4)
    kz_md = np.where(bottom, 1e-3, 0.3) * kx_md
    phi = np.clip(0.2 + 0.08 * (np.log10(kx_md) - 1.0), 0.0, 0.5)
    phi[rng.random(shape) < 0.025] = 0.0
    flip = lambda a: np.ascontiguousarray(a[:, :, ::-1])     # create_SPE10_slice.py:31: array[i,j,Nz-1-k]
    Kx = flip(kx_md) * MD_TO_MM2
    return flip(phi), Kx, Kx.copy(), flip(kz_md) * MD_TO_MM2


def spe10_synthetic_layer(Nx=60, Ny=120, seed=10, layer=0):
    """2-D slice (phi, Kx, Ky) indexed [i, j] for the 60x120 configs (tests/test_60x120_wells.py)."""
    phi, Kx, Ky, _ = spe10_synthetic(Nx, Ny, 4, seed)
    return phi[:, :, layer], Kx[:, :, layer], Ky[:, :, layer]
