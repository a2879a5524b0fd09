"""1-D slab partition of a structured grid along its slowest axis (z in 3-D, y in 2-D) - the layout
libtpb200 expects for multi-GPU runs (include/tpb200.h: tpb_grid.has_lo/has_hi).  Stands in for the DMPlex
partition Firedrake makes over mesh.comm (singlephase.py:13, twophase.py:14); one rank = one GPU."""
from __future__ import annotations

import numpy as np


def slab_range(nl, world, rank):
    """planes [k0, k1) of the slab axis owned by `rank` (as even as possible, remainder to the low ranks)."""
    base, rem = divmod(int(nl), int(world))
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


class Slab:
    """what one rank holds of a geo: plane range, cell range, local grid sizes."""

    def __init__(self, geo, world, rank):
        self.dim = geo.dim
        self.nx, self.ny = geo.Nx, geo.Ny
        self.nl = geo.Nz if geo.dim == 3 else geo.Ny
        self.np = geo.Nx * geo.Ny if geo.dim == 3 else geo.Nx
        self.world, self.rank = world, rank
        self.k0, self.k1 = slab_range(self.nl, world, rank)
        if self.k1 <= self.k0:
            raise ValueError("rank %d of %d owns no plane of a %d-plane grid" % (rank, world, self.nl))
        self.c0, self.c1 = self.k0 * self.np, self.k1 * self.np
        self.has_lo, self.has_hi = rank > 0, rank < world - 1

    @property
    def ncell(self):
        return self.c1 - self.c0

    def local_dims(self):
        """(nx, ny, nz) of the local slab."""
        if self.dim == 3:
            return self.nx, self.ny, self.k1 - self.k0
        return self.nx, self.k1 - self.k0, 1

    def take(self, field):
        """owned part of a global cell field (or of every row of a (nf, ncell) state)."""
        a = np.asarray(field)
        return np.ascontiguousarray(a[..., self.c0:self.c1])

    def localize_sources(self, entries):
        """global (cell, kind, weight, bhp, max_rate, const) records -> the ones in this slab, local cell index."""
        return [(c - self.c0,) + tuple(r) for (c, *r) in entries if self.c0 <= c < self.c1]
