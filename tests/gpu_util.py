"""Build a thermalporous_b200 Engine from an oracle Problem (test helper)."""
import numpy as np

from oracle import tp_oracle as orc


def engine_from_problem(pb, device=0):
    from thermalporous_b200.engine import Engine
    from thermalporous_b200 import _lib as L
    g = pb.grid
    eng = Engine(g.dim, g.nx, g.ny, g.nz, g.dx, g.dy, g.dz, pb.nphase, pb.prm, device=device, gravity=pb.gravity)
    eng.set_field(L.TPB_PHI, pb.phi)
    eng.set_field(L.TPB_KX, pb.Kx)
    eng.set_field(L.TPB_KY, pb.Ky)
    if g.dim == 3:
        eng.set_field(L.TPB_KZ, pb.Kz)
    if pb.nphase == 1:
        eng.set_field(L.TPB_KT, pb.kT)
    eng.set_sources([(s.cell, s.kind, s.weight, s.bhp, s.max_rate, s.const_rate) for s in pb.sources])
    return eng


def random_problem(dim, nphase, shape, seed=0, nsrc=4, spread=1.0):
    """seeded heterogeneous problem + state (SURVEY.md 8d micro-benchmark states)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    g = orc.Grid(nx, ny, nz, 6.096, 3.048, 0.6096 if dim == 3 else 1.0, dim)
    n = g.n
    logk = rng.normal(1.0, 1.3, n)
    Kx = 10.0 ** logk * 9.869233e-10
    Ky = Kx * 10.0 ** rng.normal(0, 0.2, n)
    Kz = Kx * 10.0 ** rng.normal(-1, 0.5, n)
    phi = np.clip(0.2 + 0.08 * (logk - 1.0), 0.0, 0.5) + 1e-10
    prm = orc.Params(S_o=0.9, rate=1.0)
    kT = phi * prm.ko + (1 - phi) * prm.kr
    srcs = []
    cells = rng.choice(n, size=min(nsrc, n), replace=False)
    for q, c in enumerate(cells):
        kind = q % 3
        srcs.append(orc.Source(int(c), kind, float(rng.uniform(0.3, 1.0)),
                               prm.p_prod if kind == 0 else prm.p_inj,
                               -prm.rate if kind == 0 else prm.rate, False))
    pb = orc.Problem(g, nphase, prm, phi, Kx, Ky, Kz if dim == 3 else None, kT, srcs)
    nf = pb.nf

    def state():
        p = prm.p_ref + spread * rng.uniform(-5, 5, n)
        T = rng.uniform(288.7, 422.0, n)
        rows = [p, T]
        if nf == 3:
            rows.append(rng.uniform(0.05, 0.95, n))
        return np.stack(rows)
    return pb, state(), state()
