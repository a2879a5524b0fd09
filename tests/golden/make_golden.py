#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference form code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It puts `tests/golden/fd_shim` (a minimal DG0 evaluator that quacks like the
slice of Firedrake the reference touches) and /root/reference on sys.path,
imports thermalporous.{physicalparameters,singlephase,twophase,wellcase,
heatercase,wellheatercase,sourceterms,homogeneousgeo,homogeneousboxgeo,
rectanglegeo,boxgeo,thermalmodel}, lets the reference build its own residual
form `model.F`, and evaluates it (and its complex-step derivative) on seeded
states.  Each fixture stores the inputs (grid, fields, state, the reference's
delta fields flattened into a sparse source list) and the outputs (F, J, and
for the time-loop fixtures the converged fields after each reference timestep).

The fixtures travel to the GPU box; this script and the shim do not need to.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "fd_shim"))
sys.path.insert(0, os.environ.get("TPB_REFERENCE", "/root/reference"))

import firedrake as fd  # noqa: E402  (the shim)
from thermalporous.physicalparameters import PhysicalParameters  # noqa: E402
from thermalporous.rectanglegeo import RectangleGeo  # noqa: E402
from thermalporous.boxgeo import BoxGeo  # noqa: E402
from thermalporous.homogeneousgeo import HomogeneousGeo  # noqa: E402
from thermalporous.homogeneousboxgeo import HomogeneousBoxGeo  # noqa: E402
from thermalporous.wellcase import WellCase  # noqa: E402
from thermalporous.heatercase import HeaterCase  # noqa: E402
from thermalporous.wellheatercase import WellHeaterCase  # noqa: E402
from thermalporous.sourceterms import SourceTerms  # noqa: E402
from thermalporous.singlephase import SinglePhase  # noqa: E402
from thermalporous.twophase import TwoPhase  # noqa: E402

PROD, INJ, HEATER = 0, 1, 2
TMP = "/tmp/tpb_golden_results.txt"


def sp(max_it):
    """thermalmodel.py:37-40 deletes these three keys when verbosity is False."""
    return {"snes_max_it": max_it, "snes_monitor": None, "snes_converged_reason": None,
            "ksp_converged_reason": None}


def fresh_params(**kw):
    class P(PhysicalParameters):
        pass
    p = P()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def hetero_fields(shape, rng, zero_perm=False):
    """small SPE10-flavoured random fields (values in the reference's units: mm^2)."""
    logk = rng.normal(1.0, 1.3, size=shape)
    Kx = 10.0 ** logk * 9.869233e-10
    Ky = Kx * 10.0 ** rng.normal(0.0, 0.2, size=shape)
    Kz = Kx * 10.0 ** rng.normal(-1.0, 0.5, size=shape)
    phi = np.clip(0.2 + 0.08 * (logk - 1.0), 0.0, 0.5)
    phi[rng.random(shape) < 0.05] = 0.0
    phi = phi + 1e-10                       # SPE10model3D.py:28
    if zero_perm:
        m = rng.random(shape) < 0.08
        Kx[m] = 0.0
        Ky[m] = 0.0
        Kz[m] = 0.0
    return phi, Kx, Ky, Kz


class HeteroGeo2D(RectangleGeo):
    """RectangleGeo (reference) with in-memory fields instead of data/slice_*.npy."""

    def __init__(self, nx, ny, params, fields, dx=6.096, dy=3.048, name="SPE10"):
        self._fields = fields
        self.name = name
        RectangleGeo.__init__(self, nx, ny, params, Length=nx * dx, Length_y=ny * dy)

    def generate_geo_fields(self):
        phi, Kx, Ky, _ = self._fields
        self.phi = fd.Function(self.V)
        self.K_x = fd.Function(self.V)
        self.K_y = fd.Function(self.V)
        self.phi.arr[0] = phi
        self.K_x.arr[0] = Kx
        self.K_y.arr[0] = Ky
        p = self.params
        self.kT = fd.project(self.phi * p.ko + (1 - self.phi) * p.kr, self.V)  # SPE10model.py:64


class HeteroGeo3D(BoxGeo):
    def __init__(self, nx, ny, nz, params, fields, dx=6.096, dy=3.048, dz=0.6096, name="SPE10 3D"):
        self._fields = fields
        self.name = name
        BoxGeo.__init__(self, nx, ny, nz, params, Length=nx * dx, Length_y=ny * dy, Length_z=nz * dz)

    def generate_geo_fields(self):
        phi, Kx, Ky, Kz = self._fields
        self.phi = fd.Function(self.V)
        self.K_x = fd.Function(self.V)
        self.K_y = fd.Function(self.V)
        self.K_z = fd.Function(self.V)
        self.phi.arr[0] = phi
        self.K_x.arr[0] = Kx
        self.K_y.arr[0] = Ky
        self.K_z.arr[0] = Kz
        p = self.params
        self.kT = fd.project(self.phi * p.ko + (1 - self.phi) * p.kr, self.V)  # SPE10model3D.py:72


def as_array(x, shape):
    if isinstance(x, fd.Function):
        return np.real(x.arr[0]).reshape(-1).copy()
    v = fd._evaluate(fd._as_node(x), {})
    return (np.zeros(shape) + np.real(v.d)).reshape(-1)


def sources_of(case, params, geo, constant_rate):
    """flatten the reference's delta Functions into (cell, kind, weight, bhp, max_rate, const)."""
    mesh = geo.mesh
    V = mesh.vol()
    out = []

    def add(delta, kind, bhp, max_rate):
        d = np.real(delta.arr[0]).reshape(-1)
        for c in np.nonzero(d)[0]:
            out.append((int(c), kind, float(d[c] * V), float(bhp), float(max_rate), int(constant_rate)))

    if case.name.startswith("Sources"):
        add(case.deltas_prod, PROD, params.p_prod, -params.prod_rate)
        add(case.deltas_inj, INJ, params.p_inj, params.inj_rate)
        add(case.deltas_heaters, HEATER, 0.0, 0.0)
    for w in getattr(case, "prod_wells", []):
        add(w["delta"], PROD, w["bhp"], w["max_rate"])
    for w in getattr(case, "inj_wells", []):
        add(w["delta"], INJ, w["bhp"], w["max_rate"])
    for h in getattr(case, "heaters", []):
        add(h["delta"], HEATER, 0.0, 0.0)
    return np.array(out, dtype=np.float64).reshape(-1, 6)


def eval_F(model):
    return np.real(fd.assemble(model.F))


def eval_J(model, h=1e-30):
    """complex-step derivative of the reference form, block-stencil layout."""
    u = model.u
    mesh = u.V.mesh()
    nf = u.arr.shape[0]
    N = mesh.nx * mesh.ny * mesh.nz
    offs = [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0)]
    if mesh.dim == 3:
        offs += [(0, 0, -1), (0, 0, 1)]
    kk, jj, ii = np.meshgrid(np.arange(mesh.nz), np.arange(mesh.ny), np.arange(mesh.nx), indexing="ij")
    colour = ((ii + 2 * jj + 3 * kk) % 7).ravel()
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    base = np.real(u.arr).copy()
    J = np.zeros((len(offs), nf, nf, N))
    for col in range(7):
        mask = colour == col
        if not mask.any():
            continue
        for c in range(nf):
            up = base.astype(complex)
            up[c].reshape(-1)[mask] += 1j * h
            u.arr = up
            dF = np.imag(fd.assemble(model.F)) / h
            for s, (di, dj, dk) in enumerate(offs):
                ni, nj, nk = ii + di, jj + dj, kk + dk
                ok = (ni >= 0) & (ni < mesh.nx) & (nj >= 0) & (nj < mesh.ny) & (nk >= 0) & (nk < mesh.nz)
                nb = np.where(ok, ni + mesh.nx * (nj + mesh.ny * nk), 0)
                sel = ok & mask[nb]
                J[s, :, c, sel] = dF[:, sel].T
    u.arr = base
    return J


def random_state(rng, params, nphase, shape, spread=1.0):
    p = params.p_ref + spread * rng.uniform(-5.0, 5.0, size=shape)
    T = rng.uniform(288.7, 422.0, size=shape)
    out = [p, T]
    if nphase == 2:
        out.append(rng.uniform(0.05, 0.95, size=shape))
    return np.stack(out)


def save(name, model, geo, case, params, nphase, constant_rate, extra=None):
    mesh = geo.mesh
    shape = mesh.shape
    meta = dict(name=name, nphase=nphase, dim=mesh.dim, nx=mesh.nx, ny=mesh.ny, nz=mesh.nz,
                dx=mesh.dx, dy=mesh.dy, dz=(mesh.dz or 1.0), dt=float(model.dt.values()[0]),
                case=case.name, constant_rate=bool(constant_rate),
                params={k: float(getattr(params, k)) for k in
                        ("ko", "kw", "kr", "c_v_w", "c_v_o", "c_r", "rho_r", "p_inj", "p_prod", "T_inj",
                         "T_prod", "API", "p_ref", "g", "S_o", "U", "rate", "well_radius")})
    arrays = dict(
        meta=np.array(json.dumps(meta)),
        u=np.real(model.u.arr).reshape(model.u.arr.shape[0], -1),
        u_old=np.real(model.u_.arr).reshape(model.u_.arr.shape[0], -1),
        phi=as_array(geo.phi, shape), Kx=as_array(geo.K_x, shape), Ky=as_array(geo.K_y, shape),
        Kz=as_array(geo.K_z, shape) if mesh.dim == 3 else np.zeros(0),
        kT=as_array(geo.kT, shape),
        sources=sources_of(case, params, geo, constant_rate),
        F=eval_F(model), J=eval_J(model))
    if extra:
        arrays.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, "F", arrays["F"].shape, "|F|max", np.abs(arrays["F"]).max(),
          "nsrc", len(arrays["sources"]))


def set_state(model, rng, params, nphase, spread=1.0):
    shape = model.u.V.mesh().shape
    model.u.arr = random_state(rng, params, nphase, shape, spread)
    model.u_.arr = random_state(rng, params, nphase, shape, spread)


def main():
    rng = np.random.default_rng(20261018)

    # G1: single-phase 2-D homogeneous, test0 wells, constant rate (tests/test_homo_wells.py)
    prm = fresh_params(rate=1e-6, T_prod=320.0)
    geo = HomogeneousGeo(10, 8, prm, 20.0, 20.0)
    case = WellCase(prm, geo, well_case="test0", constant_rate=True)
    m = SinglePhase(geo, case, prm, end=2.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                    solver_parameters=sp(15))
    set_state(m, rng, prm, 1)
    save("g1_sp2d_homo_const", m, geo, case, prm, 1, True)

    # G2: single-phase 2-D heterogeneous, Peaceman wells (tests/test_60x120_wells.py shape)
    prm = fresh_params()
    fields = hetero_fields((1, 10, 12), rng)
    geo = HeteroGeo2D(12, 10, prm, fields)
    pts_p = [[2.3 * 6.096, 4.6 * 3.048]]
    pts_i = [[9.4 * 6.096, 7.2 * 3.048]]
    case = WellCase(prm, geo, prod_points=pts_p, inj_points=pts_i)
    m = SinglePhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                    solver_parameters=sp(15))
    set_state(m, rng, prm, 1)
    # make the wells active in both regimes: producer below cap, injector capped
    save("g2_sp2d_hetero_peaceman", m, geo, case, prm, 1, False)

    # G3: two-phase 2-D heterogeneous, Peaceman wells (tests_twophase/test_60x120_wells_default.py shape)
    prm = fresh_params(rate=2e-4, S_o=0.9)
    fields = hetero_fields((1, 10, 12), rng)
    geo = HeteroGeo2D(12, 10, prm, fields)
    case = WellCase(prm, geo, prod_points=pts_p, inj_points=pts_i)
    m = TwoPhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    set_state(m, rng, prm, 2)
    save("g3_tp2d_hetero_peaceman", m, geo, case, prm, 2, False)

    # G3b: same but pressures near bhp so the un-capped Peaceman branch is exercised
    prm = fresh_params(rate=1.0, S_o=0.9)
    geo = HeteroGeo2D(12, 10, prm, fields)
    case = WellCase(prm, geo, prod_points=pts_p, inj_points=pts_i)
    m = TwoPhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    set_state(m, rng, prm, 2)
    save("g3b_tp2d_hetero_peaceman_uncapped", m, geo, case, prm, 2, False)

    # G4: single-phase 3-D heterogeneous with gravity, wells + heaters
    prm = fresh_params(rate=1.0)
    fields = hetero_fields((4, 5, 6), rng)
    geo = HeteroGeo3D(6, 5, 4, prm, fields)
    pp = [[1.5 * 6.096, 2.5 * 3.048, 0.5 * 0.6096]]
    ip = [[4.5 * 6.096, 1.5 * 3.048, 3.5 * 0.6096]]
    case = WellHeaterCase(prm, geo, prod_points=pp, inj_points=ip)
    m = SinglePhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                    solver_parameters=sp(15))
    set_state(m, rng, prm, 1, spread=0.01)   # small dp so gravity decides some upwind directions
    save("g4_sp3d_hetero_wellheater", m, geo, case, prm, 1, False)

    # G5: two-phase 3-D heterogeneous with gravity, wells + heaters
    prm = fresh_params(rate=1.0, S_o=0.9)
    fields = hetero_fields((4, 5, 6), rng)
    geo = HeteroGeo3D(6, 5, 4, prm, fields)
    case = WellHeaterCase(prm, geo, prod_points=pp, inj_points=ip)
    m = TwoPhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    set_state(m, rng, prm, 2, spread=0.001)
    save("g5_tp3d_hetero_wellheater", m, geo, case, prm, 2, False)

    # G5b: same geometry, large pressure differences (pressure-driven upwinding)
    m2 = TwoPhase(geo, case, prm, end=1.0, maxdt=0.25, small_dt_start=False, filename=TMP, verbosity=False,
                  solver_parameters=sp(25))
    set_state(m2, rng, prm, 2, spread=1.0)
    save("g5b_tp3d_hetero_wellheater_dp", m2, geo, case, prm, 2, False)

    # G6: two-phase 3-D, SourceTerms (summed deltas, 'Sources' branch twophase.py:362-385)
    prm = fresh_params(rate=2e-4, S_o=0.8)
    fields = hetero_fields((4, 5, 6), rng)
    geo = HeteroGeo3D(6, 5, 4, prm, fields)
    case = SourceTerms(prm, geo, prod_points=pp + pp, inj_points=ip, heater_points=pp + ip)
    m = TwoPhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    set_state(m, rng, prm, 2, spread=0.01)
    save("g6_tp3d_sources", m, geo, case, prm, 2, False)

    # G7: two-phase 3-D homogeneous cube + heaters incl. zero-perm cells
    #     (tests_twophase/test3D_homo_heater.py shape; fine grid so the 0.1 m bump hits cell centres)
    prm = fresh_params(rate=1e-7, T_inj=373.15, S_o=0.9)
    geo = HomogeneousBoxGeo(8, 8, 8, prm, Length=1.2, Length_y=1.2, Length_z=4.0)
    L = 1.2
    hp = [[L / 4 + 0.01, L / 2 + 0.02, 0.8], [3 * L / 4, L / 2, 3.2], [L / 4 + 0.01, L / 2 + 0.02, 0.8]]
    case = HeaterCase(prm, geo, heater_points=hp)
    m = TwoPhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    set_state(m, rng, prm, 2, spread=0.001)
    save("g7_tp3d_homo_heater", m, geo, case, prm, 2, False)

    # G8: single-phase 3-D with zero-permeability cells, heaters only (harmonic-mean guard)
    prm = fresh_params()
    fields = hetero_fields((3, 4, 5), rng, zero_perm=True)
    geo = HeteroGeo3D(5, 4, 3, prm, fields)
    case = HeaterCase(prm, geo, heater_points=[[2.5 * 6.096, 1.5 * 3.048, 1.5 * 0.6096]])
    m = SinglePhase(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                    solver_parameters=sp(15))
    set_state(m, rng, prm, 1, spread=0.01)
    save("g8_sp3d_zeroperm_heater", m, geo, case, prm, 1, False)

    # ------------------------------------------------------------------ time-loop fixtures
    # the reference's own ThermalModel.solve() (thermalmodel.py:82-412) drives the shim's Newton
    fd.NonlinearVariationalSolver.rtol = 1e-12
    fd.NonlinearVariationalSolver.stol = 1e-13

    def run_loop(name, model, geo, case, prm, nphase, constant_rate):
        hist = []
        orig = model.solver.solve

        def solve_and_record():
            orig()
            hist.append((float(model.dt.values()[0]), model.solver.snes.getIterationNumber(),
                         np.real(model.u.arr).reshape(model.u.arr.shape[0], -1).copy()))
        model.solver.solve = solve_and_record
        model.solve()
        u_final = np.real(model.u.arr).reshape(model.u.arr.shape[0], -1).copy()
        # residual/Jacobian at the final state are stored too (u_ == u after the loop)
        extra = dict(loop_dts=np.array([h[0] for h in hist]), loop_nits=np.array([h[1] for h in hist]),
                     loop_u=np.stack([h[2] for h in hist]), u_final=u_final,
                     u_init=np.real(model.initial_condition.arr).reshape(u_final.shape[0], -1))
        save(name, model, geo, case, prm, nphase, constant_rate, extra)

    # L1: C1 shape (tests/test_homo_wells.py): 2 steps of dt=1 day, constant-rate wells
    prm = fresh_params(rate=1e-6, T_prod=320.0)
    geo = HomogeneousGeo(10, 10, prm, 20.0, 20.0)
    case = WellCase(prm, geo, well_case="test0", constant_rate=True)
    m = SinglePhase(geo, case, prm, end=2.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                    solver_parameters=sp(15))
    run_loop("l1_sp2d_homo_loop", m, geo, case, prm, 1, True)

    # L2: C3 shape: two-phase heterogeneous slice with Peaceman wells, small_dt_start + SPE10 dt heuristic
    prm = fresh_params(rate=2e-4, S_o=0.9)
    fields = hetero_fields((1, 10, 12), np.random.default_rng(7))
    geo = HeteroGeo2D(12, 10, prm, fields)
    case = WellCase(prm, geo, prod_points=pts_p, inj_points=pts_i)
    m = TwoPhase(geo, case, prm, end=0.02, maxdt=0.01, small_dt_start=True, dt_init_fact=2 ** (-3),
                 filename=TMP, verbosity=False, solver_parameters=sp(25))
    run_loop("l2_tp2d_hetero_loop", m, geo, case, prm, 2, False)

    # L3: C4 shape: 3-D two-phase homogeneous heaters, 3 steps
    prm = fresh_params(rate=1e-7, T_inj=373.15, S_o=0.9)
    geo = HomogeneousBoxGeo(6, 6, 6, prm, Length=50.0, Length_y=50.0, Length_z=50.0)
    hp = [[50 / 8, 25.0, 10.0], [25.0, 25.0, 10.0], [25.0, 12.5, 40.0]]
    case = HeaterCase(prm, geo, heater_points=hp)
    m = TwoPhase(geo, case, prm, end=3.0, maxdt=1.0, small_dt_start=False, filename=TMP, verbosity=False,
                 solver_parameters=sp(25))
    run_loop("l3_tp3d_heater_loop", m, geo, case, prm, 2, False)


if __name__ == "__main__":
    main()
