"""Research harness (CPU): BASELINE configs C1-C4 (small) through oracle/cport; prints Newton / Krylov counts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cport
from thermalporous_b200 import cases as CS, geo as G, options as O
from thermalporous_b200.model import run_time_loop
from thermalporous_b200.physicalparameters import PhysicalParameters
from tools.run_c4 import heater_points
import bench

def params(**kw):
    class P(PhysicalParameters):
        pass
    p = P()
    for k, v in kw.items():
        setattr(p, k, v)
    return p

def run(name, geo, case, prm, nphase, pc, **loop):
    eng = cport.CpuEngine(geo.dim, geo.Nx, geo.Ny, getattr(geo, "Nz", 1), geo.Dx, geo.Dy, getattr(geo, "Dz", 1.0), nphase, prm)
    eng.set_field(cport.PHI, geo.phi); eng.set_field(cport.KX, geo.K_x); eng.set_field(cport.KY, geo.K_y)
    if geo.dim == 3: eng.set_field(cport.KZ, geo.K_z)
    if nphase == 1: eng.set_field(cport.KT, geo.kT)
    eng.set_sources(CS.source_entries(case, prm, geo))
    opts, _, _ = O.resolve(pc, nphase)
    eng.set_solver_opts(**opts)
    n = geo.ncell
    u = np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod)] + ([np.full(n, prm.S_o)] if nphase == 2 else []))
    t0 = time.time()
    res = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), bench.NpOps(), u, u.copy(), two_phase=nphase == 2, i_S=2,
                        spe10=geo.name.startswith("SPE10"), **loop)
    print("%-28s nits %s lits %s failed %d  %.1fs" % (name + " " + str(pc), res.nits_vec, res.lits_vec, res.failed_solves, time.time() - t0), flush=True)

which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
if "c1" in which:
    prm = params(rate=1e-6, T_prod=320.0)
    geo = G.HomogeneousGeo(100, 100, prm, 20.0, 20.0)
    case = CS.WellCase(prm, geo, well_case="test0", constant_rate=True)
    for pc in ("pc_fieldsplit_cd", "pc_cpr"):
        run("C1", geo, case, prm, 1, pc, end=2.0, maxdt=1.0, small_dt_start=False, dt_init_fact=2 ** -10)
if "c2" in which:
    prm = params()
    geo = G.SPE10Model(60, 120, prm, fields=G.spe10_synthetic_layer(60, 120))
    case = CS.WellCase(prm, geo, well_case="SPE10_60x120")
    run("C2", geo, case, prm, 1, "pc_cpr_QI", end=0.02, maxdt=0.01, small_dt_start=True, dt_init_fact=2 ** -5)
if "c3" in which:
    prm = params(rate=2e-4, S_o=0.9)
    geo = G.SPE10Model(60, 120, prm, fields=G.spe10_synthetic_layer(60, 120))
    case = CS.WellCase(prm, geo, well_case="SPE10_60x120")
    for pc in ("pc_cptr", "pc_cpr_TI"):
        run("C3", geo, case, prm, 2, pc, end=0.004, maxdt=0.002, small_dt_start=True, dt_init_fact=2 ** -4)
if "c4" in which:
    N = int(os.environ.get("C4N", "40"))
    prm = params(rate=1e-7, T_inj=373.15, S_o=0.9)
    geo = G.HomogeneousBoxGeo(N, N, N, prm, 50.0, 50.0, 50.0)
    case = CS.HeaterCase(prm, geo, heater_points=heater_points(50.0))
    run("C4 N=%d" % N, geo, case, prm, 2, "pc_cptr", end=3.0, maxdt=1.0, small_dt_start=False, dt_init_fact=2 ** -10)
