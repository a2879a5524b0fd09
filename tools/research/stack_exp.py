"""Research harness (CPU): the 2x stacked bench workload as ONE domain; KSP its with an emulated slab cut in the multigrid."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import cport
from thermalporous_b200 import cases as CS, options as O
st = np.load(sys.argv[1]); dt = float(sys.argv[2]) * 86400 if len(sys.argv) > 2 else float(st["dt"])
prm = bench.make_params()
geo = bench.make_geo(prm, bench.NZ, 2, "stack")
base = bench.make_geo(prm, bench.NZ, 1)
base_ent = CS.source_entries(CS.WellCase(prm, base, well_case="default"), prm, base)
ent = [(c + r * base.ncell,) + tuple(rest) for r in range(2) for (c, *rest) in base_ent]
eng = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
eng.set_field(cport.PHI, geo.phi); eng.set_field(cport.KX, geo.K_x); eng.set_field(cport.KY, geo.K_y); eng.set_field(cport.KZ, geo.K_z)
eng.set_sources(ent)
opts, _, _ = O.resolve("pc_cptr", 2); eng.set_solver_opts(**opts)
u = np.concatenate([st["u"], st["u"]], axis=1); uo = np.concatenate([st["uo"], st["uo"]], axis=1)
F, J = eng.assemble(u, uo, dt)
eng.pc_setup(J, u, dt)
t0 = time.time()
x, its, reason, rn = eng.ksp_solve(J, F)
print("cut %s lump %s: its %d reason %d (%.0fs)" % (os.environ.get("TPC_CUT"), os.environ.get("TPC_LUMP"), its, reason, time.time() - t0))
