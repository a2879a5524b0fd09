"""debug: where does the GPU assembly differ from the CPU restatement on the bench workload mid-run state?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import cport, tp_oracle as orc
from thermalporous_b200 import _lib as L, cases as CS, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.model import run_time_loop, _TorchOps
prm = bench.make_params(); geo = bench.make_geo(prm)
ent = CS.source_entries(CS.WellCase(prm, geo, well_case="default"), prm, geo)
eng = Engine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
for fid, arr in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
    eng.set_field(fid, arr)
eng.set_sources(ent)
opts, _, _ = O.resolve(bench.PC, 2); eng.set_solver_opts(**opts)
n = eng.n
u = eng.tensor(np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)])); uo = u.clone()
kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
rw = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), _TorchOps(eng), u, uo, max_steps=9, **kw)
dt = rw.next_dt
un = u.clone(); eng.set_solver_opts(snes_max_it=1); eng.newton_solve(un, uo, dt)
cpu = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
cpu.set_field(cport.PHI, geo.phi); cpu.set_field(cport.KX, geo.K_x); cpu.set_field(cport.KY, geo.K_y); cpu.set_field(cport.KZ, geo.K_z)
cpu.set_sources(ent)
F, J = eng.assemble(un, uo, dt)
Fg, Jg = F.cpu().numpy(), J.cpu().numpy()
uh, uoh = un.cpu().numpy(), uo.cpu().numpy()
Fc, Jc = cpu.assemble(uh, uoh, dt)
g = orc.Grid(geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 3)
pp = orc.Params(**{k: getattr(prm, k) for k in orc.Params().__dict__ if hasattr(prm, k)})
srcs = [orc.Source(int(s[0]), int(s[1]), float(s[2]), float(s[3]), float(s[4]), bool(s[5])) for s in ent]
pb = orc.Problem(grid=g, nphase=2, prm=pp, phi=geo.phi, Kx=geo.K_x, Ky=geo.K_y, Kz=geo.K_z, kT=None, sources=srcs)
Fo = orc.residual(pb, uh, uoh, dt)
print("dt", dt, "S range", uh[2].min(), uh[2].max(), "T range", uh[1].min(), uh[1].max(), "sources", [(s[0], s[1]) for s in ent])
for name, A, B in (("gpu-vs-cport", Fg, Fc), ("gpu-vs-oracle", Fg, Fo), ("cport-vs-oracle", Fc, Fo)):
    d = np.abs(A - B) / np.abs(B).max(axis=1, keepdims=True)
    for f in range(3):
        c = int(d[f].argmax())
        i, j, k = c % geo.Nx, (c // geo.Nx) % geo.Ny, c // (geo.Nx * geo.Ny)
        print(name, "field", f, "max rel", d[f, c], "cell", c, (i, j, k), "A", A[f, c], "B", B[f, c], "u", uh[:, c], "phi", geo.phi[c], "K", geo.K_x[c], geo.K_z[c], "count>1e-12", int((d[f] > 1e-12).sum()))
dJ = np.abs(Jg - Jc).reshape(63, -1) / np.abs(Jc).reshape(63, -1).max(axis=1, keepdims=True).clip(1e-300)
print("J gpu-vs-cport max", dJ.max(), "slot", np.unravel_index(dJ.argmax(), dJ.shape))
