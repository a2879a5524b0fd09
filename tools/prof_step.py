#!/usr/bin/env python
"""Driver for ncu: the bench workload's time loop for W un-profiled steps, then ONE profiled time step
(cudaProfilerStart/Stop around it; run ncu with --profile-from-start off).   python tools/prof_step.py [W=12] [opt=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from thermalporous_b200 import _lib as L, cases as CS, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.model import run_time_loop, _TorchOps

W = int(sys.argv[1]) if len(sys.argv) > 1 else 12
prm = bench.make_params()
geo = bench.make_geo(prm)
eng = Engine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
for fid, a in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
    eng.set_field(fid, a)
eng.set_sources(CS.source_entries(CS.WellCase(prm, geo, well_case="default"), prm, geo))
opts, _, _ = O.resolve(bench.PC, 2)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
eng.set_solver_opts(**opts)
n = eng.n
u = eng.tensor(np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)]))
uo = u.clone()
kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
newton = lambda a, b, dt: eng.newton_solve(a, b, dt)
rw = run_time_loop(newton, _TorchOps(eng), u, uo, max_steps=W, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
l0 = eng.launch_count()
res = run_time_loop(newton, _TorchOps(eng), u, uo, max_steps=1, dt0=rw.next_dt, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = res.stats[0]
print("profiled step: dt %.4g d nits %d lits %d launches %d  assemble %.1f pc_setup %.1f ksp %.1f ms" % (
    res.dt_vec[0] / 86400, st.nits, st.lits, eng.launch_count() - l0, st.t_assemble_ms, st.t_pcsetup_ms, st.t_ksp_ms))
