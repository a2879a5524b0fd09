#!/usr/bin/env python
"""Multi-GPU consistency check (run under torchrun, one rank per GPU): a slab-partitioned assembly, SpMV
and Newton solve must reproduce the single-domain CPU restatement.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import cport
from thermalporous_b200 import _lib as L, cases as CS, geo as G, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.partition import Slab
from thermalporous_b200.physicalparameters import PhysicalParameters

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prm = PhysicalParameters(); prm.S_o = 0.9; prm.rate = 2e-4
geo = G.SPE10Model3D(12, 16, 12, prm, fields=G.spe10_synthetic(12, 16, 12, seed=5))
case = CS.WellCase(prm, geo, well_case="default")
ent = CS.source_entries(case, prm, geo)
slab = Slab(geo, world, rank)
nx, ny, nz = slab.local_dims()
eng = Engine(3, nx, ny, nz, geo.Dx, geo.Dy, geo.Dz, 2, prm, device=local, has_lo=slab.has_lo, has_hi=slab.has_hi)
for fid, a in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
    eng.set_field(fid, slab.take(a))
eng.set_sources(slab.localize_sources(ent))
uid = [eng.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(uid[0], rank, world)
eng.exchange_static()
rng = np.random.default_rng(7)
n = geo.ncell
u = np.stack([prm.p_ref + rng.uniform(-0.05, 0.05, n), rng.uniform(288.7, 300.0, n), rng.uniform(0.8, 0.95, n)])
uo = np.stack([prm.p_ref + rng.uniform(-0.05, 0.05, n), rng.uniform(288.7, 300.0, n), rng.uniform(0.8, 0.95, n)])
x = rng.normal(size=(3, n))
cpu = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
for fid, a in ((cport.PHI, geo.phi), (cport.KX, geo.K_x), (cport.KY, geo.K_y), (cport.KZ, geo.K_z)):
    cpu.set_field(fid, a)
cpu.set_sources(ent)
Fc, Jc = cpu.assemble(u, uo, 864.0)
yc = cpu.spmv(Jc, x)
F, J = eng.assemble(slab.take(u), slab.take(uo), 864.0)
y = eng.spmv(J, slab.take(x))
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
eF = max(rel(F.cpu().numpy()[f], slab.take(Fc)[f]) for f in range(3))
eJ = rel(J.cpu().numpy(), Jc[..., slab.c0:slab.c1])
ey = max(rel(y.cpu().numpy()[f], slab.take(yc)[f]) for f in range(3))
# Newton from the uniform state, one small step
opts, _, _ = O.resolve("pc_cptr", 2)
# mg_dd_stop acts on slabs as on a single domain unless TPB_MG_DD_DIST=0: the single-domain CPU run follows the slabs
DD = {} if os.environ.get("TPB_MG_DD_DIST", "1") != "0" else {"mg_dd_stop": 0.0}
opts.update(snes_rtol=1e-11, snes_stol=1e-13, ksp_rtol=1e-6, snes_max_it=40, **DD)
eng.set_solver_opts(**opts); cpu.set_solver_opts(**opts)
u0 = np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, 0.9)])
ug = eng.tensor(slab.take(u0)); st = eng.newton_solve(ug, ug.clone(), 20.0)
uc = u0.copy(); sc = cpu.newton_solve(uc, u0.copy(), 20.0)
eu = max(rel(ug.cpu().numpy()[f], slab.take(uc)[f]) for f in range(3))
smin, smax = eng.field_minmax(ug, 2)
ok = eF < 1e-12 and eJ < 1e-12 and ey < 1e-12 and st.reason > 0 and eu < 1e-8 and abs(smax - uc[2].max()) < 1e-8
# the multi-rank preconditioner (z-lines and block ILU cut at the slab faces, no gathered coarse level for a
# line-smoothed hierarchy) must stay a usable one: a bounded increase of the Krylov count over the single domain
ok = ok and st.lits <= 1.6 * sc.lits + 12 * (world - 1)
# the exchanges must have gone the way the environment asked for (TPB_P2P unset = mailboxes)
want_peer = int(os.environ.get("TPB_P2P", "15"))
ok = ok and eng.peer_mode() == want_peer
print("rank %d/%d: peer mode %d (asked %d) F %.1e J %.1e spmv %.1e | newton nits %d (cpu %d) lits %d (cpu %d) reason %d fields %.1e | %s"
      % (rank, world, eng.peer_mode(), want_peer, eF, eJ, ey, st.nits, sc.nits, st.lits, sc.lits, st.reason, eu, "OK" if ok else "FAIL"), flush=True)
eng.close()
# the mirrored model class under torchrun: each rank owns a slab; same time loop as the single-domain CPU run
from thermalporous_b200.model import TwoPhase, run_time_loop
model = TwoPhase(geo, case, prm, end=0.002, maxdt=0.001, small_dt_start=True, dt_init_fact=2 ** -2,
                 solver_parameters="pc_cptr", verbosity=False, device=local)
model.engine.set_solver_opts(snes_rtol=1e-11, snes_stol=1e-13, ksp_rtol=1e-10, snes_max_it=40)   # both sides converged tightly
res = model.solve()
class NpOps:
    def copy(self, d, s): d[...] = s
    def minmax(self, u, f): return float(u[f].min()), float(u[f].max())
    def clip(self, u, f, lo, hi): np.clip(u[f], lo, hi, out=u[f])
# a Newton count at the tolerance edge steers the SPE10 dt heuristic, so the CPU run follows the dt sequence the
# slab run took; both sides are converged far below the default tolerances -> fields at the north_star's 1e-8
o2, _, _ = O.resolve("pc_cptr", 2)
o2.update(snes_rtol=1e-11, snes_stol=1e-13, ksp_rtol=1e-10, snes_max_it=40, **DD)
cpu.set_solver_opts(**o2)
uc2 = u0.copy()
class RC: pass
rc = RC(); rc.nits_vec = []
for dt in res.dt_vec:
    st2 = cpu.newton_solve(uc2, uc2.copy(), dt)
    rc.nits_vec.append(st2.nits)
    np.clip(uc2[2], 0.0, 1.0, out=uc2[2])
em = max(rel(model.fields()[f], slab.take(uc2)[f]) for f in range(3))
ok2 = res.failed_solves == 0 and abs(res.t - 0.002 * 86400.0) < 1e-6 and em < 1e-8
print("rank %d/%d: model.solve() %d steps nits %s (cpu %s) fields %.1e | %s" % (rank, world, len(res.dt_vec), res.nits_vec,
      rc.nits_vec, em, "OK" if ok2 else "FAIL"), flush=True)
ok = ok and ok2
t = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
