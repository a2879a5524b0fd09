#!/usr/bin/env python
"""Slab run vs single-domain run ON THE BENCH WORKLOAD (run under torchrun, one rank per GPU): the N-times stacked
60x220x85 reservoir of `bench.py --gpus N`, a few steps of the time loop with both sides converged far below the default
tolerances.  Rank 0 also solves the whole stacked grid on its own GPU as ONE domain (no slabs, no exchanges).
Checks: same dt sequence, converged fields within 1e-8 relative, no Newton solve of the slab run needing more than
10 iterations or 1.5x the single-domain Krylov iterations (the r1 scaling run had a 16-iteration / 547-Krylov step).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_bench_check.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from thermalporous_b200 import _lib as L, cases as CS, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.model import run_time_loop, _TorchOps
from thermalporous_b200.partition import Slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prm = bench.make_params()
geo = bench.make_geo(prm, bench.NZ, world, "stack")
base = bench.make_geo(prm, bench.NZ, 1)
base_ent = CS.source_entries(CS.WellCase(prm, base, well_case="default"), prm, base)
all_ent = [(c + r * base.ncell,) + tuple(rest) for r in range(world) for (c, *rest) in base_ent]
TIGHT = dict(snes_rtol=1e-11, snes_stol=1e-13, ksp_rtol=1e-10, snes_max_it=40)
kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)


def build(slab, ent, comm):
    nx, ny, nz = slab.local_dims() if slab else (geo.Nx, geo.Ny, geo.Nz)
    eng = Engine(3, nx, ny, nz, geo.Dx, geo.Dy, geo.Dz, 2, prm, device=local, has_lo=slab.has_lo if slab else False,
                 has_hi=slab.has_hi if slab else False)
    take = slab.take if slab else (lambda a: a)
    for fid, a in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
        eng.set_field(fid, take(a))
    eng.set_sources(ent)
    if comm:
        uid = [eng.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
        eng.exchange_static()
    opts, _, _ = O.resolve(bench.PC, 2)
    opts.update(TIGHT)
    eng.set_solver_opts(**opts)
    return eng


def run(eng, dts=None):
    n = eng.n
    u = eng.tensor(np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)]))
    uo = u.clone()
    if dts is None:
        res = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), _TorchOps(eng), u, uo, max_steps=steps, **kw)
        return u, res.dt_vec, res.nits_vec, res.lits_vec, res.failed_solves
    nits, lits = [], []
    for dt in dts:      # the single-domain run follows the slab run's dt sequence
        st = eng.newton_solve(u, uo, dt)
        eng.clip_field(u, 2, 0.0, 1.0)
        uo.copy_(u)
        nits.append(st.nits)
        lits.append(st.lits)
    return u, dts, nits, lits, 0


slab = Slab(geo, world, rank)
eng = build(slab, slab.localize_sources(all_ent), True)
u, dts, nits, lits, failed = run(eng)
us = u.cpu().numpy()
eng.close()
ok = failed == 0 and max(nits) <= 10
msg = "rank %d/%d slabs: dt %s nits %s lits %s failed %d" % (rank, world, ["%.3g" % (d / 86400) for d in dts], nits, lits, failed)
if rank == 0:
    one = build(None, all_ent, False)
    u1, _, nits1, lits1, _ = run(one, dts)
    ref = u1.cpu().numpy()
    one.close()
    holder = [ref, nits1, lits1]
else:
    holder = [None, None, None]
dist.broadcast_object_list(holder, src=0)
ref, nits1, lits1 = holder
mine = slab.take(ref)
err = max(float(np.abs(us[f] - mine[f]).max() / np.abs(mine[f]).max()) for f in range(3))
ok = ok and err < 1e-8 and all(a <= 1.5 * b + 3 for a, b in zip(lits, lits1)) and all(abs(a - b) <= 1 for a, b in zip(nits, nits1))
print(msg + " | single domain nits %s lits %s | fields %.1e | %s" % (nits1, lits1, err, "OK" if ok else "FAIL"), flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
