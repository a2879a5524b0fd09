#!/usr/bin/env python
"""Diagnose a stalled Krylov solve: runs the bench time loop on one GPU (--mult stacked copies), and when a Newton
solve fails re-plays that step by hand (assemble / pc_setup / ksp_solve) printing operator diagnostics.
   python tools/dbg_stall.py --mult 4 --steps 9"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from thermalporous_b200 import _lib as L, cases as CS, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.model import run_time_loop, _TorchOps

ap = argparse.ArgumentParser()
ap.add_argument("--mult", type=int, default=4)
ap.add_argument("--scale", default="stack")
ap.add_argument("--steps", type=int, default=9)
ap.add_argument("--nz", type=int, default=B.NZ)
args = ap.parse_args()

prm = B.make_params()
geo = B.make_geo(prm, args.nz, args.mult, args.scale)
case = CS.WellCase(prm, geo, well_case="default")
ent = CS.source_entries(case, prm, geo)
eng = Engine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
for fid, arr in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
    eng.set_field(fid, arr)
eng.set_sources(ent)
print("sources:", [(int(e[0]), int(e[1])) for e in ent], "nz", geo.Nz, flush=True)
opts, _, desc = O.resolve(B.PC, 2)
eng.set_solver_opts(**opts)
n = eng.n
u = eng.tensor(np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)]))
uo = u.clone()
kw = dict(end=1e9, maxdt=B.MAXDT, small_dt_start=True, dt_init_fact=B.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)


def stats(name, t):
    t = t.double()
    print("   %-10s min %.6e max %.6e" % (name, float(t.min()), float(t.max())), flush=True)


def op_diag(name, a):
    d = a[0]
    off = a[1:]
    possum = torch.clamp(off, min=0).sum(0)
    abssum = off.abs().sum(0)
    print("   %s: diag min %.3e max %.3e  #diag<=0 %d  #rows with positive off-diag %d (max pos/diag %.3e)  "
          "max sum|off|/diag %.4f  #rows sum|off|>diag %d" % (name, float(d.min()), float(d.max()), int((d <= 0).sum()),
          int((possum > 0).sum()), float((possum / d.abs().clamp(min=1e-300)).max()), float((abssum / d.abs().clamp(min=1e-300)).max()),
          int((abssum > d.abs() * (1 + 1e-12)).sum())), flush=True)


def replay(u0, uold, dt):
    print("=== replay of the failed step, dt %.3f s" % dt, flush=True)
    uu = u0.clone()
    for it in range(6):
        F, J = eng.assemble(uu, uold, dt)
        fn = float(F.norm())
        print(" newton %d |F| %.6e" % (it, fn), flush=True)
        stats("p", uu[0]); stats("T", uu[1]); stats("S", uu[2])
        print("   #S<0 %d  #S>1 %d" % (int((uu[2] < 0).sum()), int((uu[2] > 1).sum())), flush=True)
        eng.set_solver_opts(**opts)
        eng.pc_setup(J, uu, dt)
        op_diag("App", eng.mg_level_op(0, 0))
        op_diag("AT ", eng.mg_level_op(1, 0))
        print("   levels p:", eng.mg_levels(0), flush=True)
        x, its, reason, rn = eng.ksp_solve(J, F)
        print("   ksp its %d reason %d rnorm %.3e" % (its, reason, rn), flush=True)
        if reason < 0:
            for label, over in (("verbose history", dict(verbose=2, ksp_max_it=40)),
                                ("stage2 only (ILU)", dict(stage1=0, ksp_max_it=400, ksp_restart=200)),
                                ("cptr, 2 V-cycles", dict(mg_cycles=2)),
                                ("cptr QI decoupling", dict(decoup=1)),
                                ("cpr", dict(stage1=1)),
                                ("cptr schur a11", dict(schur_pre=1)),
                                ("cptr bjacobi stage 2", dict(stage2=2)),
                                ("cptr theta 0", dict(mg_semi_theta=0.0)),
                                ("cptr 4+4 sweeps", dict(mg_pre=4, mg_post=4))):
                o2 = dict(opts); o2.update(over)
                eng.set_solver_opts(**o2)
                eng.pc_setup(J, uu, dt)
                x2, its2, r2, rn2 = eng.ksp_solve(J, F)
                print("   [%s] its %d reason %d rnorm %.3e" % (label, its2, r2, rn2), flush=True)
            # where does the residual of the stalled solve live?
            eng.set_solver_opts(**opts)
            eng.pc_setup(J, uu, dt)
            x, its, reason, rn = eng.ksp_solve(J, F)
            r = F - eng.spmv(J, x)
            for f, nm in enumerate("pTS"):
                rf = r[f].abs()
                c = int(rf.argmax())
                k, rem = divmod(c, geo.Nx * geo.Ny); j, i = divmod(rem, geo.Nx)
                print("   residual field %s: norm %.3e  max %.3e at cell (%d,%d,%d)  phi %.3e Kx %.3e Kz %.3e S %.4f p %.4f T %.3f"
                      % (nm, float(r[f].norm()), float(rf.max()), i, j, k, float(geo.phi.reshape(-1)[c]) if hasattr(geo.phi, "reshape") else -1,
                         float(np.asarray(geo.K_x).reshape(-1)[c]), float(np.asarray(geo.K_z).reshape(-1)[c]), float(uu[2, c]), float(uu[0, c]), float(uu[1, c])), flush=True)
            return
        uu -= x
    print("   replay converged?!", flush=True)


state = {}


def newton(a, b, dt):
    ua, ub = a.clone(), b.clone()
    st = eng.newton_solve(a, b, dt)
    print("step dt %.4f s: nits %d lits %d reason %d" % (dt, st.nits, st.lits, st.reason), flush=True)
    if st.reason < 0 and "done" not in state:
        state["done"] = True
        replay(ua, ub, dt)
        eng.set_solver_opts(**opts)
    return st


res = run_time_loop(newton, _TorchOps(eng), u, uo, max_steps=args.steps, **kw)
print("nits", res.nits_vec, "lits", res.lits_vec, "failed", res.failed)
