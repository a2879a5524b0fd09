"""Research harness (CPU): one linear solve of the bench workload's Newton system at a saved state, with solver-option
variants.  usage: ksp_exp.py state.npz [dt_days] key=value ..."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.research.save_state import make_engine

def main():
    st = np.load(sys.argv[1])
    args = sys.argv[2:]
    dt = float(st["dt"])
    variants = []
    for a in args:
        if "=" not in a:
            dt = float(a) * 86400.0
        else:
            variants.append(a)
    u, uo = st["u"].copy(), st["uo"].copy()
    nz = u.shape[1] // (60 * 220)
    eng, prm, geo = make_engine(nz)
    # one Newton iteration first so that the system is not the trivial first one (u == u_old)
    F, J = eng.assemble(u, uo, dt)
    print("dt %.4g d  |F| %.4e" % (dt / 86400, np.linalg.norm(F)))
    sets = [dict()] + [dict(kv.split("=") for kv in v.split(",")) for v in variants]
    for sset in sets:
        kw = {k: (float(v) if ("." in v or "e" in v) else int(v)) for k, v in sset.items()}
        base = dict(mg_pre=2, mg_post=2, mg_cycles=1, mg_semi_theta=0.5, mg_overcorrection=1.0, mg_dd_stop=0.1, verbose=0)
        base.update(kw)
        eng.set_solver_opts(**base)
        t0 = time.time()
        eng.pc_setup(J, u, dt)
        t1 = time.time()
        x, its, reason, rn = eng.ksp_solve(J, F)
        t2 = time.time()
        print("%-60s its %3d reason %d  setup %.2fs solve %.2fs  levels p=%d T=%d" % (sset, its, reason, t1 - t0, t2 - t1, len(eng.mg_levels(0)), len(eng.mg_levels(1))), flush=True)

main()
