"""Research harness (CPU): continue the bench time loop from a saved state for a few steps. usage: loop_exp.py state.npz nsteps"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from tools.research.save_state import make_engine
from thermalporous_b200.model import run_time_loop
st = np.load(sys.argv[1]); nsteps = int(sys.argv[2])
u, uo = st["u"].copy(), st["uo"].copy()
nz = u.shape[1] // (60 * 220)
eng, prm, geo = make_engine(nz)
kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
t0 = time.time()
def newton(a, b, dt):
    s = eng.newton_solve(a, b, dt)
    print("dt %.4g d nits %d lits %d reason %d  t=%.0fs" % (dt / 86400, s.nits, s.lits, s.reason, time.time() - t0), flush=True)
    return s
res = run_time_loop(newton, bench.NpOps(), u, uo, max_steps=nsteps, dt0=float(st["dt"]), **kw)
print("total nits %d lits %d  %.1f lits/nit" % (res.total_nits, res.total_lits, res.total_lits / res.total_nits))
if len(sys.argv) > 3:
    np.savez(sys.argv[3], u=u, uo=uo, dt=res.next_dt)
