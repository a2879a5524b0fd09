"""Research harness (CPU): run the bench workload with oracle/cport and save (u, u_old, dt) at chosen steps, so that
preconditioner variants can be compared on a late-step linear system without re-running the time loop."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import cport
from thermalporous_b200 import cases as CS, options as O
from thermalporous_b200.model import run_time_loop

def make_engine(nz=bench.NZ, pc=bench.PC):
    prm = bench.make_params()
    geo = bench.make_geo(prm, nz)
    case = CS.WellCase(prm, geo, well_case="default")
    eng = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
    eng.set_field(cport.PHI, geo.phi); eng.set_field(cport.KX, geo.K_x)
    eng.set_field(cport.KY, geo.K_y); eng.set_field(cport.KZ, geo.K_z)
    eng.set_sources(CS.source_entries(case, prm, geo))
    opts, _, _ = O.resolve(pc, 2)
    eng.set_solver_opts(**opts)
    eng.set_num_threads(len(os.sched_getaffinity(0)))
    return eng, prm, geo

if __name__ == "__main__":
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else bench.NZ
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/tpstate_%d" % nz
    eng, prm, geo = make_engine(nz)
    n = geo.ncell
    u = np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)])
    uo = u.copy()
    kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
    saved = {}
    step = [0]
    t0 = time.time()
    def newton(a, b, dt):
        step[0] += 1
        np.savez(out + "_s%d.npz" % step[0], u=a, uo=b, dt=dt)
        st = eng.newton_solve(a, b, dt)
        print("step %d dt %.4g d nits %d lits %d reason %d  t=%.0fs" % (step[0], dt / 86400, st.nits, st.lits, st.reason, time.time() - t0), flush=True)
        return st
    run_time_loop(newton, bench.NpOps(), u, uo, max_steps=nsteps, **kw)
