"""Research harness (CPU): stationary convergence factor of the V-cycle on the pressure / temperature operators of a saved
state.  usage: mg_exp.py state.npz [dt_days] [key=value,...]..."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.research.save_state import make_engine

def stencil_mv(a, x, nx, ny, nz):
    n = x.size
    y = a[0] * x
    X = x.reshape(nz, ny, nx)
    Y = y.reshape(nz, ny, nx)
    A = a.reshape(7, nz, ny, nx)
    Y[:, :, 1:] += A[1][:, :, 1:] * X[:, :, :-1]
    Y[:, :, :-1] += A[2][:, :, :-1] * X[:, :, 1:]
    Y[:, 1:, :] += A[3][:, 1:, :] * X[:, :-1, :]
    Y[:, :-1, :] += A[4][:, :-1, :] * X[:, 1:, :]
    Y[1:, :, :] += A[5][1:, :, :] * X[:-1, :, :]
    Y[:-1, :, :] += A[6][:-1, :, :] * X[1:, :, :]
    return y

def main():
    st = np.load(sys.argv[1])
    dt = float(st["dt"])
    variants = []
    for a in sys.argv[2:]:
        if "=" not in a:
            dt = float(a) * 86400.0
        else:
            variants.append(a)
    u, uo = st["u"].copy(), st["uo"].copy()
    nz = u.shape[1] // (60 * 220)
    eng, prm, geo = make_engine(nz)
    F, J = eng.assemble(u, uo, dt)
    rng = np.random.default_rng(0)
    sets = [dict()] + [dict(kv.split("=") for kv in v.split(",")) for v in variants]
    for sset in sets:
        kw = {k: (float(v) if ("." in v or "e" in v) else int(v)) for k, v in sset.items()}
        base = dict(mg_pre=2, mg_post=2, mg_cycles=1, mg_semi_theta=0.5, mg_overcorrection=1.0, mg_dd_stop=0.1)
        base.update(kw)
        eng.set_solver_opts(**base)
        eng.pc_setup(J, u, dt)
        for which in (0, 1):
            levs = eng.mg_levels(which)
            a = eng.mg_level_op(which, 0)
            nx, ny, nzz = levs[0][:3]
            b = rng.standard_normal(a.shape[1])
            x = np.zeros_like(b)
            r = b.copy()
            hist = [np.linalg.norm(r)]
            t0 = time.time()
            for it in range(12):
                x += eng.mg_apply(which, r)
                r = b - stencil_mv(a, x, nx, ny, nzz)
                hist.append(np.linalg.norm(r))
            t1 = time.time()
            f = [hist[i + 1] / hist[i] for i in range(len(hist) - 1)]
            print("%-50s %s: levels %2d  factors %s  (%.2fs/cycle) dims %s" % (sset, "pT"[which], len(levs), " ".join("%.2f" % v for v in f[:3] + f[-3:]), (t1 - t0) / 12, [l[:3] for l in levs[:6]]), flush=True)

main()
