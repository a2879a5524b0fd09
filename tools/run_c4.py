#!/usr/bin/env python
"""BASELINE config 4: tests_twophase/test3D_homo_heater.py scaled up - N^3 homogeneous cube, L = 50 m, two-phase,
42 heaters (rate 1e-7, T_inj 373.15, S_o 0.9), 5 steps of dt = 1 day, small_dt_start False, `pc_cptr`.
   python tools/run_c4.py [N=216] [steps=5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thermalporous_b200.physicalparameters import PhysicalParameters
from thermalporous_b200.geo import HomogeneousBoxGeo
from thermalporous_b200.cases import HeaterCase
from thermalporous_b200.model import TwoPhase


def heater_points(L):
    xs = [L / 8, L / 4, 3 * L / 8, L / 2, 5 * L / 8, 3 * L / 4, 7 * L / 8]
    def rows(z):   # test3D_homo_heater.py:82-83 (the last point of the third row repeats (7L/8, L/4))
        return [[x, L / 2, z] for x in xs] + [[x, L / 4, z] for x in xs] + [[x, 3 * L / 4, z] for x in xs[:-1]] + [[7 * L / 8, L / 4, z]]
    return rows(0.2 * L) + rows(0.8 * L)


def build(N, steps=5, pc="pc_cptr", verbosity=False):
    prm = PhysicalParameters()
    prm.rate, prm.T_inj, prm.S_o = 1e-7, 373.15, 0.9
    geo = HomogeneousBoxGeo(N, N, N, prm, 50.0, 50.0, 50.0)
    case = HeaterCase(prm, geo, heater_points=heater_points(50.0))
    return TwoPhase(geo, case, prm, end=float(steps), maxdt=1.0, small_dt_start=False, solver_parameters=pc, verbosity=verbosity)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 216
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    t0 = time.time()
    model = build(N, steps)
    t1 = time.time()
    res = model.solve()
    sec = sum(res.timings)
    n = model.geo.ncell
    p, T, S = model.fields()
    print("C4 N=%d (%d cells, %d dofs): setup %.1f s; %d steps, nits %s, lits %s, failed %d; solve %.2f s -> %.1f Mcell-Newton-iters/s; "
          "T max %.2f K, S in [%.4f, %.4f]" % (N, n, 3 * n, t1 - t0, len(res.dt_vec), res.nits_vec, res.lits_vec, res.failed_solves,
                                               sec, n * res.total_nits / sec / 1e6, T.max(), S.min(), S.max()))
