#!/usr/bin/env python
"""Short driver for ncu captures: SPE10-sized two-phase engine, one assembly, one PC set-up, two PC applies,
two SpMVs (random heterogeneous state so that both upwind branches are taken)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thermalporous_b200 import _lib as L, options as O
from thermalporous_b200.engine import Engine
from thermalporous_b200.physicalparameters import PhysicalParameters
from thermalporous_b200 import geo as G

prm = PhysicalParameters(); prm.S_o = 0.9; prm.rate = 2e-4
geo = G.SPE10Model3D(60, 220, 85, prm, fields=G.spe10_synthetic(60, 220, 85))
eng = Engine(3, 60, 220, 85, geo.Dx, geo.Dy, geo.Dz, 2, prm)
for fid, a in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
    eng.set_field(fid, a)
opts, _, _ = O.resolve("pc_cptr", 2)
eng.set_solver_opts(**opts)
n = eng.n
gen = torch.Generator(device="cuda").manual_seed(1234)
rnd = lambda lo, hi: torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * (hi - lo) + lo
u = torch.stack([prm.p_ref + rnd(-0.05, 0.05), rnd(288.7, 300.0), rnd(0.85, 0.95)])
uo = torch.stack([prm.p_ref + rnd(-0.05, 0.05), rnd(288.7, 300.0), rnd(0.85, 0.95)])
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
for _ in range(2):
    F, J = eng.assemble(u, uo, 864.0)
eng.pc_setup(J, u, 864.0)
x = torch.randn(3, n, generator=gen, device="cuda", dtype=torch.float64)
if mode in ("all", "spmv"):
    for _ in range(2):
        z = eng.spmv(J, x)
if mode in ("all", "pc"):
    # component by component, so that a short `--launch-count` reaches every kernel family: ILU(0) half sweeps,
    # one pressure V-cycle (smoother passes of every level, restrictions, the single-CTA tail), one multi-dot
    r = eng.stage2_apply(x)
    v = eng.mg_apply(0, x[0].contiguous())
    eng.set_solver_opts(ksp_max_it=4)
    eng.pc_setup(J, u, 864.0)   # (set_solver_opts invalidates the set-up)
    d = eng.ksp_solve(J, x)     # four Arnoldi steps: multi-dot / multi-axpy
print("ok", mode)
