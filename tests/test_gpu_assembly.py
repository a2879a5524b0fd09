"""GPU parity: assembly (K1/K2) and SpMV (K9) through the C-ABI vs the reference-form golden
vectors and vs the oracle on larger seeded problems.  Tolerance 1e-12 relative (north_star),
measured per equation row / per Jacobian entry class against that row's infinity norm."""
import numpy as np
import pytest

from oracle import tp_oracle as orc
from tests.golden_util import golden_names, load, rel_err_rows
from tests.gpu_util import engine_from_problem, random_problem

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("name", golden_names())
def test_assembly_matches_reference_golden(name):
    meta, pb, z = load(name)
    eng = engine_from_problem(pb)
    F, J = eng.assemble(z["u"], z["u_old"], meta["dt"], jacobian=True)
    assert rel_err_rows(F.cpu().numpy(), z["F"]) < TOL
    assert rel_err_rows(J.cpu().numpy(), z["J"]) < TOL
    F2 = eng.assemble(z["u"], z["u_old"], meta["dt"], jacobian=False)
    # the residual-only kernel is a separately compiled instantiation: same formulas, FMA contraction may differ
    assert rel_err_rows(F2.cpu().numpy(), F.cpu().numpy()) < 1e-14
    eng.close()


@pytest.mark.parametrize("dim,nphase,shape,spread", [
    (3, 2, (9, 11, 37), 1.0), (3, 2, (8, 7, 33), 1e-3), (3, 1, (6, 9, 35), 1e-2),
    (2, 2, (1, 23, 41), 1.0), (2, 1, (1, 17, 33), 1.0), (3, 2, (1, 1, 5), 1.0), (3, 2, (5, 1, 1), 1e-3),
])
def test_assembly_matches_oracle_random(dim, nphase, shape, spread):
    pb, u, uo = random_problem(dim, nphase, shape, seed=sum(shape) + nphase, spread=spread)
    dt = 8640.0
    eng = engine_from_problem(pb)
    F, J = eng.assemble(u, uo, dt)
    Fo = orc.residual(pb, u, uo, dt)
    Jo = orc.jacobian(pb, u, uo, dt)
    assert rel_err_rows(F.cpu().numpy(), Fo) < TOL
    assert rel_err_rows(J.cpu().numpy(), Jo) < TOL
    # SpMV on the GPU Jacobian vs the oracle's stencil product
    x = np.random.default_rng(3).normal(size=(pb.nf, pb.grid.n))
    y = eng.spmv(J, x).cpu().numpy()
    yo = orc.spmv(J.cpu().numpy(), pb.grid, x)
    assert rel_err_rows(y, yo) < 1e-13
    eng.close()


def test_structural_identities_at_scale():
    """size-independent properties on a grid the oracle would not finish quickly:
    uniform state => F == 0; Jacobian row sums of the pressure column vanish for no-gravity
    incompressible limit is not available, so check J*e_const consistency with a finite difference."""
    import torch
    pb, u, uo = random_problem(3, 2, (40, 64, 96), seed=5)
    pb.sources = []
    eng = engine_from_problem(pb)
    n = pb.grid.n
    uni = np.stack([np.full(n, pb.prm.p_ref), np.full(n, pb.prm.T_prod), np.full(n, 0.9)])
    pb_nog = pb
    # with gravity a uniform state is not an equilibrium: compare against gravity-free engine
    from thermalporous_b200.engine import Engine
    eng.close()
    pb.gravity = False
    eng = engine_from_problem(pb)
    F = eng.assemble(uni, uni, 3600.0, jacobian=False)
    # exact up to FMA contraction of rho*S*T - rho_*S_*T_ (accumulation scale ~ V*phi*c_v*rho*T/dt)
    scale = pb.grid.vol * 0.5 * pb.prm.c_v_w * 1e3 * pb.prm.T_prod / 3600.0
    assert float(F.abs().max()) < 1e-13 * scale
    # directional derivative: J(u) d  ~  (F(u + eps d) - F(u - eps d)) / (2 eps)
    F0, J = eng.assemble(u, uo, 3600.0)
    rng = np.random.default_rng(0)
    d = np.stack([rng.normal(size=n) * 1e-3, rng.normal(size=n), rng.normal(size=n) * 1e-3])
    eps = 1e-6
    Fp = eng.assemble(u + eps * d, uo, 3600.0, jacobian=False)
    Fm = eng.assemble(u - eps * d, uo, 3600.0, jacobian=False)
    fd = (Fp - Fm) / (2 * eps)
    Jd = eng.spmv(J, d)
    num = (fd - Jd).abs().amax(dim=1)
    den = Jd.abs().amax(dim=1)
    assert float((num / den).max()) < 1e-5   # limited by upwind switches + fd truncation
    # mass conservation: interior fluxes telescope, sum of the oil row = accumulation only
    Wo = pb.prm.T_prod * (pb.prm.c_v_w * 0.1 + pb.prm.c_v_o * 0.9)
    acc = pb.grid.vol * Wo * pb.phi * (orc.oil_rho(pb.prm, u[0], u[1]) * u[2] - orc.oil_rho(pb.prm, uo[0], uo[1]) * uo[2]) / 3600.0
    assert float(F0[2].sum()) == pytest.approx(acc.sum(), rel=1e-8)
    eng.close()
