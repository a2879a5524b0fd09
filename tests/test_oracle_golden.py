"""Pin the NumPy oracle against vectors produced by the reference's own form code
(singlephase.py:60-273, twophase.py:67-411 executed through tests/golden/fd_shim)."""
import numpy as np
import pytest

from oracle import tp_oracle as orc
from tests.golden_util import golden_names, load, rel_err_rows

TOL = 1e-12  # north_star: assembled residual and Jacobian within 1e-12 relative in fp64


@pytest.mark.parametrize("name", golden_names())
def test_residual_matches_reference_form(name):
    meta, pb, z = load(name)
    F = orc.residual(pb, z["u"], z["u_old"], meta["dt"])
    assert F.shape == z["F"].shape
    assert rel_err_rows(F, z["F"]) < TOL


@pytest.mark.parametrize("name", golden_names())
def test_jacobian_matches_reference_form(name):
    meta, pb, z = load(name)
    J = orc.jacobian(pb, z["u"], z["u_old"], meta["dt"])
    assert J.shape == z["J"].shape
    assert rel_err_rows(J, z["J"]) < TOL


def test_known_property_values():
    """SURVEY.md appendix 9 (evaluated from physicalparameters.py:37-90, API = 10)."""
    p = orc.Params()
    assert orc.oil_rho(p, 41.369, 288.706) == pytest.approx(1021.9337041262304, rel=1e-14)
    assert orc.oil_mu(p, 288.706) == pytest.approx(115.31170207724243, rel=1e-13)
    assert orc.water_rho(p, 41.369, 288.706) == pytest.approx(1011.4771001495221, rel=1e-14)
    assert orc.water_mu(p, 288.706) == pytest.approx(1.4575065381664668e-3, rel=1e-14)
    assert orc.oil_mu(p, 422.039) == pytest.approx(6.592876975968961e-3, rel=1e-13)
    assert orc.water_rho(p, 41.369, 373.15) == pytest.approx(970.0002877954852, rel=1e-14)


def test_peaceman_well_index_homogeneous():
    """wellcase.py:182-192 with K = 3e-7: 2*pi*5*K/ln(r_o/0.1), r_o = 0.28*sqrt(50)/2."""
    assert orc._peaceman_wi(3e-7, 3e-7) == pytest.approx(4.111164584967624e-06, rel=1e-13)


def test_uniform_state_zero_residual_and_hydrostatic():
    g = orc.Grid(4, 3, 5, 6.096, 3.048, 0.6096, 3)
    prm = orc.Params(S_o=0.9)
    n = g.n
    rng = np.random.default_rng(0)
    K = 10.0 ** rng.normal(-7, 1, n)
    pb = orc.Problem(g, 2, prm, np.full(n, 0.2), K, K, K, gravity=False)
    u = np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, 0.9)])
    assert np.abs(orc.residual(pb, u, u, 3600.0)).max() == 0.0
    # mass conservation: fluxes telescope, so the summed oil equation is the accumulation only
    pb.gravity = True
    u2 = u.copy()
    u2[0] += rng.uniform(-1, 1, n)
    F = orc.residual(pb, u2, u, 3600.0)
    Wo = prm.T_prod * (prm.c_v_w * 0.1 + prm.c_v_o * 0.9)
    acc = g.vol * Wo * 0.2 * 0.9 * (orc.oil_rho(prm, u2[0], u2[1]) - orc.oil_rho(prm, u[0], u[1])) / 3600.0
    assert F[2].sum() == pytest.approx(acc.sum(), rel=1e-9)


def test_spmv_and_csr_agree():
    meta, pb, z = load("g5_tp3d_hetero_wellheater")
    J = z["J"]
    x = np.random.default_rng(1).normal(size=(pb.nf, pb.grid.n))
    y = orc.spmv(J, pb.grid, x)
    A = orc.to_csr(J, pb.grid, "field")
    assert np.allclose(A @ x.ravel(), y.ravel(), rtol=1e-13, atol=1e-6 * np.abs(y).max() * 1e-7)
    Ac = orc.to_csr(J, pb.grid, "cell")
    assert np.allclose((Ac @ x.T.ravel()).reshape(-1, pb.nf).T, y, rtol=1e-13, atol=1e-13 * np.abs(y).max())
