"""The BASELINE configs C1-C3 through the mirrored model classes (SinglePhase / TwoPhase .solve()) on the GPU,
against the CPU restatement driven by the same host time loop with the same option set.  The default solver
tolerances (SNES rtol 1e-8, KSP rtol 1e-5 | 1e-8) bound the agreement of two different-rounding runs at about 1e-6,
so both sides are converged far below them (SNES rtol 1e-11, KSP rtol 1e-10) and the fields compared at the
north_star's 1e-8."""
import numpy as np
import pytest

from oracle import cport
from thermalporous_b200 import cases as CS, geo as G, options as O
from thermalporous_b200.model import SinglePhase, TwoPhase, run_time_loop
from thermalporous_b200.physicalparameters import PhysicalParameters

pytestmark = pytest.mark.gpu


def params(**kw):
    class P(PhysicalParameters):
        pass
    p = P()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class NpOps:
    def copy(self, d, s):
        d[...] = s

    def minmax(self, u, f):
        return float(u[f].min()), float(u[f].max())

    def clip(self, u, f, lo, hi):
        np.clip(u[f], lo, hi, out=u[f])


def cpu_reference(model, nphase, pc):
    geo, prm = model.geo, model.params
    eng = cport.CpuEngine(geo.dim, geo.Nx, geo.Ny, getattr(geo, "Nz", 1), geo.Dx, geo.Dy, getattr(geo, "Dz", 1.0), nphase, prm)
    eng.set_field(cport.PHI, geo.phi)
    eng.set_field(cport.KX, geo.K_x)
    eng.set_field(cport.KY, geo.K_y)
    if geo.dim == 3:
        eng.set_field(cport.KZ, geo.K_z)
    if nphase == 1:
        eng.set_field(cport.KT, geo.kT)
    eng.set_sources(CS.source_entries(model.case, prm, geo))
    opts, _, _ = O.resolve(pc, nphase)
    opts.update(TIGHT)
    eng.set_solver_opts(**opts)
    u = np.ascontiguousarray(model.initial_condition, dtype=np.float64).copy()
    res = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), NpOps(), u, u.copy(), end=model.end, maxdt=model.maxdt,
                        small_dt_start=model.small_dt_start, dt_init_fact=model.dt_init_fact, two_phase=nphase == 2, i_S=2,
                        spe10=geo.name.startswith("SPE10"))
    eng.close()
    return u, res


TIGHT = dict(snes_rtol=1e-11, snes_stol=1e-13, ksp_rtol=1e-10, snes_max_it=40)


def check(model, nphase, pc):
    model.engine.set_solver_opts(**TIGHT)
    res = model.solve()
    uc, rc = cpu_reference(model, nphase, pc)
    assert res.failed_solves == 0 and len(res.dt_vec) == len(rc.dt_vec)
    assert np.allclose(res.dt_vec, rc.dt_vec, rtol=1e-12)
    assert model.total_nits == res.total_nits > 0 and model.total_lits >= model.total_nits
    for f, a in enumerate(model.fields()):
        assert np.abs(a - uc[f]).max() <= 1e-8 * np.abs(uc[f]).max()
    return res


@pytest.mark.parametrize("pc", ["pc_fieldsplit_cd", "pc_fieldsplit_selfp", "pc_cpr", None])
def test_c1_single_phase_homogeneous_wells(pc):
    """tests/test_homo_wells.py: N x N homogeneous box, L = 20 m, wells 'test0' at constant rate, 2 steps of 1 day."""
    prm = params(rate=1e-6, T_prod=320.0)
    geo = G.HomogeneousGeo(40, 40, prm, 20.0, 20.0)
    case = CS.WellCase(prm, geo, well_case="test0", constant_rate=True)
    model = SinglePhase(geo, case, prm, end=2.0, maxdt=1.0, small_dt_start=False, solver_parameters=pc, verbosity=False)
    res = check(model, 1, pc)
    assert len(res.dt_vec) == 2
    p, T = model.fields()
    assert 320.0 < T.max() <= prm.T_inj and p.max() > prm.p_ref > p.min()   # hot injection; injectors raise p, producers lower it


def test_c2_single_phase_spe10_slice_cpr():
    """tests/test_60x120_wells.py shape: SPE10-shaped 60x120 layer, Peaceman wells, CPR."""
    prm = params()
    geo = G.SPE10Model(60, 120, prm, fields=G.spe10_synthetic_layer(60, 120))
    case = CS.WellCase(prm, geo, well_case="SPE10_60x120")
    model = SinglePhase(geo, case, prm, end=0.02, maxdt=0.01, small_dt_start=True, dt_init_fact=2 ** -5,
                        solver_parameters="pc_cpr_QI", verbosity=False)
    check(model, 1, "pc_cpr_QI")


@pytest.mark.parametrize("pc", ["pc_cptr", "pc_cpr_TI"])
def test_c3_two_phase_spe10_slice(pc):
    """tests_twophase/test_60x120_wells_default.py: rate 2e-4, S_o 0.9, named option set."""
    prm = params(rate=2e-4, S_o=0.9)
    geo = G.SPE10Model(60, 120, prm, fields=G.spe10_synthetic_layer(60, 120))
    case = CS.WellCase(prm, geo, well_case="SPE10_60x120")
    model = TwoPhase(geo, case, prm, end=0.004, maxdt=0.002, small_dt_start=True, dt_init_fact=2 ** -4,
                     solver_parameters=pc, verbosity=False, vector=(pc == "pc_cpr_TI"))   # vector: layout only
    res = check(model, 2, pc)
    p, T, S = model.fields()
    assert 0.0 <= S.min() and S.max() <= 1.0 and len(res.dt_vec) >= 3


def test_mixing_case_initial_condition_and_field_output(tmp_path, monkeypatch):
    """mixingcase.py "coldandhot": no sources, hot / cold halves as initial condition (thermalmodel.py:23-26), fields
    written every n_save-th step as the reference's save=True does (:113-133, 304-320)."""
    from thermalporous_b200 import vtkout
    monkeypatch.chdir(tmp_path)
    prm = params(S_o=0.9)
    geo = G.HomogeneousGeo(24, 20, prm, 20.0, 20.0)
    case = CS.MixingCase(prm, geo, "coldandhot")
    model = TwoPhase(geo, case, prm, end=0.03, maxdt=0.01, small_dt_start=False, solver_parameters="pc_cptr",
                     verbosity=False, save=True, n_save=2)
    ic = model.initial_condition
    assert set(np.unique(ic[1])) == {prm.T_prod, prm.T_inj} and (ic[2] == 0.9).all()
    res = check(model, 2, "pc_cptr")
    p, T, S = model.fields()
    assert prm.T_prod < T.mean() < prm.T_inj and T.max() - T.min() < prm.T_inj - prm.T_prod + 1e-9
    # initial state + steps 1 and 3 (i_plot 0 and 2)
    nsaved = 1 + len([k for k in range(len(res.dt_vec)) if k % 2 == 0])
    for name in ("pressure", "temperature", "saturation_o"):
        assert (tmp_path / "results" / (name + ".pvd")).read_text().count("<DataSet") == nsaved
    ext, sp, arr = vtkout.read_vti(str(tmp_path / "results" / ("temperature_%d.vti" % (nsaved - 1))))
    assert ext == (0, 24, 0, 20, 0, 1) and arr["temperature"].shape == (geo.ncell,)


def test_unsupported_option_set_is_loud():
    prm = params(S_o=0.9)
    geo = G.HomogeneousGeo(8, 8, prm, 20.0, 20.0)
    case = CS.WellCase(prm, geo, well_case="test0", constant_rate=True)
    with pytest.raises(O.UnsupportedOption):
        TwoPhase(geo, case, prm, solver_parameters="pc_lu", verbosity=False)


def test_c4_two_phase_homogeneous_heaters():
    """tests_twophase/test3D_homo_heater.py (BASELINE config 4) at N = 20: 42 heater points, 3 steps of one day."""
    from tools.run_c4 import build
    model = build(20, steps=3)
    res = check(model, 2, "pc_cptr")
    assert len(res.dt_vec) == 3
    p, T, S = model.fields()
    assert prm_close(T.max(), 373.15) and T.min() >= 288.7


def prm_close(tmax, t_inj):
    return 300.0 < tmax <= t_inj + 1e-6

