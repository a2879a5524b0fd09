"""Host-side mirror of the reference's geo / case / option surface (no GPU needed)."""
import os
import numpy as np
import pytest

from tests.golden_util import load
from thermalporous_b200 import cases as CS, geo as G, options as O
from thermalporous_b200.model import ConvergenceError, run_time_loop
from thermalporous_b200.physicalparameters import PhysicalParameters


def params(**kw):
    class P(PhysicalParameters):
        pass
    p = P()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _srcs(pb):
    return sorted((s.cell, s.kind, s.weight, s.bhp, s.max_rate, s.const_rate) for s in pb.sources)


def test_wellcase_reproduces_reference_deltas():
    """the golden fixtures store the delta fields the reference's WellCase built (wellcase.py:110-169)."""
    meta, pb, z = load("g1_sp2d_homo_const")
    prm = params(rate=1e-6, T_prod=320.0)
    geo = G.HomogeneousGeo(10, 8, prm, 20.0, 20.0)
    case = CS.WellCase(prm, geo, well_case="test0", constant_rate=True)
    assert sorted(CS.source_entries(case, prm, geo)) == _srcs(pb)
    assert np.allclose(geo.kT, 0.2 * prm.ko + 0.8 * prm.kr) and geo.name == "Homogeneous 10X8 grid"


def test_heatercase_bump_and_duplicates():
    meta, pb, z = load("g7_tp3d_homo_heater")
    prm = params(rate=1e-7, T_inj=373.15, S_o=0.9)
    geo = G.HomogeneousBoxGeo(8, 8, 8, prm, 1.2, 1.2, 4.0)
    L = 1.2
    hp = [[L / 4 + 0.01, L / 2 + 0.02, 0.8], [3 * L / 4, L / 2, 3.2], [L / 4 + 0.01, L / 2 + 0.02, 0.8]]
    got = sorted(CS.source_entries(CS.HeaterCase(prm, geo, heater_points=hp), prm, geo))
    ref = _srcs(pb)
    assert [g[:2] for g in got] == [r[:2] for r in ref]
    assert np.allclose([g[2] for g in got], [r[2] for r in ref], rtol=1e-13)


def test_sourceterms_sum_coincident_points():
    meta, pb, z = load("g6_tp3d_sources")
    prm = params(rate=2e-4, S_o=0.8)

    class HG(G.BoxGeo):
        def generate_geo_fields(self):
            self.phi, self.K_x, self.K_y, self.K_z, self.kT = pb.phi, pb.Kx, pb.Ky, pb.Kz, None
    geo = HG(6, 5, 4, prm, 6 * 6.096, 5 * 3.048, 4 * 0.6096)
    pp = [[1.5 * 6.096, 2.5 * 3.048, 0.5 * 0.6096]]
    ip = [[4.5 * 6.096, 1.5 * 3.048, 3.5 * 0.6096]]
    case = CS.SourceTerms(prm, geo, prod_points=pp + pp, inj_points=ip, heater_points=pp + ip)
    assert sorted(CS.source_entries(case, prm, geo)) == _srcs(pb)
    assert prm.prod_rate == prm.rate and case.name == "Sources"


def test_spe10_layout_and_synthetic_statistics():
    phi, Kx, Ky, Kz = G.spe10_synthetic(12, 20, 17, seed=10)
    assert phi.shape == Kx.shape == (12, 20, 17)
    prm = params()
    geo = G.SPE10Model3D(12, 20, 17, prm, fields=(phi, Kx, Ky, Kz))
    # slice arrays are [i, j, k]; cells are x fastest (SPE10model3D.py:30-68)
    i, j, k = 3, 7, 11
    c = i + 12 * (j + 20 * k)
    assert geo.K_x[c] == Kx[i, j, k] and geo.phi[c] == phi[i, j, k] + 1e-10
    assert (geo.Dx, geo.Dy, geo.Dz) == (6.096, 3.048, 0.6096) and geo.name.startswith("SPE10")
    ratio = Kz / Kx
    assert np.isclose(ratio.min(), 1e-3) and np.isclose(ratio.max(), 0.3)
    assert Kx.min() >= 6.65e-4 * G.MD_TO_MM2 * 0.999 and Kx.max() <= 2e4 * G.MD_TO_MM2 * 1.001
    g2 = G.SPE10Model3D(12, 20, 17, prm, fields=(phi, Kx, Ky, Kz), refine_z=3)
    assert g2.Nz == 51 and np.isclose(g2.Dz, 0.6096 / 3) and g2.K_x[c % 240 + 240 * (3 * k + 1)] == Kx[i, j, k]
    with pytest.raises(FileNotFoundError):
        G.SPE10Model3D(12, 20, 17, prm, data_dir="/nonexistent")


def test_option_sets_resolve_like_the_reference_dispatchers():
    o, dec, _ = O.resolve("pc_cptr", 2)
    assert (o["stage1"], o["stage2"], o["ksp_type"], o["ksp_rtol"], o["snes_max_it"]) == (O.S1_CPTR, O.S2_ILU0, O.KSP_FGMRES, 1e-8, 25)
    o, dec, _ = O.resolve(None, 2)            # twophase.py:929-930 -> pc_cptr_gmres
    assert o["stage1"] == O.S1_CPTR and dec == "No"
    o, dec, _ = O.resolve("pc_cpr_TI_temp", 2)
    assert o["stage1"] == O.S1_CPR and dec == "TI_temp"
    o, dec, _ = O.resolve("pc_cpr_QI", 1)
    assert (o["stage1"], o["ksp_type"], o["snes_max_it"], o["ksp_rtol"]) == (O.S1_CPR, O.KSP_GMRES, 15, 1e-5) and dec == "QI"
    o, _, _ = O.resolve("pc_fieldsplit_cd", 1)
    assert (o["stage1"], o["schur_pre"], o["stage2"]) == (O.S1_FIELDSPLIT, O.SCHUR_CONVDIFF, O.S2_NONE)
    o, _, _ = O.resolve(None, 1)              # singlephase.py:412: unmatched name -> bare GMRES + default PC
    assert o["stage1"] == O.S1_NONE and o["stage2"] == O.S2_ILU0
    for bad in ("pc_lu", "pc_hypre", "pc_cptramg", "faspardecomp"):
        with pytest.raises(O.UnsupportedOption):
            O.resolve(bad, 2)
    o, _, _ = O.resolve("pc_fieldsplit_selfp", 1)      # singlephase.py:322-329
    assert (o["stage1"], o["schur_pre"], o["stage2"]) == (O.S1_FIELDSPLIT, O.SCHUR_SELFP, O.S2_NONE)
    o, _, _ = O.resolve({"pc_type": "fieldsplit", "pc_fieldsplit_type": "schur", "pc_fieldsplit_schur_fact_type": "FULL",
                         "pc_fieldsplit_schur_precondition": "selfp"}, 1)
    assert o["schur_pre"] == O.SCHUR_SELFP
    # the raw PETSc dict the reference's pc_cptr expands to (twophase.py:531-550) plus a decoupling key
    v_cycle = {"ksp_type": "preonly", "pc_type": "hypre", "pc_hypre_type": "boomeramg", "pc_hypre_boomeramg_max_iter": 1}
    d = {"snes_type": "newtonls", "snes_max_it": 25, "ksp_type": "fgmres", "ksp_max_it": 200, "ksp_gmres_restart": 200,
         "ksp_rtol": 1e-8, "pc_type": "composite", "pc_composite_type": "multiplicative",
         "pc_composite_pcs": "python,bjacobi", "sub_0_pc_python_type": "thermalporous.preconditioners.CPTRStage1PC",
         "sub_0_cpr_stage1_pc_type": "fieldsplit", "sub_0_cpr_stage1_pc_fieldsplit_type": "schur",
         "sub_0_cpr_stage1_fieldsplit_0": v_cycle, "sub_1_sub_pc_type": "ilu", "sub_1_sub_pc_factor_levels": 0,
         "mat_type": "aij", "sub_0_cpr_decoup": "QI"}
    o, dec, _ = O.resolve(d, 2)
    assert o["stage1"] == O.S1_CPTR and dec == "QI" and o["ksp_restart"] == 200


def test_time_loop_failure_halves_dt_and_restores_state():
    """thermalmodel.py:162-181 (retry with dt/2) and :193-229 (saturation chop + clip)."""
    calls = []

    class St:
        def __init__(self, reason, nits=3):
            self.reason, self.nits, self.lits = reason, nits, 7

    class Ops:
        def copy(self, d, s):
            d[...] = s

        def minmax(self, u, f):
            return float(u[f].min()), float(u[f].max())

        def clip(self, u, f, lo, hi):
            np.clip(u[f], lo, hi, out=u[f])
    u = np.zeros((3, 4))
    u[2] = 0.5
    uo = u.copy()

    def newton(a, b, dt):
        calls.append(dt)
        if len(calls) == 1:
            a[:] = 99.0                 # garbage left behind by a failed solve must be discarded
            return St(-3)
        a[0] += 1.0
        a[2] = 1.2 if len(calls) == 2 else 0.9
        return St(3)
    res = run_time_loop(newton, Ops(), u, uo, end=1.0, maxdt=1.0, small_dt_start=False, dt_init_fact=1.0,
                        two_phase=True, i_S=2, spe10=False, max_steps=1)
    assert calls == [86400.0, 43200.0, 21600.0]            # fail -> half; S > 1 -> half again (chop)
    assert res.failed_solves == 1 and res.chops == 1 and res.dt_vec == [21600.0]
    assert u[0, 0] == 1.0 and np.all(u[2] == 0.9) and np.array_equal(u, uo)
    with pytest.raises(ConvergenceError):
        raise ConvergenceError(-3)


def test_spe10_dat_round_trip_and_slicing(tmp_path):
    """spe_*.dat layout (x fastest, layers top-down, 3 permeability blocks) and the slice arrays the reference's
    SPE10 geo classes load (data/create_SPE10_slice.py:23-71, create_SPE10_slice2D.py:11-60)."""
    from thermalporous_b200 import spe10data as D
    rng = np.random.default_rng(0)
    shp = (D.NZ, D.NY, D.NX)
    phi, kx, ky, kz = (np.round(rng.random(shp), 6) for _ in range(4))
    D.write_dat(str(tmp_path), phi, kx, ky, kz)
    got = D.read_dat(str(tmp_path))
    for a, b in zip(got, (phi, kx, ky, kz)):
        assert np.array_equal(a, b)
    s_phi, s_kx, s_ky, s_kz = D.create_slice(8, 16, 5, x_shift=3, y_shift=7, z_shift=2, fields=got, save_dir=str(tmp_path))
    # the reference's triple loop, literally
    line = phi.reshape(-1)
    kline = kz.reshape(-1) * 9.869233e-10
    for (i, j, kk) in ((0, 0, 0), (7, 15, 4), (2, 9, 3)):
        src = (i + 3) + (j + 7) * 60 + (kk + 2) * 220 * 60
        assert s_phi[i, j, 5 - 1 - kk] == line[src] and s_kz[i, j, 5 - 1 - kk] == kline[src]
    prm = params()
    geo = G.SPE10Model3D(8, 16, 5, prm, data_dir=str(tmp_path))         # loads the slice_*.npy just written
    assert geo.K_z[2 + 8 * (9 + 16 * 1)] == s_kz[2, 9, 1]
    p2, kx2, ky2 = D.create_slice2D(6, 9, x_shift=1, y_shift=2, z_shift=4, fields=got)
    assert p2[3, 5] == line[(3 + 1) + (5 + 2) * 60 + 4 * 220 * 60] and p2.shape == (6, 9)


def test_checkpoint_round_trip(tmp_path):
    from thermalporous_b200.model import load_checkpoint, save_checkpoint
    u = np.random.default_rng(1).random((3, 40))
    save_checkpoint(str(tmp_path / "chk"), u)
    assert np.array_equal(load_checkpoint(str(tmp_path / "chk"), (3, 40)), u)
    with pytest.raises(ValueError):
        load_checkpoint(str(tmp_path / "chk"), (2, 40))


def test_mixing_case_initial_conditions():
    """mixingcase.py:20-47."""
    from thermalporous_b200 import cases as CS, geo as G
    from thermalporous_b200.physicalparameters import PhysicalParameters
    prm = PhysicalParameters()
    g2 = G.HomogeneousGeo(6, 8, prm, 12.0, 16.0)
    ic = CS.MixingCase(prm, g2, "coldandhot").init_IC("Single phase")
    T = ic[1].reshape(8, 6)
    assert ic.shape == (2, 48) and (T[:4] == prm.T_inj).all() and (T[4:] == prm.T_prod).all() and (ic[0] == prm.p_ref).all()
    assert CS.MixingCase(prm, g2, "coldonhot").mixing_case == "coldandhot"        # 2-D falls back (mixingcase.py:13-15)
    g3 = G.HomogeneousBoxGeo(3, 4, 6, prm, 6.0, 8.0, 12.0)
    ic3 = CS.MixingCase(prm, g3, "coldonhot").init_IC("Two-phase")
    T3 = ic3[1].reshape(6, 4, 3)
    assert ic3.shape == (3, 72) and (T3[:3] == prm.T_inj).all() and (T3[3:] == prm.T_prod).all() and (ic3[2] == prm.S_o).all()
    prm2 = PhysicalParameters()
    hl = CS.MixingCase(prm2, g3, "heavyonlight")
    assert prm2.API == 40
    S = hl.init_IC("Two-phase")[2].reshape(6, 4, 3)
    assert (S[:3] == 1.0).all() and (S[3:] == 0.0).all()
    with pytest.raises(SystemExit):
        hl.init_IC("Single phase")
    with pytest.raises(SystemExit):
        CS.MixingCase(prm, g3, "sideways")


def test_vti_pvd_output_and_matlab_dumps(tmp_path):
    """save=True writer (thermalmodel.py:113-133) and the utils.ExportJacobian / ExportResidual dump format
    (utils.py:27-43) with its loader: the loaded sparse matrix acts like the block-stencil SpMV."""
    from oracle import tp_oracle as orc
    from thermalporous_b200 import geo as G, vtkout as V
    from thermalporous_b200.partition import Slab
    from thermalporous_b200.physicalparameters import PhysicalParameters
    prm = PhysicalParameters()
    g = G.HomogeneousBoxGeo(4, 3, 5, prm, 10.0, 10.0, 10.0)
    rng = np.random.default_rng(0)
    for world, rank in ((1, 0), (2, 1)):
        sl = Slab(g, world, rank)
        w = V.PvdWriter(str(tmp_path / ("w%d" % world)), ["pressure", "temperature"], g, sl, "" if world == 1 else "_rank1of2")
        f = rng.random((2, sl.ncell))
        w.write(0.0, f)
        w.write(0.5, 2 * f)
        w.close()
        suffix = "" if world == 1 else "_rank1of2"
        ext, sp, arr = V.read_vti(str(tmp_path / ("w%d" % world) / ("pressure%s_1.vti" % suffix)))
        assert ext == (0, 4, 0, 3, sl.k0, sl.k1) and sp == (g.Dx, g.Dy, g.Dz)
        assert np.array_equal(arr["pressure"], 2 * f[0])
        assert (tmp_path / ("w%d" % world) / ("temperature%s.pvd" % suffix)).read_text().count("<DataSet") == 2
    J = rng.normal(size=(7, 3, 3, g.ncell))
    V.export_jacobian(J, 4, 3, 5, str(tmp_path / "matrix.txt"))
    A = V.load_matlab_matrix(str(tmp_path / "matrix.txt"))
    x = rng.normal(size=(3, g.ncell))
    y = orc.spmv(J, orc.Grid(4, 3, 5, 1.0, 1.0, 1.0, 3), x)
    assert A.shape == (3 * g.ncell, 3 * g.ncell) and np.abs(A @ x.reshape(-1) - y.reshape(-1)).max() < 1e-13
    V.export_residual(x, str(tmp_path / "rhs.txt"))
    assert np.array_equal(V.load_matlab_vector(str(tmp_path / "rhs.txt")), x.reshape(-1))


def _rate_fixture_names():
    import os
    from tests.golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "pc", "decoup.npz"))
    return sorted(k.split("|")[0] for k in z.files if k.endswith("|rates|q"))


@pytest.mark.parametrize("name", _rate_fixture_names())
def test_well_totals_match_the_reference_expressions(name):
    """cases.source_rates (the per-step totals of thermalmodel.py:231-270) against the reference's own rate
    expressions evaluated at the fixture states (tests/golden/make_pc_golden.py: rates)."""
    import os
    from tests.golden_util import GOLDEN_DIR
    want = np.load(os.path.join(GOLDEN_DIR, "pc", "decoup.npz"))[name + "|rates|q"]
    meta, pb, z = load(name)
    prm = PhysicalParameters()
    for k, v in meta["params"].items():
        setattr(prm, k, v)
    ent = [(s.cell, s.kind, s.weight, s.bhp, s.max_rate, s.const_rate) for s in pb.sources]
    cells = np.array([e[0] for e in ent])
    got = CS.source_rates(ent, z["u"][:, cells], pb.Kx[cells], pb.Ky[cells], prm, pb.nphase)
    for key, ref in zip(("inj", "prod", "oil", "water"), want):
        if np.isnan(ref):
            assert got.get(key) is None
        else:
            assert got[key] == pytest.approx(ref, rel=1e-12, abs=0.0), (name, key)


def test_model_reports_the_well_totals_in_the_reference_order():
    """ThermalModel.well_totals gathers the state at the source cells (here on a CPU tensor standing in for the
    device state) and rate_lines prints what thermalmodel.py:231-270 prints, in its order."""
    import os
    import torch
    from tests.golden_util import GOLDEN_DIR
    from thermalporous_b200.model import ThermalModel, rate_lines
    name = "g5_tp3d_hetero_wellheater"
    want = np.load(os.path.join(GOLDEN_DIR, "pc", "decoup.npz"))[name + "|rates|q"]
    meta, pb, z = load(name)
    prm = PhysicalParameters()
    for k, v in meta["params"].items():
        setattr(prm, k, v)

    class Fake:
        pass
    m = Fake()
    m._entries = [(s.cell, s.kind, s.weight, s.bhp, s.max_rate, s.const_rate) for s in pb.sources]
    cells = np.array([e[0] for e in m._entries])
    m._src_K = (pb.Kx[cells], pb.Ky[cells])
    m.u = torch.from_numpy(np.ascontiguousarray(z["u"]))
    m.params, m.nphase, m.world = prm, 2, 1
    tot = ThermalModel.well_totals(m)
    for key, ref in zip(("inj", "prod", "oil", "water"), want):
        assert tot[key] == pytest.approx(ref, rel=1e-12)
    lines = rate_lines(tot, sources_case=False)
    assert [ln.split(" is ")[0] for ln in lines] == ["Total injection rate", "Total water production rate",
                                                     "Total oil production rate", "Total production rate"]
    lines = rate_lines({"inj": 1.0, "prod": -1.0, "oil": None, "water": None}, sources_case=True)
    assert lines == ["Total injection rate is 1.0", "Total production rate is -1.0"]


def test_results_file_follows_the_reference(tmp_path):
    """thermalmodel.py:28-32,78-80: resultprint goes to the screen and to the results file; nothing is written by a
    quiet model."""
    from thermalporous_b200.model import ThermalModel

    class Fake:
        pass
    for verbosity, expect in ((True, True), (False, False)):
        m = Fake()
        m.verbosity, m.rank = verbosity, 0
        m.filename, m._results_file = str(tmp_path / ("r%d" % verbosity) / "results.txt"), None
        ThermalModel.resultprint(m, "nits = ", [4, 5], ";")
        ThermalModel.resultprint(m, "Total Linear iterations: ", 9)
        if m._results_file is not None:
            m._results_file.close()
        assert os.path.exists(m.filename) == expect
        if expect:
            assert open(m.filename).read() == "nits =  [4, 5] ;\nTotal Linear iterations:  9\n"


def test_named_option_sets_equal_the_references_own_dictionaries():
    """options.resolve(name) against options.resolve(<the dict the reference's init_solver_parameters builds for that
    name>) - the dictionaries come from the reference's own dispatchers (tests/golden/make_option_golden.py).
    One deliberate difference: every two-phase set of the reference starts from `newton_fas_krylov` (twophase.py:434-
    518, 927), whose FAS nonlinear preconditioner needs a mesh hierarchy and whose l2 line search belongs to it; the
    named sets here keep the Newton-Krylov part with the basic line search (out of scope); a raw dict that
    asks for l2 is refused (no silent mapping onto another search).  The ILU(1) sets (pc_bilu, pc_cprilu1_gmres:
    sub_pc_factor_levels 1) are refused by name and by dictionary: the second stage here is ILU(0)."""
    import json
    from tests.golden_util import GOLDEN_DIR
    sets = json.load(open(os.path.join(GOLDEN_DIR, "pc", "option_sets.json")))
    ilu1 = {"1|pc_ilu", "1|pc_bilu", "2|pc_bilu", "2|pc_cprilu1_gmres"}
    assert len(sets) == len(O.SINGLE_PHASE_SETS) + len(O.TWO_PHASE_SETS) + 2 + len(ilu1)
    for key, rec in sets.items():
        nphase, name = key.split("|")
        nphase, name = int(nphase), (None if name == "None" else name)
        d = dict(rec["parameters"])
        if rec["decoup"] != "No":
            d["sub_0_cpr_decoup"] = rec["decoup"]
        if nphase == 2:
            with pytest.raises(O.UnsupportedOption):
                O.resolve(dict(d), nphase)                         # l2 line search of the FAS branch
            assert d.pop("snes_linesearch_type") == "l2"
        if key in ilu1:
            with pytest.raises(O.UnsupportedOption):
                O.resolve(name, nphase)
            with pytest.raises(O.UnsupportedOption):
                O.resolve(d, nphase)
            continue
        by_name, dec_name, _ = O.resolve(name, nphase)
        by_dict, dec_dict, _ = O.resolve(d, nphase)
        assert dec_name == dec_dict == rec["decoup"], key
        assert "linesearch" not in by_name and "linesearch" not in by_dict
        assert by_name == by_dict, key
    with pytest.raises(O.UnsupportedOption):
        O.resolve({"pc_type": "ilu", "ksp_rtoll": 1e-5}, 1)        # a typo is an error, not a dropped key


def test_bench_reference_arm_line_keeps_the_contract(monkeypatch, capsys):
    """bench.py --impl reference: one JSON line with the contract's keys, the CPU restatement's roofline, and an honest
    same_config flag (true only for the whole grid at --gpus 1).  Run here on a 5-layer sample to stay fast."""
    import argparse
    import json
    import bench
    monkeypatch.setattr(bench, "CPU_SAMPLE_NZ", 5)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=2, steps=1, warmup=1, cpu_sample=False)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"]["same_config"] is False and line["config"]["cells"] == 60 * 220 * 5
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] == line["e2e"]["value"] > 0
    assert cb["roofline"]["stream_triad_gbs"] > 0 and 0 < cb["roofline"]["spmv_frac"] < 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # only rank 0 prints under torchrun
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args)
    assert capsys.readouterr().out == ""
