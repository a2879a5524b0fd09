"""GPU parity of the solver stack (K3-K11) through the C-ABI against the CPU restatement
(oracle/cport) on the same inputs, and against the reference-form golden time loops.

Component outputs (decoupling weights, multigrid hierarchy and Galerkin operators, V-cycle, ILU(0)
apply, full PC apply) are compared at 1e-10 relative - both sides run the same algorithm and differ
only in summation order.  Converged fields are compared at 1e-8 relative (north_star)."""
import numpy as np
import pytest

from oracle import cport
from tests.golden_util import load, rel_err_rows
from tests.gpu_util import engine_from_problem, random_problem

pytestmark = pytest.mark.gpu

COMBOS = [
    # nphase, dim, shape, opts
    (2, 3, (9, 11, 14), dict(stage1=cport.S1_CPTR, decoup=0, schur_pre=cport.SCHUR_CONVDIFF)),
    (2, 3, (6, 13, 10), dict(stage1=cport.S1_CPTR, decoup=1, schur_pre=cport.SCHUR_CONVDIFF)),
    (2, 3, (7, 8, 12), dict(stage1=cport.S1_CPTR, decoup=2, schur_pre=cport.SCHUR_A11)),
    (2, 3, (5, 9, 16), dict(stage1=cport.S1_CPR, decoup=0)),
    (2, 3, (5, 9, 16), dict(stage1=cport.S1_CPR, decoup=1)),
    (2, 3, (8, 5, 11), dict(stage1=cport.S1_CPR, decoup=2)),
    (2, 3, (6, 7, 9), dict(stage1=cport.S1_CPR, decoup=3)),
    (2, 2, (1, 21, 30), dict(stage1=cport.S1_CPR, decoup=4)),
    (2, 2, (1, 24, 17), dict(stage1=cport.S1_CPTR, decoup=1)),
    (1, 3, (6, 10, 12), dict(stage1=cport.S1_CPR, decoup=1)),
    (1, 2, (1, 20, 25), dict(stage1=cport.S1_CPR, decoup=2)),
    (1, 2, (1, 16, 23), dict(stage1=cport.S1_FIELDSPLIT, schur_pre=cport.SCHUR_CONVDIFF, stage2=cport.S2_NONE)),
    (1, 3, (5, 8, 9), dict(stage1=cport.S1_FIELDSPLIT, schur_pre=cport.SCHUR_A11, stage2=cport.S2_NONE)),
    (1, 2, (1, 12, 15), dict(stage1=cport.S1_FIELDSPLIT, schur_pre=cport.SCHUR_DIAG, stage2=cport.S2_NONE)),
    (1, 2, (1, 14, 19), dict(stage1=cport.S1_FIELDSPLIT, schur_pre=cport.SCHUR_SELFP, stage2=cport.S2_NONE)),
    (1, 3, (6, 7, 8), dict(stage1=cport.S1_FIELDSPLIT, schur_pre=cport.SCHUR_SELFP, stage2=cport.S2_NONE)),
    (2, 3, (5, 6, 9), dict(stage1=cport.S1_CPTR, decoup=1, schur_pre=cport.SCHUR_SELFP)),
    (2, 3, (4, 6, 40), dict(stage1=cport.S1_NONE, stage2=cport.S2_ILU0)),
    (2, 3, (4, 6, 10), dict(stage1=cport.S1_CPR, stage2=cport.S2_BJACOBI)),
    (2, 3, (20, 24, 40), dict(stage1=cport.S1_CPTR, decoup=1, mg_pre=2, mg_post=2, mg_cycles=2)),
    # degenerate shapes: single columns / rows / cells (odd sizes exercise the ragged last aggregate)
    (2, 3, (9, 1, 1), dict(stage1=cport.S1_CPTR, decoup=1)),
    (2, 3, (1, 1, 7), dict(stage1=cport.S1_CPR, decoup=2)),
    (1, 3, (1, 5, 1), dict(stage1=cport.S1_CPR, decoup=1)),
    (2, 2, (1, 1, 1), dict(stage1=cport.S1_CPTR)),
    (2, 3, (3, 3, 3), dict(stage1=cport.S1_CPTR, mg_pre=1, mg_post=0, mg_coarse_sweeps=1)),
    (2, 3, (17, 19, 23), dict(stage1=cport.S1_CPTR, decoup=2, mg_semi_theta=0.0, mg_full_below=100)),
    # diagonal-dominance stop of the hierarchies: never | at any level whose rows pass 0.5 | always at level 0
    (2, 3, (9, 11, 14), dict(stage1=cport.S1_CPTR, decoup=0, mg_dd_stop=0.0)),
    (2, 3, (9, 11, 14), dict(stage1=cport.S1_CPTR, decoup=1, mg_dd_stop=0.5)),
    (2, 3, (6, 13, 10), dict(stage1=cport.S1_CPTR, decoup=1, mg_dd_stop=1e9)),
    # 3-D combos above run the default smoother (zebra z-line, z never coarsened, coarse couplings scaled by 0.5);
    # the point smoother / plain Galerkin paths: red-black Gauss-Seidel with z coarsened like any other axis
    (2, 3, (9, 11, 14), dict(stage1=cport.S1_CPTR, decoup=0, mg_smoother=0)),
    (2, 3, (9, 11, 14), dict(stage1=cport.S1_CPTR, decoup=1, mg_smoother=0, mg_coarse_scale=1.0)),
    (1, 3, (6, 10, 12), dict(stage1=cport.S1_CPR, decoup=1, mg_smoother=0, mg_coarse_scale=1.0)),
    (2, 3, (20, 24, 40), dict(stage1=cport.S1_CPTR, decoup=1, mg_smoother=0, mg_cycles=2)),
    (2, 2, (1, 24, 17), dict(stage1=cport.S1_CPTR, decoup=1, mg_coarse_scale=1.0)),
    # z-line tiles: more columns per colour than one 32-column tile, tall columns, odd sizes
    (2, 3, (85, 70, 33), dict(stage1=cport.S1_CPTR, decoup=0)),
    (2, 3, (130, 9, 7), dict(stage1=cport.S1_CPR, decoup=1, mg_pre=1, mg_post=1)),
    (1, 3, (300, 3, 2), dict(stage1=cport.S1_CPR, decoup=0)),
]
TOL = 1e-10


def _pair(nphase, dim, shape, opts, seed=1):
    pb, u, uo = random_problem(dim, nphase, shape, seed=seed, spread=0.05)
    g = engine_from_problem(pb)
    c = cport.engine_from_problem(pb)
    g.set_solver_opts(**opts)
    c.set_solver_opts(**opts)
    return pb, u, uo, g, c


@pytest.mark.parametrize("nphase,dim,shape,opts", COMBOS)
def test_pc_components_match_cpu_restatement(nphase, dim, shape, opts):
    pb, u, uo, g, c = _pair(nphase, dim, shape, opts)
    dt = 4000.0
    F, J = g.assemble(u, uo, dt)
    Jh = J.cpu().numpy()
    g.pc_setup(J, u, dt)
    c.pc_setup(Jh, u, dt)
    rng = np.random.default_rng(5)
    n = pb.grid.n
    s1 = opts.get("stage1", 0)
    if s1 != cport.S1_NONE:
        wf = (1, 2) if s1 == cport.S1_CPR else (0, 1)
        for f in wf:
            if f >= pb.nf:
                continue
            wg = g.weights(f).cpu().numpy()
            wc = c.weights(f)
            assert np.abs(wg - wc).max() <= TOL * max(1.0, np.abs(wc).max())
        hier = [0] + ([1] if s1 in (cport.S1_CPTR, cport.S1_FIELDSPLIT) else [])
        for which in hier:
            lg, lc = g.mg_levels(which), c.mg_levels(which)
            assert lg == lc
            for l in range(len(lg)):
                assert rel_err_rows(g.mg_level_op(which, l).cpu().numpy(), c.mg_level_op(which, l)) < TOL
            b = rng.normal(size=n)
            yg = g.mg_apply(which, b).cpu().numpy()
            yc = c.mg_apply(which, b)
            assert np.abs(yg - yc).max() < TOL * np.abs(yc).max()
    if opts.get("stage2", cport.S2_ILU0) != cport.S2_NONE:
        r = rng.normal(size=(pb.nf, n))
        assert rel_err_rows(g.stage2_apply(r).cpu().numpy(), c.stage2_apply(r)) < TOL
    x = rng.normal(size=(pb.nf, n)) * np.abs(F.cpu().numpy()).max(axis=1, keepdims=True)
    assert rel_err_rows(g.pc_apply(x).cpu().numpy(), c.pc_apply(x)) < 1e-9
    g.close()
    c.close()


@pytest.mark.parametrize("nphase,dim,shape,opts", [COMBOS[1], COMBOS[4], COMBOS[9], COMBOS[11], COMBOS[14]])
@pytest.mark.parametrize("ksp_type", [cport.KSP_GMRES, cport.KSP_FGMRES])
def test_ksp_matches_cpu_restatement(nphase, dim, shape, opts, ksp_type):
    # classical Gram-Schmidt with a second pass once the residual is below 1e-3 (both sides): unrefined CGS
    # used to stall after ~8 decades on the slowly converging ILU-only case
    rtol = 1e-10
    pb, u, uo, g, c = _pair(nphase, dim, shape, dict(opts, ksp_type=ksp_type, ksp_rtol=rtol), seed=4)
    dt = 4000.0
    F, J = g.assemble(u, uo, dt)
    Jh, Fh = J.cpu().numpy(), F.cpu().numpy()
    g.pc_setup(J, u, dt)
    c.pc_setup(Jh, u, dt)
    xg, its_g, reason_g, rn_g = g.ksp_solve(J, F)
    xc, its_c, reason_c, rn_c = c.ksp_solve(Jh, Fh)
    assert reason_g == 2 and reason_c == 2
    assert abs(its_g - its_c) <= 1
    # both satisfy the same true-residual bound, and agree far below the linear tolerance's effect
    from oracle import tp_oracle as orc
    res = np.linalg.norm(orc.spmv(Jh, pb.grid, xg.cpu().numpy()) - Fh) / np.linalg.norm(Fh)
    assert res < 5 * rtol
    assert rel_err_rows(xg.cpu().numpy(), xc) < 1e4 * rtol
    g.close()
    c.close()


def test_zero_permeability_cells_and_sources_in_the_solver():
    """cells cut off from their neighbours (K = 0: pure accumulation rows) and active wells / heaters:
    PC components, Krylov counts and the Newton update agree with the CPU restatement."""
    pb, u, uo = random_problem(3, 2, (7, 9, 11), seed=9, spread=0.02, nsrc=5)
    rng = np.random.default_rng(2)
    dead = rng.random(pb.grid.n) < 0.07
    for K in (pb.Kx, pb.Ky, pb.Kz):
        K[dead] = 0.0
    g = engine_from_problem(pb)
    c = cport.engine_from_problem(pb)
    for e in (g, c):
        e.set_solver_opts(stage1=cport.S1_CPTR, decoup=1, ksp_rtol=1e-9)
    F, J = g.assemble(u, uo, 2000.0)
    Fc, Jc = c.assemble(u, uo, 2000.0)
    assert rel_err_rows(F.cpu().numpy(), Fc) < 1e-12 and rel_err_rows(J.cpu().numpy(), Jc) < 1e-12
    g.pc_setup(J, u, 2000.0)
    c.pc_setup(Jc, u, 2000.0)
    x = rng.normal(size=(3, pb.grid.n))
    assert rel_err_rows(g.pc_apply(x).cpu().numpy(), c.pc_apply(x)) < 1e-9
    xg, its_g, reason_g, _ = g.ksp_solve(J, F)
    xc, its_c, reason_c, _ = c.ksp_solve(Jc, Fc)
    assert reason_g == reason_c == 2 and abs(its_g - its_c) <= 1
    assert rel_err_rows(xg.cpu().numpy(), xc) < 1e-5
    g.close()
    c.close()


def test_saturation_outside_0_1_does_not_break_the_preconditioner():
    """Newton iterates may leave [0,1] (basic line search): S_o > 1 gives the water phase a negative mobility and
    the pressure rows of those cells lose diagonal dominance.  The multigrid's row repair keeps the V-cycle a
    contraction; both sides repair the same rows and the Krylov solve still converges to 1e-8."""
    pb, u, uo = random_problem(3, 2, (10, 12, 14), seed=21, spread=0.02, nsrc=3)
    rng = np.random.default_rng(3)
    bad = rng.choice(pb.grid.n, size=12, replace=False)
    u[2, bad[:8]] = 1.0 + rng.uniform(0.005, 0.05, 8)
    u[2, bad[8:]] = -rng.uniform(0.005, 0.03, 4)
    g = engine_from_problem(pb)
    c = cport.engine_from_problem(pb)
    for e in (g, c):
        e.set_solver_opts(stage1=cport.S1_CPTR, decoup=0, ksp_rtol=1e-8)
    dt = 300.0
    F, J = g.assemble(u, uo, dt)
    Jh, Fh = J.cpu().numpy(), F.cpu().numpy()
    g.pc_setup(J, u, dt)
    c.pc_setup(Jh, u, dt)
    for which in (0, 1):
        a_g, a_c = g.mg_level_op(which, 0).cpu().numpy(), c.mg_level_op(which, 0)
        assert rel_err_rows(a_g, a_c) < TOL
        assert (a_c[0] >= 0.8 * np.abs(a_c[1:]).sum(axis=0) * (1 - 1e-12)).all()      # every row repaired or healthy
        b = rng.normal(size=pb.grid.n)
        yg, yc = g.mg_apply(which, b).cpu().numpy(), c.mg_apply(which, b)
        assert np.isfinite(yg).all() and np.abs(yg - yc).max() < 1e-9 * np.abs(yc).max()
    xg, its_g, reason_g, _ = g.ksp_solve(J, F)
    xc, its_c, reason_c, _ = c.ksp_solve(Jh, Fh)
    assert reason_g == reason_c == 2 and abs(its_g - its_c) <= 1 and its_g < 80
    g.close()
    c.close()


def test_restart_and_failure_reasons():
    pb, u, uo, g, c = _pair(2, 3, (6, 8, 10), dict(stage1=cport.S1_NONE, stage2=cport.S2_BJACOBI, ksp_restart=5,
                                                   ksp_max_it=60, ksp_rtol=1e-9))
    F, J = g.assemble(u, uo, 4000.0)
    g.pc_setup(J, u, 4000.0)
    c.pc_setup(J.cpu().numpy(), u, 4000.0)
    xg, its_g, reason_g, _ = g.ksp_solve(J, F)
    xc, its_c, reason_c, _ = c.ksp_solve(J.cpu().numpy(), F.cpu().numpy())
    assert reason_g == reason_c and abs(its_g - its_c) <= 2
    g.set_solver_opts(ksp_max_it=3)
    g.pc_setup(J, u, 4000.0)
    _, its, reason, _ = g.ksp_solve(J, F)
    assert reason == -3 and its == 3      # KSP_DIVERGED_ITS
    g.close()
    c.close()


LOOPS = {
    "l1_sp2d_homo_loop": dict(end=2.0, maxdt=1.0, small_dt_start=False, dt_init_fact=2 ** -10, spe10=False),
    "l2_tp2d_hetero_loop": dict(end=0.02, maxdt=0.01, small_dt_start=True, dt_init_fact=2 ** -3, spe10=True),
    "l3_tp3d_heater_loop": dict(end=3.0, maxdt=1.0, small_dt_start=False, dt_init_fact=2 ** -10, spe10=False),
}


@pytest.mark.parametrize("name", sorted(LOOPS))
@pytest.mark.parametrize("decoup", [0, 1])
def test_time_steps_converged_fields_match_reference_golden(name, decoup):
    """every time step of the reference's own ThermalModel.solve() run (through the DG0 shim, direct linear
    solver) re-solved by the GPU Newton-Krylov path with the reference's dt: converged fields within 1e-8
    per step (north_star); Newton counts are reported by the reference and may differ by one here because
    the linear solves are inexact (north_star: counts are not required to match)."""
    meta, pb, z = load(name)
    g = engine_from_problem(pb)
    g.set_solver_opts(snes_rtol=1e-12, snes_stol=1e-13, ksp_rtol=1e-10, decoup=decoup)
    u = g.tensor(z["u_init"].copy())
    uo = u.clone()
    for k, (dt, nits_ref, ref) in enumerate(zip(z["loop_dts"], z["loop_nits"], z["loop_u"])):
        st = g.newton_solve(u, uo, float(dt))
        assert st.reason > 0 and abs(st.nits - int(nits_ref)) <= 1
        got = u.cpu().numpy()
        for f in range(pb.nf):
            assert np.abs(got[f] - ref[f]).max() <= 1e-8 * np.abs(ref[f]).max(), (k, f)
        if pb.nphase == 2:
            g.clip_field(u, 2, 0.0, 1.0)       # thermalmodel.py:226-229
        uo.copy_(u)
    g.close()


@pytest.mark.parametrize("name", sorted(LOOPS))
def test_free_running_time_loop_reaches_the_reference_end_state(name):
    """the whole loop (dt heuristics included) on the GPU: same end time; the end state agrees with the
    reference's to the time-discretisation sensitivity when a Newton count (hence a dt) differs."""
    from thermalporous_b200.model import run_time_loop, _TorchOps
    meta, pb, z = load(name)
    g = engine_from_problem(pb)
    g.set_solver_opts(snes_rtol=1e-12, snes_stol=1e-13, ksp_rtol=1e-10)
    u = g.tensor(z["u_init"].copy())
    uo = u.clone()
    res = run_time_loop(lambda a, b, dt: g.newton_solve(a, b, dt), _TorchOps(g), u, uo, two_phase=pb.nphase == 2,
                        i_S=2, **LOOPS[name])
    assert res.t == pytest.approx(float(np.sum(z["loop_dts"])), rel=1e-12)
    same_path = len(res.dt_vec) == len(z["loop_dts"]) and np.allclose(res.dt_vec, z["loop_dts"], rtol=1e-13)
    final = u.cpu().numpy()
    if same_path:
        ref = z["u_final"]
    else:
        # a Newton count at the 1e-12 tolerance edge differed, so the dt heuristic took another path: compare
        # with the CPU restatement driven through the dt sequence the GPU run actually took
        c = cport.engine_from_problem(pb)
        c.set_solver_opts(snes_rtol=1e-12, snes_stol=1e-13, ksp_rtol=1e-10)
        ref = np.ascontiguousarray(z["u_init"], dtype=np.float64).copy()
        for dt in res.dt_vec:
            st = c.newton_solve(ref, ref.copy(), dt)
            assert st.reason > 0
            if pb.nphase == 2:
                np.clip(ref[2], 0.0, 1.0, out=ref[2])
        c.close()
    for f in range(pb.nf):
        assert np.abs(final[f] - ref[f]).max() <= 1e-8 * np.abs(ref[f]).max()
    g.close()


def test_newton_host_entry_point_and_stats():
    meta, pb, z = load("l3_tp3d_heater_loop")
    g = engine_from_problem(pb)
    g.set_solver_opts(snes_rtol=1e-12, snes_stol=1e-13, ksp_rtol=1e-10)
    u = np.ascontiguousarray(z["u_init"], dtype=np.float64).copy()
    uo = u.copy()
    st = g.newton_solve_host(u, uo, float(z["loop_dts"][0]))
    assert st.reason > 0 and st.nits == int(z["loop_nits"][0]) and st.lits >= st.nits
    ref = z["loop_u"][0]
    for f in range(pb.nf):
        assert np.abs(u[f] - ref[f]).max() <= 1e-8 * np.abs(ref[f]).max()
    assert st.t_total_ms > 0 and g.launch_count() > 0
    g.close()


def test_snes_failure_is_reported_not_hidden():
    meta, pb, z = load("l2_tp2d_hetero_loop")
    g = engine_from_problem(pb)
    g.set_solver_opts(snes_max_it=1, snes_rtol=1e-14, snes_stol=0.0)
    u = g.tensor(z["u_init"].copy())
    st = g.newton_solve(u, u.clone(), 86400.0)
    assert st.reason == -5       # SNES_DIVERGED_MAX_IT
    g.close()


def test_total_oil_mass_reduction():
    """thermalmodel.py:190: assemble(phi * S_o * oil_rho(p, T) * dx) as a reduction kernel."""
    from oracle import tp_oracle as orc
    from tests.gpu_util import engine_from_problem, random_problem
    pb, u, uo = random_problem(3, 2, (7, 9, 11), seed=3)
    eng = engine_from_problem(pb)
    got = eng.oil_mass(eng.tensor(u))
    g = pb.grid
    want = g.dx * g.dy * g.dz * float(np.sum(pb.phi * u[2] * orc.oil_rho(pb.prm, u[0], u[1])))
    assert abs(got - want) <= 1e-13 * abs(want)
    eng.close()
