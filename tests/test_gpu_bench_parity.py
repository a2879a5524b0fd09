"""Parity ON THE BENCH WORKLOAD (SPE10-shaped synthetic 60x220x85, wells 'default', pc_cptr) - the configuration
bench.py times - not only on the small fixtures:

* assembly at a mid-run state of the time loop: F and J against the C restatement entry by entry (1e-12), F against
  the NumPy oracle (1e-12) and J against the oracle through directional derivatives J v = Im F(u + i h v) / h
  (complex step; forming the oracle's full Jacobian of 3.4 M unknowns would take minutes) - 1e-12 relative.
  At such a state the residual is a small difference of large terms (oil row: accumulation terms of 4e6 cancel to
  1e-7 while the row's largest entry, at the well, is 0.2), so the three implementations - which round rho(p, T)
  differently - differ by a few ulp OF THE TERMS, 1e-8 of the row's largest entry.  The residual is therefore
  compared on the scale of its terms, sum_g |J_fg(c) u_g(c)| over the diagonal block (measured: 1e-15);
* one Newton solve of the bench configuration from that state, both sides converged far below the default tolerances
  (SNES rtol 1e-12, KSP rtol 1e-10): converged fields within 1e-8 relative (north_star), GPU vs C restatement.
"""
import numpy as np
import pytest

import bench
from oracle import cport, tp_oracle as orc
from tests.golden_util import rel_err_rows
from thermalporous_b200 import _lib as L, cases as CS, options as O

pytestmark = pytest.mark.gpu
WARM_STEPS = 9          # dt has left the small_dt_start plateau, saturation and temperature fronts have formed


@pytest.fixture(scope="module")
def workload():
    from thermalporous_b200.engine import Engine
    from thermalporous_b200.model import run_time_loop, _TorchOps
    prm = bench.make_params()
    geo = bench.make_geo(prm)
    ent = CS.source_entries(CS.WellCase(prm, geo, well_case="default"), prm, geo)
    eng = Engine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
    for fid, arr in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
        eng.set_field(fid, arr)
    eng.set_sources(ent)
    opts, _, _ = O.resolve(bench.PC, 2)
    eng.set_solver_opts(**opts)
    n = eng.n
    u = eng.tensor(np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)]))
    uo = u.clone()
    kw = dict(end=1e9, maxdt=bench.MAXDT, small_dt_start=True, dt_init_fact=bench.DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
    rw = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), _TorchOps(eng), u, uo, max_steps=WARM_STEPS, **kw)
    assert rw.failed_solves == 0
    # a state in the MIDDLE of a step: u_old = converged step, u = one Newton iterate of the next step (so that the
    # Jacobian is taken where u != u_old and the upwind directions have settled)
    dt = rw.next_dt
    un = u.clone()
    eng.set_solver_opts(snes_max_it=1)
    eng.newton_solve(un, uo, dt)
    eng.set_solver_opts(**opts)
    cpu = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
    cpu.set_field(cport.PHI, geo.phi)
    cpu.set_field(cport.KX, geo.K_x)
    cpu.set_field(cport.KY, geo.K_y)
    cpu.set_field(cport.KZ, geo.K_z)
    cpu.set_sources(ent)
    cpu.set_solver_opts(**opts)
    yield dict(prm=prm, geo=geo, ent=ent, eng=eng, cpu=cpu, u=un, uo=uo, dt=dt, opts=opts)
    eng.close()
    cpu.close()


def test_assembly_on_the_bench_workload_matches_both_oracles(workload):
    w = workload
    eng, cpu, geo, prm = w["eng"], w["cpu"], w["geo"], w["prm"]
    F, J = eng.assemble(w["u"], w["uo"], w["dt"])
    Fg, Jg = F.cpu().numpy(), J.cpu().numpy()
    uh, uoh = w["u"].cpu().numpy(), w["uo"].cpu().numpy()
    assert np.abs(uh - uoh).max() > 0
    # the C restatement, entry by entry
    Fc, Jc = cpu.assemble(uh, uoh, w["dt"])
    terms = np.abs(Jc[0] * uh[None, :, :]).sum(axis=1).max(axis=1, keepdims=True)     # (3, 1): size of a row's terms

    def res_err(A, B):
        return float((np.abs(A - B) / terms).max())
    assert res_err(Fg, Fc) < 1e-12
    assert rel_err_rows(Jg.reshape(7 * 9, -1), Jc.reshape(7 * 9, -1)) < 1e-12
    # the NumPy oracle: residual, and the Jacobian's action by complex-step directional derivatives
    g = orc.Grid(geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 3)
    pp = orc.Params(**{k: getattr(prm, k) for k in orc.Params().__dict__ if hasattr(prm, k)})
    srcs = [orc.Source(int(s[0]), int(s[1]), float(s[2]), float(s[3]), float(s[4]), bool(s[5])) for s in w["ent"]]
    pb = orc.Problem(grid=g, nphase=2, prm=pp, phi=geo.phi, Kx=geo.K_x, Ky=geo.K_y, Kz=geo.K_z, kT=None, sources=srcs)
    Fo = orc.residual(pb, uh, uoh, w["dt"])
    assert res_err(Fg, Fo) < 1e-12 and res_err(Fc, Fo) < 1e-12
    # ... and where the residual is not a cancellation (pressure and energy rows), on the plain scale of the row
    assert rel_err_rows(Fg[:2], Fo[:2]) < 1e-11
    rng = np.random.default_rng(7)
    x = eng.tensor(np.zeros_like(uh))
    for trial in range(2):
        v = rng.standard_normal(uh.shape) * np.array([1e-3, 1.0, 1e-2])[:, None]   # p in MPa, T in K, S
        Jv_oracle = np.imag(orc.residual(pb, uh + 1e-30j * v, uoh, w["dt"])) / 1e-30
        x.copy_(eng.tensor(v))
        Jv_gpu = eng.spmv(J, x).cpu().numpy()
        assert rel_err_rows(Jv_gpu, Jv_oracle) < 1e-12


def test_newton_step_on_the_bench_workload_gpu_vs_cpu_1e8(workload):
    w = workload
    eng, cpu = w["eng"], w["cpu"]
    tight = dict(snes_rtol=1e-12, snes_stol=1e-14, ksp_rtol=1e-10, snes_max_it=40)
    eng.set_solver_opts(**tight)
    cpu.set_solver_opts(**tight)
    ug = w["uo"].clone()
    st = eng.newton_solve(ug, w["uo"], w["dt"])
    uc = w["uo"].cpu().numpy().copy()
    sc = cpu.newton_solve(uc, w["uo"].cpu().numpy(), w["dt"])
    assert st.reason > 0 and sc.reason > 0
    assert abs(st.nits - sc.nits) <= 1
    # the hybrid smoother and the ILU are the same algorithms on both sides: Krylov counts agree closely
    assert abs(st.lits - sc.lits) <= max(3, 0.1 * sc.lits)
    ugh = ug.cpu().numpy()
    for f in range(3):
        assert np.abs(ugh[f] - uc[f]).max() <= 1e-8 * np.abs(uc[f]).max()
    eng.set_solver_opts(**w["opts"])
    cpu.set_solver_opts(**w["opts"])
