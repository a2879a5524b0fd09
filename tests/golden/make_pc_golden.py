#!/usr/bin/env python
"""Generate tests/golden/pc/decoup.npz by running the reference's OWN decoupling algebra.

Run in the build container only (needs /root/reference):

    python tests/golden/make_pc_golden.py

thermalporous/preconditioners.py builds its CPR / CPTR first stages from PETSc Mat/Vec calls on sub-blocks of the
assembled Jacobian: `CPRStage1PC.create_decoup_{QI,TI,QI_temp,TI_temp}` (preconditioners.py:684-873) and
`CPTRStage1PC.create_decoup_{QI,TI}` (:1445-1543).  Those methods are executed here UNMODIFIED: the module is imported
over tests/golden/fd_shim (whose firedrake.petsc implements the needed slice of petsc4py on scipy.sparse), an
instance is made without running `initialize` (which needs Firedrake's assembler), its `*mat` attributes are set to
the sub-blocks of a golden Jacobian (tests/golden/g*.npz, made by the reference's form code), and the method is
called.  What it leaves in `apsinvdss` / `Atildepp` (`a0sinvdss` / `Atilde00`) - the restriction weights and the
decoupled pressure operator - is stored in block-stencil layout next to the inputs' names.

The temperature Schur approximation is recorded the same way: `ConvDiffSchurPC.initialize` /
`ConvDiffSchurTwoPhasesPC.initialize` (:11-118, :165-286) build their frozen-coefficient convection-diffusion form
from the model's state and assemble it; over the shim the form is evaluated by the DG0 evaluator that produces the
residual fixtures (fd_shim/firedrake/assemble.py) and the assembled operator is stored.  The models are rebuilt from
the residual fixtures' own inputs (fields, parameters, state, dt) with the well/heater set-ups of make_golden.py.

The fixture travels to the GPU box; this script and the shim do not need to.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "fd_shim"))
sys.path.insert(0, os.environ.get("TPB_REFERENCE", "/root/reference"))
sys.path.insert(0, ROOT)

from firedrake.petsc import PETSc  # noqa: E402  (the shim)
from thermalporous import preconditioners as ref  # noqa: E402  (the reference, unmodified)
from oracle import tp_oracle as orc  # noqa: E402
from tests.golden_util import load  # noqa: E402


def blocks(J, g):
    """field-ordered sub-blocks A[f][c] (each n x n, scipy CSR) of a block-stencil Jacobian"""
    A = orc.to_csr(J, g, "field").tocsr()
    nf, n = J.shape[1], J.shape[3]
    return [[A[f * n:(f + 1) * n, c * n:(c + 1) * n] for c in range(nf)] for f in range(nf)], n


def to_stencil(M, g):
    """n x n sparse matrix with the grid's 5|7-point pattern -> a[s][cell]"""
    M = sp.csr_matrix(M)
    n = M.shape[0]
    cells = np.arange(n)
    ii, jj, kk = cells % g.nx, (cells // g.nx) % g.ny, cells // (g.nx * g.ny)
    offs = orc.stencil_offsets(g)
    out = np.zeros((len(offs), n))
    for s, (di, dj, dk) in enumerate(offs):
        ni, nj, nk = ii + di, jj + dj, kk + dk
        ok = (ni >= 0) & (ni < g.nx) & (nj >= 0) & (nj < g.ny) & (nk >= 0) & (nk < g.nz)
        nb = (ni + g.nx * (nj + g.ny * nk))[ok]
        out[s, ok] = np.asarray(M[cells[ok], nb]).ravel()
    # nothing outside the stencil
    assert abs(abs(M).sum() - np.abs(out).sum()) <= 1e-9 * abs(M).sum()
    return out


def cpr(B, n, decoup, two_phase):
    """CPRStage1PC: p = field 0, s = the last field (T single-phase, S_o two-phase); *_temp: nonp = (T, S_o)"""
    pc = object.__new__(ref.CPRStage1PC)
    pc.decoup = decoup
    pc.Appmat = PETSc.Mat(B[0][0])
    s = 2 if two_phase else 1
    if decoup.endswith("temp"):
        T, S = 1, 2
        pc.Assmat = PETSc.Mat(sp.bmat([[B[T][T], B[T][S]], [B[S][T], B[S][S]]]))
        pc.Aspmat = PETSc.Mat(sp.bmat([[B[T][0]], [B[S][0]]]))
        pc.Apsmat = PETSc.Mat(sp.bmat([[B[0][T], B[0][S]]]))
        pc.ASSmat, pc.ASTmat = PETSc.Mat(B[S][S]), PETSc.Mat(B[S][T])
        pc.ATSmat, pc.ATTmat = PETSc.Mat(B[T][S]), PETSc.Mat(B[T][T])
        pc.ApSmat, pc.ApTmat = PETSc.Mat(B[0][S]), PETSc.Mat(B[0][T])
        # W22 = V*V field index sets (preconditioners.py:443-451): T dofs first, then S_o
        pc.TT_is, pc.SS_is, pc.pp_is = PETSc.IS(np.arange(n)), PETSc.IS(np.arange(n, 2 * n)), PETSc.IS(np.arange(n))
    else:
        pc.Assmat, pc.Aspmat, pc.Apsmat = PETSc.Mat(B[s][s]), PETSc.Mat(B[s][0]), PETSc.Mat(B[0][s])
    getattr(pc, "create_decoup_" + decoup)(None)
    W = sp.csr_matrix(pc.apsinvdss.m)
    if decoup.endswith("temp"):
        w = np.stack([W[np.arange(n), np.arange(n)].A1, W[np.arange(n), n + np.arange(n)].A1])   # w_T, w_S
    else:
        w = W.diagonal()[None, :]
        assert abs(abs(W).sum() - np.abs(w).sum()) <= 1e-12 * abs(W).sum()
    return w, pc.Atildepp.m


def cptr(B, n, decoup):
    """CPTRStage1PC: '0' = (p, T) interleaved as the reference's D0s indexing (2i, 2i+1) assumes, s = S_o"""
    il = np.empty(2 * n, dtype=np.int64)
    il[0::2], il[1::2] = np.arange(n), n + np.arange(n)          # interleaved position -> field-ordered position
    P = sp.csr_matrix((np.ones(2 * n), (np.arange(2 * n), il)), shape=(2 * n, 2 * n))
    A00 = P @ sp.bmat([[B[0][0], B[0][1]], [B[1][0], B[1][1]]]).tocsr() @ P.T
    A0s = P @ sp.bmat([[B[0][2]], [B[1][2]]]).tocsr()
    As0 = sp.bmat([[B[2][0], B[2][1]]]).tocsr() @ P.T
    pc = object.__new__(ref.CPTRStage1PC)
    pc.decoup = decoup
    pc.A00mat, pc.A0smat, pc.As0mat, pc.Assmat = PETSc.Mat(A00), PETSc.Mat(A0s), PETSc.Mat(As0), PETSc.Mat(B[2][2])
    pc.A0smat.setBlockSizes(2, 1)
    getattr(pc, "create_decoup_" + decoup)(None)
    W = sp.csr_matrix(pc.a0sinvdss.m)
    w = np.stack([W[2 * np.arange(n), np.arange(n)].A1, W[2 * np.arange(n) + 1, np.arange(n)].A1])   # w_p, w_T
    At = sp.csr_matrix(pc.Atilde00.m)
    return w, At[0::2, 0::2]


class _FakePC:
    """what initialize() asks of the PETSc PC before it assembles"""

    def __init__(self, appctx):
        self.appctx = appctx

    def getOptionsPrefix(self):
        return "fieldsplit_1_"

    def getOperators(self):
        return None, None


def rebuild(name):
    """the reference model of a residual fixture, rebuilt from the fixture's own inputs (fields, parameters, state, dt)
    with the well/heater set-ups of make_golden.py; checked to reproduce the fixture's residual"""
    import firedrake as fd
    import make_golden as mg
    meta, pb, z = load(name)
    g = pb.grid
    prm = mg.fresh_params(**meta["params"])
    shape = (g.nz, g.ny, g.nx)
    fields = tuple(np.asarray(z[k]).reshape(shape) if z[k].size else None for k in ("phi", "Kx", "Ky", "Kz"))
    two = meta["nphase"] == 2
    Model = mg.TwoPhase if two else mg.SinglePhase
    if name.startswith("g1_"):
        geo = mg.HomogeneousGeo(g.nx, g.ny, prm, g.nx * g.dx, g.ny * g.dy)
        case = mg.WellCase(prm, geo, well_case="test0", constant_rate=True)
    elif g.dim == 2:
        geo = mg.HeteroGeo2D(g.nx, g.ny, prm, fields, dx=g.dx, dy=g.dy)
        pts_p, pts_i = [[2.3 * 6.096, 4.6 * 3.048]], [[9.4 * 6.096, 7.2 * 3.048]]
        case = mg.WellCase(prm, geo, prod_points=pts_p, inj_points=pts_i)
    else:
        geo = mg.HeteroGeo3D(g.nx, g.ny, g.nz, prm, fields, dx=g.dx, dy=g.dy, dz=g.dz)
        pp = [[1.5 * 6.096, 2.5 * 3.048, 0.5 * 0.6096]]
        ip = [[4.5 * 6.096, 1.5 * 3.048, 3.5 * 0.6096]]
        if meta["case"].startswith("Sources"):
            case = mg.SourceTerms(prm, geo, prod_points=pp + pp, inj_points=ip, heater_points=pp + ip)
        else:
            case = mg.WellHeaterCase(prm, geo, prod_points=pp, inj_points=ip)
    m = Model(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=mg.TMP, verbosity=False,
              solver_parameters=mg.sp(25 if two else 15))
    m.u.arr = np.asarray(z["u"]).reshape(m.u.arr.shape).copy()
    m.u_.arr = np.asarray(z["u_old"]).reshape(m.u_.arr.shape).copy()
    m.dt.assign(meta["dt"])
    F = np.real(fd.assemble(m.F)).reshape(z["F"].shape)
    assert np.abs(F - z["F"]).max() <= 1e-13 * np.abs(z["F"]).max(), name
    return m, case, two


def convdiff(name):
    """let the reference's PC class assemble its operator on the rebuilt model"""
    from firedrake.assemble import Assembled
    m, case, two = rebuild(name)
    appctx = dict(m.appctx)
    appctx["state"] = m.u
    pc = object.__new__(ref.ConvDiffSchurTwoPhasesPC if two else ref.ConvDiffSchurPC)
    try:
        pc.initialize(_FakePC(appctx))
        raise RuntimeError("initialize() returned without assembling")
    except Assembled:
        pass
    return pc.A.stencil


def rates(name):
    """the well totals of thermalmodel.py:231-270, from the rate expressions the reference's form code attached to its
    wells (singlephase.py:151-162, twophase.py:362-408); order: injection, production, oil, water (NaN = not reported)"""
    from firedrake import assemble, dx
    m, case, two = rebuild(name)
    out = [np.nan] * 4
    if case.name.startswith("Sources"):
        out[0] = assemble(case.deltas_inj * m.inj_rate * dx)
        out[1] = assemble(case.deltas_prod * m.prod_rate * dx)
        if two:
            out[2] = assemble(case.deltas_prod * m.oil_rate * dx)
            out[3] = assemble(case.deltas_prod * m.water_rate * dx)
    if case.inj_wells:
        out[0] = assemble(sum(w["delta"] * w["rate"] * dx for w in case.inj_wells))
    if case.prod_wells:
        out[1] = assemble(sum(w["delta"] * w["rate"] * dx for w in case.prod_wells))
        if two:
            out[3] = assemble(sum(w["delta"] * w["water_rate"] * dx for w in case.prod_wells))
            out[2] = assemble(sum(w["delta"] * w["oil_rate"] * dx for w in case.prod_wells))
    return np.array([float(np.real(v)) for v in out])


def main():
    out = {}
    for name in ("g1_sp2d_homo_const", "g2_sp2d_hetero_peaceman", "g3_tp2d_hetero_peaceman",
                 "g3b_tp2d_hetero_peaceman_uncapped", "g4_sp3d_hetero_wellheater", "g5_tp3d_hetero_wellheater",
                 "g5b_tp3d_hetero_wellheater_dp", "g6_tp3d_sources"):
        r = rates(name)
        out[name + "|rates|q"] = r
        print("%-48s inj %.6e  prod %.6e  oil %.6e  water %.6e" % ((name + "|rates",) + tuple(r)))
    for name in ("g2_sp2d_hetero_peaceman", "g3_tp2d_hetero_peaceman", "g4_sp3d_hetero_wellheater",
                 "g5_tp3d_hetero_wellheater", "g6_tp3d_sources"):
        A = convdiff(name)
        out[name + "|convdiff|A"] = A
        print("%-48s A_T %s  |diag|max %.3e  offdiag/diag max %.3f" % (name + "|convdiff", A.shape, np.abs(A[0]).max(),
                                                                     (np.abs(A[1:]).sum(axis=0) / np.abs(A[0])).max()))
    for name, two_phase in (("g2_sp2d_hetero_peaceman", False), ("g3_tp2d_hetero_peaceman", True),
                            ("g5_tp3d_hetero_wellheater", True)):
        meta, pb, z = load(name)
        g = pb.grid
        B, n = blocks(z["J"], g)
        todo = [("cpr", d) for d in (("QI", "TI", "QI_temp", "TI_temp") if two_phase else ("QI", "TI"))]
        if two_phase:
            todo += [("cptr", "QI"), ("cptr", "TI")]
        for kind, decoup in todo:
            w, At = cpr(B, n, decoup, two_phase) if kind == "cpr" else cptr(B, n, decoup)
            key = "%s|%s|%s" % (name, kind, decoup)
            out[key + "|w"] = w
            out[key + "|App"] = to_stencil(At, g)
            print("%-48s weights %s  |w|max %.3e   App~ %s" % (key, w.shape, np.abs(w).max(), out[key + "|App"].shape))
    os.makedirs(os.path.join(HERE, "pc"), exist_ok=True)
    path = os.path.join(HERE, "pc", "decoup.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.0f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
