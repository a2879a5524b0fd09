"""solver_parameters surface of the reference models -> libtpb200 solver options.

The reference selects its PETSc solver tree with `solver_parameters` = None | a named option set
(string) | a raw PETSc-options dict (singlephase.py:275-444, twophase.py:413-1002).  This module
keeps that surface: the same names resolve to the same two-stage structure
(stage 1 CPR / CPTR / Schur field-split with a V-cycle per scalar block, stage 2 ILU(0)),
realised by libtpb200's own multigrid and ILU kernels instead of hypre and PETSc.  Option sets that
need something outside the hot path (direct solvers, system AMG on the interleaved (p,T) block, FAS,
PatchPC) raise UnsupportedOption - there is no silent fall-back.
"""
from __future__ import annotations

KSP_GMRES, KSP_FGMRES = 0, 1
S1_NONE, S1_CPR, S1_CPTR, S1_FIELDSPLIT = 0, 1, 2, 3
DECOUP = {"No": 0, "QI": 1, "TI": 2, "QI_temp": 3, "TI_temp": 4}
SCHUR_CONVDIFF, SCHUR_A11, SCHUR_DIAG, SCHUR_SELFP = 0, 1, 2, 3
S2_NONE, S2_ILU0, S2_BJACOBI = 0, 1, 2


class UnsupportedOption(ValueError):
    pass


def _base(nphase):
    if nphase == 1:   # `newton`, singlephase.py:289-301
        return dict(snes_max_it=15, ksp_type=KSP_GMRES, ksp_max_it=200, ksp_restart=200, ksp_rtol=1e-5)
    # `newton_krylov`, twophase.py:416-433 (named sets start from newton_fas_krylov, :927; FAS is a
    # nonlinear-multigrid research branch - the Newton-Krylov part of it is what is kept)
    return dict(snes_max_it=25, ksp_type=KSP_FGMRES, ksp_max_it=200, ksp_restart=200, ksp_rtol=1e-8)


def _two_stage(stage1, decoup="No", schur=SCHUR_CONVDIFF, stage2=S2_ILU0):
    return dict(stage1=stage1, decoup=DECOUP[decoup], schur_pre=schur, stage2=stage2)


SINGLE_PHASE_SETS = {
    # name: (reference lines, options)
    "pc_fieldsplit_cd": ("singlephase.py:309-319", _two_stage(S1_FIELDSPLIT, schur=SCHUR_CONVDIFF, stage2=S2_NONE)),
    # selfp is formed on the 5|7-point stencil (products that leave it are lumped into the diagonal), include/tpb200.h
    "pc_fieldsplit_selfp": ("singlephase.py:322-329", _two_stage(S1_FIELDSPLIT, schur=SCHUR_SELFP, stage2=S2_NONE)),
    "pc_fieldsplit_a11": ("singlephase.py:331-338", _two_stage(S1_FIELDSPLIT, schur=SCHUR_A11, stage2=S2_NONE)),
    "pc_fieldsplit_diag": ("singlephase.py:371-375", _two_stage(S1_FIELDSPLIT, schur=SCHUR_DIAG, stage2=S2_NONE)),
    "pc_cpr": ("singlephase.py:341-351", _two_stage(S1_CPR)),
    "pc_cpr_QI": ("singlephase.py:353", _two_stage(S1_CPR, "QI")),
    "pc_cpr_TI": ("singlephase.py:354", _two_stage(S1_CPR, "TI")),
    "pc_cpr_gmres": ("singlephase.py:356-369", _two_stage(S1_CPR)),
    # solver_parameters=None resolves to the unmatched name "pc_fieldsplit" => bare GMRES with PETSc's
    # default PC, which is (block-Jacobi) ILU(0)  (singlephase.py:410-413)
    "pc_fieldsplit": ("singlephase.py:412-413", _two_stage(S1_NONE)),
}

TWO_PHASE_SETS = {
    "pc_cptr": ("twophase.py:531-550", _two_stage(S1_CPTR)),
    "pc_cptr_a11": ("twophase.py:599-617", _two_stage(S1_CPTR, schur=SCHUR_A11)),
    "pc_cptr_gmres": ("twophase.py:670-696", _two_stage(S1_CPTR)),
    "pc_cpr": ("twophase.py:582-592", _two_stage(S1_CPR)),
    "pc_cpr_QI": ("twophase.py:594", _two_stage(S1_CPR, "QI")),
    "pc_cpr_TI": ("twophase.py:595", _two_stage(S1_CPR, "TI")),
    "pc_cpr_QI_temp": ("twophase.py:596", _two_stage(S1_CPR, "QI_temp")),
    "pc_cpr_TI_temp": ("twophase.py:597", _two_stage(S1_CPR, "TI_temp")),
    "pc_cpr_gmres": ("twophase.py:619-634", _two_stage(S1_CPR)),
    "pc_ilu": ("twophase.py:734-737", _two_stage(S1_NONE)),
}

_UNSUPPORTED = {
    "pc_hypre": "system BoomerAMG on the coupled matrix", "pc_amg": "system BoomerAMG", "pc_ml": "ML",
    "pc_lu": "direct solve (MUMPS/PETSc LU)", "pc_mg": "geometric PCMG over a mesh hierarchy",
    "pc_cptramg": "system AMG on the interleaved (p,T) block", "pc_cptramg_QI": "system AMG on (p,T)",
    "pc_cptramg_TI": "system AMG on (p,T)", "pc_cptramg_gmres": "system AMG on (p,T)",
    "pc_cptrlu": "direct solve of the (p,T) block", "pc_cptrlu_QI": "direct solve", "pc_cptrlu_TI": "direct solve",
    "pc_cptrlu_gmres": "direct solve", "pc_cprmg_gmres": "PCMG over a mesh hierarchy",
    "pc_ilu": "ILU(1) in the single-phase model (pc_factor_levels 1, singlephase.py:388-391; ILU here is ILU(0))",
    "pc_bilu": "ILU(1) (sub_pc_factor_levels 1, singlephase.py:402-406, twophase.py:757-762; ILU here is ILU(0))",
    "pc_cprilu1_gmres": "ILU(1) second stage (twophase.py:653-668; the second stage here is ILU(0))",
    "faspardecomp": "FAS/PatchPC", "ngmresfaspardecomp": "FAS/PatchPC", "newtonaijfaspardecomp": "FAS/PatchPC",
    "newtonmgpardecomp": "FAS/PatchPC",
}

# keys of a raw PETSc dict that only switch monitoring / are implied
_IGNORED_KEYS = {"snes_monitor", "snes_converged_reason", "ksp_converged_reason", "ksp_view", "snes_view",
                 "ksp_monitor", "ksp_monitor_true_residual", "mat_type", "snes_type",
                 "ksp_pc_side", "pc_composite_type", "pc_composite_pcs", "sub_1_sub_pc_type", "sub_1_pc_type",
                 "sub_1_pc_bjacobi_blocks", "sub_pc_type", "pc_type", "sub_0_pc_type", "sub_0_pc_python_type",
                 "sub_0_cpr_decoup", "pc_fieldsplit_type", "pc_fieldsplit_schur_precondition",
                 "pc_fieldsplit_0_fields", "pc_fieldsplit_1_fields", "sub_0_pc_fieldsplit_0_fields",
                 "sub_0_pc_fieldsplit_1_fields", "sub_0_pc_fieldsplit_type"}
# keys handled explicitly below
_HANDLED_KEYS = {"ksp_type", "ksp_max_it", "ksp_gmres_restart", "ksp_rtol", "ksp_atol", "snes_max_it", "snes_rtol",
                 "snes_atol", "snes_stol", "snes_linesearch_type", "sub_1_sub_pc_factor_levels", "pc_factor_levels",
                 "sub_pc_factor_levels", "pc_fieldsplit_schur_fact_type", "sub_0_cpr_stage1_pc_fieldsplit_schur_fact_type",
                 "ksp_monitor_residuals"}
# The inner solvers of the first stage are libtpb200's own (one multigrid V-cycle per scalar block in the role of
# hypre, tpb_* knobs): their PETSc/hypre tuning keys describe solvers that do not exist here and are accepted.
_INNER_PREFIXES = ("sub_0_cpr_stage1_", "sub_0_fieldsplit_", "fieldsplit_", "pc_hypre_", "sub_0_pc_hypre_", "tpb_",
                   "sub_0_ksp_", "sub_1_ksp_", "sub_0_sub_", "sub_1_sub_ksp_", "snes_npc_", "npc_", "snes_linesearch_max",
                   "snes_linesearch_monitor", "pc_mg_", "mg_", "pc_ml_", "pc_gamg_")

# line searches: `basic` takes the full step; `bt` is the cubic back-tracking default of PETSc, realised as step halving
# on the same sufficient-decrease test.  l2 / cp are secant searches on a different merit function: not realised.
_LS = {"basic": 0, "bt": 1}


def _flatten(d, prefix=""):
    """Firedrake flattens nested dicts into PETSc prefixes ('sub_0_cpr_stage1': {...})."""
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + k + "_"))
        else:
            out[prefix + k] = v
    return out


def _from_dict(d, nphase):
    d = _flatten(d)
    o = {}
    pc_type = d.get("pc_type", None)
    py0 = str(d.get("sub_0_pc_python_type", ""))
    if pc_type == "composite":
        if py0.endswith("CPTRStage1PC"):
            if nphase != 2:
                raise UnsupportedOption("CPTRStage1PC needs the two-phase model (preconditioners.py:1258-1267)")
            inner = d.get("sub_0_cpr_stage1_pc_type", "fieldsplit")
            if inner != "fieldsplit":
                raise UnsupportedOption("CPTR inner PC %r (only the Schur field-split of pc_cptr is realised)" % inner)
            schur = {"a11": SCHUR_A11, "selfp": SCHUR_SELFP}.get(d.get("sub_0_cpr_stage1_pc_fieldsplit_schur_precondition"),
                                                                  SCHUR_CONVDIFF)
            o.update(_two_stage(S1_CPTR, d.get("sub_0_cpr_decoup", "No"), schur))
        elif py0.endswith("CPRStage1PC") or py0.endswith("CPRStage1PC_mat"):
            o.update(_two_stage(S1_CPR, d.get("sub_0_cpr_decoup", "No")))
        elif str(d.get("pc_composite_pcs", "")).startswith("fieldsplit"):
            f0 = str(d.get("sub_0_pc_fieldsplit_0_fields", "0"))
            if f0 == "0,1":
                if d.get("sub_0_fieldsplit_0_pc_type") != "fieldsplit":
                    raise UnsupportedOption("pc_cptr*_gmres variant with a non-fieldsplit (p,T) solver")
                o.update(_two_stage(S1_CPTR))
            else:
                o.update(_two_stage(S1_CPR))
        else:
            raise UnsupportedOption("composite PC without a CPR/CPTR first stage")
    elif pc_type == "fieldsplit":
        if nphase != 1:
            raise UnsupportedOption("top-level field-split option sets exist for the single-phase model only")
        if d.get("pc_fieldsplit_type", "schur") == "additive":
            o.update(_two_stage(S1_FIELDSPLIT, schur=SCHUR_DIAG, stage2=S2_NONE))
        else:
            pre = d.get("pc_fieldsplit_schur_precondition", None)
            schur = {"a11": SCHUR_A11, "selfp": SCHUR_SELFP}.get(pre, SCHUR_CONVDIFF)
            o.update(_two_stage(S1_FIELDSPLIT, schur=schur, stage2=S2_NONE))
    elif pc_type in ("ilu", "bjacobi", None):
        o.update(_two_stage(S1_NONE))
    elif pc_type == "jacobi" or pc_type == "pbjacobi":
        o.update(_two_stage(S1_NONE, stage2=S2_BJACOBI))
    elif pc_type == "none":
        o.update(_two_stage(S1_NONE, stage2=S2_NONE))
    else:
        raise UnsupportedOption("pc_type %r is outside the hot path (see DESIGN.md, out of scope)" % pc_type)
    # Krylov / Newton scalars
    if "ksp_type" in d:
        kt = d["ksp_type"]
        if kt not in ("gmres", "fgmres"):
            raise UnsupportedOption("ksp_type %r" % kt)
        o["ksp_type"] = KSP_FGMRES if kt == "fgmres" else KSP_GMRES
    for key, name, cast in (("ksp_max_it", "ksp_max_it", int), ("ksp_gmres_restart", "ksp_restart", int),
                            ("ksp_rtol", "ksp_rtol", float), ("ksp_atol", "ksp_atol", float),
                            ("snes_max_it", "snes_max_it", int), ("snes_rtol", "snes_rtol", float),
                            ("snes_atol", "snes_atol", float), ("snes_stol", "snes_stol", float)):
        if key in d:
            o[name] = cast(d[key])
    if "snes_linesearch_type" in d:
        if d["snes_linesearch_type"] not in _LS:
            raise UnsupportedOption("snes_linesearch_type %r (basic and bt are realised)" % d["snes_linesearch_type"])
        o["linesearch"] = _LS[d["snes_linesearch_type"]]
    for key in ("sub_1_sub_pc_factor_levels", "pc_factor_levels", "sub_pc_factor_levels"):
        if int(d.get(key, 0)) != 0:
            raise UnsupportedOption("%s=%s: the second stage is ILU(0); ILU with fill is not realised" % (key, d[key]))
    for key in ("pc_fieldsplit_schur_fact_type", "sub_0_cpr_stage1_pc_fieldsplit_schur_fact_type"):
        if str(d.get(key, "FULL")).upper() != "FULL":
            raise UnsupportedOption("%s=%s: only the FULL Schur factorisation is realised (twophase.py:539)" % (key, d[key]))
    if d.get("ksp_monitor_residuals") not in (None, False):
        raise UnsupportedOption("ksp_monitor_residuals: the per-field residual monitor (thermalmodel.py:44-74) is not realised; "
                                "use tpb_verbose=2 for the Krylov residual history")
    # our own multigrid / smoother knobs may be passed with a tpb_ prefix
    for k, v in d.items():
        if k.startswith("tpb_"):
            o[k[4:]] = v
    # no silent fall-back: a key that is neither handled, nor known to be implied, nor tuning of an inner solver that
    # libtpb200 replaces is an error (typos such as 'ksp_rtoll' used to be dropped)
    for k in d:
        if k in _HANDLED_KEYS or k in _IGNORED_KEYS or k.startswith(_INNER_PREFIXES):
            continue
        raise UnsupportedOption("unrecognised solver option %r" % k)
    return o


def resolve(solver_parameters, nphase):
    """-> (opts dict for Engine.set_solver_opts, decoup name, description string).

    Mirrors SinglePhase.init_solver_parameters (singlephase.py:410-444) and
    TwoPhase.init_solver_parameters (twophase.py:929-1002)."""
    table = SINGLE_PHASE_SETS if nphase == 1 else TWO_PHASE_SETS
    opts = _base(nphase)
    if solver_parameters is None:
        solver_parameters = "pc_fieldsplit" if nphase == 1 else "pc_cptr_gmres"   # singlephase.py:412, twophase.py:929
    if isinstance(solver_parameters, str):
        name = solver_parameters
        if name in _UNSUPPORTED and name not in table:
            raise UnsupportedOption("option set %r needs %s - outside the hot path" % (name, _UNSUPPORTED[name]))
        if name not in table:
            raise UnsupportedOption("unknown option set %r" % name)
        ref, o = table[name]
        opts.update(o)
        desc = "%s (%s)" % (name, ref)
    elif isinstance(solver_parameters, dict):
        opts.update(_from_dict(solver_parameters, nphase))
        desc = "PETSc options dict"
    else:
        raise TypeError("solver_parameters must be None, a name or a dict")
    inv = {v: k for k, v in DECOUP.items()}
    return opts, inv[opts.get("decoup", 0)], desc
