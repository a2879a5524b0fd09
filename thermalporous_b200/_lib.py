"""ctypes binding of libtpb200.so (include/tpb200.h).  No CPU fallback: if the shared
library or a CUDA device is missing the product path raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtpb200.so")

TPB_PHI, TPB_KX, TPB_KY, TPB_KZ, TPB_KT = 0, 1, 2, 3, 4
PROD, INJ, HEATER = 0, 1, 2
KSP_GMRES, KSP_FGMRES = 0, 1
S1_NONE, S1_CPR, S1_CPTR, S1_FIELDSPLIT = 0, 1, 2, 3
DECOUP = {"No": 0, "QI": 1, "TI": 2, "QI_temp": 3, "TI_temp": 4}
SCHUR_CONVDIFF, SCHUR_A11, SCHUR_DIAG = 0, 1, 2
S2_NONE, S2_ILU0, S2_BJACOBI = 0, 1, 2


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("has_lo", C.c_int), ("has_hi", C.c_int)]


class Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("ko", "kw", "kr", "c_v_w", "c_v_o", "c_r", "rho_r", "T_inj", "T_prod", "API", "g", "S_o", "U")] + \
               [("gravity", C.c_int)]


class Source(C.Structure):
    _fields_ = [("cell", C.c_int64), ("kind", C.c_int32), ("const_rate", C.c_int32),
                ("weight", C.c_double), ("bhp", C.c_double), ("max_rate", C.c_double)]


class SolverOpts(C.Structure):
    _fields_ = [("snes_max_it", C.c_int), ("snes_rtol", C.c_double), ("snes_atol", C.c_double),
                ("snes_stol", C.c_double), ("linesearch", C.c_int),
                ("ksp_type", C.c_int), ("ksp_max_it", C.c_int), ("ksp_restart", C.c_int),
                ("ksp_rtol", C.c_double), ("ksp_atol", C.c_double),
                ("stage1", C.c_int), ("decoup", C.c_int), ("schur_pre", C.c_int), ("stage2", C.c_int),
                ("mg_pre", C.c_int), ("mg_post", C.c_int), ("mg_coarse_sweeps", C.c_int),
                ("mg_min_cells", C.c_int), ("mg_overcorrection", C.c_double), ("mg_cycles", C.c_int),
                ("mg_semi_theta", C.c_double), ("mg_full_below", C.c_int), ("mg_dd_stop", C.c_double),
                ("mg_coarse_scale", C.c_double), ("mg_smoother", C.c_int), ("mg_tile_sweeps", C.c_int),
                ("verbose", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("nits", C.c_int), ("lits", C.c_int), ("reason", C.c_int), ("nfev", C.c_int),
                ("fnorm0", C.c_double), ("fnorm", C.c_double),
                ("t_assemble_ms", C.c_double), ("t_pcsetup_ms", C.c_double), ("t_ksp_ms", C.c_double),
                ("t_total_ms", C.c_double)]


# every symbol include/tpb200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "tpb_create", "tpb_destroy", "tpb_last_error", "tpb_version", "tpb_set_field", "tpb_set_field_ghost",
    "tpb_set_sources", "tpb_assemble", "tpb_set_state_ghost", "tpb_jacobian_size", "tpb_nstencil", "tpb_spmv",
    "tpb_solver_defaults", "tpb_set_solver_opts", "tpb_pc_setup", "tpb_pc_apply", "tpb_ksp_solve",
    "tpb_newton_solve", "tpb_newton_solve_host", "tpb_field_minmax", "tpb_clip_field", "tpb_oil_mass", "tpb_dot",
    "tpb_comm_init", "tpb_comm_unique_id", "tpb_exchange_static", "tpb_comm_peer_mode", "tpb_launch_count", "tpb_time_kernel",
    "tpb_stream", "tpb_sync", "tpb_pc_mg_nlevels", "tpb_pc_mg_level", "tpb_pc_mg_apply", "tpb_pc_stage2_apply",
    "tpb_pc_get_weights",
]

_lib = None


class TpbError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libtpb200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen libtpb200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libtpb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "thermalporous_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, dp, i, d = C.c_void_p, C.c_void_p, C.c_int, C.c_double
    lib.tpb_create.argtypes = [C.POINTER(Grid), i, C.POINTER(Params), i, C.POINTER(vp)]
    lib.tpb_destroy.argtypes = [vp]
    lib.tpb_last_error.argtypes = [vp]
    lib.tpb_last_error.restype = C.c_char_p
    lib.tpb_set_field.argtypes = [vp, i, dp, i]
    lib.tpb_set_field_ghost.argtypes = [vp, i, dp, dp, i]
    lib.tpb_set_sources.argtypes = [vp, i, C.POINTER(Source)]
    lib.tpb_assemble.argtypes = [vp, dp, dp, d, dp, dp]
    lib.tpb_set_state_ghost.argtypes = [vp, dp, dp]
    lib.tpb_jacobian_size.argtypes = [vp]
    lib.tpb_jacobian_size.restype = C.c_size_t
    lib.tpb_nstencil.argtypes = [vp]
    lib.tpb_spmv.argtypes = [vp, dp, dp, dp]
    lib.tpb_solver_defaults.argtypes = [i, C.POINTER(SolverOpts)]
    lib.tpb_set_solver_opts.argtypes = [vp, C.POINTER(SolverOpts)]
    lib.tpb_pc_setup.argtypes = [vp, dp, dp, d]
    lib.tpb_pc_apply.argtypes = [vp, dp, dp]
    lib.tpb_ksp_solve.argtypes = [vp, dp, dp, dp, C.POINTER(i), C.POINTER(i), C.POINTER(d)]
    lib.tpb_newton_solve.argtypes = [vp, dp, dp, d, C.POINTER(Stats)]
    lib.tpb_newton_solve_host.argtypes = [vp, dp, dp, d, C.POINTER(Stats)]
    lib.tpb_field_minmax.argtypes = [vp, dp, i, C.POINTER(d)]
    lib.tpb_oil_mass.argtypes = [vp, dp, C.POINTER(d)]
    lib.tpb_clip_field.argtypes = [vp, dp, i, d, d]
    lib.tpb_dot.argtypes = [vp, dp, dp, C.c_size_t, C.POINTER(d)]
    lib.tpb_comm_init.argtypes = [vp, vp, i, i]
    lib.tpb_comm_unique_id.argtypes = [vp]
    lib.tpb_exchange_static.argtypes = [vp]
    lib.tpb_comm_peer_mode.argtypes = [vp]
    lib.tpb_launch_count.argtypes = [vp]
    lib.tpb_launch_count.restype = C.c_int64
    lib.tpb_time_kernel.argtypes = [vp, i, dp, dp, d, dp, dp, dp, dp, i, C.POINTER(d)]
    lib.tpb_stream.argtypes = [vp]
    lib.tpb_stream.restype = vp
    lib.tpb_sync.argtypes = [vp]
    lib.tpb_pc_mg_nlevels.argtypes = [vp, i]
    lib.tpb_pc_mg_level.argtypes = [vp, i, i, C.POINTER(i * 6), dp]
    lib.tpb_pc_mg_apply.argtypes = [vp, i, dp, dp]
    lib.tpb_pc_stage2_apply.argtypes = [vp, dp, dp]
    lib.tpb_pc_get_weights.argtypes = [vp, i, dp]
    _lib = lib
    return lib


def check(lib, handle, code):
    if code != 0:
        msg = lib.tpb_last_error(handle)
        raise TpbError(code, msg.decode() if msg else "?")
