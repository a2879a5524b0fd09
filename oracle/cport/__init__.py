"""ctypes wrapper of the C/OpenMP CPU restatement (oracle/cport/tpb_cpu.c).

TEST INFRASTRUCTURE / CPU BASELINE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this; the product (thermalporous_b200) never does.
The struct definitions are those of include/tpb200.h, restated here (not imported from the product).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
SRC = os.path.join(_HERE, "tpb_cpu.c")
BUILD_DIR = os.path.join(os.path.dirname(_HERE), "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libtpb_cpu.so")

PHI, KX, KY, KZ, KT = 0, 1, 2, 3, 4
PROD, INJ, HEATER = 0, 1, 2
KSP_GMRES, KSP_FGMRES = 0, 1
S1_NONE, S1_CPR, S1_CPTR, S1_FIELDSPLIT = 0, 1, 2, 3
DECOUP = {"No": 0, "QI": 1, "TI": 2, "QI_temp": 3, "TI_temp": 4}
SCHUR_CONVDIFF, SCHUR_A11, SCHUR_DIAG, SCHUR_SELFP = 0, 1, 2, 3
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


def build(force=False):
    """gcc -O3 -fopenmp oracle/cport/tpb_cpu.c -> oracle/_build/libtpb_cpu.so"""
    if os.path.exists(LIB_PATH) and not force and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-std=gnu99", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "include"), "-o", LIB_PATH, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError:
        build(force=True)
        lib = C.CDLL(LIB_PATH)
    vp, dp, i, d = C.c_void_p, C.c_void_p, C.c_int, C.c_double
    lib.tpc_create.argtypes = [C.POINTER(Grid), i, C.POINTER(Params)]
    lib.tpc_create.restype = vp
    lib.tpc_destroy.argtypes = [vp]
    lib.tpc_set_field.argtypes = [vp, i, dp]
    lib.tpc_set_sources.argtypes = [vp, i, C.POINTER(Source)]
    lib.tpc_set_solver_opts.argtypes = [vp, C.POINTER(SolverOpts)]
    lib.tpc_assemble.argtypes = [vp, dp, dp, d, dp, dp]
    lib.tpc_spmv.argtypes = [vp, dp, dp, dp]
    lib.tpc_pc_setup.argtypes = [vp, dp, dp, d]
    lib.tpc_pc_apply.argtypes = [vp, dp, dp]
    lib.tpc_ksp_solve.argtypes = [vp, dp, dp, dp, C.POINTER(i), C.POINTER(i), C.POINTER(d)]
    lib.tpc_newton_solve.argtypes = [vp, dp, dp, d, C.POINTER(Stats)]
    lib.tpc_mg_nlevels.argtypes = [vp, i]
    lib.tpc_mg_level_dims.argtypes = [vp, i, i, C.POINTER(i * 6)]
    lib.tpc_mg_level_op.argtypes = [vp, i, i, dp]
    lib.tpc_mg_apply.argtypes = [vp, i, dp, dp]
    lib.tpc_stage2_apply.argtypes = [vp, dp, dp]
    lib.tpc_get_weights.argtypes = [vp, i, dp]
    lib.tpc_stream_triad_gbs.argtypes = [C.c_long, i]
    lib.tpc_stream_triad_gbs.restype = d
    _lib = lib
    return lib


def stream_triad_gbs(n=40_000_000, reps=5):
    """host STREAM triad bandwidth in GB/s with the threads the library currently uses (bench.py's CPU roofline)."""
    return float(load().tpc_stream_triad_gbs(int(n), int(reps)))


def default_opts(nphase):
    """same defaults as tpb_solver_defaults (singlephase.py:289-301, twophase.py:416-433)."""
    o = SolverOpts()
    o.snes_max_it = 15 if nphase == 1 else 25
    o.snes_rtol, o.snes_atol, o.snes_stol = 1e-8, 1e-50, 1e-8
    o.linesearch = 0
    o.ksp_type = KSP_GMRES if nphase == 1 else KSP_FGMRES
    o.ksp_max_it = o.ksp_restart = 200
    o.ksp_rtol = 1e-5 if nphase == 1 else 1e-8
    o.ksp_atol = 1e-50
    o.stage1 = S1_CPR if nphase == 1 else S1_CPTR
    o.decoup = 0
    o.schur_pre = SCHUR_CONVDIFF
    o.stage2 = S2_ILU0
    o.mg_pre = o.mg_post = 2
    o.mg_coarse_sweeps = 4
    o.mg_min_cells = 8
    o.mg_overcorrection = 1.0
    o.mg_cycles = 1
    o.mg_semi_theta = 0.5
    o.mg_full_below = 0
    o.mg_dd_stop = 0.1
    o.mg_coarse_scale = 0.5
    o.mg_smoother = 1
    o.mg_tile_sweeps = 0
    o.verbose = 0
    return o


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


class CpuEngine:
    """Same call surface as thermalporous_b200.engine.Engine, NumPy arrays in and out."""

    def __init__(self, dim, nx, ny, nz, dx, dy, dz, nphase, params, gravity=True):
        self.lib = load()
        self.dim, self.nx, self.ny, self.nz = dim, int(nx), int(ny), int(nz)
        self.nphase = nphase
        self.nf = 2 if nphase == 1 else 3
        self.ns = 5 if dim == 2 else 7
        self.n = self.nx * self.ny * self.nz
        g = Grid(dim, self.nx, self.ny, self.nz, dx, dy, dz if dim == 3 else 1.0, 0, 0)
        p = Params(params.ko, params.kw, params.kr, params.c_v_w, params.c_v_o, params.c_r, params.rho_r,
                   params.T_inj, params.T_prod, params.API, params.g, params.S_o, params.U, int(gravity))
        self.h = self.lib.tpc_create(C.byref(g), nphase, C.byref(p))
        self.opts = default_opts(nphase)
        self.lib.tpc_set_solver_opts(self.h, C.byref(self.opts))
        self._keep = None

    def close(self):
        if self.h:
            self.lib.tpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_field(self, fid, data):
        a = np.full(self.n, float(data)) if np.isscalar(data) else _arr(data).reshape(-1)
        assert a.size == self.n
        self.lib.tpc_set_field(self.h, fid, _p(a))

    def set_sources(self, sources):
        sources = list(sources)
        arr = (Source * max(len(sources), 1))()
        for k, s in enumerate(sources):
            arr[k] = Source(int(s[0]), int(s[1]), int(bool(s[5])), float(s[2]), float(s[3]), float(s[4]))
        self.lib.tpc_set_sources(self.h, len(sources), arr)

    def set_solver_opts(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opts, k):
                raise KeyError(k)
            setattr(self.opts, k, v)
        self.lib.tpc_set_solver_opts(self.h, C.byref(self.opts))

    def assemble(self, u, u_old, dt, jacobian=True):
        u = _arr(u, (self.nf, self.n))
        uo = _arr(u_old, (self.nf, self.n))
        F = np.empty((self.nf, self.n))
        J = np.empty((self.ns, self.nf, self.nf, self.n)) if jacobian else None
        self.lib.tpc_assemble(self.h, _p(u), _p(uo), float(dt), _p(F), _p(J) if jacobian else None)
        return (F, J) if jacobian else F

    def spmv(self, J, x):
        x = _arr(x, (self.nf, self.n))
        y = np.empty_like(x)
        self.lib.tpc_spmv(self.h, _p(J), _p(x), _p(y))
        return y

    def pc_setup(self, J, u, dt):
        self._keep = (_arr(J), _arr(u, (self.nf, self.n)))
        self.lib.tpc_pc_setup(self.h, _p(self._keep[0]), _p(self._keep[1]), float(dt))

    def pc_apply(self, x):
        x = _arr(x, (self.nf, self.n))
        y = np.empty_like(x)
        self.lib.tpc_pc_apply(self.h, _p(x), _p(y))
        return y

    def ksp_solve(self, J, b):
        J = _arr(J)
        b = _arr(b, (self.nf, self.n))
        x = np.empty_like(b)
        its, reason, rn = C.c_int(), C.c_int(), C.c_double()
        self.lib.tpc_ksp_solve(self.h, _p(J), _p(b), _p(x), C.byref(its), C.byref(reason), C.byref(rn))
        return x, its.value, reason.value, rn.value

    def newton_solve(self, u, u_old, dt):
        """u (nf, n) is updated in place; returns Stats."""
        assert u.dtype == np.float64 and u.flags.c_contiguous
        uo = _arr(u_old, (self.nf, self.n))
        st = Stats()
        self.lib.tpc_newton_solve(self.h, _p(u), _p(uo), float(dt), C.byref(st))
        return st

    # ---- introspection (component parity tests)
    def mg_levels(self, which=0):
        out = []
        for l in range(self.lib.tpc_mg_nlevels(self.h, which)):
            d = (C.c_int * 6)()
            self.lib.tpc_mg_level_dims(self.h, which, l, C.byref(d))
            out.append(tuple(d))
        return out

    def mg_level_op(self, which, l):
        nx, ny, nz = self.mg_levels(which)[l][:3]
        a = np.empty((self.ns, nx * ny * nz))
        self.lib.tpc_mg_level_op(self.h, which, l, _p(a))
        return a

    def mg_apply(self, which, b):
        b = _arr(b).reshape(-1)
        y = np.empty_like(b)
        self.lib.tpc_mg_apply(self.h, which, _p(b), _p(y))
        return y

    def stage2_apply(self, r):
        r = _arr(r, (self.nf, self.n))
        z = np.empty_like(r)
        self.lib.tpc_stage2_apply(self.h, _p(r), _p(z))
        return z

    def weights(self, f):
        w = np.empty(self.n)
        self.lib.tpc_get_weights(self.h, f, _p(w))
        return w

    def num_threads(self):
        return int(self.lib.tpc_num_threads())

    def set_num_threads(self, n):
        self.lib.tpc_set_num_threads(int(n))


def engine_from_problem(pb):
    """Build a CpuEngine from an oracle.tp_oracle.Problem."""
    g = pb.grid
    eng = CpuEngine(g.dim, g.nx, g.ny, g.nz, g.dx, g.dy, g.dz, pb.nphase, pb.prm, gravity=pb.gravity)
    eng.set_field(PHI, pb.phi)
    eng.set_field(KX, pb.Kx)
    eng.set_field(KY, pb.Ky)
    if g.dim == 3:
        eng.set_field(KZ, pb.Kz)
    if pb.nphase == 1:
        eng.set_field(KT, pb.kT)
    eng.set_sources([(s.cell, s.kind, s.weight, s.bhp, s.max_rate, s.const_rate) for s in pb.sources])
    return eng
