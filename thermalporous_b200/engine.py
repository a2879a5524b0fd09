"""Engine: one libtpb200 handle = one slab of the grid on one B200.

Thin host-side owner of the C-ABI handle.  Device memory is held in torch CUDA tensors
(torch is plumbing here: allocation and torch.distributed), raw pointers cross the ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class Engine:
    def __init__(self, dim, nx, ny, nz, dx, dy, dz, nphase, params, device=0, has_lo=False, has_hi=False,
                 gravity=True):
        if not torch.cuda.is_available():
            raise RuntimeError("thermalporous_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device("cuda", device)
        self.dim, self.nx, self.ny, self.nz = dim, int(nx), int(ny), int(nz)
        self.nphase = nphase
        self.nf = 2 if nphase == 1 else 3
        self.ns = 5 if dim == 2 else 7
        self.n = self.nx * self.ny * self.nz
        self.np_plane = self.nx * self.ny if dim == 3 else self.nx
        g = L.Grid(dim, self.nx, self.ny, self.nz, dx, dy, dz if dim == 3 else 1.0, int(has_lo), int(has_hi))
        p = L.Params(params.ko, params.kw, params.kr, params.c_v_w, params.c_v_o, params.c_r, params.rho_r,
                     params.T_inj, params.T_prod, params.API, params.g, params.S_o, params.U, int(gravity))
        self.h = C.c_void_p()
        rc = self.lib.tpb_create(C.byref(g), nphase, C.byref(p), device, C.byref(self.h))
        L.check(self.lib, None, rc)
        self._keep = {}
        self.opts = L.SolverOpts()
        self.lib.tpb_solver_defaults(nphase, C.byref(self.opts))

    # ------------------------------------------------------------------ helpers
    def _chk(self, rc):
        L.check(self.lib, self.h, rc)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.tpb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tensor(self, x):
        """host array / tensor -> contiguous fp64 CUDA tensor on this engine's device."""
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.float64).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.device)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def _in(self):
        torch.cuda.current_stream(self.device).synchronize()

    def sync(self):
        self._chk(self.lib.tpb_sync(self.h))

    # ------------------------------------------------------------------ problem data
    def set_field(self, fid, data):
        if np.isscalar(data):
            data = np.full(self.n, float(data))
        t = self.tensor(np.asarray(data.detach().cpu()) if isinstance(data, torch.Tensor) else data).reshape(-1)
        assert t.numel() == self.n, "field has %d values, slab has %d cells" % (t.numel(), self.n)
        self._in()
        self._chk(self.lib.tpb_set_field(self.h, fid, t.data_ptr(), 1))

    def set_field_ghost(self, fid, lo=None, hi=None):
        tl = self.tensor(lo).reshape(-1) if lo is not None else None
        th = self.tensor(hi).reshape(-1) if hi is not None else None
        self._in()
        self._chk(self.lib.tpb_set_field_ghost(self.h, fid, tl.data_ptr() if tl is not None else None,
                                               th.data_ptr() if th is not None else None, 1))

    def set_sources(self, sources):
        """sources: iterable of (cell, kind, weight, bhp, max_rate, const_rate)."""
        sources = list(sources)
        arr = (L.Source * max(len(sources), 1))()
        for k, s in enumerate(sources):
            arr[k] = L.Source(int(s[0]), int(s[1]), int(bool(s[5])), float(s[2]), float(s[3]), float(s[4]))
        self._chk(self.lib.tpb_set_sources(self.h, len(sources), arr))

    def set_state_ghost(self, lo=None, hi=None):
        self._in()
        self._chk(self.lib.tpb_set_state_ghost(self.h, lo.data_ptr() if lo is not None else None,
                                               hi.data_ptr() if hi is not None else None))
        self.sync()

    # ------------------------------------------------------------------ kernels
    def assemble(self, u, u_old, dt, jacobian=True, F=None, J=None):
        u = self.tensor(u).reshape(self.nf, self.n)
        u_old = self.tensor(u_old).reshape(self.nf, self.n)
        F = self.empty(self.nf, self.n) if F is None else F
        if jacobian and J is None:
            J = self.empty(self.ns, self.nf, self.nf, self.n)
        self._in()
        self._chk(self.lib.tpb_assemble(self.h, u.data_ptr(), u_old.data_ptr(), float(dt), F.data_ptr(),
                                        J.data_ptr() if jacobian else None))
        self.sync()
        return (F, J) if jacobian else F

    def spmv(self, J, x, y=None):
        x = self.tensor(x).reshape(self.nf, self.n)
        y = self.empty(self.nf, self.n) if y is None else y
        self._in()
        self._chk(self.lib.tpb_spmv(self.h, J.data_ptr(), x.data_ptr(), y.data_ptr()))
        self.sync()
        return y

    # ------------------------------------------------------------------ solver
    def set_solver_opts(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opts, k):
                raise KeyError(k)
            setattr(self.opts, k, v)
        self._chk(self.lib.tpb_set_solver_opts(self.h, C.byref(self.opts)))

    def solver_opts(self):
        """the solver options currently set, as a dict"""
        return {name: getattr(self.opts, name) for name, _ in self.opts._fields_}

    def pc_setup(self, J, u, dt):
        u = self.tensor(u).reshape(self.nf, self.n)
        self._keep["pc"] = (J, u)   # the handle keeps raw pointers to both
        self._in()
        self._chk(self.lib.tpb_pc_setup(self.h, J.data_ptr(), u.data_ptr(), float(dt)))
        self.sync()

    def pc_apply(self, x, y=None):
        x = self.tensor(x).reshape(self.nf, self.n)
        y = self.empty(self.nf, self.n) if y is None else y
        self._in()
        self._chk(self.lib.tpb_pc_apply(self.h, x.data_ptr(), y.data_ptr()))
        self.sync()
        return y

    def ksp_solve(self, J, b, x=None):
        b = self.tensor(b).reshape(self.nf, self.n)
        x = self.empty(self.nf, self.n) if x is None else x
        its, reason, rn = C.c_int(), C.c_int(), C.c_double()
        self._in()
        self._chk(self.lib.tpb_ksp_solve(self.h, J.data_ptr(), b.data_ptr(), x.data_ptr(), C.byref(its),
                                         C.byref(reason), C.byref(rn)))
        self.sync()
        return x, its.value, reason.value, rn.value

    def newton_solve(self, u, u_old, dt):
        """in-place Newton solve on device tensors; returns the Stats struct."""
        st = L.Stats()
        self._in()
        self._chk(self.lib.tpb_newton_solve(self.h, u.data_ptr(), u_old.data_ptr(), float(dt), C.byref(st)))
        self.sync()
        return st

    def newton_solve_host(self, u_host, u_old_host, dt):
        """u_host / u_old_host: C-contiguous fp64 numpy arrays (nf, n); u_host is overwritten."""
        st = L.Stats()
        assert u_host.dtype == np.float64 and u_host.flags.c_contiguous
        assert u_old_host.dtype == np.float64 and u_old_host.flags.c_contiguous
        self._chk(self.lib.tpb_newton_solve_host(self.h, u_host.ctypes.data, u_old_host.ctypes.data, float(dt),
                                                 C.byref(st)))
        return st

    def field_minmax(self, u, f):
        out = (C.c_double * 2)()
        self._in()
        self._chk(self.lib.tpb_field_minmax(self.h, u.data_ptr(), f, out))
        return out[0], out[1]

    def oil_mass(self, u):
        """total oil mass over all ranks' slabs (two-phase), thermalmodel.py:190"""
        out = C.c_double()
        self._in()
        self._chk(self.lib.tpb_oil_mass(self.h, u.data_ptr(), C.byref(out)))
        return out.value

    def clip_field(self, u, f, lo, hi):
        self._in()
        self._chk(self.lib.tpb_clip_field(self.h, u.data_ptr(), f, float(lo), float(hi)))
        self.sync()

    # ------------------------------------------------------------------ PC introspection
    def mg_levels(self, which=0):
        out = []
        for l in range(self.lib.tpb_pc_mg_nlevels(self.h, which)):
            d = (C.c_int * 6)()
            self._chk(self.lib.tpb_pc_mg_level(self.h, which, l, C.byref(d), None))
            out.append(tuple(d))
        return out

    def mg_level_op(self, which, l):
        nx, ny, nz = self.mg_levels(which)[l][:3]
        a = self.empty(self.ns, nx * ny * nz)
        d = (C.c_int * 6)()
        self._in()
        self._chk(self.lib.tpb_pc_mg_level(self.h, which, l, C.byref(d), a.data_ptr()))
        return a

    def mg_apply(self, which, b):
        b = self.tensor(b).reshape(-1)
        y = torch.empty_like(b)
        self._in()
        self._chk(self.lib.tpb_pc_mg_apply(self.h, which, b.data_ptr(), y.data_ptr()))
        return y

    def stage2_apply(self, r):
        r = self.tensor(r).reshape(self.nf, self.n)
        z = torch.empty_like(r)
        self._in()
        self._chk(self.lib.tpb_pc_stage2_apply(self.h, r.data_ptr(), z.data_ptr()))
        return z

    def weights(self, f):
        w = self.empty(self.n)
        self._in()
        self._chk(self.lib.tpb_pc_get_weights(self.h, f, w.data_ptr()))
        return w

    def stream_ptr(self):
        """the handle's cudaStream_t (for torch.cuda.ExternalStream / event timing on that stream)."""
        return int(self.lib.tpb_stream(self.h))

    def launch_count(self):
        return int(self.lib.tpb_launch_count(self.h))

    def time_kernel(self, which, u, u_old, dt, F, J, x, y, reps=20):
        ms = C.c_double()
        self._in()
        self._chk(self.lib.tpb_time_kernel(self.h, which, u.data_ptr(), u_old.data_ptr(), float(dt), F.data_ptr(),
                                           J.data_ptr(), x.data_ptr(), y.data_ptr(), reps, C.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self.lib.tpb_comm_init(self.h, buf, rank, nranks))

    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._chk(self.lib.tpb_comm_unique_id(buf))
        return buf.raw

    def peer_mode(self) -> int:
        """bit mask of the exchanges that go through peer memory instead of NCCL (tpb_comm_peer_mode)"""
        return int(self.lib.tpb_comm_peer_mode(self.h))

    def exchange_static(self):
        self._chk(self.lib.tpb_exchange_static(self.h))
