"""SinglePhase / TwoPhase / ThermalModel - host-side mirror of the reference's model classes
(singlephase.py:6-58, twophase.py:7-65, thermalmodel.py:7-412) on top of libtpb200.

Same constructor signatures, the same `solver_parameters` surface (options.py) and the same time
loop: `model.solve()` runs implicit-Euler steps, one Newton solve per step (`tpb_newton_solve`,
standing in for `self.solver.solve()` thermalmodel.py:165), halves dt on a ConvergenceError
(:170-180), re-solves once with dt/2 and clips when the saturation leaves [0,1] (:193-229), and
applies the SPE10 dt heuristic (:337-345).  State lives on the GPU between steps.

`run_time_loop` holds the loop logic on an abstract `newton`/`ops` pair so that it can be unit-tested
on the CPU (tests/) without a device; the models below always bind it to the CUDA Engine.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import options as O
from .cases import source_entries, source_rates

DAY = 24.0 * 3600.0


class ConvergenceError(RuntimeError):
    """Raised where Firedrake raises firedrake.exceptions.ConvergenceError (thermalmodel.py:170)."""

    def __init__(self, reason, msg=""):
        RuntimeError.__init__(self, "Nonlinear solve failed to converge: SNES reason %d %s" % (reason, msg))
        self.reason = reason


class LoopResult:
    def __init__(self):
        self.nits_vec, self.lits_vec, self.dt_vec, self.timings = [], [], [], []
        self.failed_solves = 0
        self.failed_time = 0.0
        self.failed = []          # (dt, SNES reason, nits, lits) of every failed attempt
        self.chops = 0
        self.t = 0.0
        self.stats = []

    @property
    def total_nits(self):
        return int(sum(self.nits_vec))

    @property
    def total_lits(self):
        return int(sum(self.lits_vec))


def run_time_loop(newton, ops, u, u_old, *, end, maxdt, small_dt_start, dt_init_fact, two_phase, i_S, spe10,
                  verbose=False, log=print, max_steps=None, dt0=None, after_step=None, oil_mass=False):
    """thermalmodel.py:97-348.  `newton(u, u_old, dt)` solves one step in place and returns an object with
    .nits .lits .reason (raises nothing); `ops` gives copy(dst, src), minmax(u, f) and clip(u, f, lo, hi).
    Times are in seconds except `end`/`maxdt` (days, as in the reference)."""
    res = LoopResult()
    dt = maxdt * DAY                                             # thermalmodel.py:13
    dt_init = dt_init_fact * maxdt * DAY                         # :97
    dt_inj = maxdt * DAY                                         # :98
    if small_dt_start:
        dt = dt_init                                             # :101-102
    if dt0 is not None:
        dt = dt0
    end_s = end * DAY
    t = 0.0
    i = 0

    def solve_or_raise(dt_now):
        t0 = time.perf_counter()
        st = newton(u, u_old, dt_now)
        el = time.perf_counter() - t0
        if st.reason < 0:
            res.failed_solves += 1
            res.failed_time += el
            res.failed.append((dt_now, st.reason, st.nits, st.lits))
            raise ConvergenceError(st.reason)
        return st, el

    while t < end_s and (max_steps is None or i < max_steps):
        i += 1
        if verbose:
            log("Time: %g days. Time-step %d. dt size: %g" % (t / DAY, i, dt / DAY))
        while True:                                              # :162-181
            try:
                st, el = solve_or_raise(dt)
                res.timings.append(el)
            except ConvergenceError:
                dt *= 0.5
                if verbose:
                    log("Time: %g days. Time-step %d. New dt size: %g" % (t / DAY, i, dt / DAY))
                ops.copy(u, u_old)
                continue
            break
        if two_phase:                                            # :184-229
            if oil_mass:                                         # :190-192 (a reduction over all ranks: every rank calls)
                mass_o = ops.oil_mass(u)
                if verbose:
                    log("Total oil mass in reservoir: ", mass_o)
            eps = 1e-10
            smin, smax = ops.minmax(u, i_S)
            chop = (smax - 1.0 > eps) or (smin < -eps)
            while chop:
                if verbose:
                    log("------Negative saturation! Chopping time-step---------")
                res.chops += 1
                dt *= 0.5
                ops.copy(u, u_old)
                try:
                    st, el = solve_or_raise(dt)
                except ConvergenceError:
                    dt *= 0.5
                    ops.copy(u, u_old)
                    continue
                smin = smax = None                               # the re-solved field's bounds are not known
                break                                            # :218 (the re-check below it is dead code)
            # :226-229 clips S_o to [0, 1] after every step; the bounds just computed say when that is the identity
            # (one pass over the field saved - on host buffers that is 2 % of a step)
            if smin is None or not (smin >= 0.0 and smax <= 1.0):
                ops.clip(u, i_S, 0.0, 1.0)
        ops.copy(u_old, u)                                       # :296
        t += dt
        res.dt_vec.append(dt)
        res.nits_vec.append(st.nits)                             # :327-336
        res.lits_vec.append(st.lits)
        res.stats.append(st)
        if after_step is not None:
            after_step(t)
        if verbose:
            log("Nonlinear iterations: %d\nLinear iterations: %d" % (st.nits, st.lits))
        current_dt = dt
        if spe10:                                                # :337-345
            if st.nits < 6:
                factor = 1 + min(1.0, (6 - st.nits) ** 2 / 3 ** 2)
                dt = min(dt_inj, current_dt * factor)
            elif st.nits > 9:
                factor = 1 - min(1.0, (st.nits - 9) ** 2 / 4 ** 2) / 2
                dt = current_dt * factor
        if dt > end_s - t and t < end_s:                         # :346-348
            dt = end_s - t
    res.t = t
    res.next_dt = dt
    return res


def rate_lines(rates, sources_case):
    """the lines thermalmodel.py:231-270 prints after a step, in its order: a SourceTerms case reports injection,
    production, oil, water; a case with individual wells injection, then water, oil and total production"""
    fmt = {"inj": "Total injection rate is ", "prod": "Total production rate is ",
           "oil": "Total oil production rate is ", "water": "Total water production rate is "}
    order = ("inj", "prod", "oil", "water") if sources_case else ("inj", "water", "oil", "prod")
    return [fmt[k] + repr(rates[k]) for k in order if rates.get(k) is not None]


def save_checkpoint(name, u):
    """store the solution fields (nf, ncell) - stands in for DumbCheckpoint.store (thermalmodel.py:361-364)."""
    np.savez(name if name.endswith(".npz") else name + ".npz", solution=np.asarray(u))


def load_checkpoint(name, shape=None):
    z = np.load(name if name.endswith(".npz") else name + ".npz")
    u = z["solution"]
    if shape is not None and tuple(u.shape) != tuple(shape):
        raise ValueError("checkpoint %s holds %s, the model needs %s" % (name, u.shape, tuple(shape)))
    return u


class _TorchOps:
    def __init__(self, engine):
        self.e = engine

    def copy(self, dst, src):
        dst.copy_(src)

    def minmax(self, u, f):
        return self.e.field_minmax(u, f)

    def oil_mass(self, u):
        return self.e.oil_mass(u)

    def clip(self, u, f, lo, hi):
        self.e.clip_field(u, f, lo, hi)


class ThermalModel:
    """thermalmodel.py:7-412 with the Firedrake solver replaced by a libtpb200 handle."""
    nphase = 1

    def __init__(self, end=1.0, maxdt=0.005, save=False, n_save=2, small_dt_start=True, checkpointing=None,
                 filename="results/results.txt", dt_init_fact=2 ** (-10), verbosity=True, device=None):
        from .engine import Engine
        from . import _lib as L
        # thermalmodel.py:113-133,304-320: every n_save-th step the fields go to results/<field>.pvd; here a .pvd
        # collection of VTK ImageData (.vti) files per field, written by rank 0's slab owner for its own slab
        self.save, self.n_save = bool(save), int(n_save)
        # thermalmodel.py:20-21; the reference stores the solution with DumbCheckpoint (HDF5) - here a .npz of
        # the fields in the C-ABI cell order
        self.checkpointing = {"save": False, "load": False, "savename": "initial", "loadname": "initial"}
        self.checkpointing.update(checkpointing or {})
        self.maxdt, self.dt_init_fact, self.end, self.verbosity = maxdt, dt_init_fact, end, verbosity
        self.filename = filename
        self._results_file = None
        geo, prm = self.geo, self.params
        # One process per GPU: under `torchrun` (torch.distributed initialised with the nccl backend) every rank
        # owns one z-slab (y-slab in 2-D) of the geo, as every MPI rank owns a mesh partition in the reference
        # (mesh.comm, singlephase.py:13); a single process owns the whole grid.
        from .partition import Slab
        self.rank, self.world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        # The reference constructors have no device argument: under torchrun every rank must still end up on its own
        # GPU.  device=None resolves to LOCAL_RANK (torchrun) when there are several ranks, else to torch's current
        # device; ranks of one node sharing a device are refused before ncclCommInitRank would hang on them.
        if device is None:
            import torch
            device = int(os.environ["LOCAL_RANK"]) if (self.world > 1 and "LOCAL_RANK" in os.environ) else torch.cuda.current_device()
        if self.world > 1:
            import socket
            import torch.distributed as dist
            mine = (socket.gethostname(), int(device))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            if len(set(everyone)) != len(everyone):
                raise RuntimeError("ranks share a GPU: %r - pass device=LOCAL_RANK (INTEGRATION.md)" % (everyone,))
        self.slab = slab = Slab(geo, self.world, self.rank)
        nxl, nyl, nzl = slab.local_dims()
        self.engine = Engine(geo.dim, nxl, nyl, nzl, geo.Dx, geo.Dy, getattr(geo, "Dz", 1.0), self.nphase, prm,
                             device=device, has_lo=slab.has_lo, has_hi=slab.has_hi)
        e = self.engine
        e.set_field(L.TPB_PHI, slab.take(geo.phi))
        e.set_field(L.TPB_KX, slab.take(geo.K_x))
        e.set_field(L.TPB_KY, slab.take(geo.K_y))
        if geo.dim == 3:
            e.set_field(L.TPB_KZ, slab.take(geo.K_z))
        if self.nphase == 1:
            e.set_field(L.TPB_KT, slab.take(geo.kT))
        self._entries = slab.localize_sources(source_entries(self.case, prm, geo))
        e.set_sources(self._entries)
        src_cells = np.array([s[0] for s in self._entries], dtype=np.int64)
        self._src_K = (slab.take(geo.K_x)[src_cells], slab.take(geo.K_y)[src_cells])   # for well_totals()
        if self.world > 1:
            import torch.distributed as dist
            uid = [e.unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.comm_init(uid[0], self.rank, self.world)
            e.exchange_static()
        e.set_solver_opts(**self.solver_opts)
        try:                                                             # thermalmodel.py:23-26
            ic = self.case.init_IC(phases=self.name)
        except AttributeError:
            ic = self.init_IC_uniform()
        self.initial_condition = slab.take(ic)                           # this rank's part
        self.u = e.tensor(self.initial_condition)
        self.u_ = self.u.clone()
        self.total_nits = self.total_lits = 0
        self.last_dt = None
        self.result = None

    def resultprint(self, *args):
        """thermalmodel.py:78-80: to the screen and to the results file (appended to when it is the default
        results/results.txt, overwritten otherwise, :28-32).  Rank 0 only; the file is opened on first use, so a
        quiet model (verbosity=False) never touches the file system."""
        if not (self.verbosity and self.rank == 0):
            return
        print(*args)
        if self._results_file is None:
            d = os.path.dirname(self.filename)
            if d:
                os.makedirs(d, exist_ok=True)
            self._results_file = open(self.filename, "a" if self.filename == "results/results.txt" else "w")
        print(*args, file=self._results_file)
        self._results_file.flush()

    def _rank_suffix(self):
        return "" if self.world == 1 else "_rank%dof%d" % (self.rank, self.world)

    def solve(self, max_steps=None):
        e = self.engine
        if self.checkpointing["load"]:                           # thermalmodel.py:87-91
            self.resultprint("Using as initial solution checkpoint " + self.checkpointing["loadname"])
            self.u.copy_(e.tensor(load_checkpoint(self.checkpointing["loadname"] + self._rank_suffix(), self.u.shape)))
        else:
            self.u.copy_(e.tensor(self.initial_condition))
        self.u_.copy_(self.u)
        writer = None
        if self.save:
            from .vtkout import PvdWriter
            names = ["pressure", "temperature"] + (["saturation_o"] if self.nphase == 2 else [])   # :113-133
            writer = PvdWriter("results", names, self.geo, self.slab, self._rank_suffix())
            writer.write(0.0, self.fields())
        state = {"i_plot": 0}

        def after_step(t):
            if self.verbosity:                                           # thermalmodel.py:231-270
                for line in rate_lines(self.well_totals(), self.case.name.startswith("Sources")):
                    self.resultprint(line)
            if writer is not None:                                       # thermalmodel.py:304-320
                if state["i_plot"] % self.n_save == 0:
                    writer.write(t, self.fields())
                state["i_plot"] += 1

        res = run_time_loop(lambda u, uo, dt: e.newton_solve(u, uo, dt), _TorchOps(e), self.u, self.u_,
                            end=self.end, maxdt=self.maxdt, small_dt_start=self.small_dt_start,
                            dt_init_fact=self.dt_init_fact, two_phase=self.nphase == 2, i_S=2,
                            spe10=self.geo.name.startswith("SPE10"), verbose=self.verbosity and self.rank == 0,
                            max_steps=max_steps, after_step=after_step, oil_mass=self.verbosity and self.nphase == 2)
        if writer is not None:
            writer.close()
        self.result = res
        self.total_nits, self.total_lits = res.total_nits, res.total_lits
        self.last_dt = res.dt_vec[-1] if res.dt_vec else None
        if self.checkpointing["save"]:                           # thermalmodel.py:361-364
            save_checkpoint(self.checkpointing["savename"] + self._rank_suffix(), self.u.detach().cpu().numpy())
            self.resultprint("Saving checkpoint solution in " + self.checkpointing["savename"])
        if self.verbosity and res.dt_vec and self.rank == 0:    # thermalmodel.py:367-408
            p = self.resultprint
            p("nits = ", res.nits_vec, ";")
            p("lits = ", res.lits_vec, ";")
            p("dts = ", res.dt_vec, ";")
            p("timings = ", res.timings, ";")
            p("----------------------------------------------------------------------")
            p(self.name, "thermal model")
            p("Geo model: ", self.geo.name)
            p("Test case: ", self.case.name)
            p("Max time-step: ", self.maxdt)
            p("Final time: ", res.t / DAY)
            p("Solver Parameters")
            p("-----------------")
            sp = self.solver_parameters
            if isinstance(sp, dict):
                for x in sp:
                    p(x, ":", sp[x])
            p("realised as: ", self.solver_desc)
            if self.nphase == 1 and self.solver_opts.get("stage1") == O.S1_FIELDSPLIT:
                p("pressure-temperature ordering for fieldsplit")          # singlephase.py:283
            p(" ")
            p("Solver performance")
            p("------------------")
            p("Total CPU time (s):", sum(res.timings))
            n = len(res.dt_vec)
            p("Average Nonlinear iterations per time-step:", res.total_nits / n)
            p("Average Linear iterations per time-step: ", res.total_lits / n)
            p("Average Linear iteration per Nonlinear iteration: ", res.total_lits / max(res.total_nits, 1))
            p("Total Linear iterations: ", res.total_lits)
            p("Total Nonlinear iterations: ", res.total_nits)
            p("Number of time-steps: ", n)
            p("Last Nonlinear iterations:", res.nits_vec[-1])
            p("Last Linear iterations: ", res.lits_vec[-1])
            p("----------------------------------------------------------------------")
            p(" ")
        return res

    def well_totals(self):
        """volumetric injection / production (/ oil / water) totals at the current state, over all ranks' slabs
        (thermalmodel.py:231-270; cases.source_rates).  Reads the state at the source cells only."""
        import torch
        ent = self._entries
        keys = ("inj", "prod") + (("oil", "water") if self.nphase == 2 else ())
        if ent:
            cells = np.array([e[0] for e in ent], dtype=np.int64)
            at = self.u[:, torch.as_tensor(cells, device=self.u.device)].detach().cpu().numpy()
            tot = source_rates(ent, at, self._src_K[0], self._src_K[1], self.params, self.nphase)
        else:
            tot = dict.fromkeys(keys)
        if self.world > 1:
            import torch.distributed as dist
            v = torch.tensor([[0.0 if tot[k] is None else tot[k], 0.0 if tot[k] is None else 1.0] for k in keys],
                             device=self.u.device, dtype=torch.float64)
            dist.all_reduce(v)
            tot = {k: (float(v[i, 0]) if float(v[i, 1]) > 0 else None) for i, k in enumerate(keys)}
        return tot

    def fields(self):
        """converged fields of THIS rank's slab as host arrays: (p, T[, S_o]); cells self.slab.c0 .. c1 of the geo."""
        return tuple(self.u.detach().cpu().numpy())


class SinglePhase(ThermalModel):
    """singlephase.py:6-58."""
    nphase = 1

    def __init__(self, geo, case, params, end=1.0, maxdt=0.005, save=False, n_save=2, small_dt_start=True,
                 checkpointing=None, solver_parameters=None, filename="results/results.txt",
                 dt_init_fact=2 ** (-10), vector=False, gravity2D=False, verbosity=True, device=0):
        self.name = "Single phase"
        self.geo, self.case, self.params = geo, case, params
        # vector=True makes (p,T) one VectorFunctionSpace in the reference (interleaved dofs, singlephase.py:16-17,
        # twophase.py:20-21) so that hypre can treat them as a system; the equations and the solution are the same.
        # The C-ABI layout is always field-major, so the flag is accepted and only recorded - the option sets that
        # NEED the interleaved block (pc_cptramg*, pc_cptrlu*) are rejected by options.resolve.
        if gravity2D:
            raise NotImplementedError("gravity2D is orientation-ill-defined on a non-extruded mesh (SURVEY.md appendix 5)")
        self.vector = bool(vector)
        self.small_dt_start = small_dt_start
        self.solver_parameters = solver_parameters
        self.solver_opts, self.decoup, self.solver_desc = O.resolve(solver_parameters, 1)
        ThermalModel.__init__(self, end, maxdt, save, n_save, small_dt_start, checkpointing, filename,
                              dt_init_fact, verbosity, device)

    def init_IC_uniform(self):
        n = self.geo.ncell
        return np.stack([np.full(n, self.params.p_ref), np.full(n, self.params.T_prod)])


class TwoPhase(ThermalModel):
    """twophase.py:7-65."""
    nphase = 2

    def __init__(self, geo, case, params, end=1.0, maxdt=0.005, save=False, n_save=2, small_dt_start=True,
                 checkpointing=None, solver_parameters=None, filename="results/results.txt",
                 dt_init_fact=2 ** (-10), vector=False, gravity2D=False, verbosity=True, device=0):
        self.name = "Two-phase"
        self.geo, self.case, self.params = geo, case, params
        # vector=True makes (p,T) one VectorFunctionSpace in the reference (interleaved dofs, singlephase.py:16-17,
        # twophase.py:20-21) so that hypre can treat them as a system; the equations and the solution are the same.
        # The C-ABI layout is always field-major, so the flag is accepted and only recorded - the option sets that
        # NEED the interleaved block (pc_cptramg*, pc_cptrlu*) are rejected by options.resolve.
        if gravity2D:
            raise NotImplementedError("gravity2D is orientation-ill-defined on a non-extruded mesh (SURVEY.md appendix 5)")
        self.vector = bool(vector)
        self.i_S_o = 2
        self.small_dt_start = small_dt_start
        self.solver_parameters = solver_parameters
        self.solver_opts, self.decoup, self.solver_desc = O.resolve(solver_parameters, 2)
        ThermalModel.__init__(self, end, maxdt, save, n_save, small_dt_start, checkpointing, filename,
                              dt_init_fact, verbosity, device)

    def init_IC_uniform(self):
        n = self.geo.ncell
        p = self.params
        return np.stack([np.full(n, p.p_ref), np.full(n, p.T_prod), np.full(n, p.S_o)])
