#!/usr/bin/env python
"""bench.py - SPE10 3-D two-phase thermal timestepping on B200 (BASELINE.json metric).

metric   Mcell-Newton-iters/s = Ncell * sum(nits) / seconds over EXACTLY K implicit-Euler steps of the
         reference's time loop (thermalmodel.py:151-348) after W warm-up steps.  One "step" = one
         time step = one Newton solve (tpb_newton_solve: assembly + PC set-up + (F)GMRES per Newton
         iteration), retried with dt/2 on a failed solve exactly as the reference does.
value    state resident in HBM between steps (device pointers through the C-ABI).
e2e      the same K steps through tpb_newton_solve_host with HOST buffers: u and u_old go host->device
         and the solution comes back device->host inside the timed region, every step.
workload SPE10-shaped synthetic 60x220x85 (x N ranks in z for weak scaling), TwoPhase, wells preset
         'default' (Peaceman, rate 2e-4, S_o 0.9), solver_parameters 'pc_cptr'.

`--impl reference` times the CPU restatement of the same path (oracle/cport, C + OpenMP, all host
threads): at --gpus 1 on the SAME configuration (the full 60x220x85 grid, same steps), at --gpus N > 1 on
a bounded sample (the top 17 layers) because the N-times-stacked grid would take N times as long; the
line says which (`config.same_config`).  The reference itself (Firedrake/PETSc/hypre) cannot be installed
in this image.  The port's own assembly / SpMV GB/s are reported against a STREAM triad
measured in the same run (`cpu_baseline.roofline`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SPE10 3D two-phase: Mcell-Newton-iters/sec"
UNIT = "Mcell-Newton-iters/s"
NX, NY, NZ = 60, 220, 85
RATE, S_O, MAXDT, DT_INIT_FACT = 2e-4, 0.9, 1.0, 2.0 ** -10
PC = "pc_cptr"
BYTES = {"assemble_FJ": 608, "spmv": 552}   # algorithmic B/cell, 3-D two-phase
CPU_SAMPLE_NZ = 17


def workload_name(nz, mult, mode="stack"):
    how = "" if mult == 1 else (", %d stacked copies of the 85-layer reservoir, each with its own well pair" % mult if mode == "stack" else ", z-refined x%d" % mult)
    return ("SPE10-shaped synthetic 60x220x%d (seed 10%s) TwoPhase thermal, wells 'default' Peaceman rate 2e-4, "
            "S_o=0.9, %s, small_dt_start 2^-10 of maxdt=1 day" % (nz, how, PC))


def make_params():
    from thermalporous_b200.physicalparameters import PhysicalParameters

    class P(PhysicalParameters):
        pass
    p = P()
    p.rate, p.S_o = RATE, S_O
    return p


def make_geo(prm, nz_layers=NZ, mult=1, mode="stack"):
    """mult > 1 (weak scaling, one 85-layer slab per rank): 'stack' = mult copies of the column on top of each other
    (same cells, same physics per slab), 'refine' = every layer split mult times (BASELINE config 5; smaller cells
    make the same wells a harder problem, so it mixes solver difficulty into the scaling number)."""
    from thermalporous_b200 import geo as G
    fields = G.spe10_synthetic(NX, NY, NZ, seed=10)
    if nz_layers != NZ:
        fields = [np.ascontiguousarray(f[:, :, NZ - nz_layers:]) for f in fields]   # the top nz_layers layers (z up)
    if mult > 1 and mode == "stack":
        fields = [np.concatenate([f] * mult, axis=2) for f in fields]
        return G.SPE10Model3D(NX, NY, nz_layers * mult, prm, fields=fields)
    return G.SPE10Model3D(NX, NY, nz_layers, prm, fields=fields, refine_z=mult)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class NpOps:
    """time-loop helpers on HOST fields; with a process group the bounds are reduced over the ranks' slabs"""

    def __init__(self, dist=None, device=None):
        self.dist, self.device = dist, device

    def copy(self, d, s):
        d[...] = s

    def minmax(self, u, f):
        lo, hi = float(u[f].min()), float(u[f].max())
        if self.dist is not None:
            import torch
            t = torch.tensor([-lo, hi], device=self.device, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            lo, hi = -float(t[0].item()), float(t[1].item())
        return lo, hi

    def clip(self, u, f, lo, hi):
        np.clip(u[f], lo, hi, out=u[f])


def cpu_run(nz_layers, steps, warmup, verbose=False):
    """the CPU restatement on the top `nz_layers` layers; returns (value, seconds, nits, threads, res)."""
    from oracle import cport
    from thermalporous_b200 import cases as CS, options as O
    from thermalporous_b200.model import run_time_loop
    prm = make_params()
    geo = make_geo(prm, nz_layers)
    case = CS.WellCase(prm, geo, well_case="default")
    eng = cport.CpuEngine(3, geo.Nx, geo.Ny, geo.Nz, geo.Dx, geo.Dy, geo.Dz, 2, prm)
    eng.set_field(cport.PHI, geo.phi)
    eng.set_field(cport.KX, geo.K_x)
    eng.set_field(cport.KY, geo.K_y)
    eng.set_field(cport.KZ, geo.K_z)
    eng.set_sources(CS.source_entries(case, prm, geo))
    opts, _, _ = O.resolve(PC, 2)
    eng.set_solver_opts(**opts)
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    eng.set_num_threads(ncores)
    n = geo.ncell
    u = np.stack([np.full(n, prm.p_ref), np.full(n, prm.T_prod), np.full(n, prm.S_o)])
    uo = u.copy()
    kw = dict(end=1e9, maxdt=MAXDT, small_dt_start=True, dt_init_fact=DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
    newton = lambda a, b, dt: eng.newton_solve(a, b, dt)
    dt0 = None
    if warmup > 0:
        rw = run_time_loop(newton, NpOps(), u, uo, max_steps=warmup, **kw)
        dt0 = rw.next_dt
    t0 = time.perf_counter()
    res = run_time_loop(newton, NpOps(), u, uo, max_steps=steps, dt0=dt0, **kw)
    sec = time.perf_counter() - t0
    val = n * res.total_nits / sec / 1e6
    # the port's own roofline: its assembly and SpMV (algorithmic bytes of SURVEY 8d) against the host's STREAM triad
    roof = None
    try:
        dt = res.dt_vec[-1]
        F, J = eng.assemble(u, uo, dt)
        x = np.random.default_rng(0).standard_normal(u.shape)
        ta, ts = [], []
        for _ in range(3):
            t1 = time.perf_counter()
            eng.assemble(u, uo, dt)
            ta.append(time.perf_counter() - t1)
            t1 = time.perf_counter()
            eng.spmv(J, x)
            ts.append(time.perf_counter() - t1)
        stream = cport.stream_triad_gbs(40_000_000, 5)
        ga, gs = BYTES["assemble_FJ"] * n / min(ta) / 1e9, BYTES["spmv"] * n / min(ts) / 1e9
        roof = {"stream_triad_gbs": stream, "assembly_gbs": ga, "assembly_frac": ga / stream, "spmv_gbs": gs,
                "spmv_frac": gs / stream, "note": "algorithmic bytes (608 / 552 B per cell) / best of 3, Python call overhead and "
                "the output allocation included; STREAM triad over 3 x 320 MB, best of 5, same threads"}
    except Exception as e:   # the roofline is a side note of the baseline, never a reason to lose the line
        roof = {"error": repr(e)}
    return val, sec, res, eng.num_threads(), n, roof


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    same = args.gpus <= 1 and not args.cpu_sample
    nzl = NZ if same else CPU_SAMPLE_NZ
    val, sec, res, threads, n, roof = cpu_run(nzl, args.steps, args.warmup)
    if same:
        sample = ("the whole workload: 60x220x%d = %d cells, same field, wells, physics, option set and time loop; CPU restatement "
                  "oracle/cport (C+OpenMP), %d steps after %d warm-up" % (NZ, n, args.steps, args.warmup))
    else:
        sample = ("top %d of %d layers (60x220x%d = %d cells) of the same synthetic SPE10 field, same wells, physics, "
                  "option set and time loop; CPU restatement oracle/cport (C+OpenMP), %d steps after %d warm-up"
                  % (CPU_SAMPLE_NZ, NZ, CPU_SAMPLE_NZ, n, args.steps, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(NZ, 1), "sample": sample, "same_config": same, "cells": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "roofline": roof},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "nits": res.nits_vec, "lits": res.lits_vec,
            "note": "Firedrake/PETSc/hypre are not installable here; this is the CPU restatement, not the reference"}
    print(json.dumps(line), flush=True)


def peak_gbs():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import torch
    import torch.distributed as dist
    from thermalporous_b200 import _lib as L, cases as CS, options as O
    from thermalporous_b200.engine import Engine
    from thermalporous_b200.model import run_time_loop, _TorchOps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - thermalporous_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prm = make_params()
    refine = 1 if args.scale == "strong" else max(world, args.mult)
    geo = make_geo(prm, NZ, refine, args.scale)      # global grid 60 x 220 x 85*world (strong: 60 x 220 x 85 whatever N)
    from thermalporous_b200.partition import Slab
    slab = Slab(geo, world, rank)
    if refine > 1 and args.scale == "stack":
        # every stacked copy of the reservoir keeps its own producer/injector pair (the wells of the N=1 case at the
        # same place inside each copy), so each rank's slab is the N=1 problem coupled to its neighbours
        base = make_geo(prm, NZ, 1)
        base_ent = CS.source_entries(CS.WellCase(prm, base, well_case="default"), prm, base)
        all_ent = [(c + r * base.ncell,) + tuple(rest) for r in range(refine) for (c, *rest) in base_ent]
    else:
        all_ent = CS.source_entries(CS.WellCase(prm, geo, well_case="default"), prm, geo)
    ent = slab.localize_sources(all_ent)
    nxl, nyl, nzl = slab.local_dims()
    eng = Engine(3, nxl, nyl, nzl, geo.Dx, geo.Dy, geo.Dz, 2, prm, device=local, has_lo=slab.has_lo, has_hi=slab.has_hi)
    for fid, arr in ((L.TPB_PHI, geo.phi), (L.TPB_KX, geo.K_x), (L.TPB_KY, geo.K_y), (L.TPB_KZ, geo.K_z)):
        eng.set_field(fid, slab.take(arr))
    eng.set_sources(ent)
    if world > 1:
        uid = [eng.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
        eng.exchange_static()
    opts, _, desc = O.resolve(PC, 2)
    for kv in args.opt:
        k, v = kv.split("=")
        opts[k] = float(v) if ("." in v or "e" in v) else int(v)
        desc += " %s" % kv
    eng.set_solver_opts(**opts)
    n_loc = eng.n
    n_glob = geo.ncell
    u = eng.tensor(np.stack([np.full(n_loc, prm.p_ref), np.full(n_loc, prm.T_prod), np.full(n_loc, prm.S_o)]))
    uo = u.clone()
    kw = dict(end=1e9, maxdt=MAXDT, small_dt_start=True, dt_init_fact=DT_INIT_FACT, two_phase=True, i_S=2, spe10=True)
    ops = _TorchOps(eng)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=eng.device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W steps of the time loop (also builds the multigrid hierarchies, Krylov basis ...)
    rw = run_time_loop(lambda a, b, dt: eng.newton_solve(a, b, dt), ops, u, uo, max_steps=args.warmup, **kw)
    dt0 = rw.next_dt
    snap_u, snap_uo = u.clone(), uo.clone()

    def timed(newton, uu, uuo, ops):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        with ClockSampler(local) as cs:
            e0.record(stream)
            t0 = time.perf_counter()
            res = run_time_loop(newton, ops, uu, uuo, max_steps=args.steps, dt0=dt0, **kw)
            e1.record(stream)
            e1.synchronize()
            wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=eng.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return res, ms, wall, eng.launch_count() - l0, cs.summary()

    # ---- timed region 1: state resident in HBM
    res, ms, wall, launches, clocks = timed(lambda a, b, dt: eng.newton_solve(a, b, dt), u, uo, ops)
    value = n_glob * res.total_nits / (ms * 1e-3) / 1e6

    # ---- timed region 2 (e2e): HOST buffers through tpb_newton_solve_host, same steps from the same state
    # (every rank copies its own slab in and out; the byte counts are the whole job's)
    if True:
        hu = torch.empty_like(snap_u, device="cpu").pin_memory()
        huo = torch.empty_like(snap_u, device="cpu").pin_memory()
        hu.copy_(snap_u)
        huo.copy_(snap_uo)
        hu_np, huo_np = hu.numpy(), huo.numpy()
        res2, ms2, wall2, _, _ = timed(lambda a, b, dt: eng.newton_solve_host(a, b, dt), hu_np, huo_np,
                                         NpOps(dist if world > 1 else None, eng.device))
        nbytes = hu_np.nbytes
        e2e = {"value": n_glob * res2.total_nits / (ms2 * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": 2 * nbytes * world, "d2h_bytes_per_step": nbytes * world, "ms_per_step": ms2 / args.steps,
               "nits": res2.nits_vec, "api": "tpb_newton_solve_host (host u, u_old in; host u out)"}
        # same physics from the same state: the two regions must agree on the converged fields
        dev_final = u.cpu().numpy()
        diff = max(float(np.abs(dev_final[f] - hu_np[f]).max() / np.abs(dev_final[f]).max()) for f in range(3))
        if world > 1:
            t = torch.tensor([diff], device=eng.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            diff = float(t.item())
        e2e["max_rel_diff_vs_resident_run"] = diff

    # ---- rooflines, each kernel timed alone with CUDA events on the handle's stream (burst peak applies)
    pk, pk_how = peak_gbs()
    F = eng.empty(3, n_loc)
    J = eng.empty(7, 3, 3, n_loc)
    x = torch.randn(3, n_loc, device=eng.device, dtype=torch.float64)
    y = eng.empty(3, n_loc)
    eng.assemble(u, uo, res.dt_vec[-1], F=F, J=J)
    eng.pc_setup(J, u, res.dt_vec[-1])
    roof = {}
    # per-launch DRAM traffic (dram__bytes_read + write of one `ncu --set full` capture of the same kernel at the same
    # size): never measured inside this run - read from the committed capture summary, with its provenance
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        tr = {}
    opt_now = eng.solver_opts()
    group = opt_now["mg_tile_sweeps"] if opt_now["mg_tile_sweeps"] > 0 else opt_now["mg_pre"]
    for which, name, bpc, units in ((0, "assemble_FJ", BYTES["assemble_FJ"], n_loc), (2, "spmv", BYTES["spmv"], n_loc),
                                    (3, "line_smooth_fine", 88, n_loc)):
        t_ms = eng.time_kernel(which, u, uo, res.dt_vec[-1], F, J, x, y, reps=20)
        gbs = bpc * units / t_ms / 1e6
        ent = tr.get(name, {}) if n_loc == 1122000 else {}
        roof[name] = {"bound": "hbm", "achieved": gbs, "peak": pk, "unit": "GB/s", "frac": gbs / pk, "traffic": ent.get("bytes"),
                      "traffic_from": ent.get("from"), "ms_per_launch": t_ms, "bytes_per_unit": bpc, "units_per_launch": units,
                      "peak_source": pk_how}
    roof["assemble_FJ"]["note"] = "property pre-pass + flux kernel + source kernel (3 launches)"
    roof["line_smooth_fine"]["note"] = (
        "one launch of the hybrid z-line smoother on the finest pressure level (%d sweeps per launch): the kernel with the largest "
        "share of the step (profiles/r2_launch_summary.md).  88 B per cell = a1..a4 32 + three Thomas factors 24 + b 8 + x in 8 + x out "
        "8 + the tile rims' re-reads ~8; the kernel is issue-bound, not HBM-bound (DESIGN.md)" % group)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.scale == "strong" else "weak",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(geo.Nz, refine, args.scale), "cells": n_glob, "cells_per_gpu": n_loc,
                       "solver": desc, "l2": "working set per Newton step (Jacobian 565 MB + Krylov basis) exceeds the 126 MB L2; no flush needed",
                       "parallelism": "z-slab x%d" % world,
                       "exchanges": ("single slab" if world == 1 else
                                     "peer-memory mailboxes over NVLink (mask %d: 1 Krylov all-reduce, 2 halo planes, 4 multigrid "
                                     "gather, 8 halo fused into the SpMV kernel); NCCL for the rest" % eng.peer_mode() if eng.peer_mode() else "NCCL")},
            "nits": res.nits_vec, "lits": res.lits_vec, "dt_days": [d / 86400.0 for d in res.dt_vec],
            "failed_solves": res.failed_solves, "failed": res.failed, "host_wall_ms_per_step": wall * 1e3 / args.steps,
            "phase_ms": {"assemble": sum(s.t_assemble_ms for s in res.stats), "pc_setup": sum(s.t_pcsetup_ms for s in res.stats),
                         "ksp": sum(s.t_ksp_ms for s in res.stats)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof["line_smooth_fine"],
            "roofline_spmv": roof["spmv"], "roofline_assembly": roof["assemble_FJ"]}
    if e2e is not None:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu:
        cval, csec, cres, threads, cn, croof = cpu_run(CPU_SAMPLE_NZ, 6, 2)
        line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": threads, "kind": "port", "roofline": croof,
                                "sample": "bounded sample, NOT the same configuration: top %d of %d layers (%d cells), 6 steps "
                                          "after 2 warm-up, oracle/cport C+OpenMP restatement of the same path (%.1f s); "
                                          "`--impl reference` runs the whole grid" % (CPU_SAMPLE_NZ, NZ, cn, csec)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)     # the driver's own run: 20 steps after 5 warm-up, which takes
    ap.add_argument("--warmup", type=int, default=5)     # dt from the small_dt_start phase up to ~0.1 day
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", action="store_true", help="--impl reference: the 17-layer sample also at --gpus 1")
    ap.add_argument("--scale", default="stack", choices=["stack", "refine", "strong"],
                    help="how the grid grows with --gpus: stack / refine = weak scaling (one 85-layer slab per rank), strong = "
                         "the same 60x220x85 grid cut into N z-slabs")
    ap.add_argument("--mult", type=int, default=0, help="grid multiplier when it should differ from --gpus (experiments: "
                    "the N-rank problem on fewer ranks)")
    ap.add_argument("--opt", action="append", default=[], help="solver option override key=value (experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
