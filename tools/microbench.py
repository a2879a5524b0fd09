#!/usr/bin/env python
"""Kernel micro-benchmark: assembly (F+J, F only) and SpMV at C5/C4 sizes, CUDA-event timed on
the handle's stream, reported against the measured HBM peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermalporous_b200.engine import Engine  # noqa: E402
from thermalporous_b200 import _lib as L  # noqa: E402
from thermalporous_b200.physicalparameters import PhysicalParameters  # noqa: E402

BYTES = {(3, 2): (608, 104, 552), (3, 1): (312, 88, 256), (2, 2): (456, 96, 408), (2, 1): (240, 80, 192)}


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def run(dim, nphase, nx, ny, nz, reps=20):
    prm = PhysicalParameters()
    prm.S_o = 0.9
    eng = Engine(dim, nx, ny, nz, 6.096, 3.048, 0.6096, nphase, prm)
    n = eng.n
    gen = torch.Generator(device="cuda").manual_seed(1234)
    rnd = lambda lo, hi: torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * (hi - lo) + lo
    logk = torch.randn(n, generator=gen, device="cuda", dtype=torch.float64) * 1.3 + 1.0
    Kx = 10.0 ** logk * 9.869233e-10
    phi = (0.2 + 0.08 * (logk - 1.0)).clamp(0.0, 0.5) + 1e-10
    eng.set_field(L.TPB_PHI, phi)
    eng.set_field(L.TPB_KX, Kx)
    eng.set_field(L.TPB_KY, Kx)
    eng.set_field(L.TPB_KZ, 0.1 * Kx)
    eng.set_field(L.TPB_KT, phi * prm.ko + (1 - phi) * prm.kr)
    rows = [prm.p_ref + rnd(-5, 5), rnd(288.7, 422.0)]
    rows2 = [prm.p_ref + rnd(-5, 5), rnd(288.7, 422.0)]
    if nphase == 2:
        rows.append(rnd(0.05, 0.95))
        rows2.append(rnd(0.05, 0.95))
    u, uo = torch.stack(rows), torch.stack(rows2)
    F = eng.empty(eng.nf, n)
    J = eng.empty(eng.ns, eng.nf, eng.nf, n)
    x = torch.randn(eng.nf, n, generator=gen, device="cuda", dtype=torch.float64)
    y = eng.empty(eng.nf, n)
    pk, how = peak()
    out = {}
    for which, name in ((0, "assemble_FJ"), (1, "assemble_F"), (2, "spmv")):
        ms = eng.time_kernel(which, u, uo, 8640.0, F, J, x, y, reps)
        b = BYTES[(dim, nphase)][which] * n
        gbs = b / ms / 1e6
        out[name] = dict(ms=ms, gbs=gbs, frac=gbs / pk)
        print("%-12s %dx%dx%d nphase=%d: %8.3f ms  %8.1f GB/s  %.3f of %s peak %.0f" %
              (name, nx, ny, nz, nphase, ms, gbs, gbs / pk, how, pk), flush=True)
    eng.close()
    return out


if __name__ == "__main__":
    res = {}
    res["spe10_tp"] = run(3, 2, 60, 220, 85)
    res["c4_tp"] = run(3, 2, 216, 216, 216)
    res["spe10_sp"] = run(3, 1, 60, 220, 85)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
