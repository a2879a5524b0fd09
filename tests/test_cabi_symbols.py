"""CPU-side check that the C-ABI library loads and exports every symbol include/tpb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from thermalporous_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "tpb200.h")).read()
    declared = set(re.findall(r"\b(tpb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.tpb_version() >= 100


def test_defaults_follow_reference_option_sets():
    from thermalporous_b200 import _lib as L
    lib = L.load()
    o = L.SolverOpts()
    assert lib.tpb_solver_defaults(1, ctypes.byref(o)) == 0
    assert (o.snes_max_it, o.ksp_max_it, o.ksp_restart, o.ksp_type) == (15, 200, 200, L.KSP_GMRES)  # singlephase.py:289-301
    assert lib.tpb_solver_defaults(2, ctypes.byref(o)) == 0
    assert (o.snes_max_it, o.ksp_type) == (25, L.KSP_FGMRES) and o.ksp_rtol == 1e-8              # twophase.py:416-433


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from thermalporous_b200.engine import Engine
    from oracle.tp_oracle import Params
    with pytest.raises(RuntimeError):
        Engine(2, 4, 4, 1, 1.0, 1.0, 1.0, 1, Params())
