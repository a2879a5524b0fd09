#!/usr/bin/env python
"""Generate tests/golden/pc/option_sets.json: the dictionaries the reference's OWN dispatchers
(SinglePhase/TwoPhase.init_solver_parameters, singlephase.py:275-444, twophase.py:413-1002) produce for every named
option set libtpb200 realises, plus the decoupling each one selects.  The reference models are built over
tests/golden/fd_shim exactly as make_golden.py builds them; `model.solver_parameters` is what they would hand to
Firedrake's NonlinearVariationalSolver.

    python tests/golden/make_option_golden.py        (build container only: needs /root/reference)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "fd_shim"))
sys.path.insert(0, os.environ.get("TPB_REFERENCE", "/root/reference"))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden as mg  # noqa: E402
from thermalporous_b200 import options as O  # noqa: E402


def main():
    prm = mg.fresh_params(rate=1e-6)
    geo = mg.HomogeneousGeo(4, 4, prm, 20.0, 20.0)
    case = mg.WellCase(prm, geo, well_case="test0", constant_rate=True)
    out = {}
    for nphase, names, Model in ((1, O.SINGLE_PHASE_SETS, mg.SinglePhase), (2, O.TWO_PHASE_SETS, mg.TwoPhase)):
        for name in list(names) + [None]:
            m = Model(geo, case, prm, end=1.0, maxdt=1.0, small_dt_start=False, filename=mg.TMP, verbosity=False,
                      solver_parameters=name)
            assert isinstance(m.solver_parameters, dict), (nphase, name)
            out["%d|%s" % (nphase, name)] = {"decoup": m.decoup, "vector": bool(getattr(m, "vector", False)),
                                            "parameters": m.solver_parameters}
    os.makedirs(os.path.join(HERE, "pc"), exist_ok=True)
    path = os.path.join(HERE, "pc", "option_sets.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out), "option sets")


if __name__ == "__main__":
    main()
