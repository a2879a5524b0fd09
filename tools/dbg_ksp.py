import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cport
from tests.gpu_util import engine_from_problem, random_problem
from tests.test_gpu_solver import COMBOS
nphase, dim, shape, opts = COMBOS[int(sys.argv[1])]
pb, u, uo = random_problem(dim, nphase, shape, seed=4, spread=0.05)
g = engine_from_problem(pb); c = cport.engine_from_problem(pb)
o = dict(opts, ksp_type=int(sys.argv[2]), ksp_rtol=1e-10, verbose=2, ksp_max_it=200)
g.set_solver_opts(**o); c.set_solver_opts(**o)
F, J = g.assemble(u, uo, 4000.0)
g.pc_setup(J, u, 4000.0); c.pc_setup(J.cpu().numpy(), u, 4000.0)
print(g.ksp_solve(J, F)[1:], c.ksp_solve(J.cpu().numpy(), F.cpu().numpy())[1:])
