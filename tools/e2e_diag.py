import sys, time
sys.path.insert(0, '.')
import numpy as np
from pynite_b200 import synthetic
from pynite_b200.engine import Engine
A = synthetic.quad_plate_arrays(n=500, n_combos=64)
eng = Engine(0)
opts = eng.make_opts()
D = np.empty((64, A.n_dof)); R = np.empty((64, A.n_dof))
for it in range(6):
    print('--- e2e pass', it, flush=True)
    t = [time.perf_counter()]
    eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors); t.append(time.perf_counter())
    eng.symbolic(0); t.append(time.perf_counter())
    eng.assemble_ke(); t.append(time.perf_counter())
    eng.solve_linear(opts, reactions=False, download=False); t.append(time.perf_counter())
    eng.solve_linear(opts, reactions=True, D_out=D, R_out=R); t.append(time.perf_counter())
    print('upload %.1f symbolic %.1f assemble %.1f first solve (with analyse, no download) %.1f second solve (+reactions, +download to pageable) %.1f ms' % tuple(1e3*(t[i+1]-t[i]) for i in range(5)), flush=True)
