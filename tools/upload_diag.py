import sys, time
sys.path.insert(0, '.')
import numpy as np
from pynite_b200 import synthetic
from pynite_b200.engine import Engine
A = synthetic.quad_plate_arrays(n=500, n_combos=64)
eng = Engine(0)
for it in range(4):
    t = [time.perf_counter()]
    eng.set_nodes(A); t.append(time.perf_counter())
    eng.set_springs(A); t.append(time.perf_counter())
    eng.set_members(A); t.append(time.perf_counter())
    L = eng._lib
    from pynite_b200.engine import _ptr, _c, _I32P, _F64P
    eng._check(L.pn_set_quads(eng._h, len(A.quad_ijmn), _ptr(_c(A.quad_ijmn, np.int32), _I32P), _ptr(_c(A.quad_props, np.float64), _F64P))); t.append(time.perf_counter())
    eng._check(L.pn_set_plates(eng._h, len(A.plate_ijmn), _ptr(_c(A.plate_ijmn, np.int32), _I32P), _ptr(_c(A.plate_props, np.float64), _F64P))); t.append(time.perf_counter())
    eng.n_dof = 6*A.n_nodes
    eng.set_loads(A); t.append(time.perf_counter())
    eng.set_combos(A.factors); t.append(time.perf_counter())
    print('nodes %.2f springs %.2f members %.2f quads %.2f plates %.2f loads %.2f combos %.2f ms' % tuple(1e3*(t[i+1]-t[i]) for i in range(7)), flush=True)
    eng.symbolic(0)
