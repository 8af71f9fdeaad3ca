"""Host timeline of one end-to-end step of config 3 (PN_DEBUG_TIMING laps) plus the PCIe copy rates of the box."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
from pynite_b200 import synthetic
from pynite_b200.engine import Engine, _ptr, _c, _I32P, _F64P

A = synthetic.quad_plate_arrays(n=500, n_combos=64)
for f in ['xyz', 'support', 'enf_mask', 'enf_val', 'nspring_k', 'nspring_flag', 'quad_ijmn', 'quad_props', 'nl_dof', 'nl_case', 'nl_val',
          'qp_elem', 'qp_case', 'qp_val', 'factors']:
    a = np.ascontiguousarray(getattr(A, f))
    if a.size:
        p = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True).numpy()
        p[...] = a
        setattr(A, f, p)
eng = Engine(0)
opts = eng.make_opts()
D = torch.empty((64, A.n_dof), dtype=torch.float64, pin_memory=True).numpy()
R = torch.empty((64, A.n_dof), dtype=torch.float64, pin_memory=True).numpy()
for it in range(5):
    if it == 4:
        os.environ['PN_DEBUG_TIMING'] = '1'
    t = [time.perf_counter()]
    eng.set_nodes(A); t.append(time.perf_counter())
    eng.set_springs(A); eng.set_members(A); t.append(time.perf_counter())
    L = eng._lib
    eng._check(L.pn_set_quads(eng._h, len(A.quad_ijmn), _ptr(_c(A.quad_ijmn, np.int32), _I32P), _ptr(_c(A.quad_props, np.float64), _F64P)))
    eng._check(L.pn_set_plates(eng._h, 0, _ptr(_c(A.plate_ijmn, np.int32), _I32P), _ptr(_c(A.plate_props, np.float64), _F64P)))
    eng.n_dof = 6*A.n_nodes
    eng.counts = {0: 0, 1: 0, 2: len(A.quad_ijmn), 3: 0}
    eng._check(L.pn_prefetch_analysis(eng._h)); t.append(time.perf_counter())
    eng.set_loads(A); t.append(time.perf_counter())
    eng.set_combos(A.factors); t.append(time.perf_counter())
    eng.symbolic(0); t.append(time.perf_counter())
    eng.assemble_ke(); t.append(time.perf_counter())
    eng.solve_linear(opts, reactions=True, D_out=D, R_out=R); t.append(time.perf_counter())
    names = ['set_nodes', 'springs+members', 'quads+prefetch', 'set_loads', 'set_combos', 'symbolic', 'assemble', 'solve_linear']
    print('pass %d: ' % it + '  '.join('%s %.2f' % (n, 1e3*(t[i + 1] - t[i])) for i, n in enumerate(names)) + '  total %.2f ms' % (1e3*(t[-1] - t[0])), flush=True)
os.environ.pop('PN_DEBUG_TIMING')
# PCIe rates
x = torch.empty(774*1024*1024//8, dtype=torch.float64, device='cuda')
for name, fn in (('D2H 774 MB pinned', lambda: D.__class__ and torch.from_numpy(D).copy_(x[:D.size].view(D.shape), non_blocking=True)),
                 ('H2D 774 MB pinned', lambda: x[:D.size].view(D.shape).copy_(torch.from_numpy(D), non_blocking=True))):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%s: %.2f ms  %.1f GB/s' % (name, 1e3*dt, D.nbytes/dt*1e-9), flush=True)
t0 = time.perf_counter(); R[...] = 0.0; print('numpy zero fill of 774 MB pinned: %.2f ms' % (1e3*(time.perf_counter() - t0)))
