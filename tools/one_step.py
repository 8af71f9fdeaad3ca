"""Exactly one steady-state step of config 3 (class-path assembly, right-hand sides, factorisation, two solves, residual
check) between cudaProfilerStart / cudaProfilerStop, for `ncu --profile-from-start off ...`.
   PN_FACTOR_SPLIT_MAX=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python tools/one_step.py [n] [combos]"""
import sys
sys.path.insert(0, '.')
import torch
from pynite_b200 import synthetic
from pynite_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
combos = int(sys.argv[2]) if len(sys.argv) > 2 else 64
A = synthetic.quad_plate_arrays(n=n, n_combos=combos)
eng = Engine(0)
eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors)
eng.symbolic(0); eng.assemble_ke()
opts = eng.make_opts(refine_steps=2)
for _ in range(3):
    eng.assemble_ke()
    eng.solve_linear(opts, reactions=False, download=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.reset_stats()
eng.assemble_ke()
eng.solve_linear(opts, reactions=False, download=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = eng.stats().as_dict()
print('one step: factor %.2f ms, solve %.2f ms, launches %d, counters %s' % (st['ms_factor'], st['ms_solve'], st['kernel_launches'], eng.counters()))
