import sys, time
sys.path.insert(0, '.')
from pynite_b200 import synthetic
from pynite_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
A = synthetic.quad_plate_arrays(n=n, n_combos=64)
eng = Engine(0)
eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors)
eng.symbolic(0); eng.assemble_ke()
opts = eng.make_opts(refine_steps=2)
eng.solve_linear(opts, reactions=False, download=False)
for prof in (False, True, False):
    eng.profile_enable(prof)
    for it in range(2):
        eng.reset_stats()
        t0 = time.perf_counter(); eng.assemble_ke(); t1 = time.perf_counter()
        eng.solve_linear(opts, reactions=False, download=False); t2 = time.perf_counter()
        st = eng.stats().as_dict()
        print(f'profile={prof} assemble={1e3*(t1-t0):.1f} ms solve_call={1e3*(t2-t1):.1f} ms  phases: rhs={st["ms_rhs"]:.1f} factor={st["ms_factor"]:.1f} solve={st["ms_solve"]:.1f} launches={st["kernel_launches"]}')
