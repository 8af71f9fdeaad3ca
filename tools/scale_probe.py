"""Phase timings of the GPU path on synthetic quad plates of growing size (development aid)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from pynite_b200 import synthetic
from pynite_b200.engine import Engine

sizes = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [50, 100, 200]
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
solver = sys.argv[3] if len(sys.argv) > 3 else 'direct'
eng = Engine(0)
for n in sizes:
    A = synthetic.quad_plate_arrays(n=n, n_combos=nc)
    t0 = time.perf_counter()
    eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors)
    t1 = time.perf_counter()
    eng.reset_stats()
    s = eng.symbolic(0)
    t2 = time.perf_counter()
    eng.assemble_ke()
    t3 = time.perf_counter()
    try:
        D, R, st = eng.solve_linear(eng.make_opts(solver=solver, max_iter=200000, tol=1e-10), reactions=True)
    except Exception as e:
        print('n', n, 'solve failed:', type(e).__name__, e)
        continue
    t4 = time.perf_counter()
    print(f'n={n} dof={s.n_dof} n1={s.n1} coo={s.nnz_coo} csr={s.nnz_csr} upload={t1-t0:.3f}s symbolic={t2-t1:.3f}s '
          f'assemble={t3-t2:.3f}s solve+rxn={t4-t3:.3f}s')
    print('   ', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.as_dict().items()})
    print('    max|D|', np.abs(D).max(), 'thr', s.n_dof*nc/(t4-t2), 'DOF*combos/s (incl. symbolic)')
    for srhs in (1, 8, 64):
        ms = eng.spmm_k11_time(srhs, 10)
        nb = s.nnz11*12 + (s.n1+1)*4 + 2*s.n1*srhs*8
        print(f'    spmm s={srhs}: {ms*1e3:.1f} us  {nb/ms/1e6:.0f} GB/s')
    # second solve (symbolic + analysis cached): steady-state step
    eng.reset_stats()
    t5 = time.perf_counter()
    eng.assemble_ke()
    D, R, st = eng.solve_linear(eng.make_opts(solver=solver, max_iter=200000, tol=1e-10), reactions=False)
    t6 = time.perf_counter()
    print(f'    steady step: {t6-t5:.3f}s -> {s.n_dof*nc/(t6-t5):.3e} DOF*combos/s')
    print('   ', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.as_dict().items()})
