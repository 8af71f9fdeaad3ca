"""Multi-GPU check of the subtree-to-GPU direct solve (pn_set_distributed) over NCCL.  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/dist_direct_check.py [--mesh 160] [--combos 8]
Every rank solves the model alone first, then distributed; the two must agree to 1e-12 on every rank (bit-identity is
reported: it holds at the test sizes, at config 3 the two runs differ in the last bit, 1.2e-16 relative).  Prints the
timing of both (device events) and the exchange counters."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynite_b200 import distsolve, synthetic       # noqa: E402
from pynite_b200.engine import Engine              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mesh', dest='n', type=int, default=160)
ap.add_argument('--combos', type=int, default=8)
ap.add_argument('--steps', type=int, default=3)
args = ap.parse_args()
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, ws = dist.get_rank(), dist.get_world_size()
A = synthetic.quad_plate_arrays(n=args.n, n_combos=args.combos)
eng = Engine(local)
opts = eng.make_opts()
eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors)
eng.symbolic(0); eng.assemble_ke()
D0, R0, _ = eng.solve_linear(opts)


def timed(label):
    eng.solve_linear(opts, reactions=False, download=False)
    torch.cuda.synchronize(); dist.barrier()
    eng.timer_start()
    for _ in range(args.steps):
        eng.assemble_ke()
        eng.solve_linear(opts, reactions=False, download=False)
    ms = eng.timer_stop()/args.steps
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{label}: {float(t.item()):.2f} ms per step (max over ranks)', flush=True)


timed('single (every rank alone)')
distsolve.attach(eng)
eng.reset_stats()
D1, R1, st = eng.solve_linear(opts)
ctr = eng.counters()
same_d, same_r = np.array_equal(D0, D1), np.array_equal(R0, R1)
err = float(np.abs(D0 - D1).max()/np.abs(D0).max())
print(f'rank {rank}/{ws}: cut level {ctr["dist_cut_level"]} of {ctr["levels"]}, exchanges {ctr["dist_exchanges"]}, '
      f'bit-identical D {same_d} R {same_r} (max rel diff {err:.2e}), residual {st.max_rel_residual:.2e}', flush=True)
timed(f'distributed over {ws} GPUs')
assert err < 1e-12
dist.barrier()
eng.close()
dist.destroy_process_group()
