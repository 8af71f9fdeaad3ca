"""Domain-decomposed PCG on the config-4 model shape (quad mat on Winkler springs), one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
        tools/ddpcg_run.py --size 1000 --combos 8 --max-iter 4000
Rank 0 prints one JSON line: iterations, time per iteration (CUDA events, max over ranks), halo bytes, residual, and --
when --check is given -- the difference to the multifrontal direct solve of the same system."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynite_b200 import ddpcg, synthetic          # noqa: E402
from pynite_b200.engine import Engine             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--size', dest='n', type=int, default=200)
ap.add_argument('--combos', type=int, default=8)
ap.add_argument('--tol', type=float, default=1e-8)
ap.add_argument('--max-iter', type=int, default=20000)
ap.add_argument('--ks', type=float, default=0.1)
ap.add_argument('--check', action='store_true')
ap.add_argument('--no-graph', action='store_true', help='issue every iteration call by call instead of replaying a CUDA graph')
ap.add_argument('--profile', action='store_true', help='time the pieces of one iteration separately (100 repetitions each)')
args = ap.parse_args()

local = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank = dist.get_rank() if world > 1 else 0

A = synthetic.quad_plate_arrays(n=args.n, n_combos=args.combos, t=24.0, mat_springs=args.ks)
eng = Engine(local)
eng.set_model(A); eng.set_loads(A); eng.set_combos(A.factors)
sizes = eng.symbolic(0)
eng.assemble_ke()
Xd = None
if args.check:
    D_direct, _, _ = eng.solve_linear(eng.make_opts(), reactions=False)
ops = ddpcg.EngineOps(eng, A.n_nodes)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
hist = []
e0.record()
prof_solve = {} if args.profile else None
X, it, info = ddpcg.solve(ops, 0, args.combos, tol=args.tol, max_iter=args.max_iter, check_every=50, profile=prof_solve, graph=not args.no_graph,
                          log=(lambda k, w: hist.append((k, w))) if rank == 0 else None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
ops.set_solution(0, X)
out = {'config': f'{args.n}x{args.n} quad mat on springs ks={args.ks}', 'n_gpus': world, 'n_dof': int(sizes.n_dof), 'n1': int(sizes.n1),
       'nnz11': int(sizes.nnz11), 'combos': args.combos, 'iterations': it, 'ms_total': ms, 'ms_iteration_loop': info.get('ms_iteration_loop'),
       'ms_per_iteration': (info.get('ms_iteration_loop') or ms)/max(it, 1),
       'max_rel_residual': info['max_rel_residual'], 'rows_per_rank': info['rows_per_rank'],
       'halo_bytes_per_iteration_rank0': info['halo_bytes_per_iteration'],
       'spmm_algorithmic_GBps_per_rank': ((sizes.nnz11*12 + 2*sizes.n1*args.combos*8)/world)/((info.get('ms_iteration_loop') or ms)/max(it, 1)*1e-3)/1e9,
       'residual_history_tail': hist[-3:]}
if args.check:
    D = eng.displacements(0)
    out['max_rel_diff_vs_direct_combo0'] = float(np.abs(D - D_direct[0]).max()/np.abs(D_direct[0]).max())
if args.profile:
    # the pieces of one iteration on this rank's strip, each repeated 100 times between two events (max over ranks)
    nrs = ops.node_row_start
    ncut = ddpcg.strip_cuts(nrs, world)
    node0, node1 = ncut[rank], ncut[rank + 1]
    r0, r1 = int(nrs[node0]), int(nrs[node1])
    sN = args.combos
    dev = ops.device
    Pv = torch.randn((ops.n1, sN), dtype=torch.float64, device=dev)
    APv, Xv, Rv, Zv = (torch.zeros_like(Pv) for _ in range(4))
    rz = torch.ones(sN, dtype=torch.float64, device=dev)
    pAp = torch.ones(sN, dtype=torch.float64, device=dev)
    red = torch.ones(2*sN, dtype=torch.float64, device=dev)

    def timed(fn, reps=100):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)/reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()), 4)

    out['profile_solve_ms'] = prof_solve
    out['profile_ms'] = {
        'spmm_rows': timed(lambda: ops.spmm_rows(r0, r1, Pv, APv)),
        'cg_dots': timed(lambda: ops.cg_dots(r0, r1, Pv, APv, pAp)),
        'cg_update': timed(lambda: ops.cg_update(node0, node1, False, rz, pAp, Pv, APv, Xv, Rv, Zv, red)),
        'cg_direction': timed(lambda: ops.cg_direction(r0, r1, False, rz, red, Zv, Pv)),
        'allreduce_x2': timed(lambda: (dist.all_reduce(pAp), dist.all_reduce(red))) if world > 1 else 0.0,
    }
if rank == 0:
    print(json.dumps(out), flush=True)
eng.set_stream(None)
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
