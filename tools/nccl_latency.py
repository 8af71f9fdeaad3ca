"""Latency of the three collectives of one DD-PCG iteration (tools/ddpcg_run.py) on N GPUs, eager and as a CUDA graph:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/nccl_latency.py"""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank = dist.get_rank()
s, halo_rows = 8, 6012          # config 4: 8 right-hand sides, one mesh row of 1002 nodes x 6 DOF per neighbour
small = torch.zeros(s, dtype=torch.float64, device=dev)
small2 = torch.zeros(2*s, dtype=torch.float64, device=dev)
P = torch.zeros((halo_rows*4, s), dtype=torch.float64, device=dev)


def exchange():
    ops = []
    for q, off in ((rank - 1, 0), (rank + 1, 1)):
        if 0 <= q < world:
            ops.append(dist.P2POp(dist.isend, P[off*halo_rows:(off + 1)*halo_rows], q))
            ops.append(dist.P2POp(dist.irecv, P[(2 + off)*halo_rows:(3 + off)*halo_rows], q))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def timed(fn, reps=200, graph=False):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    run = fn
    if graph:
        cs = torch.cuda.Stream()
        cs.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cs):
            fn()
        run = g.replay
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)/reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {'n_gpus': world}
for name, fn in (('allreduce_8_doubles', lambda: dist.all_reduce(small)), ('allreduce_16_doubles', lambda: dist.all_reduce(small2)),
                 ('halo_exchange_385KB_per_neighbour', exchange),
                 ('all_three', lambda: (exchange(), dist.all_reduce(small), dist.all_reduce(small2)))):
    out[name + '_eager_ms'] = round(timed(fn), 4)
    out[name + '_graph_ms'] = round(timed(fn, graph=True), 4)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
