"""Multi-GPU check of `FEModel3D.analyze_linear` with combos sharded across ranks (launch with torchrun):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/distributed_check.py
Every rank must end up with all combos, identical to a non-distributed solve on the same rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynite_b200 import FEModel3D          # noqa: E402
from tests import models                   # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, ws = dist.get_rank(), dist.get_world_size()
m = models.big_quad_plate(FEModel3D, n=40, n_combos=7)
m.analyze_linear()
sharded = {c: m._D[c].copy() for c in m.load_combos}
rx = {c: m._Rxn[c].copy() for c in m.load_combos}
own = list(m._run.engine_combo)
m.distribute_combos = False
m.analyze_linear()
err = max(np.abs(sharded[c] - m._D[c]).max()/np.abs(m._D[c]).max() for c in m.load_combos)
err_r = max(np.abs(rx[c] - m._Rxn[c]).max()/np.abs(m._Rxn[c]).max() for c in m.load_combos)
print(f'rank {rank}/{ws}: own combos {own}; max rel diff sharded vs local: D {err:.2e}, Rxn {err_r:.2e}', flush=True)
assert err < 1e-12 and err_r < 1e-12
dist.barrier()
dist.destroy_process_group()
