"""World-size-2 gloo test of the multi-GPU host logic: combo sharding, the ragged all-gather of result rows,
and max-over-ranks timing.  (The data path itself has no collective: every rank solves its own combos.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynite_b200 import sharding


def test_shard_combos_partitions():
    for n in (0, 1, 5, 64, 65):
        for ws in (1, 2, 3, 8):
            parts = [sharding.shard_combos(n, r, ws) for r in range(ws)]
            assert sorted(sum(parts, [])) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_combo, n_dof, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        mine = sharding.shard_combos(n_combo, rank, ws)
        counts = [len(sharding.shard_combos(n_combo, r, ws)) for r in range(ws)]
        # stand-in for the per-rank solve: row c of the "solution" is a known function of the combo index
        local = np.array([[c*1000.0 + k for k in range(n_dof)] for c in mine]).reshape(len(mine), n_dof)
        full = sharding.all_gather_rows(local, counts)
        expect = np.array([[c*1000.0 + k for k in range(n_dof)] for c in range(n_combo)])
        ok = full.shape == expect.shape and np.array_equal(full, expect)
        tmax = sharding.max_over_ranks(10.0 + rank)
        # a failure on one rank (singular matrix, unstable node) must be seen by every rank, also by a rank that has no
        # combos of its own (world size larger than the combo count)
        seen_fail = sharding.any_rank_failed(rank == 1)
        seen_ok = sharding.any_rank_failed(False)
        out[rank] = (ok and seen_fail and not seen_ok, tmax)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_combo', [1, 5, 64])
def test_two_rank_gather(n_combo):
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), n_combo, 7, out), nprocs=ws, join=True)
    assert all(out[r][0] for r in range(ws))
    assert all(out[r][1] == 11.0 for r in range(ws))
