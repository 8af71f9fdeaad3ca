"""World-size-2 gloo test of the domain-decomposed PCG driver (pynite_b200/ddpcg.py): strip cuts, halo plan, halo
exchange, the two all-reduces per iteration and the final strip all-gather, on CPU tensors.  The two library kernels
(`K11[strip] @ P`, node-block Jacobi) are replaced by a scipy stand-in built from a golden fixture's K11, so this test
checks the distributed logic only; the CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynite_b200 import ddpcg
from tests.util import load_fixture


class ScipyOps:
    def __init__(self, name):
        A, ref, _ = load_fixture(name)
        n1 = len(ref['d1'])
        self.K = sps.csr_matrix((ref['K11_data'], ref['K11_indices'], ref['K11_indptr']), shape=(n1, n1))
        self.K = ((self.K + self.K.T)*0.5).tocsr()
        self.K.sort_indices()
        self.n1 = n1
        self.device = torch.device('cpu')
        node_of = ref['d1']//6
        starts = np.full(A.n_nodes + 1, n1, dtype=np.int64)
        for row in range(n1 - 1, -1, -1):
            starts[node_of[row]] = row
        for nd in range(A.n_nodes - 1, -1, -1):
            starts[nd] = min(starts[nd], starts[nd + 1])
        self.node_row_start = starts
        rng = np.random.default_rng(3)
        self.B = rng.standard_normal((n1, 3))

    def pattern(self):
        return self.K.indptr, self.K.indices

    def rhs(self, combo0, s):
        return torch.from_numpy(self.B[:, :s].copy())

    def spmm_rows(self, row0, row1, X, Y):
        Y[row0:row1] = torch.from_numpy(self.K[row0:row1] @ X.numpy())

    def precond_nodes(self, node0, node1, R, Z):
        for nd in range(node0, node1):
            a, b = int(self.node_row_start[nd]), int(self.node_row_start[nd + 1])
            if b > a:
                Z[a:b] = torch.from_numpy(np.linalg.solve(self.K[a:b, a:b].toarray(), R[a:b].numpy()))


def test_strip_cuts_and_halo_plan():
    starts = np.array([0, 6, 12, 15, 21, 27, 27, 33])
    cuts = ddpcg.strip_cuts(starts, 3)
    assert cuts[0] == 0 and cuts[-1] == 7 and cuts == sorted(cuts)
    recv, send = ddpcg.halo_plan(1, 3, [0, 12, 21, 33], [0, 6, 15], [18, 27, 33])
    assert recv == [(0, 6, 12), (2, 21, 27)]
    assert send == [(0, 12, 18), (2, 15, 21)]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, name, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        ops = ScipyOps(name)
        X, it, info = ddpcg.solve(ops, 0, 3, tol=1e-11, max_iter=5000, check_every=5)
        exact = spla.splu(ops.K.tocsc()).solve(ops.B)
        err = np.abs(X.numpy() - exact).max()/np.abs(exact).max()
        out[rank] = (float(err), it, info['halo_rows'], info['rows_per_rank'])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'moment_frame'])
def test_two_rank_ddpcg(name):
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), name, out), nprocs=ws, join=True)
    for r in range(ws):
        err, it, halo, rows = out[r]
        assert err < 1e-8, (err, it)
        assert halo > 0 and sum(rows) > 0 and it > 0


def test_single_rank_matches():
    ops = ScipyOps('mat_foundation')
    X, it, info = ddpcg.solve(ops, 0, 2, tol=1e-11, max_iter=5000, check_every=5)
    exact = spla.splu(ops.K.tocsc()).solve(ops.B[:, :2])
    assert np.abs(X.numpy() - exact).max()/np.abs(exact).max() < 1e-8
