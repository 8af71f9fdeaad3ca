"""Domain-decomposed preconditioned CG over NVLink (BASELINE config 4).

One process per GPU.  The free DOFs (K11 rows) are cut into contiguous node-aligned strips, one per rank -- node IDs
of a generated mesh are row-major, so a strip is a band of mesh rows.  Every rank assembles the whole K on its own GPU
(the numeric assembly is ~0.1 ms; shipping the CSR would cost more than recomputing it) but multiplies only its strip:

    per iteration   halo exchange of the search direction  (P2P isend/irecv of the boundary rows a neighbour's strip
                    references; for a banded matrix that is the two adjacent ranks)
                    AP[strip] = K11[strip, :] P                       (library kernel, `pn_dev_spmm_k11_rows`)
                    one all-reduce of  p.Ap                           (n_rhs scalars)
                    x, r updates on the strip; z = M^-1 r with the node-block Jacobi inverse (library kernel)
                    one all-reduce of  [r.z, r.r]                     (2 n_rhs scalars)

All right-hand sides advance together (their own alpha / beta per column).  Vector algebra on the strip uses torch
tensor expressions on the CUDA stream the library is attached to; torch.distributed (NCCL) is the plumbing.  The
arithmetic back end is abstracted (`ops`) so the strip / halo / reduction logic is also exercised by a world-size-2
gloo test on CPU tensors with a scipy stand-in for the two library kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def strip_cuts(node_row_start, world):
    """Node indices [c_0 = 0, c_1, ..., c_world = N] that balance the number of K11 rows per strip."""
    n_nodes = len(node_row_start) - 1
    n1 = int(node_row_start[-1])
    cuts = [0]
    for r in range(1, world):
        target = n1*r//world
        cuts.append(int(np.searchsorted(node_row_start, target, side='left')))
    cuts.append(n_nodes)
    for r in range(1, world + 1):          # monotone even for tiny models
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts


def halo_plan(rank, world, row_cuts, col_lo, col_hi):
    """Rows to receive from / send to every other rank.

    `row_cuts[q]..row_cuts[q+1]` is rank q's strip; `col_lo[q]`, `col_hi[q]` bound the columns its strip references.
    Returns (recv, send): lists of (peer, row_begin, row_end), pieces of the peer's (recv) / my (send) strip."""
    r0, r1 = row_cuts[rank], row_cuts[rank + 1]
    recv, send = [], []
    for q in range(world):
        if q == rank:
            continue
        q0, q1 = row_cuts[q], row_cuts[q + 1]
        a, b = max(col_lo[rank], q0), min(col_hi[rank], q1)        # what I need of q's strip
        if a < b:
            recv.append((q, a, b))
        a, b = max(col_lo[q], r0), min(col_hi[q], r1)              # what q needs of mine
        if a < b:
            send.append((q, a, b))
    return recv, send


class EngineOps:
    """Arithmetic back end on the GPU: the two library kernels through the C ABI, on torch's current stream."""

    def __init__(self, engine, n_nodes):
        self.eng = engine
        engine.set_stream(torch.cuda.current_stream().cuda_stream)
        engine.dev_block_jacobi_setup()
        self.node_row_start = engine.node_row_starts(n_nodes)
        self.n1 = int(self.node_row_start[-1])
        self.device = torch.device('cuda', engine.device)

    def set_stream(self, stream):
        """Attaches the library's launches to a torch stream (the capture stream while an iteration is recorded)."""
        self.eng.set_stream(stream.cuda_stream)

    def pattern(self):
        indptr, indices, _ = self.eng.block(0, values=False)
        return indptr, indices

    def col_range(self, row0, row1):
        """Columns [lo, hi) a strip references, computed on the device."""
        return self.eng.k11_col_range(row0, row1)

    def rhs(self, combo0, s):
        B = torch.empty((self.n1, s), dtype=torch.float64, device=self.device)
        self.eng.dev_get_rhs(combo0, s, B.data_ptr())
        return B

    def spmm_rows(self, row0, row1, X, Y):
        self.eng.dev_spmm_rows(X.shape[1], row0, row1, X.data_ptr(), Y.data_ptr())

    def precond_nodes(self, node0, node1, R, Z):
        self.eng.dev_block_jacobi_nodes(R.shape[1], node0, node1, R.data_ptr(), Z.data_ptr())

    def set_solution(self, combo0, X):
        self.eng.dev_set_solution(combo0, X.shape[1], X.data_ptr())

    # fused strip kernels: the per-column scalars stay on the device and are all-reduced in place by the driver
    def cg_dots(self, row0, row1, A, B, out):
        self.eng.dev_cg_dots(A.shape[1], row0, row1, A.data_ptr(), B.data_ptr(), out.data_ptr())

    def cg_update(self, node0, node1, first, rz, pAp, P, AP, X, R, Z, out2):
        self.eng.dev_cg_update(X.shape[1], node0, node1, first, rz.data_ptr(), pAp.data_ptr(), P.data_ptr(), AP.data_ptr(),
                               X.data_ptr(), R.data_ptr(), Z.data_ptr(), out2.data_ptr())

    def cg_direction(self, row0, row1, first, rz, rz_new, Z, P):
        self.eng.dev_cg_direction(Z.shape[1], row0, row1, first, rz.data_ptr(), rz_new.data_ptr(), Z.data_ptr(), P.data_ptr())


def solve(ops, combo0, s, tol=1e-10, max_iter=200000, check_every=25, group=None, log=None, graph=True, profile=None):
    """Solves K11 X = B for combos [combo0, combo0 + s) on all ranks of `group`; returns (X on every rank, iterations,
    max relative residual).  Works without torch.distributed initialised (world size 1).

    graph: on CUDA, one iteration (halo exchange, five library kernels, two all-reduces) is recorded once as a CUDA
    graph -- NCCL's sends, receives and all-reduces included -- and replayed; issued call by call from Python an
    iteration costs ~1.8 ms of launch and binding latency whatever the strip size, which is what capped the scaling."""
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    nrs = ops.node_row_start
    ncut = strip_cuts(nrs, world)
    row_cuts = [int(nrs[c]) for c in ncut]
    r0, r1 = row_cuts[rank], row_cuts[rank + 1]
    node0, node1 = ncut[rank], ncut[rank + 1]
    n1 = ops.n1
    dev = ops.device

    # columns my strip references -> halo plan
    if hasattr(ops, 'col_range'):
        lo, hi = ops.col_range(r0, r1)
    else:
        indptr, indices = ops.pattern()
        mine = indices[indptr[r0]:indptr[r1]]
        lo = int(mine.min()) if mine.size else r0
        hi = int(mine.max()) + 1 if mine.size else r1
    if distributed:
        t = torch.tensor([lo, hi], dtype=torch.int64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t, group=group)
        col_lo = [int(x[0]) for x in allt]
        col_hi = [int(x[1]) for x in allt]
    else:
        col_lo, col_hi = [lo], [hi]
    recv, send = halo_plan(rank, world, row_cuts, col_lo, col_hi)
    halo_rows = sum(b - a for _, a, b in recv)

    def allreduce(v):
        if distributed:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        return v

    def exchange(P):
        if not distributed or (not recv and not send):
            return
        ops_list = []
        for q, a, b in send:
            ops_list.append(dist.P2POp(dist.isend, P[a:b], q, group=group))
        for q, a, b in recv:
            ops_list.append(dist.P2POp(dist.irecv, P[a:b], q, group=group))
        for w in dist.batch_isend_irecv(ops_list):
            w.wait()

    B = ops.rhs(combo0, s)
    X = torch.zeros((n1, s), dtype=torch.float64, device=dev)
    R = B.clone()
    Z = torch.zeros_like(R)
    P = torch.zeros_like(R)
    AP = torch.zeros_like(R)
    own = slice(r0, r1)
    fused = hasattr(ops, 'cg_update')
    it = 0
    worst = 0.0
    best, stalled = float('inf'), 0

    def converged(rr, bb):
        nonlocal worst, best, stalled
        worst = float(torch.sqrt((rr/torch.where(bb > 0, bb, torch.ones_like(bb))).max()))
        if log:
            log(it, worst)
        if not np.isfinite(worst) or worst <= tol:
            return True
        # the recurrence residual cannot drop below ~2^-53 cond(K11): stop when it has stopped improving
        if worst < 0.99*best:
            best, stalled = worst, 0
        else:
            stalled += 1
        return stalled >= 8

    ev0 = ev1 = None
    if dev.type == 'cuda':
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if fused:
        # five library kernels, one halo exchange and two all-reduces per iteration; no host round trip in between
        rz = torch.zeros(s, dtype=torch.float64, device=dev)
        pAp = torch.zeros(s, dtype=torch.float64, device=dev)
        red = torch.zeros(2*s, dtype=torch.float64, device=dev)
        ops.cg_update(node0, node1, True, rz, pAp, P, AP, X, R, Z, red)
        allreduce(red)
        rz.copy_(red[:s])
        bb = red[s:].clone()
        ops.cg_direction(r0, r1, True, rz, rz, Z, P)
        if float(bb.max()) == 0.0:
            return X, 0, {'iterations': 0, 'max_rel_residual': 0.0, 'rows_per_rank': [], 'halo_rows': halo_rows, 'halo_bytes_per_iteration': 0}
        def iteration():
            exchange(P)
            ops.spmm_rows(r0, r1, P, AP)
            ops.cg_dots(r0, r1, P, AP, pAp)
            allreduce(pAp)
            ops.cg_update(node0, node1, False, rz, pAp, P, AP, X, R, Z, red)
            allreduce(red)
            ops.cg_direction(r0, r1, False, rz, red, Z, P)          # red[:s] = new r.z
            rz.copy_(red[:s])

        replay = None
        if graph and dev.type == 'cuda' and hasattr(ops, 'set_stream') and max_iter > 8:
            for _ in range(3):                                      # eager iterations: communicators and work buffers exist
                iteration()
                it += 1
            torch.cuda.synchronize(dev)
            cs = torch.cuda.Stream(device=dev)
            cs.wait_stream(torch.cuda.current_stream(dev))
            ops.set_stream(cs)                                      # the library records into the capture stream
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=cs):
                    iteration()
                replay = g.replay
            finally:
                ops.set_stream(torch.cuda.current_stream(dev))
        while it < max_iter:
            if replay is not None:
                replay()
            else:
                iteration()
            it += 1
            if (it % check_every == 0 or it == max_iter) and converged(red[s:], bb):
                break
    else:
        ops.precond_nodes(node0, node1, R, Z)
        red = torch.stack([(R[own]*Z[own]).sum(0), (R[own]*R[own]).sum(0)])
        allreduce(red)
        rz, bb = red[0].clone(), red[1].clone()
        P[own] = Z[own]
        if float(bb.max()) == 0.0:
            return X, 0, {'iterations': 0, 'max_rel_residual': 0.0, 'rows_per_rank': [], 'halo_rows': halo_rows, 'halo_bytes_per_iteration': 0}
        while it < max_iter:
            exchange(P)
            ops.spmm_rows(r0, r1, P, AP)
            pAp = allreduce((P[own]*AP[own]).sum(0))
            alpha = torch.where(pAp != 0, rz/pAp, torch.zeros_like(rz))
            X[own] += alpha*P[own]
            R[own] -= alpha*AP[own]
            ops.precond_nodes(node0, node1, R, Z)
            red = torch.stack([(R[own]*Z[own]).sum(0), (R[own]*R[own]).sum(0)])
            allreduce(red)
            rz_new, rr = red[0], red[1]
            beta = torch.where(rz != 0, rz_new/rz, torch.zeros_like(rz))
            P[own] = Z[own] + beta*P[own]
            rz = rz_new.clone()
            it += 1
            if (it % check_every == 0 or it == max_iter) and converged(rr, bb):
                break
    ms_loop = None
    if ev0 is not None:
        ev1.record()
        torch.cuda.synchronize(dev)
        ms_loop = ev0.elapsed_time(ev1)
    if profile is not None and fused and dev.type == 'cuda':
        # diagnostic: the real closures of this solve, 50 repetitions each (the iterates are garbage afterwards: copies)
        keep = [t.clone() for t in (X, R, Z, P, AP, rz, pAp, red)]

        def timed(fn, reps=50):
            fn()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([a.elapsed_time(b)/reps], dtype=torch.float64, device=dev)
            allreduce_max = t.clone()
            if distributed:
                dist.all_reduce(allreduce_max, op=dist.ReduceOp.MAX, group=group)
            return round(float(allreduce_max.item()), 4)

        profile['exchange'] = timed(lambda: exchange(P))
        profile['iteration_eager'] = timed(iteration)
        if replay is not None:
            profile['iteration_graph'] = timed(replay)
        for t, k in zip((X, R, Z, P, AP, rz, pAp, red), keep):
            t.copy_(k)
    # every rank ends up with the whole solution: all-gather of the strips (ragged -> padded)
    if distributed:
        hmax = max(row_cuts[q + 1] - row_cuts[q] for q in range(world))
        buf = torch.zeros((hmax, s), dtype=torch.float64, device=dev)
        buf[:r1 - r0] = X[own]
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        for q in range(world):
            X[row_cuts[q]:row_cuts[q + 1]] = parts[q][:row_cuts[q + 1] - row_cuts[q]]
    info = {'iterations': it, 'max_rel_residual': worst, 'rows_per_rank': [row_cuts[q + 1] - row_cuts[q] for q in range(world)],
            'halo_rows': halo_rows, 'halo_bytes_per_iteration': halo_rows*s*8, 'ms_iteration_loop': ms_loop}
    return X, it, info
