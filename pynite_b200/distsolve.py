"""Multi-GPU direct solve: the host side of `pn_set_distributed` (subtree-to-GPU multifrontal).

One process per GPU.  Every rank uploads the whole model; the library cuts the elimination tree at the highest level
with at least `world` fronts, factorises and sweeps this rank's subtrees for all right-hand sides, and asks for three
all-gathers per factor + solve through a callback (update matrices of the subtree roots, their update vectors, the
solved rows).  This module supplies that callback:

  * `attach(engine)`            -- `torch.distributed` (NCCL): `all_gather_into_tensor` on views of the library's
                                   device buffers, one process per GPU;
  * `LocalGroup(world)`         -- the same exchange between `world` contexts of ONE process on ONE GPU (one Python
                                   thread per rank).  Used by the single-GPU tests: it runs the real kernels and the
                                   real pack / unpack paths of every rank, only the transport differs.

The reference has no counterpart (SuperLU inside `spsolve` is sequential, `Analysis.py:217`); results are bit-identical
to the single-GPU solve of this library.
"""
from __future__ import annotations

import ctypes as ct
import threading

from pynite_b200 import engine as _eng

_ALLGATHER_FN = ct.CFUNCTYPE(ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64)


class _DevView:
    """Zero-copy view of `n` doubles of device memory for `torch.as_tensor` (CUDA array interface, version 2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


def _tensor(ptr, n, device):
    import torch
    return torch.as_tensor(_DevView(ptr, n), device=device)


def _register(engine, rank, world, fn):
    cb = _ALLGATHER_FN(fn)
    engine._dist_cb = cb                      # keep the trampoline alive as long as the engine
    lib = _eng.load_library()
    engine._check(lib.pn_set_distributed(engine._h, rank, world, ct.cast(cb, ct.c_void_p), None))


def attach(engine, group=None):
    """Distributes `engine`'s direct solver over the ranks of the default (or given) torch.distributed process group.
    Call once per engine, on every rank, before the solve.  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return rank, world
    device = torch.device('cuda', engine.device)
    views = {}

    def allgather(_user, send_ptr, recv_ptr, seg):
        try:
            key = (send_ptr, recv_ptr, seg)
            if key not in views:
                if len(views) > 16:
                    views.clear()
                views[key] = (_tensor(send_ptr, seg, device), _tensor(recv_ptr, seg*world, device))
            send, recv = views[key]
            dist.all_gather_into_tensor(recv, send, group=group)
            torch.cuda.current_stream(device).synchronize()
            return 0
        except Exception as err:          # noqa: BLE001 -- must not propagate through the C frame
            print('pynite_b200.distsolve: all-gather failed:', err)
            return 1
    _register(engine, rank, world, allgather)
    # rank 0 computes the ordering for all ranks (share_ordering in pn_direct.cu): it gets the cores of the box less one
    # per other rank, the others keep one worker thread (result zero fill)
    import os
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', str(world)) or world)
    if local_world > 1:
        engine.set_option('host_threads', max(1, (os.cpu_count() or 1) - local_world) if rank == 0 else 1)
    return rank, world


def detach(engine):
    lib = _eng.load_library()
    engine._check(lib.pn_set_distributed(engine._h, 0, 1, None, None))
    engine._dist_cb = None


class LocalGroup:
    """`world` ranks as threads of one process sharing one GPU.

    Only one rank's kernels are in flight at any time: a rank holds `gpu` while it is inside the library and hands it
    over (after a device synchronisation) when it waits in an exchange.  (The left-looking sweep kernel keeps all its
    CTAs resident and polls flags written by its own CTAs; two such kernels from different contexts must not share
    the GPU.)"""

    def __init__(self, world, device=0):
        self.world = world
        self.device = device
        self.gpu = threading.Lock()
        self.barrier = threading.Barrier(world)
        self.sends = [None]*world
        self.errors = []

    def attach(self, engine, rank):
        import torch
        device = torch.device('cuda', self.device)
        world = self.world

        def allgather(_user, send_ptr, recv_ptr, seg):
            try:
                torch.cuda.synchronize(device)
                self.sends[rank] = _tensor(send_ptr, seg, device)
                self.gpu.release()
                try:
                    self.barrier.wait(timeout=90)
                    recv = _tensor(recv_ptr, seg*world, device)
                    for r in range(world):
                        recv[r*seg:(r + 1)*seg].copy_(self.sends[r])
                    torch.cuda.synchronize(device)
                    self.barrier.wait(timeout=90)
                finally:
                    self.gpu.acquire()
                return 0
            except Exception as err:      # noqa: BLE001
                self.errors.append(err)
                self.barrier.abort()
                return 1
        _register(engine, rank, world, allgather)

    def run(self, fn):
        """Runs fn(rank) on `world` threads, each holding the GPU lock while it computes; returns the results."""
        out = [None]*self.world
        errs = [None]*self.world

        def work(r):
            self.gpu.acquire()
            try:
                out[r] = fn(r)
            except Exception as err:      # noqa: BLE001
                errs[r] = err
                try:
                    self.barrier.abort()
                except Exception:         # noqa: BLE001
                    pass
            finally:
                self.gpu.release()
        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        return out
