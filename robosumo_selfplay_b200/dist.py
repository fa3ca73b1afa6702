"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (NVLink 5 / NVSwitch) on a box, gloo on CPU tests.

Only two exchanges exist on this path: the flat gradient(+stats) all-reduce once per minibatch, with the
3-scalar advantage-moment all-reduce in front of it, and the opponent-snapshot broadcast once per update.  Envs and their
trajectories never leave the GPU that stepped them.  (The reference has no collective on this path; the vendored pattern
this mirrors is MpiAdamOptimizer's flat all-reduce, baselines/baselines/common/mpi_adam_optimizer.py:21-46.)
"""
import os


class Comm:
    def __init__(self, backend=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
            kw = {}
            if backend == 'nccl':
                kw['device_id'] = torch.device('cuda', self.local_rank)
            dist.init_process_group(backend, **kw)

    def all_reduce_sum(self, tensor):
        if self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        return tensor

    def broadcast(self, tensor, src=0):
        if self.world > 1:
            self.dist.broadcast(tensor, src=src)
        return tensor

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def shard(self, n_global):
        """Contiguous env range of this rank: [lo, hi)."""
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per


def split_minibatch(global_idx, lo, hi):
    """Part of a GLOBAL minibatch index slice that falls in this rank's sample range [lo, hi), as local indices.
    Every rank replays the same host-side shuffle, so no sample moves between GPUs."""
    import numpy as np
    g = np.asarray(global_idx)
    m = (g >= lo) & (g < hi)
    return (g[m] - lo).astype(np.int32)


def legacy_shuffle(inds):
    """`np.random.shuffle(inds)` on NumPy's legacy global RandomState (the stream the reference's minibatch schedule draws from,
    alg_ppo.py:89,364), replayed bit-exactly by the library's host routine `rs_legacy_shuffle`: same permutation, same generator
    state afterwards; the swap partners are drawn a block ahead and prefetched (on par with NumPy at 0.5 M indices, 2.3x faster at
    the 8.4 M of T=2048), and it needs no Python-level generator object on the hot loop."""
    import ctypes
    import numpy as np
    from . import _lib
    assert isinstance(inds, np.ndarray) and inds.ndim == 1 and inds.dtype == np.int64 and inds.flags.c_contiguous
    name, key, pos, has_gauss, cached = np.random.get_state()
    assert name == 'MT19937'
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    p = ctypes.c_int(int(pos))
    _lib.check(_lib.lib().rs_legacy_shuffle(ctypes.c_void_p(key.ctypes.data), ctypes.byref(p), ctypes.c_void_p(inds.ctypes.data),
                                            ctypes.c_longlong(inds.shape[0])))
    np.random.set_state((name, key, p.value, has_gauss, cached))
    return inds


class EpochPermutations:
    """The `noptepochs` successive `np.random.shuffle(inds)` permutations of one update (alg_ppo.py:357-364), computed AHEAD on a
    helper thread: one epoch ahead by default, so that the host shuffle of 0.5-8 M indices overlaps the GPU work of the current
    epoch, or all of them (`ahead=nepochs`) when the caller creates the object before the rollout.  The helper works on a
    private copy of the legacy MT19937 state; the global `np.random` state is advanced only when a permutation is handed out,
    so an early stop (kl_threshold) leaves the stream exactly where the reference would: nothing speculative is committed.
    Nobody else may draw from `np.random` between construction and the last `next()` (the rollout and the update do not)."""

    def __init__(self, n, nepochs, ahead=1):
        import numpy as np
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        name, key, pos, self._hg, self._cg = np.random.get_state()
        assert name == 'MT19937'
        self._key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        self._pos = int(pos)
        self._inds = np.arange(n)
        self._left = int(nepochs)
        self._pool = ThreadPoolExecutor(max_workers=1)       # one worker: the shuffles continue one generator stream, in order
        self._queue = deque()
        for _ in range(max(1, int(ahead))):
            self._launch()

    def _compute(self):
        import ctypes
        from . import _lib
        p = ctypes.c_int(self._pos)
        _lib.check(_lib.lib().rs_legacy_shuffle(ctypes.c_void_p(self._key.ctypes.data), ctypes.byref(p),
                                                ctypes.c_void_p(self._inds.ctypes.data), ctypes.c_longlong(self._inds.shape[0])))
        self._pos = p.value
        return self._inds.copy(), self._key.copy(), self._pos

    def _launch(self):
        if self._left > 0:
            self._left -= 1
            self._queue.append(self._pool.submit(self._compute))

    def __iter__(self):
        return self

    def __next__(self):
        import numpy as np
        if not self._queue:
            self._pool.shutdown(wait=False)
            raise StopIteration
        perm, key, pos = self._queue.popleft().result()
        np.random.set_state(('MT19937', key, pos, self._hg, self._cg))
        self._launch()
        return perm

    def close(self):
        self._left = 0
        while self._queue:
            self._queue.popleft().result()
        self._pool.shutdown(wait=False)
