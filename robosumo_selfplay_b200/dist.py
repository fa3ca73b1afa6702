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

    def all_gather_int(self, value, device=None):
        """[value of rank 0, ..., value of rank world-1] as a list of Python ints (one small collective)."""
        if self.world == 1:
            return [int(value)]
        t = self.torch
        dev = device if device is not None else (t.device('cuda', self.local_rank) if t.cuda.is_available() and self.dist.get_backend() == 'nccl' else 'cpu')
        x = t.zeros(self.world, dtype=t.int64, device=dev)
        x[self.rank] = int(value)
        self.dist.all_reduce(x, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in x.cpu().tolist()]

    def broadcast_int(self, value, src=0, device=None):
        """rank `src`'s Python int on every rank."""
        if self.world == 1:
            return int(value)
        t = self.torch
        dev = device if device is not None else (t.device('cuda', self.local_rank) if t.cuda.is_available() and self.dist.get_backend() == 'nccl' else 'cpu')
        x = t.tensor([int(value)], dtype=t.int64, device=dev)
        self.dist.broadcast(x, src=src)
        return int(x.item())

    def make_peer(self, nfloats, device_index):
        """An rs_peer handle for the gradient all-reduce over NVLink peer memory (include/rs_b200.h), or None when this job cannot
        use it (one rank, CPU / gloo, ranks on several nodes, RS_B200_PEER=0): the caller then falls back to all_reduce_sum (NCCL).
        Collective: every rank must call it at the same point."""
        t = self.torch
        if self.world == 1 or not t.cuda.is_available() or self.dist.get_backend() != 'nccl' or os.environ.get('RS_B200_PEER', '1') == '0':
            return None
        if int(os.environ.get('LOCAL_WORLD_SIZE', str(self.world))) != self.world:
            return None                                   # CUDA IPC handles do not cross nodes
        import ctypes
        from . import _lib
        L = _lib.lib()
        h = ctypes.c_void_p()
        nb = L.rs_peer_handle_bytes()
        mine = (ctypes.c_ubyte * nb)()
        dev = t.device('cuda', device_index)
        ok = L.rs_peer_create(self.rank, self.world, int(nfloats), int(device_index), ctypes.byref(h)) == 0 and L.rs_peer_export(h, mine) == 0
        gathered = t.empty(self.world * nb, dtype=t.uint8, device=dev)
        self.dist.all_gather_into_tensor(gathered, t.tensor(list(mine), dtype=t.uint8, device=dev))
        if ok:
            ok = L.rs_peer_connect(h, gathered.cpu().numpy().tobytes()) == 0
        # all or nothing: a rank that cannot map a peer (no P2P path between two GPUs, IPC disabled in a container) sends everybody
        # back to the NCCL all-reduce; this all-reduce is also the barrier after which peers may signal each other
        agree = t.tensor([1 if ok else 0], dtype=t.int32, device=dev)
        self.dist.all_reduce(agree, op=self.dist.ReduceOp.MIN)
        if int(agree.item()) == 0:
            if h:
                L.rs_peer_destroy(h)
            return None
        return h

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


class SplitPerm:
    """One epoch's permutation already cut into this rank's minibatch parts on the host: idx [nmb][cap] local indices, counts [nmb]."""
    __slots__ = ('idx', 'counts', 'n_total', 'cap')

    def __init__(self, idx, counts, n_total):
        self.idx, self.counts, self.n_total, self.cap = idx, counts, int(n_total), int(idx.shape[1])


def host_split(perm, nbt, lo, hi):
    """The data-parallel cut of a global permutation, on the host (safe on the helper thread of EpochPermutations):
    per minibatch m the entries of perm[m*nbt:(m+1)*nbt] inside this rank's sample range [lo, hi), in order, as local indices
    (one pass of the library's host routine rs_epoch_split_host, which releases the GIL).
    With N ranks the global permutation has N x as many entries as a rank trains on; cutting it where it is drawn keeps the per-epoch
    upload at the rank's own 4 bytes per sample instead of the whole permutation (8 GPUs: 2 MB instead of 16 MB per epoch and rank)."""
    import ctypes
    import numpy as np
    from . import _lib
    n = int(perm.shape[0]); nbt = int(nbt)
    nmb = (n + nbt - 1) // nbt
    pc = np.ascontiguousarray(perm) if perm.dtype in (np.int32, np.int64) else np.ascontiguousarray(perm, dtype=np.int64)
    local = np.empty(int(hi - lo) + 1, np.int32); counts = np.empty(nmb, np.int32)
    _lib.check(_lib.lib().rs_epoch_split_host(ctypes.c_void_p(pc.ctypes.data), pc.dtype.itemsize, ctypes.c_longlong(n), nbt, ctypes.c_longlong(int(lo)),
                                             ctypes.c_longlong(int(hi)), ctypes.c_void_p(local.ctypes.data), ctypes.c_void_p(counts.ctypes.data)))
    cap = max(int(counts.max()), 1)
    idx = np.zeros((nmb, cap), np.int32)
    off = np.concatenate([[0], np.cumsum(counts)])
    for m in range(nmb):
        idx[m, :counts[m]] = local[off[m]:off[m + 1]]
    return SplitPerm(idx, counts, n)


class EpochSchedule:
    """The minibatch schedule of one epoch ON THE DEVICE (replaces the per-minibatch NumPy masks of round 1).

    Every rank uploads the same global permutation once per epoch (int32, through a page-locked staging buffer); on one GPU a
    minibatch is a slice of it, on several `rs_epoch_split` compacts, per minibatch and in order, the entries of this rank's sample
    range [lo, hi) into local indices.  `rs_adv_moments_multi` then computes the advantage moments of ALL minibatches in one
    launch and, data-parallel, they are all-reduced ONCE per epoch (returns and values are constant during an update), so a
    minibatch step costs exactly one collective: the flat gradient.  `parts()` hands out (idx, n_local, n_global, adv_sums)."""

    def __init__(self, device, n_total, nbatch_train, lo=0, hi=None, comm=None):
        import torch
        from . import _lib
        self.t, self.L, self._lib = torch, _lib.lib(), _lib
        self.device, self.comm = device, comm
        self.n_total, self.nbt = int(n_total), int(nbatch_train)
        self.lo, self.hi = int(lo), int(n_total if hi is None else hi)
        self.world = comm.world if comm is not None else 1
        self.nmb = (self.n_total + self.nbt - 1) // self.nbt
        self.perm = torch.empty(self.n_total, dtype=torch.int32, device=device)
        self.stage = torch.empty(self.n_total, dtype=torch.int32, pin_memory=torch.cuda.is_available())
        self.sums = torch.empty((self.nmb, 2), dtype=torch.float64, device=device)
        self.copied = None
        if self.world > 1:
            self.idx = torch.empty(self.nmb * self.nbt, dtype=torch.int32, device=device)
            self.counts = torch.empty(self.nmb, dtype=torch.int32, device=device)

    def _p(self, x):
        import ctypes
        return ctypes.c_void_p(x.data_ptr()) if x is not None else None

    def load(self, perm, returns, values):
        """perm: the epoch's global permutation (numpy, any integer dtype), or its host-side cut (SplitPerm, data-parallel).
        Returns the list of minibatch parts."""
        import ctypes
        import numpy as np
        t = self.t
        if isinstance(perm, SplitPerm):
            return self._load_split(perm, returns, values)
        if self.copied is not None:
            self.copied.synchronize()                       # the previous upload has left the staging buffer
        np.copyto(self.stage.numpy(), perm, casting='unsafe')
        self.perm.copy_(self.stage, non_blocking=True)
        if t.cuda.is_available():
            self.copied = t.cuda.Event(); self.copied.record(t.cuda.current_stream(self.device))
        st = ctypes.c_void_p(t.cuda.current_stream(self.device).cuda_stream)
        sizes = [min(self.nbt, self.n_total - m * self.nbt) for m in range(self.nmb)]
        if self.world == 1:
            self._lib.check(self.L.rs_adv_moments_multi(self._p(self.perm), None, self.nmb, self.nbt, self.n_total, self._p(returns), self._p(values),
                                                        self._p(self.sums), st))
            return [(self.perm[m * self.nbt:m * self.nbt + sizes[m]], sizes[m], sizes[m], self.sums[m]) for m in range(self.nmb)]
        self._lib.check(self.L.rs_epoch_split(self._p(self.perm), self.n_total, self.nbt, self.lo, self.hi, self._p(self.idx), self._p(self.counts), st))
        self._lib.check(self.L.rs_adv_moments_multi(self._p(self.idx), self._p(self.counts), self.nmb, self.nbt, self.n_total, self._p(returns),
                                                    self._p(values), self._p(self.sums), st))
        self.comm.all_reduce_sum(self.sums)                 # ONE collective per epoch for all advantage moments
        counts = self.counts.cpu().tolist()                 # one small D2H per epoch: the launch geometry of the local parts
        return [(self.idx[m * self.nbt:m * self.nbt + counts[m]], counts[m], sizes[m], self.sums[m]) for m in range(self.nmb)]


    def _load_split(self, sp, returns, values):
        """Data-parallel with the cut already made on the host: ONE upload of [counts | idx] (this rank's 4 bytes per sample), the
        advantage moments of all minibatches in one launch, one all-reduce of them per epoch; no device-to-host copy."""
        import ctypes
        t = self.t
        assert sp.n_total == self.n_total and sp.idx.shape[0] == self.nmb
        need = self.nmb + self.nmb * sp.cap
        if getattr(self, '_sbuf', None) is None or self._sbuf.numel() < need:
            self._sbuf = t.empty(need + need // 8, dtype=t.int32, device=self.device)
            self._sstage = t.empty(need + need // 8, dtype=t.int32, pin_memory=t.cuda.is_available())
        if self.copied is not None:
            self.copied.synchronize()
        h = self._sstage.numpy()
        h[:self.nmb] = sp.counts
        h[self.nmb:need] = sp.idx.ravel()
        self._sbuf[:need].copy_(self._sstage[:need], non_blocking=True)
        if t.cuda.is_available():
            self.copied = t.cuda.Event(); self.copied.record(t.cuda.current_stream(self.device))
        st = ctypes.c_void_p(t.cuda.current_stream(self.device).cuda_stream)
        counts_dev, idx_dev = self._sbuf[:self.nmb], self._sbuf[self.nmb:need]
        self._lib.check(self.L.rs_adv_moments_multi(self._p(idx_dev), self._p(counts_dev), self.nmb, sp.cap, self.n_total, self._p(returns),
                                                    self._p(values), self._p(self.sums), st))
        if self.comm is not None:
            self.comm.all_reduce_sum(self.sums)
        sizes = [min(self.nbt, self.n_total - m * self.nbt) for m in range(self.nmb)]
        counts = [int(c) for c in sp.counts]
        return [(idx_dev[m * sp.cap:m * sp.cap + counts[m]], counts[m], sizes[m], self.sums[m]) for m in range(self.nmb)]


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

    def __init__(self, n, nepochs, ahead=1, dtype=None, split=None):
        import numpy as np
        self._split = split                 # (nbatch_train, lo, hi): data-parallel, hand out the host-side cut of this rank (SplitPerm) instead of the permutation
        self._dtype = dtype                 # e.g. np.int32: the helper thread hands out the permutation already narrowed for the device upload
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        name, key, pos, self._hg, self._cg = np.random.get_state()
        assert name == 'MT19937'
        self._key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        self._pos = int(pos)
        self._inds = np.arange(n, dtype=np.int32 if n < 2**31 - 1 else np.int64)      # int32: half the bytes under the shuffle's random accesses
        self._left = int(nepochs)
        self._pool = ThreadPoolExecutor(max_workers=1)       # one worker: the shuffles continue one generator stream, in order
        self._pool2 = ThreadPoolExecutor(max_workers=1) if split is not None else None      # the cut of epoch k runs beside the shuffle of epoch k + 1
        self._queue = deque()
        for _ in range(max(1, int(ahead))):
            self._launch()

    def _compute(self):
        import ctypes
        from . import _lib
        p = ctypes.c_int(self._pos)
        fn = _lib.lib().rs_legacy_shuffle32 if self._inds.dtype.itemsize == 4 else _lib.lib().rs_legacy_shuffle
        _lib.check(fn(ctypes.c_void_p(self._key.ctypes.data), ctypes.byref(p), ctypes.c_void_p(self._inds.ctypes.data), ctypes.c_longlong(self._inds.shape[0])))
        self._pos = p.value
        if self._split is not None:
            return self._inds.copy(), self._key.copy(), self._pos
        import numpy as np
        return self._inds.astype(self._dtype if self._dtype is not None else np.int64), self._key.copy(), self._pos

    def _cut(self, fut):
        perm, key, pos = fut.result()
        return host_split(perm, *self._split), key, pos

    def _launch(self):
        if self._left > 0:
            self._left -= 1
            f = self._pool.submit(self._compute)
            self._queue.append(self._pool2.submit(self._cut, f) if self._pool2 is not None else f)

    def __iter__(self):
        return self

    def __next__(self):
        import numpy as np
        if not self._queue:
            self._pool.shutdown(wait=False)
            if self._pool2 is not None:
                self._pool2.shutdown(wait=False)
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
        if self._pool2 is not None:
            self._pool2.shutdown(wait=False)
