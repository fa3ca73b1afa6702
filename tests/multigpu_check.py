"""Data-parallel equivalence check (last row): the same global minibatches trained on 1 GPU and sharded over
G GPUs (gradient all-reduce over NCCL) give the same parameters up to fp32 summation order.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py

Rank 0 prints one JSON line.  (Needs >= 2 GPUs, so it is run by hand under `gpurun --gpus 2`, not by `pytest -m gpu`.)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.dist import Comm, split_minibatch   # noqa: E402
from robosumo_selfplay_b200.model import PPOModel               # noqa: E402

D, A, N, NMB, STEPS = 121, 8, 8192, 4, 3


def data():
    rng = np.random.RandomState(0)
    return dict(obs=rng.randn(N, D).astype(np.float32), act=(rng.randn(N, A) * 0.5).astype(np.float32), ret=(rng.randn(N) * 2).astype(np.float32),
                val=rng.randn(N).astype(np.float32), old=(8 + rng.randn(N)).astype(np.float32))


def train(model, dd, lo, hi, comm, precision):
    dev = model.device
    t = {k: torch.as_tensor(v[lo:hi], device=dev) for k, v in dd.items()}
    rng = np.random.RandomState(1)
    inds = np.arange(N)
    for _ in range(STEPS):
        rng.shuffle(inds)
        for s in range(0, N, N // NMB):
            mb = split_minibatch(inds[s:s + N // NMB], lo, hi)
            model.train_indexed(1e-3, 0.2, t['obs'], t['ret'], t['act'], t['val'], t['old'], None, torch.as_tensor(mb, device=dev), global_n=N // NMB)
    return model.get_flat()


if __name__ == '__main__':
    comm = Comm()
    torch.cuda.set_device(comm.local_rank)
    dd = data()
    out = {}
    for precision in ('fp32', 'tf32'):
        np.random.seed(3)
        ref = PPOModel(ob_dim=D, ac_dim=A, device=comm.local_rank, precision=precision)          # every rank: full-batch single-GPU reference
        init = ref.get_flat()
        p_single = train(ref, dd, 0, N, None, precision)
        np.random.seed(3)
        m = PPOModel(ob_dim=D, ac_dim=A, device=comm.local_rank, comm=comm, precision=precision)
        lo, hi = comm.shard(N)
        p_multi = train(m, dd, lo, hi, comm, precision)
        d = np.abs(p_single - p_multi)
        out[precision] = dict(max_abs_diff=float(d.max()), moved=float(np.abs(p_single - init).max()), worst_index=int(d.argmax()),
                              n_above_1e_6=int((d > 1e-6).sum()), median_abs_diff=float(np.median(d)), n_params=int(d.size))
    if comm.rank == 0:
        print(json.dumps(dict(world=comm.world, **out)))
    comm.barrier()
