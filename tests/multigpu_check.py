"""Data-parallel equivalence worker (last row), launched by tests/test_multigpu.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py

Checks, on G >= 2 GPUs over NCCL: (1) the same global minibatches trained on 1 GPU and sharded over G GPUs (device-side split of
the global permutation, one advantage-moment all-reduce per epoch, one gradient all-reduce per minibatch) give the same
parameters up to fp32 summation order; (2) ragged shards (opponent-data reuse) partition every minibatch exactly; (3) learn()
with opponent_mode='ours' picks the same opponent index on every rank and keeps the parameters replicated.
Rank 0 prints one JSON line with "ok": true/false."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.dist import Comm, EpochSchedule, host_split   # noqa: E402
from robosumo_selfplay_b200.model import PPOModel             # noqa: E402

D, A, N, NMB, STEPS = 121, 8, 8192, 4, 3


def data():
    rng = np.random.RandomState(0)
    return dict(obs=rng.randn(N, D).astype(np.float32), act=(rng.randn(N, A) * 0.5).astype(np.float32), ret=(rng.randn(N) * 2).astype(np.float32),
                val=rng.randn(N).astype(np.float32), old=(8 + rng.randn(N)).astype(np.float32))


def train(model, dd, lo, hi, comm):
    dev = model.device
    t = {k: torch.as_tensor(v[lo:hi], device=dev) for k, v in dd.items()}
    rng = np.random.RandomState(1)
    inds = np.arange(N)
    sched = EpochSchedule(dev, N, N // NMB, lo, hi, comm)
    counts = []
    for _ in range(STEPS):
        rng.shuffle(inds)
        # data-parallel: alternate between the device-side cut of the permutation (rs_epoch_split) and the host-side one (host_split)
        perm = host_split(inds, N // NMB, lo, hi) if (comm is not None and _ % 2 == 1) else inds
        for idx, n_loc, gn, sums in sched.load(perm, t['ret'], t['val']):
            counts.append(n_loc)
            model.train_indexed(1e-3, 0.2, t['obs'], t['ret'], t['act'], t['val'], t['old'], None, idx, global_n=gn, adv_sums=sums)
    return model.get_flat(), counts


if __name__ == '__main__':
    comm = Comm()
    torch.cuda.set_device(comm.local_rank)
    dev = torch.device('cuda', comm.local_rank)
    dd = data()
    out, ok = {}, True
    for precision in ('fp32', 'tf32'):
        np.random.seed(3)
        ref = PPOModel(ob_dim=D, ac_dim=A, device=comm.local_rank, precision=precision)          # every rank: full-batch single-GPU reference
        init = ref.get_flat()
        p_single, _ = train(ref, dd, 0, N, None)
        np.random.seed(3)
        m = PPOModel(ob_dim=D, ac_dim=A, device=comm.local_rank, comm=comm, precision=precision)
        # ragged shards: rank r owns [lo, hi) with unequal sizes, as under opponent-data reuse
        sizes = [N // comm.world + (37 if r % 2 == 0 else -37) for r in range(comm.world)]
        sizes[-1] += N - sum(sizes)
        lo = sum(sizes[:comm.rank]); hi = lo + sizes[comm.rank]
        p_multi, counts = train(m, dd, lo, hi, comm)
        m.check_peer()
        peer_used = m._peer_h not in (None, False)
        # the same training with the gradient all-reduce through torch.distributed / NCCL instead of the peer-memory kernel
        np.random.seed(3)
        os.environ['RS_B200_PEER'] = '0'
        m2 = PPOModel(ob_dim=D, ac_dim=A, device=comm.local_rank, comm=comm, precision=precision)
        p_nccl, _ = train(m2, dd, lo, hi, comm)
        os.environ['RS_B200_PEER'] = '1'
        assert m2._peer_h is None
        tot = torch.tensor(counts, dtype=torch.int64, device=dev)
        comm.all_reduce_sum(tot)
        ok &= bool((tot == N // NMB).all())                                   # every global minibatch is partitioned exactly
        d = np.abs(p_single - p_multi)
        moved = float(np.abs(p_single - init).max())
        dn = np.abs(p_nccl - p_multi)
        out[precision] = dict(max_abs_diff=float(d.max()), moved=moved, worst_index=int(d.argmax()),
                              n_above_1e_6=int((d > 1e-6).sum()), median_abs_diff=float(np.median(d)), n_params=int(d.size),
                              peer_allreduce=bool(peer_used), peer_vs_nccl_max_abs_diff=float(dn.max()))
        ok &= dn.max() <= (2e-6 if precision == 'fp32' else 0.03 * moved)      # rank-order sums (peer kernel) vs NCCL's order
        # fp32 pipe: summation-order differences only; tf32: the tensor core's grouping-dependent accumulation, amplified by Adam
        ok &= d.max() <= (2e-6 if precision == 'fp32' else 0.03 * moved)
    # learn() data-parallel with opponent_mode='ours' and opponent-data reuse: identical opponent choices and parameters on every rank
    from robosumo_selfplay_b200 import alg_ppo
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=32, seed=100 + comm.rank, device=comm.local_rank, device_api=True)
    model = alg_ppo.learn(env=env, total_timesteps=3 * 32 * 8 * comm.world, seed=5, nsteps=8, nminibatches=4, noptepochs=2, lr=1e-3, gamma=0.995, lam=1.0,
                          rho_bar=10., c_bar=1., log_interval=100, anneal_bound=1000, opponent_mode='ours', use_opponent_data='off_policy',
                          neglogp_threshold=11.5, comm=comm)
    opp = torch.tensor([h['opponent'] for h in model.history], dtype=torch.int64, device=dev)
    lo_, hi_ = opp.clone(), opp.clone()
    torch.distributed.all_reduce(lo_, op=torch.distributed.ReduceOp.MIN); torch.distributed.all_reduce(hi_, op=torch.distributed.ReduceOp.MAX)
    same_opp = bool((lo_ == hi_).all())
    pmin, pmax = model.params.clone(), model.params.clone()
    torch.distributed.all_reduce(pmin, op=torch.distributed.ReduceOp.MIN); torch.distributed.all_reduce(pmax, op=torch.distributed.ReduceOp.MAX)
    same_params = bool((pmin == pmax).all())
    samples = [h['samples'] for h in model.history]
    out['learn'] = dict(opponents=opp.cpu().tolist(), same_opponent_on_all_ranks=same_opp, params_replicated=same_params, samples=samples,
                        finite=bool(torch.isfinite(model.params).all()))
    ok &= same_opp and same_params and out['learn']['finite'] and all(s > 32 * 8 * comm.world for s in samples)
    env.close()
    if comm.rank == 0:
        print(json.dumps(dict(world=comm.world, ok=bool(ok), **out)))
    comm.barrier()
    torch.distributed.destroy_process_group()
