"""learn() -- drop-in for the reference's self-play PPO loop (alg_ppo.py:25-513) on one or several B200s.

Same control flow and hyper-parameter surface; what changed is where the work happens:
  * rollout, V-trace and every minibatch step run on the GPU (runner.py, model.py of this package);
  * the minibatch schedule is still produced by NumPy's legacy global RandomState in the reference's call order
    (seed -> 3 x 6 ortho_init draws -> per-update opponent draw -> per-epoch shuffle), so the permutation is bit-exact;
  * with several ranks (one process per GPU) every rank replays the same schedule, trains on the part of each global
    minibatch that lives in its own env shard, and the flat gradient is all-reduced over NCCL once per minibatch;
    the sampled opponent snapshot is broadcast from rank 0 once per update.
Out of scope here: matplotlib figures, TensorBoard summaries.
"""
import os
import os.path as osp
import random
import time
from collections import deque

import numpy as np

from .model import PPOModel
from .runner import Runner
from .dist import split_minibatch, EpochPermutations


def constfn(val):
    def f(_):
        return val
    return f


def set_global_seeds(seed):
    """baselines/baselines/common/misc_util.py:48-62 without the TF part."""
    if seed is not None:
        np.random.seed(seed)
        random.seed(seed)


def explained_variance(ypred, y):
    """baselines/baselines/common/math_util.py:25-38."""
    vary = np.var(y)
    return np.nan if vary == 0 else 1 - np.var(y - ypred) / vary


def safemean(xs):
    return np.nan if len(xs) == 0 else np.mean(xs)


class KVLogger:
    """Minimal stand-in for baselines.logger: stdout table + progress.csv (keys of alg_ppo.py:444-456)."""

    def __init__(self, log_dir, enabled=True):
        self.dir, self.enabled, self.kvs, self.rows, self.keys = log_dir, enabled, {}, [], []
        if log_dir and enabled:
            os.makedirs(log_dir, exist_ok=True)

    def logkv(self, k, v):
        self.kvs[k] = v

    def info(self, msg):
        if self.enabled:
            print(msg, flush=True)

    def dumpkvs(self):
        if not self.enabled:
            self.kvs = {}
            return
        for k in self.kvs:
            if k not in self.keys:
                self.keys.append(k)
        self.rows.append(dict(self.kvs))
        print(' | '.join('%s %.5g' % (k, v) if isinstance(v, (int, float, np.floating)) else '%s %s' % (k, v) for k, v in self.kvs.items()), flush=True)
        if self.dir:
            with open(osp.join(self.dir, 'progress.csv'), 'w') as f:
                f.write(','.join(self.keys) + '\n')
                for r in self.rows:
                    f.write(','.join(str(r.get(k, '')) for k in self.keys) + '\n')
        self.kvs = {}


def learn(*, network='mlp', env, total_timesteps, opponent_mode='random', use_opponent_data=None, seed=None, nsteps=2048, ent_coef=0.0,
          lr=3e-4, vf_coef=0.5, max_grad_norm=0.5, gamma=0.99, lam=0.95, rho_bar=1., c_bar=1., log_interval=10, nminibatches=4,
          noptepochs=4, cliprange=0.2, save_interval=1, load_path=None, nagent=2, anneal_bound=500, vgap=None, kl_threshold=None,
          neglogp_threshold=10000., log_dir=None, comm=None, update_fn=None, max_snapshots=30, **network_kwargs):
    import torch
    assert network == 'mlp' and nagent == 2
    assert network_kwargs.get('num_hidden', 64) == 64 and network_kwargs.get('value_network', 'copy') == 'copy'
    set_global_seeds(seed)
    if isinstance(lr, float): lr = constfn(lr)
    if isinstance(cliprange, float): cliprange = constfn(cliprange)
    total_timesteps = int(total_timesteps)
    world = comm.world if comm is not None else 1
    rank = comm.rank if comm is not None else 0
    logger = KVLogger(log_dir, enabled=(rank == 0))

    nenvs_local = env.num_envs
    nenvs = nenvs_local * world
    D, A = env.observation_space[0].shape[0], env.action_space[0].shape[0]
    nbatch_local = nenvs_local * nsteps
    nbatch = nenvs * nsteps
    nbatch_train = nbatch // nminibatches
    device = getattr(env, 'device', torch.device('cuda', 0))

    mk = lambda scope, trainable: PPOModel(ob_dim=D, ac_dim=A, ent_coef=ent_coef, vf_coef=vf_coef, max_grad_norm=max_grad_norm,
                                           trainable=trainable, model_scope=scope, device=device, comm=comm if trainable else None)
    model = mk('model_0', True)                       # creation order = np.random draw order (alg_ppo.py:117-133)
    models = [model, mk('model_1', False)]
    model_util = mk('model_util', False)
    if comm is not None:
        comm.broadcast(model.params, 0)
    checkdir = osp.join(log_dir, 'checkpoints') if log_dir else None
    snapshots = {}                                    # version -> flat params (host); the checkpoint directory doubles as the pool
    def save(version):
        snapshots[version] = model.get_flat()
        if checkdir and rank == 0:
            model.save(osp.join(checkdir, '%.5i' % version))
    save(0)
    if load_path is not None:
        for m in models:
            m.load(load_path)

    runner = Runner(env=env, models=models, nsteps=nsteps, nagent=nagent, gamma=gamma, lam=lam, rho_bar=rho_bar, c_bar=c_bar,
                    anneal_bound=anneal_bound, seed=(seed or 0) * 7919 + rank)
    epinfobuf = deque(maxlen=100)
    tfirststart = time.perf_counter()
    version_gap, history = [], []
    lo, hi = rank * nbatch_local, (rank + 1) * nbatch_local
    prev = None

    nupdates = total_timesteps // nbatch
    for update in range(1, nupdates + 1):
        assert nbatch % nminibatches == 0
        tstart = time.perf_counter()
        frac = 1.0 - (update - 1.0) / nupdates
        lrnow, cliprangenow = lr(frac), cliprange(frac)

        # ---- opponent (alg_ppo.py:192-247); every rank draws from the same np.random stream -> same index ----
        versions = sorted(snapshots.keys())
        if update == 1:
            idx = 0
        elif opponent_mode == 'random':
            idx = int(np.random.choice(update, 1)[0])
        elif opponent_mode == 'latest':
            idx = update - 1
        elif opponent_mode == 'ours':
            # ratio-divergence-weighted sampling over <= 30 snapshots (alg_ppo.py:228-244).  The reference call passes the
            # action positionally and the sf01-scrambled opponent_obs, which raises at HEAD; this is the intended computation.
            o_obs, o_act = prev['obs'][1], prev['actions'][1]
            base = models[1].act_model.action_probability(o_obs, given_action=o_act)
            sub = np.sort(np.random.choice(len(versions), max_snapshots, replace=False)) if len(versions) > max_snapshots else np.arange(len(versions))
            rd = []
            for i in sub:
                model_util.set_flat(snapshots[versions[i]])
                newp = model_util.act_model.action_probability(o_obs, given_action=o_act)
                rd.append(float((newp / base - 1.0).abs().mean().item()))
            rd = np.array(rd); rd = rd / rd.sum()
            idx = int(versions[sub[np.random.choice(len(rd), 1, p=rd)[0]]])
        else:
            raise NotImplementedError("opponent_mode=%r ('fix' needs the policy_zoo MLP: next row N1)" % opponent_mode)
        version_gap.append(update - 1 - idx)
        models[1].set_flat(snapshots[idx])
        if comm is not None:
            comm.broadcast(models[1].params, 0)       # opponent-snapshot broadcast over NCCL (98 KB)

        # ---- rollout (device resident) ----
        R = runner.run(update, as_numpy=False)
        prev = R
        t_roll = time.perf_counter()
        epinfobuf.extend(R['epinfos'])
        clip_ratio = rho_bar
        fix = lambda x: torch.clamp(torch.nan_to_num(x, nan=clip_ratio), 0.0, clip_ratio)      # alg_ppo.py:258-279
        off_policy_ratio, total_ratio = fix(R['off_policy_ratio']), fix(R['ratio'])
        usable = (R['neglogpacs'][1] < neglogp_threshold).nonzero().flatten()

        # ---- training set (alg_ppo.py:325-344) ----
        take0 = lambda k: R[k][0]
        if use_opponent_data is None or (vgap is not None and version_gap[-1] > vgap):
            data = {k: take0(k).contiguous() for k in ('obs', 'returns', 'actions', 'values', 'neglogpacs')}
            weights = None
        else:
            assert world == 1, "opponent-data reuse is single-GPU for now"
            data = {k: torch.cat([R[k][0], R[k][1][usable]], 0).contiguous() for k in ('obs', 'returns', 'actions', 'values', 'neglogpacs')}
            ones = torch.ones(nbatch, dtype=torch.float32, device=device)
            extra = {'direct': torch.ones(len(usable), dtype=torch.float32, device=device), 'off_policy': off_policy_ratio[usable],
                     'both': total_ratio[usable]}[use_opponent_data]
            weights = torch.cat([ones, extra]).contiguous()
        n_local = data['returns'].shape[0]
        update_sample_num = n_local * world if weights is None else n_local

        # ---- epochs x minibatches (alg_ppo.py:355-398) ----
        perms = EpochPermutations(update_sample_num, noptepochs)     # np.random.shuffle(inds) per epoch, replayed bit-exactly one epoch ahead (dist.py)
        stat_acc = []
        early_stop = False
        for epoch in range(noptepochs):
            inds = next(perms)
            starts = list(range(0, update_sample_num, nbatch_train))
            if world == 1:
                dev_inds = torch.as_tensor(inds.astype(np.int32), device=device)
                parts = [(dev_inds[s:s + nbatch_train], min(nbatch_train, update_sample_num - s)) for s in starts]
            else:
                loc = [split_minibatch(inds[s:s + nbatch_train], lo, hi) for s in starts]
                cat = torch.as_tensor(np.concatenate(loc) if len(loc) else np.zeros(0, np.int32), device=device)
                offs = np.cumsum([0] + [len(x) for x in loc])
                parts = [(cat[offs[i]:offs[i + 1]], min(nbatch_train, update_sample_num - s)) for i, s in enumerate(starts)]
            for mb_idx, gn in parts:
                stats, _ = model.train_indexed(lrnow, cliprangenow, data['obs'], data['returns'], data['actions'], data['values'],
                                               data['neglogpacs'], weights, mb_idx, global_n=gn)
                stat_acc.append(stats)
                if kl_threshold is not None and float(stats[3].item()) > kl_threshold * 1.5:
                    early_stop = True
                    break
            if early_stop:
                break
        perms.close()
        lossvals = torch.stack(stat_acc).double().mean(0).cpu().numpy().tolist()          # np.mean(mblossvals, axis=0)
        tnow = time.perf_counter()
        history.append(dict(update=update, opponent=idx, rollout_s=t_roll - tstart, update_s=tnow - t_roll, losses=lossvals))
        if update_fn is not None:
            update_fn(update)
        if update % log_interval == 0 or update == 1:
            ev = explained_variance(data['values'].cpu().numpy(), data['returns'].cpu().numpy())
            logger.logkv("misc/serial_timesteps", update * nsteps)
            logger.logkv("misc/nupdates", update)
            logger.logkv("misc/total_timesteps", update * nbatch)
            logger.logkv("misc/explained_variance", float(ev))
            logger.logkv('eprewmean', safemean([e['r'] for e in epinfobuf]))
            logger.logkv('epdenserewmean', safemean([e['r'] for e in epinfobuf]))     # sic: the reference logs 'r' here too (alg_ppo.py:449)
            logger.logkv('eplenmean', safemean([e['l'] for e in epinfobuf]))
            logger.logkv('misc/time_elapsed', tnow - tfirststart)
            logger.logkv('misc/rollout_s', t_roll - tstart)
            logger.logkv('misc/update_s', tnow - t_roll)
            logger.logkv('misc/opponent_version', idx)
            for lossval, lossname in zip(lossvals, model.loss_names):
                logger.logkv('loss/' + lossname, float(lossval))
            logger.dumpkvs()
        if save_interval and (update % save_interval == 0 or update == 1):
            save(update)
    model.history = history
    return model
