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
from .dist import EpochPermutations, EpochSchedule


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


class EpisodeMonitor:
    """`monitor.csv` (baselines/bench/monitor.py:95-121): a JSON header line, then one `r,l,t` row per finished episode.  The
    reference wraps every env process in its own Monitor and writes one file per env; the vectorised env writes ONE file for all
    E pairs, rows in (step, env) order of completion."""

    def __init__(self, log_dir, env_id):
        import json
        self.f = None
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
            self.f = open(osp.join(log_dir, 'monitor.csv'), 'wt')
            self.f.write('# %s \n' % json.dumps({'t_start': time.time(), 'env_id': env_id}))
            self.f.write('r,l,t\n')
            self.f.flush()

    def write(self, epinfos):
        if self.f is not None and epinfos:
            self.f.write(''.join('%s,%s,%s\n' % (e['r'], e['l'], e['t']) for e in epinfos))
            self.f.flush()

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


def ratio_figure(log_dir, update, idx, raw, nlp, clip_ratio, neglogp_threshold):
    """The per-update IS-ratio diagnostics of alg_ppo.py:292-318: five 100-bin histograms (log off-policy / off-env / total ratio
    clipped below at -100, both agents' neglogp clipped to +-threshold) computed ON THE DEVICE and saved as fig/ratio_<update>.npz;
    the same five panels are also rendered to fig/ratio_<update>.png when matplotlib is importable (it is not in this image)."""
    import torch
    out = {}
    for name, x in (('off_policy', raw[0]), ('off_env', raw[1]), ('total', raw[2])):
        lx = torch.clamp(torch.log(x.double()), min=-100.0)
        lo, hi = float(lx.min()), float(lx.max())
        hi = hi if hi > lo else lo + 1.0
        out[name + '_log_hist'] = torch.histc(lx, bins=100, min=lo, max=hi).cpu().numpy()
        out[name + '_log_range'] = np.array([lo, hi])
        out[name + '_clip_frac'] = np.array(float((x > clip_ratio).double().mean()))
    for a in range(2):
        v = torch.clamp(nlp[a].double().flatten(), -neglogp_threshold, neglogp_threshold)
        lo, hi = float(v.min()), float(v.max())
        hi = hi if hi > lo else lo + 1.0
        out['neglogp%d_hist' % a] = torch.histc(v, bins=100, min=lo, max=hi).cpu().numpy()
        out['neglogp%d_range' % a] = np.array([lo, hi])
    out['opponent_version'] = np.array(idx)
    fig_dir = osp.join(log_dir, 'fig')
    os.makedirs(fig_dir, exist_ok=True)
    np.savez(osp.join(fig_dir, 'ratio_%d.npz' % update), **out)
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except Exception:
        return out
    plt.figure(figsize=(16, 9))
    panels = [((2, 3, 1), 'off_policy', 'off-policy ratio (log scale): %.2f%% clipped'), ((2, 3, 2), 'off_env', 'off-env ratio (log scale): %.2f%% clipped'),
              ((2, 3, 3), 'total', 'off-policy-env ratio (log scale): %.2f%% clipped')]
    for pos, name, title in panels:
        plt.subplot(*pos)
        edges = np.linspace(out[name + '_log_range'][0], out[name + '_log_range'][1], 101)
        plt.bar(edges[:-1], out[name + '_log_hist'], width=edges[1] - edges[0], align='edge')
        plt.title(title % (100.0 * float(out[name + '_clip_frac'])))
    for a, pos in ((0, (2, 2, 3)), (1, (2, 2, 4))):
        plt.subplot(*pos)
        edges = np.linspace(out['neglogp%d_range' % a][0], out['neglogp%d_range' % a][1], 101)
        plt.bar(edges[:-1], out['neglogp%d_hist' % a], width=edges[1] - edges[0], align='edge')
        plt.title('-log pi_1(a^%d|o^%d)' % (a + 1, a + 1))
    plt.suptitle('opponent version: %d' % idx)
    plt.savefig(osp.join(fig_dir, 'ratio_%d.png' % update))
    plt.close()
    return out


class SnapshotRing:
    """The opponent pool ON THE DEVICE (row N2): every saved version of the learner as one row of an [n_snap, P] float32 tensor in
    HBM (98 KB each; 2 000 snapshots = 196 MB).  Every rank holds the same rows (parameters are replicated), so choosing an
    opponent moves an INDEX between ranks, never a parameter vector, and loading it is a device-to-device row copy.  The
    reference's pool is the checkpoint directory (alg_ppo.py:122-123,217-244,459-464: joblib.load + 13 assign ops per update)."""

    def __init__(self, P, device, capacity=64):
        import torch
        self.torch, self.P, self.device = torch, P, device
        self.stride = (P + 31) & ~31                    # rows are read in place by kernels that issue 16-byte loads: keep them 128-byte aligned
        self.buf = torch.empty((capacity, self.stride), dtype=torch.float32, device=device)
        self.row = {}                                   # version -> row

    def put(self, version, params):
        if version not in self.row:
            if len(self.row) == self.buf.shape[0]:
                nb = self.torch.empty((2 * self.buf.shape[0], self.stride), dtype=self.torch.float32, device=self.device)
                nb[:self.buf.shape[0]].copy_(self.buf)
                self.buf = nb
            self.row[version] = len(self.row)
        self.buf[self.row[version], :self.P].copy_(params)

    def get(self, version):
        return self.buf[self.row[version], :self.P]

    def versions(self):
        return sorted(self.row)

    def __len__(self):
        return len(self.row)


def select_training_set(R, use_opponent_data, vgap, last_version_gap, neglogp_threshold, rho_bar, nbatch_local):
    """alg_ppo.py:258-344 on device tensors: the IS ratios with NaN -> rho_bar and clipping to [0, rho_bar], the usable opponent
    samples (neglogp of the opponent's action under the learner below the threshold), and the training set = agent-0 samples
    followed by the usable agent-1 samples with their importance weights.  Returns (data dict, weights or None, n_usable)."""
    import torch
    fix = lambda x: torch.clamp(torch.nan_to_num(x, nan=rho_bar), 0.0, rho_bar)
    keys = ('obs', 'returns', 'actions', 'values', 'neglogpacs')
    usable = (R['neglogpacs'][1] < neglogp_threshold).nonzero().flatten()
    if use_opponent_data is None or (vgap is not None and last_version_gap > vgap):
        return {k: R[k][0].contiguous() for k in keys}, None, int(usable.numel())
    if use_opponent_data not in ('direct', 'off_policy', 'both'):
        raise ValueError("use_opponent_data must be None, 'direct', 'off_policy' or 'both'")
    data = {k: torch.cat([R[k][0], R[k][1][usable]], 0).contiguous() for k in keys}
    dev = data['returns'].device
    ones = torch.ones(nbatch_local, dtype=torch.float32, device=dev)
    if use_opponent_data == 'direct':
        extra = torch.ones(int(usable.numel()), dtype=torch.float32, device=dev)
    elif use_opponent_data == 'off_policy':
        extra = fix(R['off_policy_ratio'])[usable]
    else:
        extra = fix(R['ratio'])[usable]
    return data, torch.cat([ones, extra]).contiguous(), int(usable.numel())


def ratio_divergence_weights(policy, ring, versions, base_neglogp, o_obs, o_act):
    """The sampling weights of opponent_mode='ours' (alg_ppo.py:228-244): mean |p_new / p_base - 1| over the opponent's last batch
    for every candidate snapshot, where the reference's `action_probability` returns a NEGLOGP (so the ratio is one of
    neglogps, as in the reference).  The candidate parameters are read in place from the device ring."""
    keep = policy.params
    rd = []
    try:
        for v in versions:
            policy.params = ring.get(v)
            newp = policy.action_probability(o_obs, given_action=o_act)
            rd.append((newp / base_neglogp - 1.0).abs().double().mean())
    finally:
        policy.params = keep
    import torch
    return torch.stack(rd)


_NETWORK_KWARGS = {'num_hidden': 64, 'value_network': 'copy', 'num_layers': 2}


def learn(*, network='mlp', env, total_timesteps, eval_env=None, opponent_mode='ours', use_opponent_data=None, seed=None, nsteps=2048,
          ent_coef=0.0, lr=3e-4, vf_coef=0.5, max_grad_norm=0.5, gamma=0.99, lam=0.95, rho_bar=1., c_bar=1., log_interval=10,
          nminibatches=4, noptepochs=4, cliprange=0.2, save_interval=1, load_path=None, nagent=2, anneal_bound=500, vgap=None,
          kl_threshold=None, neglogp_threshold=10000., fix_opponent_path=None, log_dir=None, comm=None, update_fn=None, max_snapshots=30,
          strict_status=True, precision='tf32', **network_kwargs):
    """Signature and defaults of the reference's learn (alg_ppo.py:25-29) with these differences, all deliberate:
    `nagent` defaults to 2 (the reference's signature says 1 but run.py:181-187 always passes 2, and the loop only works for 2);
    `model_fn` / `init_fn` / `mpi_rank_weight` are not accepted (one model class; `comm` is this package's dist.Comm);
    `log_dir`, `update_fn`, `max_snapshots`, `strict_status`, `precision` ('tf32' tensor cores | 'fp32') are additions; `eval_env` is accepted and, as in the reference
    (alg_ppo.py:346-351 only touches it when set), unused by the default path; network kwargs other than the 2x64 ReLU 'copy'
    architecture of defaults.py:8-26 raise instead of being ignored."""
    import torch
    if network != 'mlp' or nagent != 2:
        raise ValueError("this path implements network='mlp' with nagent=2 (run.py --algo ppo on RoboSumo)")
    for k, v in network_kwargs.items():
        if k == 'activation':
            if getattr(v, '__name__', str(v)) not in ('relu', 'relu6_not', 'ReLU'):
                raise ValueError("only activation=relu is implemented (defaults.py:25)")
        elif k not in _NETWORK_KWARGS:
            raise TypeError("learn() got an unexpected network keyword %r" % k)
        elif v != _NETWORK_KWARGS[k]:
            raise ValueError("%s=%r is not implemented (this path: %s=%r)" % (k, v, k, _NETWORK_KWARGS[k]))
    if opponent_mode not in ('fix', 'random', 'latest', 'ours'):
        raise ValueError("opponent_mode must be 'fix', 'random', 'latest' or 'ours'")
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

    make_model = lambda scope, trainable: PPOModel(ob_dim=D, ac_dim=A, ent_coef=ent_coef, vf_coef=vf_coef, max_grad_norm=max_grad_norm,
                                                   trainable=trainable, model_scope=scope, device=device, comm=comm if trainable else None,
                                                   precision=precision)
    model = make_model('model_0', True)               # creation order = np.random draw order (alg_ppo.py:117-133)
    models = [model, make_model('model_1', False)]
    model_util = make_model('model_util', False)
    if comm is not None:
        comm.broadcast(model.params, 0)
    checkdir = osp.join(log_dir, 'checkpoints') if log_dir else None
    ring = SnapshotRing(model.P, device)              # version -> parameters, device resident; the checkpoint files mirror it (rank 0)
    def save(version):
        ring.put(version, model.params)
        if checkdir and rank == 0:
            model.save(osp.join(checkdir, '%.5i' % version))
    save(0)
    if load_path is not None:
        for m in models:
            m.load(load_path)
    if opponent_mode == 'fix':
        # alg_ppo.py:194-206: the opponent is a pretrained policy_zoo MLP for the whole run (the reference passes the MLPPolicy object to
        # build_policy, which cannot call it -- SURVEY 8f N1; this is the computation that code intends)
        from .policy_zoo import ZooMLPPolicy, ZooOpponentModel
        if fix_opponent_path is None:
            raise ValueError("opponent_mode='fix' needs fix_opponent_path (a policy_zoo agent-params-v*.npy)")
        models[1] = ZooOpponentModel(ZooMLPPolicy.load(fix_opponent_path, D - 1, A, device), seed=(seed or 0) + 17 * rank)

    runner = Runner(env=env, models=models, nsteps=nsteps, nagent=nagent, gamma=gamma, lam=lam, rho_bar=rho_bar, c_bar=c_bar,
                    anneal_bound=anneal_bound, seed=(seed or 0) * 7919 + rank)
    epinfobuf = deque(maxlen=100)
    tfirststart = time.perf_counter()
    version_gap, history = [], []
    # observability: monitor.csv, fig/ratio_<update>.npz(+png), ratio_summary.pkl -- rank 0 only
    monitor = EpisodeMonitor(log_dir if rank == 0 else None, getattr(env, 'env_id', 'RoboSumo'))
    ratio_log = dict(off_policy_ratio_mean=[], off_policy_ratio_clip_frac=[], off_env_ratio_mean=[], off_env_ratio_clip_frac=[],
                     total_ratio_mean=[], total_ratio_clip_frac=[], ppo_clip_frac=[], approxkl=[], useful_ratio=[])
    prev = None
    sched = None

    nupdates = total_timesteps // nbatch
    for update in range(1, nupdates + 1):
        assert nbatch % nminibatches == 0
        tstart = time.perf_counter()
        frac = 1.0 - (update - 1.0) / nupdates
        lrnow, cliprangenow = lr(frac), cliprange(frac)

        # ---- opponent (alg_ppo.py:192-247); every rank draws from the same np.random stream ----
        versions = ring.versions()
        if update == 1 or opponent_mode == 'fix':
            idx = 0
        elif opponent_mode == 'random':
            idx = int(np.random.choice(update, 1)[0])
        elif opponent_mode == 'latest':
            idx = update - 1
        else:
            # 'ours': ratio-divergence-weighted sampling over <= 30 snapshots (alg_ppo.py:228-244).  The reference call passes the
            # action positionally and the sf01-scrambled opponent_obs, which raises at HEAD; this is the intended computation.
            o_obs, o_act = prev['obs'][1], prev['actions'][1]
            base = models[1].act_model.action_probability(o_obs, given_action=o_act)
            sub = np.sort(np.random.choice(len(versions), max_snapshots, replace=False)) if len(versions) > max_snapshots else np.arange(len(versions))
            rd = ratio_divergence_weights(model_util.act_model, ring, [versions[i] for i in sub], base, o_obs, o_act)
            if comm is not None:                  # the batch is rank-local: average the divergences so that every rank holds the same weights
                comm.all_reduce_sum(rd); rd = rd / world
            rd = rd.cpu().numpy(); rd = rd / rd.sum()
            idx = int(versions[sub[np.random.choice(len(rd), 1, p=rd)[0]]])
            if comm is not None:                  # belt and braces: the index, not a 98 KB vector, is what crosses ranks
                idx = comm.broadcast_int(idx, 0, device)
        version_gap.append(update - 1 - idx)
        if opponent_mode != 'fix':
            models[1].params.copy_(ring.get(idx))        # device-to-device row copy; every rank holds the same ring

        # the update's permutations depend on the generator stream only: when the sample count is known up front they are all
        # computed on a helper thread WHILE the rollout runs (nothing else draws from np.random until they are consumed)
        split = (nbatch_train, rank * nbatch_local, (rank + 1) * nbatch_local) if world > 1 else None      # data-parallel: the helper thread also cuts out this rank's part
        perms = EpochPermutations(nbatch, noptepochs, ahead=noptepochs, dtype=np.int32, split=split) if use_opponent_data is None else None
        # ---- rollout (device resident) ----
        R = runner.run(update, as_numpy=False)
        prev = R
        if hasattr(env, 'check_status'):
            env.check_status(strict=strict_status)          # NaN / contact-buffer overflow anywhere in the rollout raises here (builder.py:351-369)
        t_roll = time.perf_counter()
        epinfobuf.extend(R['epinfos'])
        monitor.write(R['epinfos'])
        clip_ratio = rho_bar
        raw3 = [torch.nan_to_num(R[k], nan=clip_ratio) for k in ('off_policy_ratio', 'off_env_ratio', 'ratio')]      # alg_ppo.py:258-279
        rstat = torch.stack([torch.stack([x.double().mean(), (x > clip_ratio).double().mean()]) for x in raw3]).cpu().numpy()
        for (mkey, ckey), (mean_, frac_) in zip((('off_policy_ratio_mean', 'off_policy_ratio_clip_frac'), ('off_env_ratio_mean', 'off_env_ratio_clip_frac'),
                                                 ('total_ratio_mean', 'total_ratio_clip_frac')), rstat):
            ratio_log[mkey].append(float(mean_)); ratio_log[ckey].append(float(frac_))
        if rank == 0 and log_dir:
            ratio_figure(log_dir, update, idx, raw3, R['neglogpacs'], clip_ratio, neglogp_threshold)

        # ---- training set (alg_ppo.py:325-344) ----
        data, weights, n_usable = select_training_set(R, use_opponent_data, vgap, version_gap[-1], neglogp_threshold, rho_bar, nbatch_local)
        ratio_log['useful_ratio'].append(n_usable / float(nbatch_local))
        n_local = int(data['returns'].shape[0])
        # data-parallel: rank r owns the contiguous global sample range [lo, hi) (rank-major; with the default agent-0 data this is the
        # env-major order of one process over all envs).  With opponent data the ranges are ragged: one small all-gather of the sizes.
        sizes = comm.all_gather_int(n_local, device) if comm is not None else [n_local]
        lo = int(sum(sizes[:rank])); hi = lo + n_local
        update_sample_num = int(sum(sizes))

        # ---- epochs x minibatches (alg_ppo.py:355-398) ----
        if perms is None:
            perms = EpochPermutations(update_sample_num, noptepochs, dtype=np.int32,     # np.random.shuffle(inds) per epoch, replayed bit-exactly one epoch ahead (dist.py)
                                      split=(nbatch_train, lo, hi) if world > 1 else None)
        assert perms._inds.shape[0] == update_sample_num
        if sched is None or sched.n_total != update_sample_num or sched.lo != lo or sched.hi != hi:
            sched = EpochSchedule(device, update_sample_num, nbatch_train, lo, hi, comm)
        stat_acc = []
        early_stop = False
        for epoch in range(noptepochs):
            parts = sched.load(next(perms), data['returns'], data['values'])
            for mb_idx, n_mb, gn, sums in parts:
                stats, _ = model.train_indexed(lrnow, cliprangenow, data['obs'], data['returns'], data['actions'], data['values'],
                                               data['neglogpacs'], weights, mb_idx, global_n=gn, adv_sums=sums)
                stat_acc.append(stats)
                if kl_threshold is not None and float(stats[3].item()) > kl_threshold * 1.5:
                    early_stop = True
                    break
            if early_stop:
                break
        perms.close()
        lossvals = torch.stack(stat_acc).double().mean(0).cpu().numpy().tolist()          # np.mean(mblossvals, axis=0)
        if hasattr(model, 'check_peer'):
            model.check_peer()                              # a rank that missed a peer all-reduce (time-out) means diverged replicas: raise
        ratio_log['approxkl'].append(float(lossvals[3])); ratio_log['ppo_clip_frac'].append(float(lossvals[4]))
        if rank == 0 and log_dir and opponent_mode == 'random' and (update % 100 == 0 or update == 1):
            import pickle
            with open(osp.join(log_dir, 'ratio_summary.pkl'), 'wb') as f:      # same nine lists, same order (incl. the repeated
                pickle.dump([version_gap, ratio_log['off_policy_ratio_mean'], ratio_log['off_env_ratio_clip_frac'],      # clip-frac entry) as alg_ppo.py:468-472
                             ratio_log['off_env_ratio_mean'], ratio_log['off_env_ratio_clip_frac'], ratio_log['total_ratio_mean'],
                             ratio_log['total_ratio_clip_frac'][-1] if ratio_log['total_ratio_clip_frac'] else 0.0, ratio_log['ppo_clip_frac'], ratio_log['approxkl']], f)
        tnow = time.perf_counter()
        history.append(dict(update=update, opponent=idx, rollout_s=t_roll - tstart, update_s=tnow - t_roll, losses=lossvals,
                            samples=update_sample_num, usable=n_usable))
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
    monitor.close()
    model.history = history
    model.ratio_log = ratio_log
    model.snapshots = ring
    return model
