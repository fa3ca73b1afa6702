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
    # observability: monitor.csv, fig/ratio_<update>.npz(+png), ratio_summary.pkl -- rank 0 only
    monitor = EpisodeMonitor(log_dir if rank == 0 else None, getattr(env, 'env_id', 'RoboSumo'))
    ratio_log = dict(off_policy_ratio_mean=[], off_policy_ratio_clip_frac=[], off_env_ratio_mean=[], off_env_ratio_clip_frac=[],
                     total_ratio_mean=[], total_ratio_clip_frac=[], ppo_clip_frac=[], approxkl=[])
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

        # the update's permutations depend on the generator stream only: when the sample count is known up front they are all
        # computed on a helper thread WHILE the rollout runs (nothing else draws from np.random until they are consumed)
        perms = EpochPermutations(nbatch, noptepochs, ahead=noptepochs) if use_opponent_data is None else None
        # ---- rollout (device resident) ----
        R = runner.run(update, as_numpy=False)
        prev = R
        t_roll = time.perf_counter()
        epinfobuf.extend(R['epinfos'])
        monitor.write(R['epinfos'])
        clip_ratio = rho_bar
        raw3 = [torch.nan_to_num(R[k], nan=clip_ratio) for k in ('off_policy_ratio', 'off_env_ratio', 'ratio')]      # alg_ppo.py:258-279
        rstat = torch.stack([torch.stack([x.double().mean(), (x > clip_ratio).double().mean()]) for x in raw3]).cpu().numpy()
        for (mk, ck), (mean_, frac_) in zip((('off_policy_ratio_mean', 'off_policy_ratio_clip_frac'), ('off_env_ratio_mean', 'off_env_ratio_clip_frac'),
                                             ('total_ratio_mean', 'total_ratio_clip_frac')), rstat):
            ratio_log[mk].append(float(mean_)); ratio_log[ck].append(float(frac_))
        if rank == 0 and log_dir:
            ratio_figure(log_dir, update, idx, raw3, R['neglogpacs'], clip_ratio, neglogp_threshold)
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
        if perms is None:
            perms = EpochPermutations(update_sample_num, noptepochs)     # np.random.shuffle(inds) per epoch, replayed bit-exactly one epoch ahead (dist.py)
        assert perms._inds.shape[0] == update_sample_num
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
        ratio_log['approxkl'].append(float(lossvals[3])); ratio_log['ppo_clip_frac'].append(float(lossvals[4]))
        if rank == 0 and log_dir and opponent_mode == 'random' and (update % 100 == 0 or update == 1):
            import pickle
            with open(osp.join(log_dir, 'ratio_summary.pkl'), 'wb') as f:      # same nine lists, same order (incl. the repeated
                pickle.dump([version_gap, ratio_log['off_policy_ratio_mean'], ratio_log['off_env_ratio_clip_frac'],      # clip-frac entry) as alg_ppo.py:468-472
                             ratio_log['off_env_ratio_mean'], ratio_log['off_env_ratio_clip_frac'], ratio_log['total_ratio_mean'],
                             ratio_log['total_ratio_clip_frac'][-1] if ratio_log['total_ratio_clip_frac'] else 0.0, ratio_log['ppo_clip_frac'], ratio_log['approxkl']], f)
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
    monitor.close()
    model.history = history
    model.ratio_log = ratio_log
    return model
