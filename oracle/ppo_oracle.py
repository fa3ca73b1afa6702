"""ORACLE (test infrastructure): CPU restatement of the learner-side arithmetic of the reference.

  mlp / heads / DiagGaussianPd   baselines/baselines/common/models.py:93-101, a2c/utils.py:58-63,
                                 policies.py:50,70-71, common/distributions.py:227-251
  ortho_init + RNG stream        baselines/baselines/a2c/utils.py:20-35, common/misc_util.py:48-62, alg_ppo.py:89,117-133
  minibatch schedule             alg_ppo.py:355-398
  PPOModel.train                 model.py:51-139,179-213 (TF1 graph) -- restated with torch float64 autograd;
                                 tf.train.AdamOptimizer(lr, epsilon=1e-5) and tf.clip_by_global_norm semantics [TF, not in tree]
  Runner post-processing         runner.py:127-200,251-267 -- restated in numpy; PINNED against the reference's own runner.py,
                                 imported under shims by tests/golden/make_runner_golden.py (tests/golden/runner_*.npz)

TensorFlow is absent from the build container, so the PPO part is UNPINNED against TF output; the V-trace / curriculum /
flatten part is pinned against the reference source itself.
"""
import numpy as np

H = 64


def param_shapes(D, A):
    return [(D, H), (H,), (H, H), (H,), (D, H), (H,), (H, H), (H,), (H, A), (A,), (1, A), (H, 1), (1,)]


def unflatten(flat, D, A):
    out, o = [], 0
    for shp in param_shapes(D, A):
        n = int(np.prod(shp))
        out.append(np.asarray(flat[o:o + n]).reshape(shp))
        o += n
    assert o == len(flat)
    return out


def flatten(arrs):
    return np.concatenate([np.asarray(a, dtype=np.float32).ravel() for a in arrs])


def ortho_init(shape, scale, rng=np.random):
    """a2c/utils.py:20-35 (2-D case)."""
    a = rng.normal(0.0, 1.0, shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == shape else v
    return (scale * q[:shape[0], :shape[1]]).astype(np.float32)


def init_params(D, A, rng=np.random):
    """Variable creation order of one PPOModel (policies.py:156-190, 14-71): six ortho draws."""
    w = {}
    w['pi0'] = ortho_init((D, H), np.sqrt(2), rng)
    w['pi1'] = ortho_init((H, H), np.sqrt(2), rng)
    w['vf0'] = ortho_init((D, H), np.sqrt(2), rng)
    w['vf1'] = ortho_init((H, H), np.sqrt(2), rng)
    w['pih'] = ortho_init((H, A), 0.01, rng)
    w['vfh'] = ortho_init((H, 1), 1.0, rng)
    z = lambda *s: np.zeros(s, dtype=np.float32)
    return flatten([w['pi0'], z(H), w['pi1'], z(H), w['vf0'], z(H), w['vf1'], z(H), w['pih'], z(A), z(1, A), w['vfh'], z(1)])


def forward(flat, obs, D, A):
    """mean [n, A], value [n], logstd [A] in float64."""
    p = [np.asarray(x, dtype=np.float64) for x in unflatten(flat, D, A)]
    x = np.asarray(obs, dtype=np.float64)
    h = np.maximum(x @ p[0] + p[1], 0); h = np.maximum(h @ p[2] + p[3], 0)
    mean = h @ p[8] + p[9]
    g = np.maximum(x @ p[4] + p[5], 0); g = np.maximum(g @ p[6] + p[7], 0)
    value = (g @ p[11] + p[12])[:, 0]
    return mean, value, p[10][0]


def neglogp(act, mean, logstd):
    act = np.asarray(act, dtype=np.float64)
    return 0.5 * np.sum(np.square((act - mean) / np.exp(logstd)), axis=-1) + 0.5 * np.log(2.0 * np.pi) * act.shape[-1] + np.sum(logstd)


def sf01(arr):
    s = arr.shape
    return arr.swapaxes(1, 2).reshape(s[0], s[1] * s[2], *s[3:])


def sf0(arr):
    return arr.swapaxes(0, 1).ravel()


def runner_postprocess(update, anneal_bound, gamma, lam, rho_bar, c_bar, shaping, main, values, neglogpacs, opp_neglogpacs,
                       dones, last_values, last_dones):
    """runner.py:127-200.  Inputs [2][T][E] (dones = mb_dones: done flags BEFORE each step; last_dones [E][2]);
    shaping/main are the info values (float64).  Returns rewards f32, returns f32, (off_policy, off_env, ratio)."""
    T, E = values.shape[1:]
    alpha = 0
    if update <= anneal_bound:
        alpha = np.linspace(1, 0, anneal_bound)[update - 1]
    rewards = (alpha * shaping + (1 - alpha) * main).astype(np.float32)
    values = values.astype(np.float32); neglogpacs = neglogpacs.astype(np.float32); opp = opp_neglogpacs.astype(np.float32)
    returns = np.zeros_like(rewards)
    off_policy = np.exp(opp[1] - neglogpacs[1])
    off_env = np.exp(neglogpacs[0] - opp[0])
    ratio = off_policy * off_env
    for agt in range(2):
        if agt == 0:
            rho_clip = np.ones_like(ratio); c_clip = np.ones_like(ratio)
        else:
            rho_clip = np.clip(ratio, None, rho_bar); c_clip = np.clip(ratio, None, c_bar)
        c_clip = c_clip * lam
        acc = np.zeros(E)
        for t in reversed(range(T)):
            if t == T - 1:
                nnt = 1.0 - last_dones[:, agt]; nv = last_values[agt]
            else:
                nnt = 1.0 - dones[agt, t + 1]; nv = values[agt, t + 1]
            delta = rho_clip[t] * (rewards[agt, t] + gamma * nv * nnt - values[agt, t])
            acc = delta + gamma * nnt * c_clip[t] * acc
            returns[agt, t] = values[agt, t] + acc
    return rewards, returns, (off_policy, off_env, ratio)


def ppo_train_step(flat, m, v, t, D, A, obs, returns, actions, values, old_nlp, weights, lr, cliprange, ent_coef=0.0, vf_coef=0.5,
                   max_grad_norm=0.5, b1=0.9, b2=0.999, eps=1e-5):
    """One PPOModel.train call in float64 (model.py:179-213).  Returns new (flat, m, v), stats[5], log_ratio, grads, gnorm."""
    import torch
    f64 = torch.float64
    theta = torch.tensor(np.asarray(flat, dtype=np.float64), dtype=f64, requires_grad=True)
    ps, o = [], 0
    for shp in param_shapes(D, A):
        n = int(np.prod(shp)); ps.append(theta[o:o + n].reshape(shp)); o += n
    x = torch.tensor(np.asarray(obs, dtype=np.float64)); a = torch.tensor(np.asarray(actions, dtype=np.float64))
    R = torch.tensor(np.asarray(returns, dtype=np.float64)); V = np.asarray(values, dtype=np.float64)
    old = torch.tensor(np.asarray(old_nlp, dtype=np.float64)); w = torch.tensor(np.asarray(weights, dtype=np.float64))
    advs = np.asarray(returns, dtype=np.float64) - V
    advs = (advs - advs.mean()) / (advs.std() + 1e-8)
    ADV = torch.tensor(advs)
    h = torch.relu(x @ ps[0] + ps[1]); h = torch.relu(h @ ps[2] + ps[3]); mean = h @ ps[8] + ps[9]
    g = torch.relu(x @ ps[4] + ps[5]); g = torch.relu(g @ ps[6] + ps[7]); vpred = (g @ ps[11] + ps[12])[:, 0]
    logstd = ps[10][0]
    nlp = 0.5 * torch.sum(((a - mean) / torch.exp(logstd)) ** 2, dim=-1) + 0.5 * np.log(2.0 * np.pi) * A + torch.sum(logstd)
    entropy = torch.sum(logstd + 0.5 * np.log(2.0 * np.pi * np.e))
    vf_loss = 0.5 * torch.mean((vpred - R) ** 2)
    ratio = torch.exp(old - nlp)
    ratio = torch.where(torch.isnan(ratio), torch.full_like(ratio, 2.0), ratio)
    pg1 = -ADV * ratio; pg2 = -ADV * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    pg_loss = torch.mean(w * torch.maximum(pg1, pg2))
    approxkl = torch.mean(nlp - old)
    clipfrac = torch.mean((torch.abs(ratio - 1.0) > cliprange).double())
    loss = pg_loss - entropy * ent_coef + vf_loss * vf_coef
    loss.backward()
    grad = theta.grad.numpy().copy()
    gnorm = float(np.sqrt(np.sum(grad ** 2)))
    if max_grad_norm is not None:
        grad = grad * (max_grad_norm / max(gnorm, max_grad_norm))
    m = b1 * np.asarray(m, dtype=np.float64) + (1 - b1) * grad
    v = b2 * np.asarray(v, dtype=np.float64) + (1 - b2) * grad * grad
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    new = np.asarray(flat, dtype=np.float64) - lr_t * m / (np.sqrt(v) + eps)
    stats = [float(x.detach()) for x in (pg_loss, vf_loss, entropy, approxkl, clipfrac)]
    return new, m, v, stats, (old - nlp).detach().numpy(), theta.grad.numpy().copy(), gnorm


def minibatch_schedule(seed, D, A, nupdates, N, nminibatches, noptepochs, opponent_mode='random'):
    """The legacy-RandomState stream of alg_ppo.learn: seed -> 3 models x 6 ortho draws -> per update (>= 2) the opponent draw ->
    per epoch one shuffle.  Returns (opponent_idx per update, list of minibatch index arrays per update)."""
    np.random.seed(seed)
    for _ in range(3):
        init_params(D, A)
    opp, sched = [], []
    nbt = N // nminibatches
    for update in range(1, nupdates + 1):
        if update == 1:
            opp.append(0)
        elif opponent_mode == 'random':
            opp.append(int(np.random.choice(update, 1)[0]))
        else:
            opp.append(update - 1)
        inds = np.arange(N)
        mbs = []
        for _ in range(noptepochs):
            np.random.shuffle(inds)
            for start in range(0, N, nbt):
                mbs.append(inds[start:start + nbt].copy())
        sched.append(mbs)
    return opp, sched


def zoo_mlp_act(flat, obs, D, A):
    """Deterministic action and vpred of policy_zoo's MLPPolicy (robosumo/robosumo/policy_zoo/policy.py:39-80, utils.py:8-30,70-82)
    from its flat parameter vector, float64."""
    flat = np.asarray(flat, dtype=np.float64).ravel()
    o = [0]

    def take(*shp):
        n = int(np.prod(shp)) if shp else 1
        v = flat[o[0]:o[0] + n]
        o[0] += n
        return v.reshape(shp) if shp else float(v[0])
    rs, rss, rc = take(), take(), take()
    os_, oss, oc = take(D), take(D), take()
    vf = [take(D, H), take(H), take(H, H), take(H), take(H, 1), take(1)]
    pi = [take(D, H), take(H), take(H, H), take(H), take(H, A), take(A)]
    take(1, A)
    assert o[0] == flat.size
    mean = os_ / oc
    std = np.sqrt(np.maximum(oss / oc - mean ** 2, 1e-2))
    obz = np.clip((np.asarray(obs, dtype=np.float64) - mean) / std, -5.0, 5.0)
    h = np.tanh(np.tanh(obz @ vf[0] + vf[1]) @ vf[2] + vf[3])
    vz = (h @ vf[4] + vf[5])[:, 0]
    rmean = rs / rc
    rstd = np.sqrt(max(rss / rc - rmean ** 2, 1e-2))
    g = np.tanh(np.tanh(obz @ pi[0] + pi[1]) @ pi[2] + pi[3])
    return g @ pi[4] + pi[5], vz * rstd + rmean


def select_training_set(batch, use_opponent_data, vgap, last_version_gap, neglogp_threshold, rho_bar, nbatch):
    """NumPy restatement of the reference's training-set selection, alg_ppo.py:258-344 (TEST INFRASTRUCTURE).

    `batch`: dict of the Runner outputs as [2, N, ...] arrays (obs, returns, actions, values, neglogpacs) plus the flat
    off_policy_ratio / ratio [N].  Returns (data dict, weights, usable_index)."""
    clip_ratio = rho_bar
    opr = np.array(batch['off_policy_ratio'], dtype=np.float32); tr = np.array(batch['ratio'], dtype=np.float32)
    opr[np.isnan(opr)] = clip_ratio; opr = np.clip(opr, 0., clip_ratio)                      # alg_ppo.py:262-266
    tr[np.isnan(tr)] = clip_ratio; tr = np.clip(tr, 0., clip_ratio)                          # alg_ppo.py:276-280
    usable = np.where(batch['neglogpacs'][1] < neglogp_threshold)[0]                          # alg_ppo.py:286
    keys = ('obs', 'returns', 'actions', 'values', 'neglogpacs')
    if use_opponent_data is None or (vgap is not None and last_version_gap > vgap):           # alg_ppo.py:325-330
        data = {k: batch[k][0] for k in keys}
    else:
        data = {k: np.concatenate([batch[k][0], batch[k][1, usable]], axis=0) for k in keys}  # alg_ppo.py:334-335
    if use_opponent_data is None:
        weights = np.ones(nbatch, dtype=np.float32)                                           # alg_ppo.py:337-338
    elif use_opponent_data == 'direct':
        weights = np.ones(data['obs'].shape[0], dtype=np.float32)
    elif use_opponent_data == 'off_policy':
        weights = np.concatenate([np.ones(nbatch, dtype=np.float32), opr[usable]])
    else:
        weights = np.concatenate([np.ones(nbatch, dtype=np.float32), tr[usable]])
    # the epoch loop indexes weights[mbinds] with mbinds < len(data) (alg_ppo.py:378-381), so only this prefix is ever read
    return data, weights[:data['obs'].shape[0]], usable


def ratio_divergence(flat_candidates, flat_current, obs, act, D, A):
    """alg_ppo.py:228-241: RD_i = mean | action_probability_i / action_probability_current - 1 | where action_probability is the
    NEGLOGP of the given actions (policies.py:107-108), normalised to sum 1."""
    m0, _, ls0 = forward(flat_current, obs, D, A)
    base = neglogp(act, m0, ls0)
    rd = []
    for f in flat_candidates:
        m, _, ls = forward(f, obs, D, A)
        rd.append(np.abs(neglogp(act, m, ls) / base - 1.).mean())
    rd = np.array(rd)
    return rd / rd.sum()
