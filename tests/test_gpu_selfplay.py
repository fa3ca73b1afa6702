"""GPU tests of the self-play loop around the kernels: device-side minibatch schedule, opponent pool and sampling modes
(SURVEY 8f N2), opponent-data reuse (N3), the fixed policy_zoo opponent in training (N1), status latch and seeding surface."""
import ctypes
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
D, A = 121, 8


def _p(x):
    return ctypes.c_void_p(x.data_ptr()) if x is not None else None


def test_epoch_split_and_moments_match_host():
    """rs_epoch_split (stable device compaction of a global permutation into per-rank local minibatch indices) against the host
    routine dist.split_minibatch, ragged last minibatch included: bit-exact; rs_adv_moments_multi against float64 NumPy."""
    import torch
    from robosumo_selfplay_b200 import _lib
    from robosumo_selfplay_b200.dist import split_minibatch
    L = _lib.lib()
    rng = np.random.RandomState(0)
    for N, nbt, world in ((8192, 2048, 4), (10000, 3000, 3), (5000, 5000, 2), (777, 100, 8)):
        perm = rng.permutation(N).astype(np.int32)
        pd = torch.as_tensor(perm, device='cuda')
        nmb = (N + nbt - 1) // nbt
        bounds = np.linspace(0, N, world + 1).astype(int)
        bounds[1:-1] += rng.randint(-3, 4, world - 1)                       # ragged ranges (opponent-data reuse makes them unequal)
        ret = torch.as_tensor(rng.randn(N).astype(np.float32), device='cuda'); val = torch.as_tensor(rng.randn(N).astype(np.float32), device='cuda')
        total = np.zeros((nmb, 2))
        for r in range(world):
            lo, hi = int(bounds[r]), int(bounds[r + 1])
            idx = torch.full((nmb * nbt,), -7, dtype=torch.int32, device='cuda'); cnt = torch.zeros(nmb, dtype=torch.int32, device='cuda')
            _lib.check(L.rs_epoch_split(_p(pd), N, nbt, lo, hi, _p(idx), _p(cnt), None))
            sums = torch.zeros((nmb, 2), dtype=torch.float64, device='cuda')
            _lib.check(L.rs_adv_moments_multi(_p(idx), _p(cnt), nmb, nbt, N, _p(ret[lo:hi].contiguous()), _p(val[lo:hi].contiguous()), _p(sums), None))
            idx, cnt, sums = idx.cpu().numpy(), cnt.cpu().numpy(), sums.cpu().numpy()
            for m in range(nmb):
                want = split_minibatch(perm[m * nbt:(m + 1) * nbt], lo, hi)
                assert cnt[m] == len(want) and np.array_equal(idx[m * nbt:m * nbt + cnt[m]], want)
                adv = (ret[lo:hi].cpu().numpy().astype(np.float64) - val[lo:hi].cpu().numpy().astype(np.float64))[want]
                np.testing.assert_allclose(sums[m], [adv.sum(), (adv * adv).sum()], rtol=1e-12, atol=1e-12)
            total += sums
        adv = ret.cpu().numpy().astype(np.float64) - val.cpu().numpy().astype(np.float64)
        for m in range(nmb):
            a = adv[perm[m * nbt:(m + 1) * nbt]]
            np.testing.assert_allclose(total[m], [a.sum(), (a * a).sum()], rtol=1e-11, atol=1e-11)   # what the all-reduce over ranks yields
        # single-GPU form: idx = the permutation itself, no counts
        sums = torch.zeros((nmb, 2), dtype=torch.float64, device='cuda')
        _lib.check(L.rs_adv_moments_multi(_p(pd), None, nmb, nbt, N, _p(ret), _p(val), _p(sums), None))
        np.testing.assert_allclose(sums.cpu().numpy(), total, rtol=1e-11, atol=1e-11)


def test_status_latch_seed_and_mixed_batch_sizes():
    """(a) a NaN state is latched on the device and raised by check_status() even after auto-reset wiped the env's own status word
    (mujoco-py raises MujocoException from its warning callback, builder.py:351-369); (b) env.seed(s) re-keys the reset streams;
    (c) a small env created after a large one does not lower the large one's shared-memory limit (both keep stepping)."""
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    big = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=4096, seed=1, device_api=True)
    big.reset()
    small = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=16, seed=2, device_api=True, timestep_limit=3)
    small.reset()
    act_b = torch.randn(4096, 2, 8, device='cuda'); act_s = torch.randn(16, 2, 8, device='cuda')
    for _ in range(3):
        big.step(act_b); small.step(act_s)
    torch.cuda.synchronize()
    assert big.check_status() == (0, 0) and small.check_status()[0] & 1 == 0
    # (a) poison one pair, step past its episode end (timestep_limit=3 -> auto-reset clears the per-env word), the latch still raises
    q, v, _, _ = small.get_state()
    q[5, 0] = float('nan')
    small.set_state(q, v)
    for _ in range(5):
        small.step(act_s)
    _, _, _, status = small.get_state()
    with pytest.raises(RuntimeError, match='NaN'):
        small.check_status()
    assert small.check_status(strict=False)[0] & 1 == 0                       # the raising call cleared the latch; the pair was auto-reset since
    # (b) seeding surface
    e1 = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=8, seed=11, device_api=True)
    e2 = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=8, seed=22, device_api=True)
    assert e1.seed(123) == [123 + i for i in range(8)]
    e2.seed(123)
    o1 = e1.reset().clone(); o2 = e2.reset().clone()
    assert torch.equal(o1, o2)
    e2.seed(456)
    assert not torch.equal(e2.reset(), o1)
    for e in (big, small, e1, e2):
        e.close()


def _learn(tmp_path, **kw):
    from robosumo_selfplay_b200 import alg_ppo
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    E, T = kw.pop('E', 64), kw.pop('T', 10)
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=7, device_api=True)
    args = dict(env=env, total_timesteps=kw.pop('updates', 3) * E * T, seed=11, nsteps=T, nminibatches=4, noptepochs=2, lr=1e-3, gamma=0.995, lam=1.0,
                rho_bar=10., c_bar=1., log_interval=1, anneal_bound=1000, log_dir=str(tmp_path), precision='fp32')
    args.update(kw)
    try:
        return alg_ppo.learn(**args)
    finally:
        env.close()


def test_opponent_modes_latest_and_ours(tmp_path):
    """alg_ppo.py:217-244.  'latest' plays version update-1.  'ours' samples among the stored versions with probability
    proportional to the ratio divergence of each snapshot on the opponent's last batch: the weights are recomputed on the CPU
    from the device ring's rows (float64 restatement, rel 2e-3) and the drawn index is replayed from the recorded np.random state."""
    import torch
    from oracle import ppo_oracle as po
    from robosumo_selfplay_b200 import alg_ppo
    m = _learn(tmp_path / 'latest', opponent_mode='latest', updates=4)
    assert [h['opponent'] for h in m.history] == [0, 1, 2, 3]
    assert m.snapshots.versions() == [0, 1, 2, 3, 4]
    np.testing.assert_array_equal(m.snapshots.get(4).cpu().numpy(), m.get_flat())
    rec = []
    orig = alg_ppo.ratio_divergence_weights
    def spy(policy, ring, versions, base, o_obs, o_act):
        rd = orig(policy, ring, versions, base, o_obs, o_act)
        rec.append(dict(versions=list(versions), rd=rd.cpu().numpy().copy(), state=np.random.get_state(), obs=o_obs.cpu().numpy(), act=o_act.cpu().numpy(),
                        rows=[ring.get(v).cpu().numpy().copy() for v in versions]))
        return rd
    alg_ppo.ratio_divergence_weights = spy
    cur = []
    try:
        m = _learn(tmp_path / 'ours', opponent_mode='ours', updates=4, update_fn=lambda u: cur.append(None))
    finally:
        alg_ppo.ratio_divergence_weights = orig
    assert len(rec) == 3                                                    # updates 2..4
    opp = [h['opponent'] for h in m.history]
    assert opp[0] == 0
    for k, r in enumerate(rec):
        update = k + 2
        assert r['versions'] == list(range(update))                        # every stored version is a candidate (<= 30 of them)
        # the opponent that generated the batch is the previous update's choice; its parameters are a row of the ring
        prev_opp = opp[update - 2]
        want = po.ratio_divergence(r['rows'], r['rows'][prev_opp], r['obs'], r['act'], D, A)
        got = r['rd'] / r['rd'].sum()
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-6)
        st = np.random.get_state()
        np.random.set_state(r['state'])
        idx = r['versions'][np.random.choice(len(got), 1, p=got)[0]]
        np.random.set_state(st)
        assert idx == opp[update - 1]


def test_opponent_data_reuse_selection_and_training(tmp_path):
    """Row N3 (alg_ppo.py:258-344,378-381).  (a) select_training_set on a synthetic batch with NaN / huge ratios and unusable
    samples equals the NumPy restatement for None / direct / off_policy / both and for the vgap switch; (b) learn() with
    use_opponent_data='both' trains on agent-0 + usable agent-1 samples: the minibatches are the legacy shuffle over the ragged
    sample count (last minibatch short), the weights are 1 on the learner's samples and the clipped total ratio on the opponent's."""
    import torch
    from oracle import ppo_oracle as po
    from robosumo_selfplay_b200 import alg_ppo
    from robosumo_selfplay_b200.model import PPOModel
    rng = np.random.RandomState(3)
    N = 500
    batch = dict(obs=rng.randn(2, N, 7).astype(np.float32), returns=rng.randn(2, N).astype(np.float32), actions=rng.randn(2, N, 3).astype(np.float32),
                 values=rng.randn(2, N).astype(np.float32), neglogpacs=(8 + 4 * rng.randn(2, N)).astype(np.float32),
                 off_policy_ratio=np.exp(3 * rng.randn(N)).astype(np.float32), ratio=np.exp(4 * rng.randn(N)).astype(np.float32))
    batch['off_policy_ratio'][::17] = np.nan; batch['ratio'][::13] = np.nan; batch['ratio'][5] = np.inf
    R = {k: torch.as_tensor(v, device='cuda') for k, v in batch.items()}
    for mode in (None, 'direct', 'off_policy', 'both'):
        for vgap, gap in ((None, 0), (2, 1), (2, 5)):
            data, w, nus = alg_ppo.select_training_set(R, mode, vgap, gap, 10.0, 10.0, N)
            rd, rw, rus = po.select_training_set(batch, mode, vgap, gap, 10.0, 10.0, N)
            assert nus == len(rus)
            for k in rd:
                assert np.array_equal(data[k].cpu().numpy(), rd[k]), (mode, vgap, gap, k)
            got_w = np.ones(len(rw), np.float32) if w is None else w.cpu().numpy()
            np.testing.assert_array_equal(got_w, rw)
    # (b) end to end: spy on the minibatches
    seen = []
    orig = PPOModel.train_indexed
    def spy(self, lr, clip, obs, ret, act, val, nl, w, idx, **kw):
        seen.append((idx.cpu().numpy().copy(), None if w is None else w.cpu().numpy().copy(), int(ret.shape[0]), kw.get('global_n')))
        return orig(self, lr, clip, obs, ret, act, val, nl, w, idx, **kw)
    PPOModel.train_indexed = spy
    E, T = 64, 10
    try:
        m = _learn(tmp_path, opponent_mode='random', use_opponent_data='both', neglogp_threshold=11.5, updates=2, E=E, T=T)
    finally:
        PPOModel.train_indexed = orig
    nb = E * T
    nbt = nb // 4
    k = 0
    np.random.seed(11)
    for _ in range(3):
        po.init_params(D, A)
    for u, h in enumerate(m.history):
        if u >= 1:
            np.random.choice(u + 1, 1)
        n_tot = h['samples']
        assert n_tot == nb + h['usable'] and 0 < h['usable'] < nb           # the threshold cuts some but not all opponent samples
        inds = np.arange(n_tot)
        for ep in range(2):
            np.random.shuffle(inds)
            for s in range(0, n_tot, nbt):
                idx, w, n_data, gn = seen[k]; k += 1
                assert n_data == n_tot and np.array_equal(idx, inds[s:s + nbt]) and gn == len(inds[s:s + nbt])
                assert w is not None and (w[:nb] == 1).all() and (w[nb:] >= 0).all() and (w[nb:] <= 10.0).all()
        assert n_tot % nbt != 0                                             # the last minibatch of an epoch is ragged
    assert k == len(seen)
    assert np.isfinite(m.get_flat()).all()
    # vgap: version gaps above the bound fall back to the learner's own samples
    seen.clear()
    PPOModel.train_indexed = spy
    try:
        m = _learn(tmp_path / 'vgap', opponent_mode='latest', use_opponent_data='direct', vgap=-1, updates=2)
    finally:
        PPOModel.train_indexed = orig
    assert all(s[2] == nb and s[1] is None for s in seen)


def test_fix_opponent_mode_trains_against_zoo_policy(tmp_path):
    """opponent_mode='fix' (alg_ppo.py:194-206): agent 1 is a pretrained policy_zoo MLP (tanh units, observation filter, 120-d
    observation = the learner's without the timestep feature) for the whole run.  The opponent's recorded neglogp of its own
    actions equals the float64 restatement of policy_zoo/policy.py on the recorded observations."""
    import torch
    from oracle import ppo_oracle as po
    from robosumo_selfplay_b200.policy_zoo import ZooMLPPolicy, ZooOpponentModel
    from robosumo_selfplay_b200.model import PPOModel
    from robosumo_selfplay_b200.runner import Runner
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    path = '/root/reference/robosumo/robosumo/policy_zoo/assets/ant/mlp/agent-params-v3.npy'
    if not os.path.exists(path):                                            # the GPU box has no reference checkout: synthetic parameters
        rng = np.random.RandomState(0)
        n = 3 + 2 * 120 + 1 + 2 * (120 * 64 + 64 + 64 * 64 + 64) + 64 + 1 + 64 * 8 + 8 + 8
        flat = (rng.randn(n) * 0.1).astype(np.float32)
        flat[0:3] = [5.0, 80.0, 10.0]; flat[3:123] = rng.randn(120); flat[123:243] = 20 + rng.rand(120) * 30; flat[243] = 10.0
        path = str(tmp_path / 'zoo.npy'); np.save(path, flat)
    flat = np.load(path)
    E, T = 32, 6
    np.random.seed(1)
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=3, device_api=True)
    zoo = ZooMLPPolicy(flat, 120, 8)
    models = [PPOModel(ob_dim=D, ac_dim=A, precision='fp32'), ZooOpponentModel(zoo, seed=5)]
    r = Runner(env=env, models=models, nsteps=T, gamma=0.995, lam=1.0, rho_bar=10.0, c_bar=1.0, anneal_bound=1000)
    R = r.run(1, as_numpy=False)
    obs1 = R['obs'][1].cpu().numpy(); act1 = R['actions'][1].cpu().numpy()
    mean, _ = po.zoo_mlp_act(flat, obs1[:, :-1], 120, 8)
    logstd = flat[-8:].astype(np.float64)
    want = 0.5 * (((act1 - mean) / np.exp(logstd)) ** 2).sum(1) + 0.5 * np.log(2 * np.pi) * 8 + logstd.sum()
    np.testing.assert_allclose(R['opponent_neglogpacs'][1].cpu().numpy(), want, rtol=2e-4, atol=2e-4)
    # the learner's neglogp of the opponent's actions and its values on the opponent's observations (runner.py:89-90)
    m0, v0, ls0 = po.forward(models[0].get_flat(), obs1, D, A)
    np.testing.assert_allclose(R['neglogpacs'][1].cpu().numpy(), po.neglogp(act1, m0, ls0), rtol=2e-4)
    np.testing.assert_allclose(R['values'][1].cpu().numpy(), v0, atol=2e-5)
    env.close()
    m = _learn(tmp_path / 'fix', opponent_mode='fix', fix_opponent_path=path, updates=2, E=32, T=8)
    assert [h['opponent'] for h in m.history] == [0, 0] and np.isfinite(m.get_flat()).all()
    with pytest.raises(ValueError):
        _learn(tmp_path / 'bad', opponent_mode='fix', updates=1)
    with pytest.raises(TypeError):
        _learn(tmp_path / 'bad2', opponent_mode='latest', updates=1, normalize_observations=True)
