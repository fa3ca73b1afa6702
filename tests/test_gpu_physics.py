"""GPU parity: CUDA physics (through the C ABI / B200SumoVecEnv) against the CPU oracle on the same
seeded inputs.  Tolerances are fp32-vs-fp64 and stated per test."""
import numpy as np
import pytest

from tests.helpers import reset_like_state, settled_states

pytestmark = pytest.mark.gpu


def make_env(E, name='Ant', **kw):
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    return B200SumoVecEnv('RoboSumo-%s-vs-%s-v0' % (name, name), num_envs=E, seed=1, **kw)


def test_forward_free_flight_and_limits(oracle_models):
    """qacc with no contacts (M, bias, actuation, joint-limit rows): rel 1e-5."""
    om = oracle_models('ant')
    rng = np.random.RandomState(0)
    E = 16
    qs, vs, cs = [], [], []
    for _ in range(E):
        q, v = reset_like_state(om, rng, spread=2.0, z=3.0)
        q[3:7] += rng.uniform(-.5, .5, 4); q[7:15] += rng.uniform(-.6, .6, 8)
        qs.append(q); vs.append(rng.randn(om.nv) * 2.0); cs.append(rng.uniform(-1.5, 1.5, om.nu))
    env = make_env(E)
    env.set_state(np.array(qs), np.array(vs))
    qacc, ncon, nit = env.forward_debug(np.array(cs))
    qacc = qacc.cpu().numpy()
    for e in range(E):
        ref = om.forward(qs[e], vs[e], cs[e])
        assert abs(qacc[e] - ref).max() <= 1e-5 * abs(ref).max(), e
    assert int(ncon.sum()) == 0


def test_forward_with_contacts(oracle_models):
    """qacc with ground and inter-agent contacts: same contact count, |err| <= 2e-4 * max(1, |qacc|)."""
    om = oracle_models('ant')
    rng = np.random.RandomState(5)
    E = 12
    qs, vs = settled_states(om, rng, E, steps=40, action_scale=0.15, spread=0.6)
    cs = rng.uniform(-1, 1, (E, om.nu))
    env = make_env(E)
    env.set_state(qs, vs)
    qacc, ncon, nit = env.forward_debug(cs)
    qacc, ncon = qacc.cpu().numpy(), ncon.cpu().numpy()
    total = 0
    for e in range(E):
        r = om.forward(qs[e], vs[e], cs[e], full=True)
        assert ncon[e] == r['ncon'], e
        total += r['ncon']
        assert abs(qacc[e] - r['qacc']).max() <= 2e-4 * max(1.0, abs(r['qacc']).max()), e
    assert total > 20


def test_trajectory_parity_20_env_steps(oracle_models):
    """20 env steps (= 400 forward evaluations) from reset-like states with an N(0,1) action tape:
    |dqpos| < 2e-4, |dqvel| < 5e-3 at every step (fp32 kernel vs fp64 oracle)."""
    om = oracle_models('ant')
    rng = np.random.RandomState(7)
    E = 8
    st = [reset_like_state(om, rng) for _ in range(E)]
    q = np.array([s[0] for s in st]); v = np.array([s[1] for s in st]); w = np.zeros((E, om.nv))
    env = make_env(E, device_api=True, auto_reset=False)
    env.set_state(q, v)
    import torch
    for t in range(20):
        a = rng.randn(E, 2, 8)
        env.step(torch.as_tensor(a, dtype=torch.float32, device='cuda'))
        gq, gv, _, status = env.get_state()
        for e in range(E):
            om.step(q[e], v[e], a[e].ravel(), 5, w[e])
        assert abs(gq.cpu().numpy() - q).max() < 2e-4, t
        assert abs(gv.cpu().numpy() - v).max() < 5e-3, t
        assert int((status & 7).max()) == 0


def test_full_size_invariants_4096():
    """Size-independent properties at BASELINE config 2 (E=4096): unit quaternions, finite state,
    joint angles near their ranges, torsos above the floor, no status flags, obs layout."""
    import torch
    E = 4096
    env = make_env(E, device_api=True)
    obs = env.reset()
    torch.manual_seed(0)
    for t in range(30):
        a = torch.randn(E, 2, 8, device='cuda')
        obs, rew, done, (info, epi) = env.step(a)
    q, v, step, status = env.get_state()
    assert torch.isfinite(q).all() and torch.isfinite(v).all() and torch.isfinite(obs).all()
    for o in (3, 18):
        n = q[:, o:o + 4].norm(dim=1)
        assert (n - 1).abs().max() < 1e-5
    assert int((status & 1).sum()) == 0
    assert q[:, 2].min() > 0.0 and q[:, 17].min() > 0.0
    hinge = torch.cat([q[:, 7:15], q[:, 22:30]], 1)
    assert hinge.abs().max() < 1.6
    # obs = [qpos_a | qvel_a | 78 zeros | opp qpos[:7] | 6 zeros | t]
    assert torch.equal(obs[:, 0, :15], q[:, :15]) and torch.equal(obs[:, 1, :15], q[:, 15:30])
    assert torch.equal(obs[:, 0, 15:29], v[:, :14]) and (obs[:, :, 29:107] == 0).all()
    assert torch.equal(obs[:, 0, 107:114], q[:, 15:22]) and torch.equal(obs[:, 1, 107:114], q[:, 0:7])
    want = (-1.0 + 2.0 * step.double() / 500.0).float()
    assert torch.equal(obs[:, 0, -1], want) and torch.equal(obs[:, 1, -1], want)


@pytest.mark.parametrize('name,nq,nv,na', [('Bug', 38, 36, 12), ('Spider', 46, 44, 16)])
def test_trajectory_parity_bug_spider(name, nq, nv, na, oracle_models):
    """BASELINE config 3 morphologies (self-collisions included): 12 env steps, same tolerances as Ant."""
    import torch
    om = oracle_models(name.lower())
    rng = np.random.RandomState(3)
    E = 6
    st = [reset_like_state(om, rng) for _ in range(E)]
    q = np.array([s[0] for s in st]); v = np.array([s[1] for s in st]); w = np.zeros((E, om.nv))
    env = make_env(E, name, device_api=True, auto_reset=False)
    assert env.nq == nq and env.nv == nv and env.obs_dim == {'Bug': 165, 'Spider': 209}[name]
    env.set_state(q, v)
    for t in range(12):
        a = rng.randn(E, 2, na)
        env.step(torch.as_tensor(a, dtype=torch.float32, device='cuda'))
        gq, gv, _, status = env.get_state()
        for e in range(E):
            om.step(q[e], v[e], a[e].ravel(), 5, w[e])
        assert abs(gq.cpu().numpy() - q).max() < 2e-4, t
        assert abs(gv.cpu().numpy() - v).max() < 5e-3, t
        assert int((status & 7).max()) == 0
