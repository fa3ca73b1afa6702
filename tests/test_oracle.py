"""The CPU oracle itself: physical invariants and closed-form checks (the reference holds no golden
vectors for this path -- SURVEY 8c -- so the oracle is pinned by physics, not by MuJoCo output)."""
import numpy as np
import pytest

from tests.helpers import reset_like_state


@pytest.mark.parametrize('name', ['ant', 'bug', 'spider'])
def test_mass_matrix_symmetric_positive_definite(name, oracle_models):
    om = oracle_models(name)
    rng = np.random.RandomState(0)
    q, v = reset_like_state(om, rng, z=3.0)
    r = om.forward(q, v, np.zeros(om.nu), full=True)
    assert abs(r['M'] - r['M'].T).max() < 1e-12
    assert np.linalg.eigvalsh(r['M']).min() > 0.05


def test_free_flight_com_accelerates_at_g(oracle_models):
    om = oracle_models('ant')
    M = om.M
    mass = np.array(M['body_mass']); gb = np.array(M['geom_bodyid'])
    rng = np.random.RandomState(1)
    q, v = reset_like_state(om, rng, z=5.0)
    v = rng.randn(om.nv) * 1.0
    ctrl = rng.uniform(-1, 1, om.nu)

    def com(qq):
        gx = om.forward(qq, np.zeros(om.nv), np.zeros(om.nu), full=True)['geom_xpos']
        idx = np.arange(6, 19)
        return (mass[gb[idx], None] * gx[idx]).sum(0) / mass[gb[idx]].sum()
    cs = [com(q)]
    for _ in range(4):
        om.step(q, v, ctrl, 1)
        cs.append(com(q))
    cs = np.array(cs)
    acc = (cs[2:] - 2 * cs[1:-1] + cs[:-2]) / 1e-4
    np.testing.assert_allclose(acc, np.tile([0, 0, -9.81], (3, 1)), atol=2e-2)   # RK4 second difference at dt=0.01


def test_ant_settles_on_tatami_with_soft_contacts(oracle_models):
    om = oracle_models('ant')
    q = om.qpos0.copy(); v = np.zeros(om.nv)
    q[0], q[15] = 1.0, -1.0
    for _ in range(60):
        om.step(q, v, np.zeros(om.nu), 5)
    assert abs(v).max() < 1e-6
    # ankles rest on their joint limit (+-30 deg), torso above the 0.5 m tatami top
    assert abs(abs(q[8]) - np.radians(30)) < 2e-3 and 0.8 < q[2] < 0.95
    r = om.forward(q, v, np.zeros(om.nu), full=True)
    assert r['ncon'] == 8 and (r['contacts'][:, 0] < 0.01).all() and (r['contacts'][:, 0] > 0).all()


def test_ant_never_self_collides(oracle_models):
    """Justifies skipping intra-agent pairs for Ant in the CUDA path."""
    om = oracle_models('ant')
    rng = np.random.RandomState(3)
    nga = (om.ngeom - 6) // 2
    for ep in range(2):
        q, v = reset_like_state(om, rng)
        w = np.zeros(om.nv)
        for t in range(40):
            ctrl = rng.randn(om.nu)
            om.step(q, v, ctrl, 5, w)
            for c in om.forward(q, v, ctrl, full=True)['contacts']:
                g1, g2 = int(c[7]) // 1000, int(c[7]) % 1000
                assert not (g1 >= 6 and g2 >= 6 and (g1 - 6) // nga == (g2 - 6) // nga)


def test_env_oracle_rewards_and_timeout(oracle_models):
    from oracle.env_oracle import OracleVecEnv
    from oracle.physics import load_model_json
    env = OracleVecEnv(load_model_json('ant_ant'), 1, seed=0, timestep_limit=3)
    obs = env.reset()
    assert obs.shape == (1, 2, 121) and obs[0, 0, -1] == -1.0
    for t in range(4):
        obs, rew, done, infos = env.step(np.zeros((1, 2, 8)))
        if t < 3:
            assert not done.any() and abs(obs[0, 0, -1] - (-1 + 2 * (t + 1) / 500)) < 1e-12
    assert done.all() and infos[0][0]['main_reward'] == -1000 and infos[0][0]['timeout'] and infos[0][1]['timeout']
    assert obs[0, 0, -1] == -1.0 and infos[0][0]['episode']['l'] == 4     # auto-reset replaced the terminal obs
