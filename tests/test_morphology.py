"""Product-side model compiler (morphology.py) against the golden models compiled from the
reference's MJCF files (tests/golden/model_*.json, made by tests/golden/make_model_golden.py)."""
import ctypes
import json
import os

import numpy as np
import pytest

from robosumo_selfplay_b200.morphology import AgentSpec, PairSpec, rs_agent_model

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ANCHORS = {'ant': (0.85085, 2.3683, 121), 'bug': (0.65450, 2.3097, 165), 'spider': (2.55254, 2.8291, 209)}


@pytest.mark.parametrize('name', ['ant', 'bug', 'spider'])
def test_masses_axes_ranges_match_reference_xml(name):
    M = json.load(open(os.path.join(GOLD, 'model_%s_%s.json' % (name, name))))
    a = AgentSpec(name)
    nb = a.nbody
    mine = [a.m_torso] + [x for l in range(a.L) for x in (a.aux[l][0], a.hip[l][0], a.ank[l][0])]
    np.testing.assert_allclose(mine, M['body_mass'][1:1 + nb], rtol=1e-12)
    torso, total, obs = ANCHORS[name]
    assert abs(a.m_torso - torso) < 1e-4 and abs(a.total_mass - total) < 1e-4 and a.obs_dim == obs
    jax = np.array(M['jnt_axis'][1:1 + 2 * a.L]).reshape(a.L, 2, 3)
    np.testing.assert_allclose(jax[:, 0], a.ax_hip, atol=1e-12)
    np.testing.assert_allclose(jax[:, 1], a.ax_ank, atol=1e-12)
    rng = np.array(M['jnt_range'][1:1 + 2 * a.L]).reshape(a.L, 2, 2)
    np.testing.assert_allclose(rng[:, 0, 0], a.lo_hip); np.testing.assert_allclose(rng[:, 0, 1], a.hi_hip)
    np.testing.assert_allclose(rng[:, 1, 0], a.lo_ank); np.testing.assert_allclose(rng[:, 1, 1], a.hi_ank)
    # body positions: hip body at the aux capsule end, ankle body at the hip capsule end
    bp = np.array(M['body_pos'][1:1 + nb])
    for l in range(a.L):
        np.testing.assert_allclose(bp[2 + 3 * l], a.r_hip[l], atol=1e-12)
        np.testing.assert_allclose(bp[3 + 3 * l], a.r_ank[l], atol=1e-12)
    # inertias of capsules
    bi = np.array(M['body_inertia'][1:1 + nb])
    for l in range(a.L):
        np.testing.assert_allclose(bi[3 + 3 * l], [a.ank[l][1], a.ank[l][1], a.ank[l][2]], rtol=1e-12)
    assert M['nq'] == 2 * a.nq and M['nv'] == 2 * a.nv and M['nu'] == 2 * a.nu
    # actuators act on the joints in dof order (hip_1, ankle_1, hip_2, ...)
    assert M['act_jntid'][:a.nu] == list(range(1, 1 + a.nu))


@pytest.mark.parametrize('name', ['ant', 'bug', 'spider'])
def test_invweights_match_oracle(name, oracle_models):
    om = oracle_models(name)
    a = AgentSpec(name)
    np.testing.assert_allclose(a.body_invweight0, om.body_invweight0[1:1 + a.nbody], rtol=1e-10)
    np.testing.assert_allclose(a.dof_invweight0, om.dof_invweight0[:a.nv], rtol=1e-10)


def test_pack_layout_and_qpos0():
    p = PairSpec('ant', 'ant')
    arr = p.pack()
    assert ctypes.sizeof(rs_agent_model) == 1064 and arr[0].L == 4 and arr[1].nq == 15
    M = json.load(open(os.path.join(GOLD, 'model_ant_ant.json')))
    np.testing.assert_allclose(p.qpos0(), M['qpos0'], atol=1e-15)
