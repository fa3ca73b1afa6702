"""Kernel source (rs_core.h) compiled as a host emulation vs the independent double-precision oracle.
CPU-only check of the math the CUDA kernels execute; the GPU parity tests are in test_gpu_*.py."""
import ctypes

import numpy as np
import pytest

from robosumo_selfplay_b200.morphology import PairSpec
from tests.emu.build import build
from tests.helpers import reset_like_state, settled_states


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def emu():
    return ctypes.CDLL(build())


def emu_forward(L, ps, q, v, ctrl):
    nv = ps.nv
    q = np.array(q, np.float32); v = np.array(v, np.float32); ctrl = np.array(ctrl, np.float32)
    qacc = np.zeros(nv, np.float32); M = np.zeros((nv, nv), np.float32); tau = np.zeros(nv, np.float32)
    ncon, nit = ctypes.c_int(), ctypes.c_int()
    st = L.emu_forward(ps.pack(), ctypes.c_float(0.01), 8, P(q), P(v), P(ctrl), P(qacc), P(M), P(tau),
                       ctypes.byref(ncon), ctypes.byref(nit), None, None)
    return qacc, M, ncon.value, nit.value, st


@pytest.mark.parametrize('name', ['ant', 'bug', 'spider'])
def test_free_flight_inertia_bias_limits(name, emu, oracle_models):
    om = oracle_models(name); ps = PairSpec(name, name)
    rng = np.random.RandomState(0)
    for _ in range(4):
        q, v = reset_like_state(om, rng, spread=2.0, z=3.0)
        q[3:7] += rng.uniform(-.5, .5, 4); q[7:7 + 2 * ps.agents[0].L] += rng.uniform(-.6, .6, 2 * ps.agents[0].L)
        v = rng.randn(om.nv) * 2.0
        ctrl = rng.uniform(-1.5, 1.5, om.nu)
        r = om.forward(q, v, ctrl, full=True)
        qacc, M, ncon, nit, st = emu_forward(emu, ps, q, v, ctrl)
        assert ncon == r['ncon']            # bug/spider legs may self-collide here: intra-agent pairs are part of the path
        np.testing.assert_allclose(M, r['M'], atol=2e-6)
        assert abs(qacc - r['qacc']).max() <= 1e-5 * abs(r['qacc']).max()


def test_contact_forward_ant(emu, oracle_models):
    om = oracle_models('ant'); ps = PairSpec('ant', 'ant')
    rng = np.random.RandomState(5)
    qs, vs = settled_states(om, rng, 6, steps=40, action_scale=0.15, spread=0.6)
    seen = 0
    for q, v in zip(qs, vs):
        ctrl = rng.uniform(-1, 1, om.nu)
        r = om.forward(q, v, ctrl, full=True)
        qacc, M, ncon, nit, st = emu_forward(emu, ps, q, v, ctrl)
        assert ncon == r['ncon'] and st == 0
        seen += ncon
        assert abs(qacc - r['qacc']).max() <= 2e-4 * max(1.0, abs(r['qacc']).max())
    assert seen > 10


def test_trajectory_ant_20_steps(emu, oracle_models):
    om = oracle_models('ant'); ps = PairSpec('ant', 'ant')
    rng = np.random.RandomState(7)
    q, v = reset_like_state(om, rng)
    qf, vf, wf = q.astype(np.float32), v.astype(np.float32), np.zeros(om.nv, np.float32)
    w = np.zeros(om.nv)
    for t in range(20):
        ctrl = rng.randn(om.nu)
        om.step(q, v, ctrl, 5, w)
        st = emu.emu_step(ps.pack(), ctypes.c_float(0.01), 8, P(qf), P(vf), P(wf), P(ctrl.astype(np.float32)), 5)
        assert st == 0
        assert abs(q - qf).max() < 2e-4 and abs(v - vf).max() < 5e-3, t


@pytest.mark.parametrize('name', ['bug', 'spider'])
def test_trajectory_bug_spider_with_self_collisions(name, emu, oracle_models):
    """Six- and eight-legged bodies: intra-agent leg-leg contacts occur; 15 env steps stay within fp32 tolerance."""
    om = oracle_models(name); ps = PairSpec(name, name)
    rng = np.random.RandomState(1)
    q, v = reset_like_state(om, rng)
    qf, vf, wf = q.astype(np.float32), v.astype(np.float32), np.zeros(om.nv, np.float32)
    w = np.zeros(om.nv)
    seen = 0
    for t in range(15):
        ctrl = rng.randn(om.nu)
        seen = max(seen, om.step(q, v, ctrl, 5, w))
        st = emu.emu_step(ps.pack(), ctypes.c_float(0.01), 8, P(qf), P(vf), P(wf), P(ctrl.astype(np.float32)), 5)
        assert st == 0
        assert abs(q - qf).max() < 2e-4 and abs(v - vf).max() < 5e-3, t
    assert seen >= 2


@pytest.mark.parametrize('na,nb', [('ant', 'bug'), ('spider', 'ant')])
def test_trajectory_mixed_morphology_pairs(na, nb, emu):
    """Mixed pairs of robosumo/__init__.py (different dof counts per agent -> unequal diagonal blocks in the solver)."""
    from oracle.physics import OracleModel, load_model_json
    om = OracleModel(load_model_json('%s_%s' % (na, nb))); ps = PairSpec(na, nb)
    rng = np.random.RandomState(2)
    q = om.qpos0.copy(); phi = rng.uniform(0, 2 * np.pi)
    for a, o in ((0, 0), (1, ps.agents[0].nq)):
        q[o] = 1.15 * np.cos(phi + a * np.pi); q[o + 1] = 1.15 * np.sin(phi + a * np.pi); q[o + 2] = 1.25
    q += rng.uniform(-.1, .1, om.nq); v = 0.1 * rng.randn(om.nv); om.normalize_qpos(q); w = np.zeros(om.nv)
    qf, vf, wf = q.astype(np.float32), v.astype(np.float32), np.zeros(om.nv, np.float32)
    for t in range(12):
        ctrl = rng.randn(om.nu)
        om.step(q, v, ctrl, 5, w)
        st = emu.emu_step(ps.pack(), ctypes.c_float(0.01), 8, P(qf), P(vf), P(wf), P(ctrl.astype(np.float32)), 5)
        assert st == 0 and abs(q - qf).max() < 2e-4 and abs(v - vf).max() < 5e-3, t


def test_line_search_regression_cases_and_no_iteration_cap(emu, oracle_models):
    """States (tests/golden/newton_linesearch_cases.npz, found by a soak of the host emulation) where a row with a tiny J.d puts a
    line-search breakpoint at alpha ~ 1e9: the one-pass exact line search must evaluate phi' near the SMALL end of the bracketing
    linear piece (a midpoint there loses the root to cancellation and the solver then stalls at the iteration cap).  The step must
    converge (status 0) and agree with the float64 oracle; a short soak must never raise RS_STATUS_NEWTON_MAXIT."""
    import os
    om = oracle_models('ant'); ps = PairSpec('ant', 'ant')
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'newton_linesearch_cases.npz'))
    for q0, v0, w0, ctrl in zip(z['q'], z['v'], z['w'], z['ctrl']):
        qf, vf, wf = q0.copy(), v0.copy(), w0.copy()
        st = emu.emu_step(ps.pack(), ctypes.c_float(0.01), 16, P(qf), P(vf), P(wf), P(ctrl), 5)
        assert st == 0
        q, v, w = q0.astype(np.float64), v0.astype(np.float64), w0.astype(np.float64)
        om.normalize_qpos(q)
        om.step(q, v, ctrl.astype(np.float64), 5, w)
        assert abs(q - qf).max() < 2e-4 and abs(v - vf).max() < 5e-3
    rng = np.random.RandomState(11)
    for env in range(25):
        q, v = reset_like_state(om, rng)
        qf, vf, wf = q.astype(np.float32), v.astype(np.float32), np.zeros(om.nv, np.float32)
        for t in range(100):
            st = emu.emu_step(ps.pack(), ctypes.c_float(0.01), 16, P(qf), P(vf), P(wf), P(rng.randn(om.nu).astype(np.float32)), 5)
            assert st & 5 == 0, (env, t, st)


def test_leg_straddling_the_tatami_edge(emu, oracle_models):
    """Arena geometry at the box edge, |x| in [2.1, 2.3] (tatami.xml:21, utils.py:64-68): agent 0 hangs over the edge in random
    poses, so that legs cross it with neither endpoint touching.  Kernel source and oracle must find the same contacts (endpoint
    spheres + the closest edge against the capsule interior) and the same accelerations; at least some poses must produce a
    contact ON the edge whose normal is neither vertical nor horizontal (i.e. an edge, not a face, contact)."""
    om = oracle_models('ant'); ps = PairSpec('ant', 'ant')
    rng = np.random.RandomState(0)
    n_edge = n_checked = 0
    for trial in range(300):
        q = om.qpos0.copy()
        q[0:3] = [rng.uniform(1.7, 2.2), rng.uniform(-1, 1), rng.uniform(0.6, 1.0)]
        quat = rng.randn(4) * [1, .3, .3, .5]; q[3:7] = quat / np.linalg.norm(quat)
        q[7:15] += rng.uniform(-.5, .5, 8)
        q[15:18] = [-1.0, 0, 0.9]                 # agent 1 well clear of the mat
        v = 0.1 * rng.randn(om.nv); ctrl = rng.uniform(-1, 1, om.nu)
        om.normalize_qpos(q)
        r = om.forward(q, v, ctrl, full=True)
        edge = [c for c in r['contacts'] if abs(abs(c[1]) - 2.3) < 0.02 and abs(c[3] - 0.5) < 0.05 and 0.05 < abs(c[6]) < 0.98]
        if not edge:
            continue
        n_edge += 1
        qacc, M, ncon, nit, st = emu_forward(emu, ps, q, v, ctrl)
        assert ncon == r['ncon'] and st == 0, (trial, ncon, r['ncon'])
        assert abs(qacc - r['qacc']).max() <= 3e-4 * max(1.0, abs(r['qacc']).max()), trial
        n_checked += 1
    assert n_edge >= 10 and n_checked == n_edge


def test_inter_agent_contacts_low_rank_and_dense_paths(emu, oracle_models):
    """Agents pressed against each other: one or two inter-agent contacts take the low-rank (Woodbury) correction of the arrowhead
    solve, three or more the dense elimination.  Both must reproduce the float64 oracle's accelerations, and the sample must
    exercise both paths."""
    om = oracle_models('ant'); ps = PairSpec('ant', 'ant')
    rng = np.random.RandomState(17)
    seen = {1: 0, 2: 0, 3: 0}
    for spread in (0.3, 0.4, 0.5, 0.6):
        qs, vs = settled_states(om, rng, 6, steps=25, action_scale=0.4, spread=spread)
        for q, v in zip(qs, vs):
            ctrl = rng.uniform(-1, 1, om.nu)
            r = om.forward(q, v, ctrl, full=True)
            qf = np.array(q, np.float32); vf = np.array(v, np.float32); cf = np.array(ctrl, np.float32)
            qacc = np.zeros(ps.nv, np.float32); con = np.zeros((64, 8), np.float32)
            ncon, nit = ctypes.c_int(), ctypes.c_int()
            st = emu.emu_forward(ps.pack(), ctypes.c_float(0.01), 16, P(qf), P(vf), P(cf), P(qacc), None, None,
                                 ctypes.byref(ncon), ctypes.byref(nit), P(con), None)
            assert st == 0 and ncon.value == r['ncon']
            m = int((con[:ncon.value, 7] >= 0).sum())            # bA * 100 + bB with bA = -1 for the world
            if m:
                seen[min(m, 3)] += 1
            assert abs(qacc - r['qacc']).max() <= 2e-4 * max(1.0, abs(r['qacc']).max()), (spread, m)
    assert seen[1] > 0 and seen[2] > 0 and seen[3] > 0, seen
