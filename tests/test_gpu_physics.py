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


def test_full_size_oracle_anchor_4096(oracle_models):
    """BASELINE.json's full size (4096 pairs, one persistent-kernel launch per step) anchored on the oracle: 64 pairs sampled from
    the batch -- first, last, the block boundaries of the 28-pairs-per-block layout, and random ones -- must follow the float64
    oracle stepped from the same reset states with the same actions (tolerances of the small-batch trajectory test).

    The dynamics has discrete events (a contact entering the 0.01 margin with a closing velocity changes the force by a finite
    amount), so a pair that passes within float32 rounding of such an event can take the other branch: with seed 1, pair 3541 lands
    |dv| = 0.12 away from the oracle at step 6 -- and so does the ORACLE ITSELF when its initial state is perturbed by 1e-6
    (half of the perturbed runs).  Such a pair is accepted only if the oracle reaches the kernel's state from an initial state
    within 1e-6 of the same start; at most 3 of the 64 pairs may need that."""
    import torch
    om = oracle_models('ant')
    E = 4096
    env = make_env(E, device_api=True, auto_reset=False)
    env.reset()
    q0, v0, _, _ = env.get_state()
    rng = np.random.RandomState(4)
    pick = np.unique(np.concatenate([[0, 1, 27, 28, 29, 55, 56, 4087, 4088, 4094, 4095], rng.choice(E, 53, replace=False)]))
    qi = q0.double().cpu().numpy()[pick]; vi = v0.double().cpu().numpy()[pick]
    T = 6
    acts = 0.7 * rng.randn(T, E, 2, 8)

    def oracle_run(q, v, e):
        q = q.copy(); v = v.copy(); w = np.zeros(om.nv)
        for t in range(T):
            om.step(q, v, acts[t, e].ravel(), 5, w)
        return q, v

    for t in range(T):
        env.step(torch.as_tensor(acts[t], dtype=torch.float32, device='cuda'))
    gq, gv, _, status = env.get_state()
    assert int(status.max()) & 7 == 0
    gq = gq.double().cpu().numpy()[pick]; gv = gv.double().cpu().numpy()[pick]
    branch = 0
    for i, e in enumerate(pick):
        q, v = oracle_run(qi[i], vi[i], e)
        if abs(gq[i] - q).max() < 2e-4 and abs(gv[i] - v).max() < 5e-3:
            continue
        branch += 1
        prng = np.random.RandomState(100 + i)
        best = np.inf
        for k in range(24):
            qp = qi[i] + 1e-6 * prng.randn(om.nq); om.normalize_qpos(qp)
            q, v = oracle_run(qp, vi[i], e)
            if abs(gq[i] - q).max() < 2e-4:
                best = min(best, abs(gv[i] - v).max())
        assert best < 5e-3, (int(e), best)
    assert branch <= 3, branch
    env.close()


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


def _rotz90(q, v):
    """The arena (square tatami, four rails) maps onto itself under a quarter turn about z: rotate both free joints."""
    import torch
    q2, v2 = q.clone(), v.clone()
    h = 0.5 ** 0.5
    for qo, vo in ((0, 0), (15, 14)):
        q2[:, qo] = -q[:, qo + 1]; q2[:, qo + 1] = q[:, qo]                     # (x, y) -> (-y, x)
        w, x, y, z = q[:, qo + 3], q[:, qo + 4], q[:, qo + 5], q[:, qo + 6]       # rz(90 deg) * quat
        q2[:, qo + 3] = h * (w - z); q2[:, qo + 4] = h * (x - y); q2[:, qo + 5] = h * (y + x); q2[:, qo + 6] = h * (z + w)
        v2[:, vo] = -v[:, vo + 1]; v2[:, vo + 1] = v[:, vo]                     # world-frame linear velocity; angular is body-frame
    return q2, v2


def test_full_size_symmetries_and_batch_invariance_4096():
    """Size-independent properties at E=4096 after 40 steps of contact-rich motion, one further step each:
    (a) an env's result does not depend on which other envs share its launch (bit-exact, 4096 vs a 96-env subset);
    (b) exchanging the two (identical) ants exchanges observations and rewards;
    (c) a quarter turn of the whole state about z turns the result by a quarter turn (the arena is a square)."""
    import torch
    E = 4096
    env = make_env(E, device_api=True)
    env.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    for t in range(40):
        env.step(torch.randn(E, 2, 8, device='cuda', generator=g))
    q, v, step, status = env.get_state()
    a = torch.randn(E, 2, 8, device='cuda', generator=g)
    base = make_env(E, device_api=True, auto_reset=False); base.reset(); base.set_state(q, v)
    ob, rw, dn, _ = base.step(a)
    ob, rw, dn = ob.clone(), rw.clone(), dn.clone()
    qn, vn, _, _ = base.get_state()
    # (a) batch composition
    sel = torch.randperm(E, device='cuda', generator=g)[:96]
    sub = make_env(96, device_api=True, auto_reset=False); sub.reset(); sub.set_state(q[sel], v[sel])
    so, sr, sd, _ = sub.step(a[sel])
    assert torch.equal(so[:, :, :-1], ob[sel][:, :, :-1]) and torch.equal(sr, rw[sel]) and torch.equal(sd, dn[sel])
    # (b) agent exchange
    sw = make_env(E, device_api=True, auto_reset=False); sw.reset()
    sw.set_state(torch.cat([q[:, 15:], q[:, :15]], 1), torch.cat([v[:, 14:], v[:, :14]], 1))
    wo, wr, wd, _ = sw.step(a.flip(1))
    # fp32 summation order differs between the mirrored problems and an env in a stiff contact transition amplifies that
    # within the 20 evaluations of the step: the median must agree to rounding, the 99th percentile must stay small
    # (measured: median 1e-7, p99 7e-4 in qpos)
    def bulk_and_tail(err, median_tol, p99_tol):
        srt = err.sort().values
        return float(srt[len(srt) // 2]) < median_tol and float(srt[int(0.99 * (len(srt) - 1))]) < p99_tol
    eo = (wo.flip(1) - ob)[:, :, :-1].abs().amax((1, 2)); er = (wr.flip(1) - rw).abs().amax(1)
    print('exchange: obs err p50 %.1e p99 %.1e max %.1e ; rew err p99 %.1e max %.1e' % (float(eo.median()), float(eo.sort().values[int(.99 * (E - 1))]), float(eo.max()), float(er.sort().values[int(.99 * (E - 1))]), float(er.max())))
    assert bulk_and_tail(eo, 2e-6, 1e-4) and bulk_and_tail(er, 2e-5, 1e-3)
    assert (wd.flip(1) != dn).sum() <= 2
    # (c) quarter turn
    q2, v2 = _rotz90(q, v)
    rt = make_env(E, device_api=True, auto_reset=False); rt.reset(); rt.set_state(q2, v2)
    rt.step(a)
    qr, vr, _, _ = rt.get_state()
    qe, ve = _rotz90(qn, vn)
    sgn = torch.ones_like(qe)
    for qo in (3, 18):                            # q and -q are the same rotation
        sgn[:, qo:qo + 4] = torch.sign((qr[:, qo:qo + 4] * qe[:, qo:qo + 4]).sum(1, keepdim=True))
    eq = (qr - qe * sgn).abs().amax(1); ev = (vr - ve).abs().amax(1)
    print('quarter turn: qpos err p50 %.1e p99 %.1e max %.1e ; qvel err p99 %.1e max %.1e' % (float(eq.median()), float(eq.sort().values[int(.99 * (E - 1))]), float(eq.max()), float(ev.sort().values[int(.99 * (E - 1))]), float(ev.max())))
    assert bulk_and_tail(eq, 2e-6, 5e-3) and bulk_and_tail(ev, 1e-4, 0.5)
    assert (rt.d_done != dn).sum() <= 2
