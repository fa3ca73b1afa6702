"""Closed-form anchors of MuJoCo's documented soft-constraint model (solref (0.02, 1), solimp (0.9, 0.95, 0.001, 0.5, 2), pyramidal
cones, Newton), checked on the CPU oracle with tiny hand-made models -- numbers neither the oracle nor the kernels were tuned to.

Until a MuJoCo golden dump exists (tools/dump_mujoco_golden.py, tests/test_mujoco_golden.py) these are the strongest pins of the
constraint arithmetic the CUDA path is compared with: [M] MuJoCo computation docs, "Solver parameters": for an active row
    aref = -b v - k d(r) r,  b = 2 / (dmax tc),  k = 1 / (dmax^2 tc^2 zeta^2),  tc = max(solref[0], 2 h),
    R = (1 - d) / d * diagApprox,  force = -(J qacc - aref) / R,
pyramidal contacts: 4 rows n +- mu t_k with R_row = 2 mu^2 R(diagApprox = invweight (1 + mu^2)); the impedance d(r) rises from
dmin to dmax over |r| in [0, width] with the midpoint-0.5 / power-2 sigmoid."""
import numpy as np
import pytest

from oracle.physics import OracleModel

DMIN, DMAX, WIDTH, TC = 0.9, 0.95, 0.001, 0.02
K = 1.0 / (DMAX * DMAX * TC * TC)
G = 9.81


def impedance(r):
    x = abs(r) / WIDTH
    y = 1.0 if x >= 1 else (2 * x * x if x <= 0.5 else 1 - 2 * (1 - x) ** 2)
    return DMIN + y * (DMAX - DMIN)


def fixed_point(f, x0=1e-4):
    x = x0
    for _ in range(200):
        x = f(x)
    return x


def make_model(bodies, joints, geoms, acts=(), gravity=(0, 0, -G), timestep=0.01):
    """bodies: dict(parent, pos, mass, inertia[3]); joints: dict(body, type, axis, range|None); geoms: dict(body, type, pos, size, margin,
    friction, contact 0/1).  World body 0 is implicit.  Layout = tests/golden/model_*.json (oracle/mjcf_compile.py)."""
    nb = len(bodies) + 1
    M = dict(body_parent=[0], body_pos=[[0, 0, 0]], body_quat=[[1, 0, 0, 0]], body_ipos=[[0, 0, 0]], body_iquat=[[1, 0, 0, 0]], body_mass=[0.0],
             body_inertia=[[0, 0, 0]], body_jntadr=[-1], body_jntnum=[0], body_dofadr=[-1], body_dofnum=[0], body_weldid=[0])
    jt, qadr, dadr = [], 0, 0
    M.update(jnt_type=[], jnt_qposadr=[], jnt_dofadr=[], jnt_bodyid=[], jnt_pos=[], jnt_axis=[], jnt_range=[], jnt_limited=[], jnt_margin=[],
             dof_bodyid=[], dof_jntid=[], dof_armature=[], dof_damping=[], qpos0=[])
    for bi, b in enumerate(bodies, 1):
        M['body_parent'].append(b.get('parent', 0)); M['body_pos'].append(list(b['pos'])); M['body_quat'].append([1, 0, 0, 0])
        M['body_ipos'].append([0, 0, 0]); M['body_iquat'].append([1, 0, 0, 0]); M['body_mass'].append(b['mass']); M['body_inertia'].append(list(b['inertia']))
        js = [j for j in joints if j['body'] == bi]
        assert len(js) <= 1
        if js:
            j = js[0]
            ji = len(M['jnt_type'])
            free = j['type'] == 'free'
            M['body_jntadr'].append(ji); M['body_jntnum'].append(1); M['body_dofadr'].append(dadr); M['body_dofnum'].append(6 if free else 1)
            M['body_weldid'].append(bi)
            M['jnt_type'].append(0 if free else 3); M['jnt_qposadr'].append(qadr); M['jnt_dofadr'].append(dadr); M['jnt_bodyid'].append(bi)
            M['jnt_pos'].append([0, 0, 0]); M['jnt_axis'].append(list(j.get('axis', (0, 0, 1))))
            rng = j.get('range')
            M['jnt_range'].append(list(rng) if rng else [0, 0]); M['jnt_limited'].append(1 if rng else 0); M['jnt_margin'].append(0.0)
            nd = 6 if free else 1
            for _ in range(nd):
                M['dof_bodyid'].append(bi); M['dof_jntid'].append(ji)
                M['dof_armature'].append(0.0 if free else j.get('armature', 0.0)); M['dof_damping'].append(0.0 if free else j.get('damping', 0.0))
            M['qpos0'] += (list(b['pos']) + [1, 0, 0, 0]) if free else [0.0]
            qadr += 7 if free else 1; dadr += nd
        else:
            M['body_jntadr'].append(-1); M['body_jntnum'].append(0); M['body_dofadr'].append(-1); M['body_dofnum'].append(0)
            M['body_weldid'].append(M['body_weldid'][b.get('parent', 0)])
    T = dict(plane=0, sphere=2, capsule=3)
    M.update(geom_type=[T[g['type']] for g in geoms], geom_bodyid=[g['body'] for g in geoms], geom_pos=[list(g.get('pos', (0, 0, 0))) for g in geoms],
             geom_quat=[[1, 0, 0, 0] for _ in geoms], geom_size=[list(g['size']) for g in geoms], geom_margin=[g.get('margin', 0.0) for g in geoms],
             geom_friction=[list(g.get('friction', (1, 0.005, 0.0001))) for g in geoms], geom_contype=[g.get('contact', 1) for g in geoms],
             geom_conaffinity=[g.get('contact', 1) for g in geoms], geom_condim=[3 for _ in geoms])
    M.update(act_jntid=[a['joint'] for a in acts], act_gear=[a['gear'] for a in acts], act_ctrlrange=[[-1, 1] for _ in acts])
    M.update(nq=qadr, nv=dadr, nu=len(acts), nbody=nb, njnt=len(M['jnt_type']), ngeom=len(geoms), timestep=timestep, gravity=list(gravity))
    return M


def settle(om, q, v, ctrl, steps):
    w = np.zeros(om.nv)
    for _ in range(steps):
        om.step(q, v, ctrl, 1, w)
    return q, v


@pytest.mark.parametrize('mass,margin', [(1.0, 0.01), (7.5, 0.01), (2.0, 0.0)])
def test_sphere_at_rest_hovers_at_the_closed_form_height(mass, margin):
    """A free sphere resting on the floor plane: 4 pyramid rows share the load, r = dist - margin solves
    |r| = g (1 - d(r)) / (d(r)^2 k) -- independent of the mass because diagApprox = body_invweight0 = 1 / m for a free body."""
    r_s = 0.25
    I = 0.4 * mass * r_s ** 2
    M = make_model([dict(pos=(0, 0, 0.3), mass=mass, inertia=(I, I, I))], [dict(body=1, type='free')],
                   [dict(body=0, type='plane', size=(20, 20, 0.1)), dict(body=1, type='sphere', size=(r_s, 0, 0), margin=margin)])
    om = OracleModel(M)
    np.testing.assert_allclose(om.body_invweight0[1, 0], 1.0 / mass, rtol=1e-12)
    q = om.qpos0.copy(); v = np.zeros(6)
    settle(om, q, v, np.zeros(0), 400)
    assert abs(v).max() < 1e-9
    pen = fixed_point(lambda x: G * (1 - impedance(x)) / (impedance(x) ** 2 * K))
    assert 1e-5 < pen < WIDTH                       # inside the impedance ramp: the anchor exercises the sigmoid too
    np.testing.assert_allclose(q[2] - r_s, margin - pen, rtol=0, atol=2e-9)


def test_joint_limit_steady_state_under_motor_torque():
    """A single limited hinge (armature 1, as tatami.xml:6) pushed against its upper limit by a motor: the limit row is one-sided,
    force = torque at rest, so the overshoot is  delta = tau (1 - d) / (d^2 k) * dof_invweight0,  dof_invweight0 = 1 / (I + armature)."""
    Iy, arm, gear = 0.1, 1.0, 150.0
    M = make_model([dict(pos=(0, 0, 1), mass=1.0, inertia=(0.1, Iy, 0.1))], [dict(body=1, type='hinge', axis=(0, 1, 0), range=(-0.5, 0.5), armature=arm, damping=1.0)],
                   [dict(body=1, type='sphere', size=(0.05, 0, 0), contact=0)], acts=[dict(joint=0, gear=gear)], gravity=(0, 0, 0))
    om = OracleModel(M)
    iwd = 1.0 / (Iy + arm)
    np.testing.assert_allclose(om.dof_invweight0[0], iwd, rtol=1e-12)
    for ctrl in (1.0, 0.3, 2.0):                   # 2.0 is clipped to ctrlrange 1 (ant.xml:4)
        tau = gear * min(ctrl, 1.0)
        q = np.array([0.45]); v = np.zeros(1)
        settle(om, q, v, np.array([ctrl]), 600)
        assert abs(v[0]) < 1e-9
        delta = fixed_point(lambda x: tau * (1 - impedance(x)) * iwd / (impedance(x) ** 2 * K), 1e-3)
        np.testing.assert_allclose(q[0] - 0.5, delta, rtol=1e-7)


@pytest.mark.parametrize('theta_deg,slips', [(35.0, False), (42.0, False), (48.0, True), (55.0, True)])
def test_friction_cone_slip_onset_at_mu_one(theta_deg, slips):
    """A four-footed sled (cannot roll) on the plane with gravity tilted by theta about y: with friction 1 it creeps at a bounded
    speed for tan(theta) < mu and accelerates without bound for tan(theta) > mu, towards g (sin theta - mu cos theta)."""
    th = np.radians(theta_deg)
    feet = [dict(body=1, type='sphere', pos=(sx * 0.3, sy * 0.3, 0.0), size=(0.1, 0, 0), margin=0.01, friction=(1, 0.005, 0.0001)) for sx in (-1, 1) for sy in (-1, 1)]
    M = make_model([dict(pos=(0, 0, 0.11), mass=2.0, inertia=(0.2, 0.2, 0.3))], [dict(body=1, type='free')],
                   [dict(body=0, type='plane', size=(50, 50, 0.1), friction=(1, 0.1, 0.1))] + feet, gravity=(G * np.sin(th), 0, -G * np.cos(th)))
    om = OracleModel(M)
    q = om.qpos0.copy(); v = np.zeros(6)
    settle(om, q, v, np.zeros(0), 150)
    v1 = v[0]
    settle(om, q, v, np.zeros(0), 50)
    acc = (v[0] - v1) / 0.5
    if slips:
        want = G * (np.sin(th) - np.cos(th))
        assert acc > 0.5 * want and acc < 1.05 * want, (acc, want)
        assert v[0] > 0.3
    else:
        assert abs(acc) < 2e-3 * G and 0 <= v[0] < 0.1, (acc, v[0])
