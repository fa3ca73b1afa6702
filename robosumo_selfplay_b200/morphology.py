"""Morphology tables and scene constants of the RoboSumo arena, compiled into the constant
pack the CUDA kernels consume.

This is the product-side model compiler: it does NOT parse MJCF.  The three bodies of the
reference (robosumo/robosumo/envs/assets/{ant,bug,spider}.xml) all share one topology --
a spherical torso on a free joint carrying L legs, each leg = a capsule welded to the torso,
a `hip` hinge body with one capsule, and an `ankle` hinge body with one capsule -- so each
morphology is a small table of leg vectors, joint axes and ranges.  The scene merge rules
of robosumo/robosumo/envs/utils.py:46-183 (class density, margins, tatami resize) and the
registration constants of robosumo/robosumo/__init__.py:8-105 are applied here as data.

tests/test_morphology.py checks every derived quantity (masses, inertias, body/joint order,
invweights) against tests/golden/model_*.json, which is produced from the reference's own
XML files by the oracle-side compiler.
"""
import ctypes
import math

import numpy as np

MAXL = 8

# --- scene (assets/tatami.xml + utils.py:64-88 + sumo.py:47-55) ---------------------------------
TIMESTEP = 0.01
FRAME_SKIP = 5
GRAVITY_Z = -9.81
FLOOR_Z = -0.025
TATAMI_SIZE = 2.0
TATAMI_BOX_HALF = (TATAMI_SIZE + 0.3, TATAMI_SIZE + 0.3, 0.25)
TATAMI_BOX_Z = 0.25
BORDER_Z = 0.5
BORDER_R = 0.03
RING_LIMIT = TATAMI_SIZE + 0.1         # sumo.py:55
TIMESTEP_LIMIT = 500
MARGIN = 0.01                          # agent geom margin (ant.xml:3); world geoms have 0 -> max = 0.01
FRICTION = 1.0                         # sliding friction: max(agent 1, world 1)
HINGE_ARMATURE = 1.0                   # tatami.xml:6
HINGE_DAMPING = 1.0
GEAR = 150.0
INIT_RADIUS, INIT_Z = 1.5, 0.75        # utils.py:108-115 (qpos0)
RESET_RADIUS, RESET_Z = 1.15, 1.25     # sumo.py:235
LEG_DENSITY_OVERRIDE = {'spider': 5.0}  # spider.xml leg geoms carry density="5.0" explicitly
AGENT_DENSITY = {'ant': 13.0, 'bug': 10.0, 'spider': 39.0}

_S = 1.0  # readability


def _legs_ant():
    legs = []
    for (sx, sy, ax, rng) in ((-1, 1, (1, 1, 0), (-70, -30)), (1, 1, (-1, 1, 0), (30, 70)),
                              (-1, -1, (-1, 1, 0), (-70, -30)), (1, -1, (1, 1, 0), (30, 70))):
        legs.append(dict(hip=(0.2 * sx, 0.2 * sy, 0.0), ank=(0.2 * sx, 0.2 * sy, 0.0), tip=(0.4 * sx, 0.4 * sy, 0.0),
                         ank_axis=ax, hip_range=(-30, 30), ank_range=rng))
    return legs


def _legs_bug():
    legs = []
    spec = (
        ((0.18, 0.215), (0.18, 0.215), (0.324, 0.387), (-1, 1, 0), (30, 70)),
        ((-0.18, 0.215), (-0.18, 0.215), (-0.324, 0.387), (1, 1, 0), (-70, -30)),
        ((-0.2, 0.0), (-0.275, 0.0), (-0.55, 0.0), (0, 1, 0), (-70, -30)),
        ((0.2, 0.0), (0.275, 0.0), (0.55, 0.0), (0, 1, 0), (30, 70)),
        ((-0.18, -0.215), (-0.18, -0.215), (-0.324, -0.387), (-1, 1, 0), (-70, -30)),
        ((0.18, -0.215), (0.18, -0.215), (0.324, -0.387), (1, 1, 0), (30, 70)),
    )
    for hip, ank, tip, ax, rng in spec:
        legs.append(dict(hip=hip + (0.0,), ank=ank + (0.0,), tip=tip + (0.0,), ank_axis=ax,
                         hip_range=(-30, 30), ank_range=rng))
    return legs


def _legs_spider():
    legs = []
    spec = (
        ((-0.056, 0.209), (-0.050, 0.188), (-0.112, 0.418), (0.97, 0.26), 1),
        ((0.056, 0.209), (0.050, 0.188), (0.112, 0.418), (-0.97, 0.26), -1),
        ((-0.188, 0.108), (-0.170, 0.097), (-0.376, 0.216), (0.50, 0.87), 1),
        ((0.188, 0.108), (0.170, 0.097), (0.376, 0.216), (-0.50, 0.87), -1),
        ((-0.209, -0.056), (-0.188, -0.050), (-0.418, -0.112), (-0.26, 0.97), 1),
        ((0.209, -0.056), (0.188, -0.050), (0.418, -0.112), (0.26, 0.97), -1),
        ((-0.108, -0.188), (-0.097, -0.170), (-0.216, -0.376), (-0.87, 0.50), 1),
        ((0.108, -0.188), (0.097, -0.170), (0.216, -0.376), (0.87, 0.50), -1),
    )
    for hip, ank, tip, ax, s in spec:
        legs.append(dict(hip=hip + (0.125,), ank=ank + (0.113,), tip=tip + (-0.600,), ank_axis=ax + (0.0,),
                         hip_range=(-20, 20), ank_range=(-35, 45) if s > 0 else (-45, 35)))
    return legs


MORPHOLOGIES = {
    'ant': dict(torso_r=0.25, leg_r=0.08, legs=_legs_ant()),
    'bug': dict(torso_r=0.25, leg_r=0.08, legs=_legs_bug()),
    'spider': dict(torso_r=0.25, leg_r=0.04, legs=_legs_spider()),
}


class rs_agent_model(ctypes.Structure):
    """Mirror of `rs_agent_model` in include/rs_b200.h (field order and sizes must match)."""
    _fields_ = [
        ('L', ctypes.c_int), ('nq', ctypes.c_int), ('nv', ctypes.c_int), ('nu', ctypes.c_int),
        ('torso_r', ctypes.c_float), ('leg_r', ctypes.c_float), ('armature', ctypes.c_float),
        ('damping', ctypes.c_float), ('gear', ctypes.c_float), ('adjust_z', ctypes.c_float),
        ('reach', ctypes.c_float), ('pad2', ctypes.c_float),
        ('mT', ctypes.c_float), ('cT', ctypes.c_float * 3), ('IT', ctypes.c_float * 9),
        ('iw_torso', ctypes.c_float), ('iw_aux', ctypes.c_float * MAXL), ('iw_hip', ctypes.c_float * MAXL),
        ('iw_ank', ctypes.c_float * MAXL), ('iwd_hip', ctypes.c_float * MAXL), ('iwd_ank', ctypes.c_float * MAXL),
        ('r_hip', ctypes.c_float * (3 * MAXL)), ('ax_hip', ctypes.c_float * (3 * MAXL)),
        ('r_ank', ctypes.c_float * (3 * MAXL)), ('ax_ank', ctypes.c_float * (3 * MAXL)),
        ('e_ank', ctypes.c_float * (3 * MAXL)),
        ('m_hip', ctypes.c_float * MAXL), ('ip_hip', ctypes.c_float * MAXL), ('ia_hip', ctypes.c_float * MAXL),
        ('m_ank', ctypes.c_float * MAXL), ('ip_ank', ctypes.c_float * MAXL), ('ia_ank', ctypes.c_float * MAXL),
        ('lo_hip', ctypes.c_float * MAXL), ('hi_hip', ctypes.c_float * MAXL),
        ('lo_ank', ctypes.c_float * MAXL), ('hi_ank', ctypes.c_float * MAXL),
    ]


def _capsule(density, r, length):
    """mass and principal inertias (perpendicular, axial) of a capsule of total axis length `length`."""
    h = length
    mass = density * (math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3)
    ms = mass * 4 * r / (4 * r + 3 * h)
    mc = mass - ms
    ip = mc * (3 * r * r + h * h) / 12 + 2 * ms * r * r / 5 + ms * h * (3 * r + 2 * h) / 8
    ia = mc * r * r / 2 + 2 * ms * r * r / 5
    return mass, ip, ia


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


class AgentSpec:
    """All per-morphology constants in float64; `.pack()` gives the float32 C struct."""

    def __init__(self, name, adjust_z=0.0):
        mo = MORPHOLOGIES[name]
        self.name = name
        self.adjust_z = adjust_z
        self.L = L = len(mo['legs'])
        self.nq, self.nv, self.nu = 7 + 2 * L, 6 + 2 * L, 2 * L
        self.nbody = 1 + 3 * L
        self.obs_dim = self.nq + self.nv + 6 * self.nbody + 7 + 6 + 1     # agents.py:190-214
        self.torso_r, self.leg_r = mo['torso_r'], mo['leg_r']
        rho = AGENT_DENSITY[name]
        rho_leg = LEG_DENSITY_OVERRIDE.get(name, rho)
        deg = math.pi / 180.0
        self.r_hip = np.array([l['hip'] for l in mo['legs']], dtype=np.float64)
        self.r_ank = np.array([l['ank'] for l in mo['legs']], dtype=np.float64)      # in the hip frame
        self.e_ank = np.array([l['tip'] for l in mo['legs']], dtype=np.float64)      # in the ankle frame
        self.ax_hip = np.tile(np.array([0.0, 0.0, 1.0]), (L, 1))
        ax = np.array([l['ank_axis'] for l in mo['legs']], dtype=np.float64)
        self.ax_ank = ax / np.linalg.norm(ax, axis=1, keepdims=True)
        self.lo_hip = np.array([l['hip_range'][0] * deg for l in mo['legs']])
        self.hi_hip = np.array([l['hip_range'][1] * deg for l in mo['legs']])
        self.lo_ank = np.array([l['ank_range'][0] * deg for l in mo['legs']])
        self.hi_ank = np.array([l['ank_range'][1] * deg for l in mo['legs']])
        # body masses / inertias (inertiafromgeom, one geom per body)
        self.m_torso = rho * 4.0 / 3.0 * math.pi * self.torso_r ** 3
        self.i_torso = 0.4 * self.m_torso * self.torso_r ** 2
        self.aux = [_capsule(rho_leg, self.leg_r, np.linalg.norm(self.r_hip[l])) for l in range(L)]
        self.hip = [_capsule(rho_leg, self.leg_r, np.linalg.norm(self.r_ank[l])) for l in range(L)]
        self.ank = [_capsule(rho_leg, self.leg_r, np.linalg.norm(self.e_ank[l])) for l in range(L)]
        # torso rigid group = torso sphere + welded aux capsules
        mT = self.m_torso + sum(a[0] for a in self.aux)
        c = sum(self.aux[l][0] * self.r_hip[l] / 2 for l in range(L)) / mT
        I = self.i_torso * np.eye(3) + self.m_torso * (c @ c * np.eye(3) - np.outer(c, c))
        for l in range(L):
            m, ip, ia = self.aux[l]
            u = self.r_hip[l] / np.linalg.norm(self.r_hip[l])
            d = self.r_hip[l] / 2 - c
            I += ip * np.eye(3) + (ia - ip) * np.outer(u, u) + m * (d @ d * np.eye(3) - np.outer(d, d))
        self.mT, self.cT, self.IT = mT, c, I
        self.total_mass = mT + sum(h[0] for h in self.hip) + sum(a[0] for a in self.ank)
        self._invweights()

    # body ordering of the reference model inside one agent: torso, then per leg (aux, hip, ankle)
    def _bodies_qpos0(self):
        """(mass, com, inertia3x3, chain) of every MuJoCo body of this agent at qpos0, torso frame."""
        out = [(self.m_torso, np.zeros(3), self.i_torso * np.eye(3), ())]
        for l in range(self.L):
            for kind in range(3):
                m, ip, ia = (self.aux, self.hip, self.ank)[kind][l]
                a = (np.zeros(3), self.r_hip[l], self.r_hip[l] + self.r_ank[l])[kind]
                b = a + (self.r_hip[l], self.r_ank[l], self.e_ank[l])[kind]
                u = (b - a) / np.linalg.norm(b - a)
                I = ip * np.eye(3) + (ia - ip) * np.outer(u, u)
                chain = ((), (('hip', l),), (('hip', l), ('ank', l)))[kind]
                out.append((m, (a + b) / 2, I, chain))
        return out

    def _jac(self, com, chain):
        """6 x nv Jacobian [lin; ang] of a body point at qpos0 (identity torso orientation)."""
        J = np.zeros((6, self.nv))
        J[0:3, 0:3] = np.eye(3)
        J[0:3, 3:6] = -_skew(com)          # e_k x com
        J[3:6, 3:6] = np.eye(3)
        for kind, l in chain:
            dof = 6 + 2 * l + (0 if kind == 'hip' else 1)
            axis = self.ax_hip[l] if kind == 'hip' else self.ax_ank[l]
            anchor = self.r_hip[l] if kind == 'hip' else self.r_hip[l] + self.r_ank[l]
            J[0:3, dof] = np.cross(axis, com - anchor)
            J[3:6, dof] = axis
        return J

    def _invweights(self):
        """body_invweight0 / dof_invweight0 of MuJoCo's mj_setConst at qpos0 [M]."""
        nv = self.nv
        M = np.zeros((nv, nv))
        bodies = self._bodies_qpos0()
        Js = []
        for m, com, I, chain in bodies:
            J = self._jac(com, chain)
            Js.append(J)
            M += m * J[0:3].T @ J[0:3] + J[3:6].T @ I @ J[3:6]
        M[np.arange(6, nv), np.arange(6, nv)] += HINGE_ARMATURE
        self.M0 = M
        Minv = np.linalg.inv(M)
        iw = []
        for J in Js:
            A = J @ Minv @ J.T
            iw.append((np.trace(A[0:3, 0:3]) / 3, np.trace(A[3:6, 3:6]) / 3))
        self.body_invweight0 = np.array(iw)         # rows: torso, (aux, hip, ank) per leg
        d = np.diag(Minv).copy()
        d[0:3] = d[0:3].mean()
        d[3:6] = d[3:6].mean()
        self.dof_invweight0 = d

    def pack(self):
        s = rs_agent_model()
        s.L, s.nq, s.nv, s.nu = self.L, self.nq, self.nv, self.nu
        s.torso_r, s.leg_r = self.torso_r, self.leg_r
        s.armature, s.damping, s.gear, s.adjust_z = HINGE_ARMATURE, HINGE_DAMPING, GEAR, self.adjust_z
        s.reach = max(np.linalg.norm(self.r_hip[l]) + np.linalg.norm(self.r_ank[l]) + np.linalg.norm(self.e_ank[l])
                      for l in range(self.L)) + max(self.leg_r, self.torso_r)
        s.mT = self.mT
        for k in range(3):
            s.cT[k] = self.cT[k]
        for k in range(9):
            s.IT[k] = self.IT.ravel()[k]
        s.iw_torso = self.body_invweight0[0, 0]
        for l in range(self.L):
            s.iw_aux[l] = self.body_invweight0[1 + 3 * l, 0]
            s.iw_hip[l] = self.body_invweight0[2 + 3 * l, 0]
            s.iw_ank[l] = self.body_invweight0[3 + 3 * l, 0]
            s.iwd_hip[l] = self.dof_invweight0[6 + 2 * l]
            s.iwd_ank[l] = self.dof_invweight0[7 + 2 * l]
            for k in range(3):
                s.r_hip[3 * l + k] = self.r_hip[l, k]
                s.ax_hip[3 * l + k] = self.ax_hip[l, k]
                s.r_ank[3 * l + k] = self.r_ank[l, k]
                s.ax_ank[3 * l + k] = self.ax_ank[l, k]
                s.e_ank[3 * l + k] = self.e_ank[l, k]
            s.m_hip[l], s.ip_hip[l], s.ia_hip[l] = self.hip[l]
            s.m_ank[l], s.ip_ank[l], s.ia_ank[l] = self.ank[l]
            s.lo_hip[l], s.hi_hip[l] = self.lo_hip[l], self.hi_hip[l]
            s.lo_ank[l], s.hi_ank[l] = self.lo_ank[l], self.hi_ank[l]
        return s


class PairSpec:
    """Two agents in one arena (RoboSumo-<A>-vs-<B>-v0, robosumo/__init__.py)."""

    def __init__(self, name_a, name_b, adjust_z=0.0):
        self.agents = [AgentSpec(name_a, adjust_z), AgentSpec(name_b, adjust_z)]
        self.nq = sum(a.nq for a in self.agents)
        self.nv = sum(a.nv for a in self.agents)
        self.nu = sum(a.nu for a in self.agents)
        self.obs_dims = [a.obs_dim for a in self.agents]
        self.act_dims = [a.nu for a in self.agents]

    def qpos0(self):
        q = []
        for i, a in enumerate(self.agents):
            ang = i * math.pi
            q += [INIT_RADIUS * math.cos(ang), INIT_RADIUS * math.sin(ang), INIT_Z, 1.0, 0.0, 0.0, 0.0] + [0.0] * (2 * a.L)
        return np.array(q)

    def pack(self):
        arr = (rs_agent_model * 2)()
        arr[0] = self.agents[0].pack()
        arr[1] = self.agents[1].pack()
        return arr


def parse_env_id(env_id):
    """'RoboSumo-Ant-vs-Bug-v0' -> ('ant', 'bug')."""
    parts = env_id.split('-')
    assert parts[0] == 'RoboSumo' and parts[2] == 'vs', env_id
    return parts[1].lower(), parts[3].lower()
