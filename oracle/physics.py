"""ORACLE (test infrastructure): ctypes front-end of oracle/physics_oracle.c.

Loads a generic model (dict made by oracle/mjcf_compile.py, normally read back from
tests/golden/model_*.json so that nothing touches /root/reference at run time) and
exposes mj_forward / mj_step equivalents.  See physics_oracle.c for citations and
for the parity status (UNPINNED against MuJoCo).
"""
import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NATIVE = False
NATIVE_FLAGS = '-O3 -march=native -funroll-loops'
PORTABLE_FLAGS = '-O3'


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def use_native_build():
    """bench.py's CPU legs: compile the oracle on THIS machine with -O3 -march=native (oracle/Makefile `native`) and load that
    build instead of the portable one.  Returns the flag string in use.  Must be called before the first lib()."""
    global _NATIVE
    assert _LIB is None, "the oracle library is already loaded"
    try:
        subprocess.check_call(['make', '-s', '-C', _HERE, 'native'])
        _NATIVE = True
        return NATIVE_FLAGS
    except Exception:
        return PORTABLE_FLAGS


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libphysics_oracle_native.so' if _NATIVE else 'libphysics_oracle.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_model_create.restype = ctypes.c_void_p
        L.orc_model_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_model_free.argtypes = [ctypes.c_void_p]
        L.orc_set_const.argtypes = [ctypes.c_void_p]
        L.orc_get_invweight.argtypes = [ctypes.c_void_p] * 3
        L.orc_forward.restype = ctypes.c_int
        L.orc_forward.argtypes = [ctypes.c_void_p] * 13
        L.orc_step.restype = ctypes.c_int
        L.orc_step.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
        L.orc_step_batch.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_normalize_qpos.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def load_model_json(name):
    """name: 'ant_ant' | 'bug_bug' | 'spider_spider' -> generic model dict (tests/golden)."""
    path = os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'model_%s.json' % name)
    with open(path) as f:
        return json.load(f)


class OracleModel:
    INT_KEYS = ['body_parent', 'body_jntadr', 'body_jntnum', 'body_dofadr', 'body_dofnum', 'body_weldid',
                'jnt_type', 'jnt_qposadr', 'jnt_dofadr', 'jnt_bodyid', 'jnt_limited',
                'dof_bodyid', 'dof_jntid',
                'geom_type', 'geom_bodyid', 'geom_contype', 'geom_conaffinity', 'geom_condim', 'act_jntid']
    DBL_KEYS = ['body_pos', 'body_quat', 'body_ipos', 'body_iquat', 'body_mass', 'body_inertia',
                'jnt_pos', 'jnt_axis', 'jnt_range', 'jnt_margin', 'dof_armature', 'dof_damping',
                'geom_pos', 'geom_quat', 'geom_size', 'geom_margin', 'geom_friction',
                'act_gear', 'act_ctrlrange', 'qpos0']

    def __init__(self, M):
        self.M = M
        self.nq, self.nv, self.nu = M['nq'], M['nv'], M['nu']
        self.nbody, self.njnt, self.ngeom = M['nbody'], M['njnt'], M['ngeom']
        ints = [M['nq'], M['nv'], M['nu'], M['nbody'], M['njnt'], M['ngeom']]
        for k in self.INT_KEYS:
            ints += list(np.asarray(M[k], dtype=np.int64).ravel())
        dbls = [M['timestep']] + list(M['gravity'])
        for k in self.DBL_KEYS:
            dbls += list(np.asarray(M[k], dtype=np.float64).ravel())
        self._ints = np.asarray(ints, dtype=np.int32)
        self._dbls = np.asarray(dbls, dtype=np.float64)
        self.h = lib().orc_model_create(_p(self._ints), _p(self._dbls))
        assert self.h, "oracle rejected the model"
        lib().orc_set_const(self.h)
        self.body_invweight0 = np.zeros((self.nbody, 2))
        self.dof_invweight0 = np.zeros(self.nv)
        lib().orc_get_invweight(self.h, _p(self.body_invweight0), _p(self.dof_invweight0))
        self.qpos0 = np.asarray(M['qpos0'], dtype=np.float64)

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def forward(self, qpos, qvel, ctrl, full=False):
        qpos = np.array(qpos, dtype=np.float64); qvel = np.array(qvel, dtype=np.float64)
        ctrl = np.array(ctrl, dtype=np.float64)
        qacc = np.zeros(self.nv); Mm = np.zeros((self.nv, self.nv)); bias = np.zeros(self.nv)
        qs = np.zeros(self.nv); con = np.zeros((256, 8)); nefc = ctypes.c_int(0); it = ctypes.c_int(0)
        gx = np.zeros((self.ngeom, 3)); bx = np.zeros((self.nbody, 3))
        ncon = lib().orc_forward(self.h, _p(qpos), _p(qvel), _p(ctrl), _p(qacc), _p(Mm), _p(bias), _p(qs), _p(con),
                                 ctypes.addressof(nefc), ctypes.addressof(it), _p(gx), _p(bx))
        if not full:
            return qacc
        return dict(qacc=qacc, M=Mm, bias=bias, qacc_smooth=qs, contacts=con[:ncon], ncon=ncon, nefc=nefc.value,
                    iters=it.value, geom_xpos=gx, body_xpos=bx, qpos=qpos)

    def step(self, qpos, qvel, ctrl, nsub=5, warm=None):
        """In-place nsub x mj_step. Returns max ncon."""
        assert qpos.dtype == np.float64 and qvel.dtype == np.float64
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        return lib().orc_step(self.h, _p(qpos), _p(qvel), _p(ctrl), nsub, _p(warm))

    def step_batch(self, qpos, qvel, ctrl, nsub=5, warm=None, nthreads=1):
        E = qpos.shape[0]
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        lib().orc_step_batch(self.h, E, _p(qpos), _p(qvel), _p(ctrl), nsub, _p(warm), nthreads)

    def normalize_qpos(self, qpos):
        lib().orc_normalize_qpos(self.h, _p(qpos))
