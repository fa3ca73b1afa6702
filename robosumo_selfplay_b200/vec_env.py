"""B200SumoVecEnv -- drop-in for the reference's vectorised env stack on one GPU.

Replaces, behind the same gym-style surface, the whole chain

    SubprocVecEnv (subproc_vec_env.py:35-116, ABC baselines/baselines/common/vec_env/vec_env.py:29-138)
      -> bench.Monitor (baselines/baselines/bench/monitor.py:51-78, episode info only)
      -> sumo_env.SumoEnv wrapper (sumo_env.py:23-72)
      -> robosumo SumoEnv (robosumo/robosumo/envs/sumo.py:120-253) / Agent (agents.py:156-223)
      -> MuJoCo via mujoco-py (mujoco_env.py:104-129)

with one CUDA kernel launch per `step` over all `num_envs` env pairs (csrc/rs_api.cu: k_step).

Two call styles:
  * host style (default, what `Runner` in the reference expects): numpy in, numpy out, `infos` is
    the reference's tuple[E] of tuple[2] of dict.  Copies go through `rs_step_host`.
  * device style (`device_api=True`): torch CUDA tensors in/out, zero-copy views of persistent
    buffers that the next `step_wait` overwrites (cf. ShmemVecEnv's persistent obs buffer,
    baselines/baselines/common/vec_env/shmem_vec_env.py:20-75); `infos` is the raw
    `[E, 2, 8]` info tensor plus the `[E, 3]` episode tensor.
"""
import ctypes
import time
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .morphology import PairSpec, parse_env_id, FRAME_SKIP, TIMESTEP, TIMESTEP_LIMIT, RING_LIMIT

INFO_KEYS = ('ctrl_reward', 'lose_penalty', 'win_reward', 'main_reward', 'move_to_opp_reward', 'push_opp_reward',
             'shaping_reward')


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)


class Tuple(tuple):
    """gym.spaces.Tuple stand-in: indexable, len() == number of agents."""

    def __new__(cls, spaces):
        return super().__new__(cls, spaces)

    @property
    def spaces(self):
        return tuple(self)


class VecEnv(ABC):
    """Same contract as baselines' VecEnv ABC (vec_env.py:29-138)."""
    closed = False
    viewer = None

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode='human'):
        raise NotImplementedError("rendering is out of scope of the B200 hot path")

    def get_images(self):
        raise NotImplementedError("rendering is out of scope of the B200 hot path")

    @property
    def unwrapped(self):
        return self


class _AgentView:
    """`env.agents[i]._adjust_z` as used by eval_robosumo_against_fix.py:111-112."""

    def __init__(self, spec):
        self._adjust_z = spec.adjust_z
        self.obs_dim = spec.obs_dim
        self.action_dim = spec.nu


class LazyInfos:
    """Sequence of E entries; entry e = (info dict of agent 0, info dict of agent 1), built when indexed."""

    def __init__(self, info, done, epi, t):
        self._info, self._done, self._epi, self._t = info, done, epi, t

    def __len__(self):
        return self._info.shape[0]

    def __getitem__(self, e):
        if isinstance(e, slice):
            return tuple(self[i] for i in range(*e.indices(len(self))))
        if e < 0:
            e += len(self)
        if not 0 <= e < len(self):
            raise IndexError(e)
        pair = []
        for a in range(2):
            row = self._info[e, a]
            d = {k: float(row[i]) for i, k in enumerate(INFO_KEYS)}
            flags = int(row[7])
            if flags & 1:
                d['winner'] = True
            if flags & 2:
                d['timeout'] = True
            if a == 0 and self._done[e, 0]:
                d['episode'] = {'r': round(float(self._epi[e, 0]), 6), 'dr': round(float(self._epi[e, 1]), 6),
                                'l': int(self._epi[e, 2]), 't': self._t}
            pair.append(d)
        return tuple(pair)

    def __iter__(self):
        return (self[e] for e in range(len(self)))

    # vectorised accessors for consumers that want a column without E dict constructions
    def column(self, key):
        """[E, 2] array of one info key (e.g. 'shaping_reward')."""
        return self._info[:, :, INFO_KEYS.index(key)].astype(np.float64)

    def finished(self):
        """Indices of the envs whose episode ended this step (agent 0 done)."""
        return np.nonzero(self._done[:, 0])[0]


class B200SumoVecEnv(VecEnv):
    def __init__(self, env_id='RoboSumo-Ant-vs-Ant-v0', num_envs=8, seed=42, device=0, adjust_z=0.0,
                 device_api=False, auto_reset=True, newton_iters=16, timestep_limit=TIMESTEP_LIMIT):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("B200SumoVecEnv needs a CUDA device; there is no CPU fallback")
        self.env_id = env_id
        self.spec = env_id
        names = parse_env_id(env_id)
        self.pair = PairSpec(names[0], names[1], adjust_z)
        self.agents = [_AgentView(a) for a in self.pair.agents]
        self.device_api = device_api
        self.auto_reset = auto_reset
        self.device = torch.device('cuda', device)
        L = _lib.lib()
        cfg = _lib.rs_config(num_envs=num_envs, frame_skip=FRAME_SKIP, timestep_limit=timestep_limit,
                             newton_iters=newton_iters, timestep=TIMESTEP, ring_limit=RING_LIMIT,
                             init_pos_noise=0.1, init_vel_noise=0.1, seed=seed, device=device, reserved=0)
        self._pack = self.pair.pack()
        h = ctypes.c_void_p()
        _lib.check(L.rs_create(ctypes.byref(cfg), self._pack, ctypes.byref(h)))
        self._h = h
        self._L = L
        oa, ob = self.pair.obs_dims
        aa, ab = self.pair.act_dims
        self.mixed = (oa != ob)            # mixed morphologies: ragged obs / actions, returned per agent
        self.obs_dims, self.act_dims = (oa, ob), (aa, ab)
        self.obs_dim, self.act_dim = oa, aa
        self.nq, self.nv, self.nu = self.pair.nq, self.pair.nv, self.pair.nu
        ospace = Tuple([Box(-np.inf, np.inf, (oa,)), Box(-np.inf, np.inf, (ob,))])
        aspace = Tuple([Box(-1.0, 1.0, (aa,)), Box(-1.0, 1.0, (ab,))])      # agents.py:92-115
        VecEnv.__init__(self, num_envs, ospace, aspace)
        E = num_envs
        dev = self.device
        self.d_obs = torch.zeros((E, oa + ob) if self.mixed else (E, 2, oa), dtype=torch.float32, device=dev)
        self.d_rew = torch.zeros((E, 2), dtype=torch.float32, device=dev)
        self.d_done = torch.zeros((E, 2), dtype=torch.uint8, device=dev)
        self.d_info = torch.zeros((E, 2, 8), dtype=torch.float32, device=dev)
        self.d_epi = torch.zeros((E, 3), dtype=torch.float32, device=dev)
        # host-style result buffers are page-locked, so rs_step_host copies device -> caller memory with no staging hop
        def pinned(shape, dtype):
            tt = torch.zeros(shape, dtype=dtype, pin_memory=True)
            self._pinned.append(tt)               # keeps the allocation alive behind the numpy view
            return tt.numpy()
        self._pinned = []
        self.h_obs = pinned((E, oa + ob) if self.mixed else (E, 2, oa), torch.float32)
        self.h_rew = pinned((E, 2), torch.float32)
        self.h_done = pinned((E, 2), torch.uint8)
        self.h_info = pinned((E, 2, 8), torch.float32)
        self.h_epi = pinned((E, 3), torch.float32)
        self.h_act = pinned((E, self.nu), torch.float32)
        self.waiting = False
        self.closed = False
        self._pending = None
        self.tstart = time.time()

    # ---- helpers -------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _np(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def _assert_not_closed(self):
        assert not self.closed, "Trying to operate on a B200SumoVecEnv after calling close()"

    # ---- VecEnv surface ------------------------------------------------------------------
    def reset(self, mask=None):
        self._assert_not_closed()
        m = None
        if mask is not None:
            m = self.torch.as_tensor(mask, dtype=self.torch.uint8, device=self.device).contiguous()
        _lib.check(self._L.rs_reset(self._h, ctypes.c_void_p(m.data_ptr()) if m is not None else None,
                                    ctypes.c_void_p(self.d_obs.data_ptr()), self._stream()))
        return self._obs_out(self.d_obs if self.device_api else self.d_obs.cpu().numpy().astype(np.float64))

    def _obs_out(self, obs):
        """Same-morphology pairs: [E, 2, D].  Mixed pairs: tuple (obs_a [E, Da], obs_b [E, Db])."""
        if not self.mixed:
            return obs
        return obs[:, :self.obs_dims[0]], obs[:, self.obs_dims[0]:]

    def _flat_actions(self, actions):
        if isinstance(actions, (tuple, list)) and len(actions) == 2 and getattr(actions[0], 'ndim', 0) == 2:
            if self.torch.is_tensor(actions[0]):
                return self.torch.cat([actions[0], actions[1]], 1)
            return np.concatenate([np.asarray(actions[0]), np.asarray(actions[1])], 1)
        return actions.reshape(self.num_envs, self.nu)

    def step_async(self, actions):
        self._assert_not_closed()
        self._pending = actions
        self.waiting = True

    def step_wait(self):
        self._assert_not_closed()
        actions = self._pending
        self.waiting = False
        if self.device_api:
            a = self._flat_actions(actions)
            if a.dtype != self.torch.float32 or not a.is_contiguous():
                a = a.to(self.torch.float32).contiguous()
            _lib.check(self._L.rs_step(self._h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(self.d_obs.data_ptr()),
                                       ctypes.c_void_p(self.d_rew.data_ptr()), ctypes.c_void_p(self.d_done.data_ptr()),
                                       ctypes.c_void_p(self.d_info.data_ptr()), ctypes.c_void_p(self.d_epi.data_ptr()),
                                       1 if self.auto_reset else 0, self._stream()))
            return self._obs_out(self.d_obs), self.d_rew, self.d_done, (self.d_info, self.d_epi)
        self.h_act[...] = np.asarray(self._flat_actions(actions), dtype=np.float32).reshape(self.num_envs, self.nu)
        a = self.h_act
        self.torch.cuda.current_stream(self.device).synchronize()
        _lib.check(self._L.rs_step_host(self._h, self._np(a), self._np(self.h_obs), self._np(self.h_rew),
                                        self._np(self.h_done), self._np(self.h_info), self._np(self.h_epi),
                                        1 if self.auto_reset else 0))
        return (self._obs_out(self.h_obs.astype(np.float64)), self.h_rew.astype(np.float64), self.h_done.astype(bool),
                self._infos_as_dicts())

    def _infos_as_dicts(self):
        """The reference's infos: tuple[E] of tuple[2] of dict (sumo.py:131-186, sumo_env.py:48-65) -- built LAZILY: the step
        returns a sequence view over a snapshot of the info / done / episode arrays and `infos[e]` materialises that env's two
        dicts on access, so the O(E) Python loop of round 1 (2 E dicts per step, several ms at E = 4096) leaves the step.
        Runner-style consumers that only read `infos[e][a]['shaping_reward']` or scan for 'episode' work unchanged."""
        return LazyInfos(self.h_info.copy(), self.h_done.copy(), self.h_epi.copy(), round(time.time() - self.tstart, 6))

    # ---- state access (parity hooks; MjSim.get_state / set_state) -------------------------
    def set_state(self, qpos, qvel):
        """qpos [E, nq], qvel [E, nv] (numpy or torch).  Returns the observation of the new state."""
        t = self.torch
        q = t.as_tensor(np.asarray(qpos, dtype=np.float32) if not t.is_tensor(qpos) else qpos, dtype=t.float32, device=self.device).contiguous()
        v = t.as_tensor(np.asarray(qvel, dtype=np.float32) if not t.is_tensor(qvel) else qvel, dtype=t.float32, device=self.device).contiguous()
        assert q.shape == (self.num_envs, self.nq) and v.shape == (self.num_envs, self.nv)
        _lib.check(self._L.rs_set_state(self._h, ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(v.data_ptr()),
                                        ctypes.c_void_p(self.d_obs.data_ptr()), self._stream()))
        t.cuda.current_stream(self.device).synchronize()
        return self._obs_out(self.d_obs if self.device_api else self.d_obs.cpu().numpy().astype(np.float64))

    def get_state(self):
        t = self.torch
        E = self.num_envs
        q = t.empty((E, self.nq), dtype=t.float32, device=self.device)
        v = t.empty((E, self.nv), dtype=t.float32, device=self.device)
        step = t.empty((E,), dtype=t.int32, device=self.device)
        status = t.empty((E,), dtype=t.int32, device=self.device)
        _lib.check(self._L.rs_get_state(self._h, ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(v.data_ptr()),
                                        ctypes.c_void_p(step.data_ptr()), ctypes.c_void_p(status.data_ptr()), self._stream()))
        return q, v, step, status

    def forward_debug(self, ctrl):
        """qacc, ncon, niter of one forward evaluation at the stored state (sim.forward())."""
        t = self.torch
        E = self.num_envs
        c = t.as_tensor(np.asarray(ctrl, dtype=np.float32) if not t.is_tensor(ctrl) else ctrl, dtype=t.float32, device=self.device).contiguous()
        qacc = t.empty((E, self.nv), dtype=t.float32, device=self.device)
        ncon = t.empty((E,), dtype=t.int32, device=self.device)
        nit = t.empty((E,), dtype=t.int32, device=self.device)
        _lib.check(self._L.rs_forward_debug(self._h, ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(qacc.data_ptr()),
                                            ctypes.c_void_p(ncon.data_ptr()), ctypes.c_void_p(nit.data_ptr()), self._stream()))
        return qacc, ncon, nit

    def diagnostics(self):
        """int32 [E, 4]: Newton iterations, coupled evaluations, contacts (summed over the last step's 20 evaluations), max iterations."""
        t = self.torch
        out = t.empty((self.num_envs, 4), dtype=t.int32, device=self.device)
        _lib.check(self._L.rs_get_diag(self._h, ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def check_status(self, strict=True, clear=True):
        """What mujoco-py's warning callback does inside the worker (builder.py:351-369: any MuJoCo warning becomes a
        MujocoException), checked once per rollout on a device-side latch that auto-reset does not clear: raises on NaN/Inf state
        (mjWARN_BADQPOS/QVEL/QACC) and, when `strict`, on a full contact buffer (mjWARN_CONTACTFULL: contacts beyond the per-pair
        capacity were dropped).  A Newton solve that hit its iteration cap is reported as a warning (MuJoCo does not warn there).
        Returns (bits, env_steps_flagged)."""
        import warnings
        out = (ctypes.c_int * 2)()
        _lib.check(self._L.rs_status_latch(self._h, out, 1 if clear else 0, self._stream()))
        bits, n = int(out[0]), int(out[1])
        if bits & 1:
            raise RuntimeError("B200SumoVecEnv: NaN/Inf in the physics state of an env pair (%d env-steps flagged since the last check)" % n)
        if bits & 2:
            msg = "B200SumoVecEnv: contact buffer full, contacts were dropped (%d env-steps flagged since the last check)" % n
            if strict:
                raise RuntimeError(msg)
            warnings.warn(msg)
        if bits & 4:
            warnings.warn("B200SumoVecEnv: a constraint solve stopped at the iteration cap (%d env-steps flagged since the last check)" % n)
        return bits, n

    def seed(self, seed=None):
        """env.seed(s) of the reference (run.py:73-83 seeds env i with seed + i; mujoco_env.py:82-84).  Here env e draws its reset
        states from the Philox stream (seed, e); the new key takes effect at the next reset / auto-reset."""
        if seed is None:
            seed = int(time.time() * 1e6) & 0xFFFFFFFF
        _lib.check(self._L.rs_seed(self._h, ctypes.c_ulonglong(int(seed) & 0xFFFFFFFFFFFFFFFF)))
        return [int(seed) + e for e in range(self.num_envs)]

    def close_extras(self):
        if getattr(self, '_h', None):
            self._L.rs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if not self.closed:
                self.close()
        except Exception:
            pass
