"""ORACLE (test infrastructure): sequential CPU restatement of the RoboSumo env stack.

Layers restated (one class per reference layer, same method names):

  OracleSumoCore   robosumo/robosumo/envs/sumo.py:120-253   (_step, reward/done block,
                   _comp_move_reward, _comp_push_reward, reset_model)
                   robosumo/robosumo/envs/agents.py:156-223  (get_qpos, get_obs,
                   before_step/after_step)
                   robosumo/robosumo/envs/mujoco_env.py:104-129 (_reset, set_state,
                   do_simulation)
  OracleSumoWrap   sumo_env.py:23-72                         (episode bookkeeping,
                   timestep feature, timeout flag)
  OracleVecEnv     subproc_vec_env.py:6-32,65-82             (worker auto-reset, stacking)

Physics comes from oracle/physics.py (restated MuJoCo, parity UNPINNED).  The trusted
sequential reference in the sense of baselines' test_vec_env.py:14-44.
"""
import numpy as np

from .mjcf_compile import agent_slices
from .physics import OracleModel

WIN_REWARD = 2000.0
DRAW_PENALTY = -1000.0
MOVE_TO_OPP_COEF = 0.1
PUSH_OUT_COEF = 10.0
CTRL_COEF = 0.1
CFRC_CLIP = 100.0


class OracleSumoCore:
    def __init__(self, model_dict, frame_skip=5, tatami_size=2.0, timestep_limit=500,
                 init_pos_noise=0.1, init_vel_noise=0.1, adjust_z=0.0, seed=None):
        self.M = model_dict
        self.om = OracleModel(model_dict)
        self.sl = agent_slices(model_dict)
        self.frame_skip = frame_skip
        self._tatami_size = tatami_size + 0.1          # sumo.py:55
        self._timestep_limit = timestep_limit
        self._init_pos_noise = init_pos_noise
        self._init_vel_noise = init_vel_noise
        self._adjust_z = adjust_z
        self.dt = model_dict['timestep'] * frame_skip   # mujoco_env.py:121-123
        self.np_random = np.random.RandomState(seed)
        self.qpos = self.om.qpos0.copy()
        self.qvel = np.zeros(self.om.nv)
        self.warm = np.zeros(self.om.nv)
        self._num_steps = 0
        self.max_ncon = 0
        self.act_dims = [sum(1 for j in model_dict['act_jntid'] if model_dict['jnt_name'][j].startswith(s))
                         for s in model_dict['scopes']]

    # agents.py:156-161
    def _agent_qpos(self, i):
        q = self.qpos[self.sl[i]['q0']:self.sl[i]['q1']].copy()
        q[2] += self._adjust_z
        return q

    def _agent_qvel(self, i):
        return self.qvel[self.sl[i]['v0']:self.sl[i]['v1']]

    # agents.py:190-214 ; cfrc_ext stays zero under MuJoCo >= 2.0 without force sensors [M]
    def _get_obs(self):
        out = []
        for i in range(2):
            opp = 1 - i
            nb = len(self.sl[i]['bodies'])
            obs = [self._agent_qpos(i), self._agent_qvel(i), np.zeros(6 * nb),
                   self._agent_qpos(opp)[:7], np.zeros(6), np.array([-1.0])]
            out.append(np.concatenate(obs))
        return tuple(out)

    def set_state(self, qpos, qvel):
        self.qpos = np.array(qpos, dtype=np.float64)
        self.qvel = np.array(qvel, dtype=np.float64)
        self.om.normalize_qpos(self.qpos)              # mj_forward -> mj_kinematics normalises quats [M]

    def reset(self):
        # mujoco_env.py:104-108 ; mj_resetData also clears qacc_warmstart
        self.qpos = self.om.qpos0.copy()
        self.qvel = np.zeros(self.om.nv)
        self.warm = np.zeros(self.om.nv)
        return self.reset_model()

    def reset_model(self):
        # sumo.py:232-253
        self._num_steps = 0
        r, z = 1.15, 1.25
        delta = (2.0 * np.pi) / 2
        phi = self.np_random.uniform(0.0, 2.0 * np.pi)
        for i in range(2):
            angle = phi + i * delta
            x, y = r * np.cos(angle), r * np.sin(angle)
            s = self.sl[i]['q0']
            if x: self.qpos[s] = x           # agents.py:117-125 (`if xyz[k]:` guards)
            if y: self.qpos[s + 1] = y
            if z: self.qpos[s + 2] = z
        pos_noise = self.np_random.uniform(size=self.om.nq, low=-self._init_pos_noise, high=self._init_pos_noise)
        vel_noise = self._init_vel_noise * self.np_random.randn(self.om.nv)
        self.set_state(self.qpos + pos_noise, self.qvel + vel_noise)
        return self._get_obs()

    def step(self, actions):
        # sumo.py:120-192
        posbefore = [self._agent_qpos(i)[:2].copy() for i in range(2)]
        ctrl = np.concatenate([np.asarray(a, dtype=np.float64) for a in actions])
        ncon = self.om.step(self.qpos, self.qvel, ctrl, self.frame_skip, self.warm)
        self.max_ncon = max(self.max_ncon, ncon)
        posafter = [self._agent_qpos(i)[:2].copy() for i in range(2)]
        infos = [{}, {}]
        for i in range(2):
            infos[i]['ctrl_reward'] = -CTRL_COEF * np.square(np.asarray(actions[i], dtype=np.float64)).sum()
        obs = self._get_obs()
        self._num_steps += 1
        dones = [False, False]
        rewards = [0.0, 0.0]

        def out_of_ring(xyz):
            return bool(xyz[2] < 0.29 or np.max(np.abs(xyz[:2])) >= self._tatami_size)

        for i in range(2):
            opp = 1 - i
            infos[i]['lose_penalty'] = 0.0
            if out_of_ring(self._agent_qpos(i)[:3]):
                infos[i]['lose_penalty'] = -WIN_REWARD
                dones[i] = True
            infos[i]['win_reward'] = 0.0
            if out_of_ring(self._agent_qpos(opp)[:3]):
                infos[i]['win_reward'] += WIN_REWARD
                infos[i]['winner'] = True
                dones[i] = True
            infos[i]['main_reward'] = infos[i]['win_reward'] + infos[i]['lose_penalty']
            if self._num_steps > self._timestep_limit:
                infos[i]['main_reward'] += DRAW_PENALTY
                dones[i] = True
            move_vec = (posafter[i] - posbefore[i]) / self.dt
            direction = posafter[opp] - posbefore[i]
            direction = direction / np.linalg.norm(direction)
            infos[i]['move_to_opp_reward'] = max(np.sum(move_vec * direction), 0.0) * MOVE_TO_OPP_COEF
            infos[i]['push_opp_reward'] = -PUSH_OUT_COEF * np.exp(-np.linalg.norm(posafter[opp]))
            infos[i]['shaping_reward'] = infos[i]['ctrl_reward'] + infos[i]['push_opp_reward'] + infos[i]['move_to_opp_reward']
            rewards[i] = infos[i]['main_reward'] + infos[i]['shaping_reward']
        return obs, tuple(rewards), tuple(dones), tuple(infos)


class OracleSumoWrap:
    """sumo_env.py:6-72."""

    def __init__(self, core):
        self.env = core
        self.needs_reset = True
        self.episode_step = 0
        self.rewards = None
        self.dense_rewards = None

    def reset(self):
        self.rewards = []
        self.dense_rewards = []
        self.needs_reset = False
        obs = self.env.reset()
        self.episode_step = 0
        return obs

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        obs, rew, done, info = self.env.step(action)
        self.rewards.append(rew[0])
        self.dense_rewards.append(info[0]['shaping_reward'])
        if done[0]:
            self.needs_reset = True
            info[0]['episode'] = {"r": round(sum(self.rewards), 6), "dr": round(sum(self.dense_rewards), 6),
                                  "l": len(self.rewards)}
            if info[0]['main_reward'] == -1000:
                for agt in info:
                    agt['timeout'] = True
        self.episode_step += 1
        for ob in obs:
            ob[-1] += 2.0 * self.episode_step / 500.0
        return obs, rew, done, info


class OracleVecEnv:
    """Sequential stand-in for SubprocVecEnv (subproc_vec_env.py): one wrapped env per slot,
    auto-reset when agent 0 is done (terminal obs replaced; reward/done kept)."""

    def __init__(self, model_dict, num_envs, seed=0, **kw):
        self.envs = [OracleSumoWrap(OracleSumoCore(model_dict, seed=seed + i, **kw)) for i in range(num_envs)]
        self.num_envs = num_envs
        self.reset_hook = None      # optional callable(env_index, core) -> None, used by parity tests to inject states

    def reset(self):
        obs = []
        for i, e in enumerate(self.envs):
            o = e.reset()
            if self.reset_hook:
                self.reset_hook(i, e.env)
                o = e.env._get_obs()
            obs.append(np.stack(o))
        return np.stack(obs)

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for i, e in enumerate(self.envs):
            o, r, d, info = e.step(actions[i])
            if d[0]:
                o = e.reset()
                if self.reset_hook:
                    self.reset_hook(i, e.env)
                    o = e.env._get_obs()
            obs.append(np.stack(o)); rews.append(r); dones.append(d); infos.append(info)
        return np.stack(obs), np.stack(rews), np.stack(dones), tuple(infos)
