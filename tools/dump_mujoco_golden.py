#!/usr/bin/env python
"""Runs OFF-BOX, on a machine that has the reference's stack installed (mujoco-py 2.1.2.14 + MuJoCo 2.1.0 + gym 0.15.4, see
/root/reference/requirements.txt and Dockerfile:21-24) -- this container has neither MuJoCo nor a network, so the file this tool
writes is the one piece of parity evidence that cannot be produced here (SURVEY 7 "Hard parts", 8c).

    PYTHONPATH=/path/to/robosumo-selfplay/robosumo python tools/dump_mujoco_golden.py \
        --env RoboSumo-Ant-vs-Ant-v0 --seeds 8 --steps 50 --out tests/golden/mujoco_ant_ant.npz [--time]

For every seed: reset the reference env (robosumo.envs.SumoEnv, robosumo/robosumo/envs/sumo.py), then T times: record the MuJoCo
state (`sim.get_state().flatten()` = [time, qpos, qvel], mujoco-py/mujoco_py/mjsimstate.pyx:31-39 -- the layout rs_set_state
takes), draw actions ~ N(0, 1) from a seeded RandomState, call env.step and record what comes back and what MuJoCo holds:
qpos', qvel', both observations (120-d upstream layout; this fork's wrapper appends the timestep feature, sumo_env.py:68-70),
rewards, dones, ncon, cfrc_ext, qacc_warmstart, and once per file the compiled-model constants the oracle's model compiler must
reproduce (body_mass, body_inertia, body_invweight0, dof_invweight0, geom_size, jnt_range, qpos0, opt.*).
`tests/test_mujoco_golden.py` loads every tests/golden/mujoco_*.npz it finds and replays it through the CPU oracle and the CUDA
kernels.  `--time` adds the single-process env-steps/s of the reference on that machine (the true CPU baseline)."""
import argparse
import time

import numpy as np

FORMAT_VERSION = 1


def make_env(env_id):
    import gym
    import robosumo.envs  # noqa: F401  (registers the RoboSumo ids, robosumo/robosumo/__init__.py:8-105)
    env = gym.make(env_id)
    return env.unwrapped if hasattr(env, 'unwrapped') else env


def dump(env_id, seeds, steps, do_time):
    env = make_env(env_id)
    sim, m = env.sim, env.sim.model
    out = dict(format_version=FORMAT_VERSION, env_id=env_id, source='mujoco-py', nq=m.nq, nv=m.nv, nu=m.nu,
               body_mass=np.array(m.body_mass), body_inertia=np.array(m.body_inertia), body_invweight0=np.array(m.body_invweight0),
               dof_invweight0=np.array(m.dof_invweight0), dof_armature=np.array(m.dof_armature), dof_damping=np.array(m.dof_damping),
               geom_size=np.array(m.geom_size), geom_pos=np.array(m.geom_pos), geom_type=np.array(m.geom_type), geom_bodyid=np.array(m.geom_bodyid),
               geom_margin=np.array(m.geom_margin), geom_friction=np.array(m.geom_friction), jnt_range=np.array(m.jnt_range), qpos0=np.array(m.qpos0),
               opt_timestep=m.opt.timestep, opt_integrator=int(m.opt.integrator), opt_solver=int(m.opt.solver), opt_cone=int(m.opt.cone),
               opt_iterations=int(m.opt.iterations), opt_tolerance=float(m.opt.tolerance), opt_gravity=np.array(m.opt.gravity),
               frame_skip=int(env.frame_skip))
    try:
        import mujoco_py
        out['mujoco_py_version'] = mujoco_py.__version__
    except Exception:
        pass
    S, T = seeds, steps
    A = m.nu // 2
    rec = dict(state0=np.zeros((S, T, 1 + m.nq + m.nv)), actions=np.zeros((S, T, 2, A)), qpos1=np.zeros((S, T, m.nq)), qvel1=np.zeros((S, T, m.nv)),
               rew=np.zeros((S, T, 2)), done=np.zeros((S, T, 2), bool), ncon=np.zeros((S, T), np.int32), cfrc_ext=np.zeros((S, T, m.nbody, 6)),
               qacc_warmstart=np.zeros((S, T, m.nv)), valid=np.zeros((S, T), bool))
    obs_rec = None
    for s in range(S):
        env.seed(1000 + s)
        env.reset()
        rng = np.random.RandomState(s)
        for t in range(T):
            rec['state0'][s, t] = sim.get_state().flatten()[:1 + m.nq + m.nv]
            a = rng.randn(2, A)
            rec['actions'][s, t] = a
            obs, rew, done, info = env.step((a[0], a[1]))
            if obs_rec is None:
                obs_rec = np.zeros((S, T, 2, len(obs[0])))
            obs_rec[s, t, 0], obs_rec[s, t, 1] = obs[0], obs[1]
            rec['qpos1'][s, t] = sim.data.qpos; rec['qvel1'][s, t] = sim.data.qvel
            rec['rew'][s, t] = rew; rec['done'][s, t] = done; rec['ncon'][s, t] = sim.data.ncon
            rec['cfrc_ext'][s, t] = sim.data.cfrc_ext; rec['qacc_warmstart'][s, t] = sim.data.qacc_warmstart
            rec['valid'][s, t] = True
            if any(done):
                break
    out.update(rec); out['obs'] = obs_rec
    if do_time:
        env.reset()
        rng = np.random.RandomState(0)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 10.0:
            _, _, done, _ = env.step((rng.randn(A), rng.randn(A)))
            n += 1
            if any(done):
                env.reset()
        out['env_steps_per_s_single_process'] = n / (time.perf_counter() - t0)
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--env', default='RoboSumo-Ant-vs-Ant-v0')
    ap.add_argument('--seeds', type=int, default=8)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--out', required=True)
    ap.add_argument('--time', action='store_true')
    args = ap.parse_args()
    np.savez_compressed(args.out, **dump(args.env, args.seeds, args.steps, args.time))
    print('wrote', args.out)
