"""TEST / BENCH INFRASTRUCTURE ONLY (see oracle/README header rules): the reference's process-per-env execution model around
the oracle env, used by `bench.py --impl reference` to time the CPU path the way the reference runs it.

Protocol restated from subproc_vec_env.py:6-116 (not copied): one spawned daemon process per env, a duplex pipe per worker,
pickled ('step', action) / ('reset', None) / ('close', None) messages, the worker auto-resets when agent 0 reports done
(subproc_vec_env.py:12-16) and the parent gathers the replies in env order (subproc_vec_env.py:71-76)."""
import multiprocessing as mp

import numpy as np


def _serve(conn, parent_conn, pair, seed):
    parent_conn.close()
    from oracle.env_oracle import OracleSumoCore, OracleSumoWrap
    from oracle.physics import load_model_json
    env = OracleSumoWrap(OracleSumoCore(load_model_json(pair), seed=seed))
    try:
        while True:
            cmd, payload = conn.recv()
            if cmd == 'step':
                ob, rew, done, info = env.step(payload)
                if done[0]:
                    ob = env.reset()
                conn.send((np.stack(ob), np.asarray(rew), np.asarray(done), info))
            elif cmd == 'reset':
                conn.send(np.stack(env.reset()))
            elif cmd == 'close':
                break
            else:
                raise NotImplementedError(cmd)
    finally:
        conn.close()


class SubprocOracleVecEnv:
    def __init__(self, pair='ant_ant', num_envs=8, seed=0):
        ctx = mp.get_context('spawn')
        self.num_envs = num_envs
        pipes = [ctx.Pipe() for _ in range(num_envs)]
        self.remotes = [p[0] for p in pipes]
        self.procs = []
        for i, (parent_end, child_end) in enumerate(pipes):
            p = ctx.Process(target=_serve, args=(child_end, parent_end, pair, seed + i), daemon=True)
            p.start()
            child_end.close()
            self.procs.append(p)
        self.closed = False

    def reset(self):
        for r in self.remotes:
            r.send(('reset', None))
        return np.stack([r.recv() for r in self.remotes])

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(('step', a))
        res = [r.recv() for r in self.remotes]
        obs, rews, dones, infos = zip(*res)
        return np.stack(obs), np.stack(rews), np.stack(dones), infos

    def close(self):
        if self.closed:
            return
        for r in self.remotes:
            r.send(('close', None))
        for p in self.procs:
            p.join(timeout=5)
        self.closed = True
