"""Generates tests/golden/runner_*.npz by running the reference's OWN runner.py (imported from /root/reference under two
test-only shims: a matplotlib stub because of runner.py:4 and `np.bool = bool` because of runner.py:159) on fake env / models
that replay pre-generated tensors.  Build-container only.

    python tests/golden/make_runner_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference_runner():
    for name in ('matplotlib', 'matplotlib.pyplot', 'tqdm'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.tqdm = lambda x, **k: x
            sys.modules[name] = m
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if not hasattr(np, 'bool'):
        np.bool = bool
    sys.path.insert(0, '/root/reference')
    import runner as ref_runner
    sys.path.pop(0)
    return ref_runner


class Tape:
    """Pre-generated rollout: everything the runner consumes, indexed by step."""

    def __init__(self, seed, T, E, D=121, A=8, done_p=0.05):
        r = np.random.RandomState(seed)
        self.T, self.E, self.D, self.A = T, E, D, A
        self.obs = r.randn(T + 1, E, 2, D)
        self.actions = r.randn(T, 2, E, A).astype(np.float32)
        self.values = r.randn(T + 1, 2, E).astype(np.float32) * 3          # V0 on agent obs (index T = bootstrap)
        self.nlp_own = (8 + r.randn(T, 2, E)).astype(np.float32)            # model[agt].step neglogp
        self.nlp_cross = (8 + r.randn(T, 2, E)).astype(np.float32)          # the other model's neglogp of the same action
        self.shaping = r.randn(T, E, 2) * 2
        self.main = np.where(r.rand(T, E, 2) < 0.03, 2000.0, 0.0) * np.sign(r.randn(T, E, 2))
        d0 = r.rand(T, E) < done_p
        self.dones = np.stack([d0, d0 | (r.rand(T, E) < 0.01)], axis=-1)
        self.rew = r.randn(T, E, 2)


class FakeEnv:
    def __init__(self, tape):
        self.tape, self.t = tape, 0
        self.num_envs = tape.E
        sp = types.SimpleNamespace(shape=(tape.D,))
        self.observation_space = (sp, sp)

    def reset(self):
        return self.tape.obs[0]

    def step(self, actions):
        tp, t = self.tape, self.t
        infos = tuple(tuple({'shaping_reward': tp.shaping[t, e, a], 'main_reward': tp.main[t, e, a]} for a in range(2)) for e in range(tp.E))
        for e in range(tp.E):
            if tp.dones[t, e, 0]:
                infos[e][0]['episode'] = {'r': float(t), 'l': e}
        self.t += 1
        return tp.obs[t + 1], tp.rew[t], tp.dones[t], infos


class FakeModel:
    """model index k; replays values / neglogps by (step, agent)."""

    def __init__(self, tape, k, clock):
        self.tape, self.k, self.clock = tape, k, clock
        self.initial_state = None
        self.train_model = types.SimpleNamespace(X=types.SimpleNamespace(dtype=types.SimpleNamespace(name='float32')))
        self.act_model = self

    def _agent_of(self, obs):
        t = min(self.clock.t, self.tape.T)
        for a in range(2):
            if np.array_equal(obs, self.tape.obs[t][:, a, :].astype(np.float32)):
                return a
        raise AssertionError("unknown obs")

    def step(self, obs, S=None, M=None):
        a, t = self._agent_of(obs), self.clock.t
        return self.tape.actions[t, a], self.tape.values[t, a], None, self.tape.nlp_own[t, a]

    def value(self, obs, S=None, M=None):
        a, t = self._agent_of(obs), self.clock.t
        return self.tape.values[t, a]

    def action_probability(self, obs, given_action=None):
        a, t = self._agent_of(obs), self.clock.t
        return self.tape.nlp_cross[t, a]


def run_case(ref_runner, seed, T, E, update, anneal_bound, gamma, lam, rho_bar, c_bar):
    tape = Tape(seed, T, E)
    env = FakeEnv(tape)
    models = [FakeModel(tape, 0, env), FakeModel(tape, 1, env)]
    R = ref_runner.Runner(env=env, models=models, nsteps=T, nagent=2, gamma=gamma, lam=lam, rho_bar=rho_bar, c_bar=c_bar,
                          anneal_bound=anneal_bound)
    out = R.run(update)
    names = ['obs', 'returns', 'dones', 'actions', 'values', 'neglogpacs', 'rewards', 'opp_neglogpacs', 'opponent_obs',
             'opponent_actions']
    d = {n: np.asarray(o) for n, o in zip(names, out[:10])}
    d['off_policy_ratio'], d['off_env_ratio'], d['ratio'] = out[12], out[13], out[14]
    d['n_epinfos'] = np.array(len(out[11]))
    # inputs, in the layout our kernels take ([2][T][E])
    d['in_shaping'] = tape.shaping.transpose(2, 0, 1); d['in_main'] = tape.main.transpose(2, 0, 1)
    d['in_values'] = tape.values[:T].transpose(1, 0, 2); d['in_last_values'] = tape.values[T]
    d['in_nlp'] = np.stack([tape.nlp_own[:, 0], tape.nlp_cross[:, 1]])          # mb_neglogpacs: agt0 own, agt1 = model0 on a1
    d['in_opp_nlp'] = np.stack([tape.nlp_cross[:, 0], tape.nlp_own[:, 1]])      # mb_opponent_neglogpacs
    prev = np.concatenate([np.zeros((1, E, 2), bool), tape.dones[:-1]])         # mb_dones[t] = dones before step t
    d['in_dones'] = prev.transpose(2, 0, 1); d['in_last_dones'] = tape.dones[-1]
    d['in_obs'] = tape.obs[:T].transpose(2, 0, 1, 3).astype(np.float32); d['in_actions'] = tape.actions.transpose(1, 0, 2, 3)
    d['params'] = np.array([update, anneal_bound, gamma, lam, rho_bar, c_bar], dtype=np.float64)
    return d


if __name__ == '__main__':
    rr = import_reference_runner()
    cases = {'a': (0, 16, 5, 1, 1000, 0.995, 1.0, 10.0, 1.0),       # defaults.py values, update 1 (alpha = 1)
             'b': (1, 24, 7, 400, 1000, 0.995, 0.95, 1.5, 1.0),     # mid-anneal, lam < 1, active rho clip
             'c': (2, 8, 3, 1500, 1000, 0.99, 1.0, 10.0, 0.5)}      # past the anneal bound (alpha = 0)
    for k, c in cases.items():
        d = run_case(rr, *c)
        np.savez_compressed(os.path.join(HERE, 'runner_%s.npz' % k), **d)
        print(k, {n: v.shape for n, v in d.items() if n in ('obs', 'returns', 'ratio', 'opponent_obs')})
