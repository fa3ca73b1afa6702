"""Fixed pretrained opponents of the reference's policy_zoo (row N1 of SURVEY 8f).

`MLPPolicy` of robosumo/robosumo/policy_zoo/policy.py:23-86: observation filter `clip((o - mean) / std, -5, 5)` with
`std = sqrt(max(sumsq/count - mean^2, 1e-2))` (utils.py:8-30), two separate 64-64 tanh MLPs, linear heads, state-independent
logstd.  Parameters come as one flat vector in TF global-variable creation order (utils.py:70-82):
    retfilter sum, sumsq, count | obsfilter sum[D], sumsq[D], count | vffc1 w,b | vffc2 w,b | vffinal w,b |
    polfc1 w,b | polfc2 w,b | polfinal w,b | logstd[1,A]
The `.npy` files are reference assets (robosumo/robosumo/policy_zoo/assets/<morph>/mlp/agent-params-v{1,2,3}.npy) and are
loaded from the path the caller gives; the forward pass runs in the CUDA MLP kernel (tanh activation) of csrc/rs_learn.cuh.
"""
import ctypes

import numpy as np

from . import _lib
from .policies import flatten_params, H


def split_zoo_params(flat, D, A):
    flat = np.asarray(flat, dtype=np.float32).ravel()
    shapes = [('ret_sum', ()), ('ret_sumsq', ()), ('ret_count', ()), ('ob_sum', (D,)), ('ob_sumsq', (D,)), ('ob_count', ()),
              ('vf0w', (D, H)), ('vf0b', (H,)), ('vf1w', (H, H)), ('vf1b', (H,)), ('vfhw', (H, 1)), ('vfhb', (1,)),
              ('pi0w', (D, H)), ('pi0b', (H,)), ('pi1w', (H, H)), ('pi1b', (H,)), ('pihw', (H, A)), ('pihb', (A,)), ('logstd', (1, A))]
    out, o = {}, 0
    for name, shp in shapes:
        n = int(np.prod(shp)) if shp else 1
        out[name] = flat[o:o + n].reshape(shp) if shp else float(flat[o])
        o += n
    assert o == flat.size, "zoo parameter vector has %d entries, expected %d (obs %d, act %d)" % (flat.size, o, D, A)
    return out


class ZooMLPPolicy:
    def __init__(self, flat_params, ob_dim, ac_dim, device=0):
        import torch
        self.torch = torch
        self.D, self.A = ob_dim, ac_dim
        self.device = torch.device('cuda', device) if not isinstance(device, torch.device) else device
        z = split_zoo_params(flat_params, ob_dim, ac_dim)
        mean = z['ob_sum'] / z['ob_count']
        std = np.sqrt(np.maximum(z['ob_sumsq'] / z['ob_count'] - mean ** 2, 1e-2))
        rmean = z['ret_sum'] / z['ret_count']
        self.ret_mean = float(rmean)
        self.ret_std = float(np.sqrt(max(z['ret_sumsq'] / z['ret_count'] - rmean ** 2, 1e-2)))
        self.ob_mean = torch.as_tensor(mean.astype(np.float32), device=self.device)
        self.ob_std = torch.as_tensor(std.astype(np.float32), device=self.device)
        # same flat layout as PPOModel (policies.param_shapes) so that the same kernel evaluates it
        self.params = torch.as_tensor(flatten_params([z['pi0w'], z['pi0b'], z['pi1w'], z['pi1b'], z['vf0w'], z['vf0b'], z['vf1w'], z['vf1b'],
                                                      z['pihw'], z['pihb'], z['logstd'], z['vfhw'], z['vfhb']]), device=self.device)
        self.logstd = torch.as_tensor(z['logstd'].ravel().astype(np.float32), device=self.device)
        self._L = _lib.lib()

    @classmethod
    def load(cls, path, ob_dim, ac_dim, device=0):
        return cls(np.load(path), ob_dim, ac_dim, device)

    def mean_value(self, ob):
        """ob [n, D] device float32 -> (mean [n, A], un-normalised value prediction [n]) (policy.py:39-70)."""
        t = self.torch
        obz = t.clamp((ob - self.ob_mean) / self.ob_std, -5.0, 5.0).contiguous()
        n = obz.shape[0]
        mean = t.empty((n, self.A), dtype=t.float32, device=self.device)
        value = t.empty((n,), dtype=t.float32, device=self.device)
        job = (_lib.rs_mlp_job * 1)()
        job[0].params = self.params.data_ptr(); job[0].obs = obz.data_ptr(); job[0].obs_row_stride = obz.stride(0)
        job[0].mean = mean.data_ptr(); job[0].value = value.data_ptr(); job[0].activation = 1
        st = ctypes.c_void_p(t.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._L.rs_mlp_forward_multi(job, 1, self.D, self.A, n, 0, st))
        return mean, value * self.ret_std + self.ret_mean

    def act(self, observation, stochastic=False):
        """observation [n, D] (numpy or torch).  Returns (actions [n, A], {'vpred': [n]}) like policy.py:72-80."""
        t = self.torch
        is_t = t.is_tensor(observation)
        ob = observation.to(self.device, t.float32) if is_t else t.as_tensor(np.asarray(observation, dtype=np.float32), device=self.device)
        mean, vpred = self.mean_value(ob)
        act = mean + t.exp(self.logstd) * t.randn_like(mean) if stochastic else mean
        if is_t:
            return act, {'vpred': vpred}
        return act.cpu().numpy(), {'vpred': vpred.cpu().numpy()}


class _ZooActModel:
    """PolicyWithValue surface (policies.py:84-128) over a zoo MLP for the learner-sized (D + 1)-wide observation: the timestep
    feature this fork appends (sumo_env.py:68-70) is cut off, as eval_robosumo_against_fix.py:207 does."""

    def __init__(self, zoo, seed=0):
        self.zoo, self.torch = zoo, zoo.torch
        self.D, self.A, self.device = zoo.D + 1, zoo.A, zoo.device
        self.initial_state = None
        self._seed, self._tick = seed, 0
        self._L = zoo._L

    def _mean_value(self, observation):
        t = self.torch
        ob = observation.to(self.device, t.float32) if t.is_tensor(observation) else t.as_tensor(np.asarray(observation, dtype=np.float32), device=self.device)
        return self.zoo.mean_value(ob[:, :self.zoo.D])

    def _nlp(self, act, mean):
        t = self.torch
        act = act.contiguous(); mean = mean.contiguous()
        out = t.empty((act.shape[0],), dtype=t.float32, device=self.device)
        st = ctypes.c_void_p(t.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._L.rs_neglogp(act.shape[0], self.A, ctypes.c_void_p(act.data_ptr()), ctypes.c_void_p(mean.data_ptr()),
                                      ctypes.c_void_p(self.zoo.logstd.data_ptr()), ctypes.c_void_p(out.data_ptr()), st))
        return out

    def step(self, observation, deterministic=False, **_):
        t = self.torch
        mean, value = self._mean_value(observation)
        if deterministic:
            act = mean.clone()
        else:
            g = t.Generator(device=self.device); g.manual_seed(self._seed * 1000003 + self._tick)
            self._tick += 1
            act = mean + t.exp(self.zoo.logstd) * t.randn(mean.shape, generator=g, device=self.device, dtype=t.float32)
        return act, value, None, self._nlp(act, mean)

    def value(self, ob, *a, **k):
        return self._mean_value(ob)[1]

    def action_probability(self, observation, given_action=None, **_):
        t = self.torch
        act = given_action.to(self.device, t.float32) if t.is_tensor(given_action) else t.as_tensor(np.asarray(given_action, dtype=np.float32), device=self.device)
        return self._nlp(act, self._mean_value(observation)[0])


class ZooOpponentModel:
    """What `runner.models[1]` is under opponent_mode='fix' (alg_ppo.py:194-206): a non-trainable model whose act_model is the
    pretrained policy_zoo MLP.  `generic = True` makes Runner step it through the step / action_probability surface instead of
    the fused four-job MLP launch (different width, tanh units and an observation filter)."""
    generic = True
    trainable = False

    def __init__(self, zoo, seed=0):
        self.act_model = _ZooActModel(zoo, seed)
        self.D, self.A, self.device = self.act_model.D, zoo.A, zoo.device
        self.step, self.value, self.initial_state = self.act_model.step, self.act_model.value, None


def evaluate_against_fixed(env, model, opponent, rounds, max_steps=100000):
    """The evaluation loop of eval_robosumo_against_fix.py:198-229: deterministic learner (agent 0) against a fixed zoo opponent
    (agent 1, which sees its observation without the timestep feature); counts win / draw / lose over `rounds` finished episodes.
    `env` is a device-style B200SumoVecEnv (the reference builds it with `_adjust_z = -0.5`, i.e. adjust_z=-0.5)."""
    import torch
    obs = env.reset()
    win = draw = lose = done_rounds = steps = 0
    while done_rounds < rounds and steps < max_steps:
        o0, o1 = (obs[0], obs[1]) if env.mixed else (obs[:, 0, :], obs[:, 1, :])
        a0, _, _, _ = model.step(o0, deterministic=True)
        a1, _ = opponent.act(o1[:, :-1], stochastic=False)
        act = (a0, a1) if env.mixed else torch.stack([a0, a1], 1)
        obs, rew, done, (info, epi) = env.step(act)
        fin = done[:, 0].bool()
        if bool(fin.any()):
            w0 = (info[:, 0, 7].int() & 1).bool() & fin
            w1 = (info[:, 1, 7].int() & 1).bool() & fin & ~w0
            win += int(w0.sum()); lose += int(w1.sum()); draw += int((fin & ~w0 & ~w1).sum())
            done_rounds += int(fin.sum())
        steps += 1
    n = max(done_rounds, 1)
    return dict(rounds=done_rounds, win=win / n, draw=draw / n, lose=lose / n, steps=steps)
