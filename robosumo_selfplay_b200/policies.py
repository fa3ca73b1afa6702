"""policies.py MLP semantics on a flat device-resident parameter vector.

Mirrors the reference's `build_policy` / `PolicyWithValue` for `network='mlp', value_network='copy', num_hidden=64,
num_layers=2, activation=relu` (policies.py:14-193, defaults.py:8-26): two separate 2x64 ReLU trunks, a linear mean head
(init_scale 0.01), a state-independent logstd, a linear value head.  The arithmetic runs in the CUDA kernels of
csrc/rs_learn.cuh (rs_mlp_forward / rs_neglogp); this file holds the parameter layout, the reference's initialisation
(ortho_init driven by NumPy's legacy global RandomState, in variable-creation order) and the step/value/
action_probability surface.
"""
import ctypes

import numpy as np

from . import _lib

H = 64
LOG2PI_HALF = 0.9189385332046727


def param_shapes(D, A):
    """Order of tf.trainable_variables(scope) = the joblib checkpoint order (model.py:153-161; verified on model.ckpt)."""
    return [(D, H), (H,), (H, H), (H,), (D, H), (H,), (H, H), (H,), (H, A), (A,), (1, A), (H, 1), (1,)]


def param_count(D, A):
    return int(sum(int(np.prod(s)) for s in param_shapes(D, A)))


def _ortho(shape, scale):
    # baselines/baselines/a2c/utils.py:20-35 -- consumes np.random (legacy global RandomState) exactly like the reference
    a = np.random.normal(0.0, 1.0, shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == shape else v
    return (scale * q[:shape[0], :shape[1]]).astype(np.float32)


def init_params(D, A):
    """Fresh parameters with the reference's initialisers and its np.random draw order (six ortho draws per model)."""
    pi0 = _ortho((D, H), np.sqrt(2)); pi1 = _ortho((H, H), np.sqrt(2))
    vf0 = _ortho((D, H), np.sqrt(2)); vf1 = _ortho((H, H), np.sqrt(2))
    pih = _ortho((H, A), 0.01); vfh = _ortho((H, 1), 1.0)
    z = lambda *s: np.zeros(s, dtype=np.float32)
    return flatten_params([pi0, z(H), pi1, z(H), vf0, z(H), vf1, z(H), pih, z(A), z(1, A), vfh, z(1)])


def flatten_params(arrs):
    return np.concatenate([np.asarray(a, dtype=np.float32).ravel() for a in arrs])


def unflatten_params(flat, D, A):
    out, o = [], 0
    for shp in param_shapes(D, A):
        n = int(np.prod(shp))
        out.append(np.asarray(flat[o:o + n], dtype=np.float32).reshape(shp).copy())
        o += n
    return out


def logstd_offset(D, A):
    return 2 * (D * H + H + H * H + H) + H * A + A


class PolicyWithValue:
    """step / value / action_probability over a flat parameter tensor (torch, CUDA, float32 [P])."""

    def __init__(self, params, ob_dim, ac_dim, seed=0, precision='tf32'):
        import torch
        self.torch = torch
        self.params = params
        self.D, self.A = ob_dim, ac_dim
        self.device = params.device
        self._L = _lib.lib()
        self._tick = 0
        self._seed = seed
        self.initial_state = None
        self._ls0 = logstd_offset(ob_dim, ac_dim)
        self.precision = precision                  # 'tf32': tcgen05 tensor cores; 'fp32': FP32 pipe (numerics reference)

    # ---- device primitives --------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x):
        t = self.torch
        if t.is_tensor(x):
            return x.to(device=self.device, dtype=t.float32)
        return t.as_tensor(np.asarray(x, dtype=np.float32), device=self.device)

    def logstd(self):
        return self.params[self._ls0:self._ls0 + self.A]

    def forward(self, obs, want_mean=True, want_value=True):
        """obs [n, D] (rows may be strided).  Returns (mean [n, A] | None, value [n] | None) on the device."""
        t = self.torch
        obs = self._dev(obs)
        assert obs.dim() == 2 and obs.shape[1] == self.D and obs.stride(1) == 1
        n = obs.shape[0]
        mean = t.empty((n, self.A), dtype=t.float32, device=self.device) if want_mean else None
        value = t.empty((n,), dtype=t.float32, device=self.device) if want_value else None
        _lib.check(self._L.rs_mlp_forward(ctypes.c_void_p(self.params.data_ptr()), self.D, self.A, ctypes.c_void_p(obs.data_ptr()),
                                          obs.stride(0), n, ctypes.c_void_p(mean.data_ptr()) if want_mean else None,
                                          ctypes.c_void_p(value.data_ptr()) if want_value else None,
                                          1 if self.precision == 'tf32' else 0, self._stream()))
        return mean, value

    def neglogp_of(self, actions, mean):
        t = self.torch
        actions = self._dev(actions).contiguous()
        n = actions.shape[0]
        out = t.empty((n,), dtype=t.float32, device=self.device)
        ls = self.logstd().contiguous()
        _lib.check(self._L.rs_neglogp(n, self.A, ctypes.c_void_p(actions.data_ptr()), ctypes.c_void_p(mean.data_ptr()),
                                      ctypes.c_void_p(ls.data_ptr()), ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    # ---- reference surface (policies.py:84-128) -------------------------------------------
    def _out(self, x, like):
        return x if self.torch.is_tensor(like) else x.cpu().numpy()

    def step(self, observation, deterministic=False, **_):
        t = self.torch
        mean, value = self.forward(observation)
        if deterministic:
            act = mean.clone()
        else:
            g = t.Generator(device=self.device); g.manual_seed(self._seed * 1000003 + self._tick)
            self._tick += 1
            act = mean + t.exp(self.logstd()) * t.randn(mean.shape, generator=g, device=self.device, dtype=t.float32)
        nlp = self.neglogp_of(act, mean)
        return self._out(act, observation), self._out(value, observation), None, self._out(nlp, observation)

    def value(self, ob, *args, **kwargs):
        _, v = self.forward(ob, want_mean=False)
        return self._out(v, ob)

    def action_probability(self, observation, given_action=None, **_):
        mean, _ = self.forward(observation, want_value=False)
        return self._out(self.neglogp_of(given_action, mean), observation)

    def value_and_neglogp(self, observation, given_action=None, **_):
        mean, v = self.forward(observation)
        return self._out(v, observation), self._out(self.neglogp_of(given_action, mean), observation)

    def entropy(self):
        return float((self.logstd().double() + 0.5 * np.log(2.0 * np.pi * np.e)).sum().item())
