"""PPOModel -- drop-in for the reference's model.py:9-213 (TF1 graph + session) on one B200.

The learner parameters, Adam moments and all minibatch arithmetic stay on the device; `train` runs the hand-written
kernels of csrc/rs_learn.cuh through the C ABI (rs_adv_moments -> rs_ppo_grad -> [NCCL all-reduce] -> rs_adam_step).
Checkpoints are the reference's: `joblib.dump(list_of_13_float32_arrays)` in tf.trainable_variables order
(model.py:153-177), so `/root/reference/model.ckpt`-style files load directly.
"""
import ctypes
import os
import types
import zlib

import numpy as np

from . import _lib
from .policies import PolicyWithValue, init_params, param_count, param_shapes, flatten_params, unflatten_params


class PPOModel:
    loss_names = ['policy_loss', 'value_loss', 'policy_entropy', 'approxkl', 'clipfrac']     # model.py:140

    def __init__(self, *, ob_dim, ac_dim, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, trainable=True, model_scope="",
                 device=0, max_minibatch=1 << 20, comm=None, policy=None, ob_space=None, ac_space=None, precision='tf32', **_ignored):
        import torch
        self.torch = torch
        self.D, self.A = ob_dim, ac_dim
        self.P = param_count(ob_dim, ac_dim)
        self.scope = model_scope
        self.device = torch.device('cuda', device) if not isinstance(device, torch.device) else device
        if self.device.index is not None:
            torch.cuda.set_device(self.device)      # one process per GPU: the library's launches go to the calling thread's current device
        self.ent_coef, self.vf_coef, self.max_grad_norm = float(ent_coef), float(vf_coef), max_grad_norm
        self.trainable = trainable
        self.comm = comm                        # robosumo_selfplay_b200.dist.Comm or None
        self._peer_h = False                    # rs_peer handle of the gradient all-reduce: False = not set up yet, None = not available
        self._L = _lib.lib()
        assert self._L.rs_param_count(ob_dim, ac_dim) == self.P
        self.params = torch.as_tensor(init_params(ob_dim, ac_dim), device=self.device)       # consumes np.random like ortho_init
        self.precision = precision
        self.act_model = PolicyWithValue(self.params, ob_dim, ac_dim, seed=zlib.crc32(model_scope.encode()) % 9973, precision=precision)
        self.train_model = types.SimpleNamespace(X=types.SimpleNamespace(dtype=types.SimpleNamespace(name='float32')))
        self.step = self.act_model.step
        self.value = self.act_model.value
        self.initial_state = None
        if trainable:
            self.m = torch.zeros(self.P, dtype=torch.float32, device=self.device)
            self.v = torch.zeros(self.P, dtype=torch.float32, device=self.device)
            self.t = 0
            self.grad_stats = torch.zeros(self.P + 8, dtype=torch.float32, device=self.device)
            self.adv_sums = torch.zeros(2, dtype=torch.float64, device=self.device)
            self.gnorm = torch.zeros(1, dtype=torch.float32, device=self.device)
            self._ws = None
            self._ws_mb = 0

    # ---- helpers ---------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _p(self, x):
        return ctypes.c_void_p(x.data_ptr()) if x is not None else None

    def _workspace(self, n):
        if self._ws is None or n > self._ws_mb:
            self._ws_mb = max(n, 1024)
            self._ws = self.torch.empty(int(self._L.rs_ppo_workspace_floats(self.D, self.A, self._ws_mb)), dtype=self.torch.float32,
                                        device=self.device)
        return self._ws

    def _dev(self, x, dtype=None):
        t = self.torch
        dtype = dtype or t.float32
        if t.is_tensor(x):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return t.as_tensor(np.ascontiguousarray(x), device=self.device).to(dtype).contiguous()

    # ---- training ---------------------------------------------------------------------------
    def train_indexed(self, lr, cliprange, obs, returns, actions, values, neglogpacs, weights, idx, global_n=None, want_log_ratio=False,
                      adv_sums=None):
        """One minibatch step on device-resident flat sample arrays; the minibatch is `idx` (int32 device tensor of LOCAL
        sample indices).  With a communicator, every rank passes its local part of the global minibatch and `global_n`.
        `adv_sums` (device double[2]): the minibatch's GLOBAL advantage moments when the caller already has them
        (dist.EpochSchedule computes those of a whole epoch in one launch and all-reduces them once); otherwise they are computed
        (and all-reduced) here.  Launches per step: 2 (gradient tiles + reduction) + [NCCL all-reduce] + 1 (clip + Adam + stats)."""
        t = self.torch
        n = int(idx.numel()) if idx is not None else int(returns.numel())
        gn = int(global_n) if global_n is not None else n
        st = self._stream()
        L = self._L
        self.t += 1
        mgn = float(self.max_grad_norm) if self.max_grad_norm is not None else 0.0
        prec = 1 if self.precision == 'tf32' else 0
        log_ratio = t.empty(n, dtype=t.float32, device=self.device) if want_log_ratio else None
        stats = t.empty(5, dtype=t.float64, device=self.device)       # loss_names order; entropy with the pre-update logstd (model.py:70)
        ready = adv_sums is not None
        sums = adv_sums if ready else self.adv_sums
        if self.comm is None and gn == n:          # single GPU: the whole minibatch step is one library call
            _lib.check(L.rs_ppo_minibatch_step(self._p(self.params), self._p(self.m), self._p(self.v), self.D, self.A, self._p(obs),
                                               self._p(actions), self._p(returns), self._p(values), self._p(neglogpacs), self._p(weights),
                                               self._p(idx), n, float(cliprange), self.ent_coef, self.vf_coef, mgn, float(lr), self.t,
                                               self._p(self._workspace(n)), self._p(self.grad_stats), self._p(sums), 1 if ready else 0,
                                               self._p(self.gnorm), self._p(stats), self._p(log_ratio), prec, st))
            return stats, log_ratio
        if not ready:
            _lib.check(L.rs_adv_moments(self._p(idx), n, self._p(returns), self._p(values), self._p(sums), st))
            if self.comm is not None:
                self.comm.all_reduce_sum(sums)
        peer = self._peer()
        if peer is not None:                        # data-parallel on one node: the whole step is one library call (gradient all-reduce = rs_peer_allreduce,
                                                    # ONE kernel of ours over NVLink peer memory on this stream)
            _lib.check(L.rs_ppo_minibatch_step_peer(peer, self._p(self.params), self._p(self.m), self._p(self.v), self.D, self.A, self._p(obs),
                                                    self._p(actions), self._p(returns), self._p(values), self._p(neglogpacs), self._p(weights),
                                                    self._p(idx), n, gn, float(cliprange), self.ent_coef, self.vf_coef, mgn, float(lr), self.t,
                                                    self._p(self._workspace(n)), self._p(self.grad_stats), self._p(sums), self._p(self.gnorm),
                                                    self._p(stats), self._p(log_ratio), prec, st))
            return stats, log_ratio
        _lib.check(L.rs_ppo_grad(self._p(self.params), self.D, self.A, self._p(obs), self._p(actions), self._p(returns), self._p(values),
                                 self._p(neglogpacs), self._p(weights), self._p(idx), n, gn, self._p(sums), float(cliprange),
                                 self.ent_coef, self.vf_coef, self._p(self._workspace(n)), self._p(self.grad_stats), self._p(log_ratio),
                                 self._p(stats), prec, st))
        if self.comm is not None:                   # ranks on several nodes / gloo / RS_B200_PEER=0: torch.distributed (NCCL) all-reduce of the
            self.comm.all_reduce_sum(self.grad_stats)          # flat [grads | 4 stat sums] (98 KB, latency-bound)
        _lib.check(L.rs_adam_step(self._p(self.params), self._p(self.m), self._p(self.v), self._p(self.grad_stats), self.D, self.A,
                                  self.ent_coef, mgn, float(lr), self.t, 0.9, 0.999, 1e-5, self._p(self.gnorm), gn, self._p(stats), st))
        return stats, log_ratio

    def _peer(self):
        """The peer all-reduce handle of this model (created on first use: a collective, so every rank gets here together)."""
        if self.comm is None or self.comm.world == 1:
            return None
        if self._peer_h is False:
            self._peer_h = self.comm.make_peer(self.P + 4, self.device.index if self.device.index is not None else 0)
        return self._peer_h

    def check_peer(self):
        """Raises if a peer all-reduce of this model ever timed out waiting for a rank (synchronises the device)."""
        if self._peer_h not in (None, False) and self._L.rs_peer_error(self._peer_h) != 0:
            raise RuntimeError("rs_peer_allreduce: a rank did not deliver its gradient within the time-out; parameters are no longer in sync")

    def stats_to_list(self, stats_dev):
        return [float(x) for x in stats_dev.double().cpu().numpy()]

    def train(self, lr, cliprange, obs, returns, masks, actions, values, neglogpacs, rewards, IS_weight, states=None):
        """Reference signature (model.py:179-213): numpy minibatch in, [pg_loss, vf_loss, entropy, approxkl, clipfrac,
        log_ratio, summary] out."""
        obs_d = self._dev(obs); ret_d = self._dev(returns); act_d = self._dev(actions)
        val_d = self._dev(values); nlp_d = self._dev(neglogpacs); w_d = self._dev(IS_weight)
        stats, log_ratio = self.train_indexed(lr, cliprange, obs_d, ret_d, act_d, val_d, nlp_d, w_d, None, want_log_ratio=True)
        return self.stats_to_list(stats) + [log_ratio.cpu().numpy(), None]

    # ---- checkpoints (model.py:153-177) -----------------------------------------------------
    def get_params_list(self):
        return unflatten_params(self.params.cpu().numpy(), self.D, self.A)

    def save(self, save_path):
        import joblib
        dirname = os.path.dirname(save_path)
        if dirname:
            os.makedirs(dirname, exist_ok=True)
        joblib.dump(self.get_params_list(), save_path)

    def load(self, load_path):
        import joblib
        loaded = joblib.load(os.path.expanduser(load_path))
        self.load_list(loaded)

    def load_list(self, loaded):
        shapes = param_shapes(self.D, self.A)
        if isinstance(loaded, dict):
            loaded = [loaded[k] for k in sorted(loaded.keys())]
        assert len(loaded) == len(shapes), 'number of variables loaded mismatches len(variables)'
        for a, s in zip(loaded, shapes):
            assert tuple(np.shape(a)) == tuple(s), (np.shape(a), s)
        self.set_flat(flatten_params(loaded))

    def set_flat(self, flat):
        self.params.copy_(self.torch.as_tensor(np.asarray(flat, dtype=np.float32), device=self.device))

    def get_flat(self):
        return self.params.cpu().numpy().copy()
