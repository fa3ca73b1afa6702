"""Runner -- drop-in for the reference's runner.py:27-252 with the whole rollout resident on the GPU.

Per env step the reference does 5 `sess.run` calls and 2*nenv pickled pipe messages (runner.py:62-104,
subproc_vec_env.py:65-76); here one step is 4 launches of the MLP kernel, 1 sampling/neglogp kernel and 1 physics kernel,
all stream-ordered with no host synchronisation.  The trajectory is stored as [agent][t][env] (time-major, env contiguous),
which is what the V-trace kernel scans and what `sf01` (runner.py:255-260) flattens to env-major `e * T + t`.
"""
import ctypes
import time

import numpy as np

from . import _lib


def sf01(arr):
    """swap and then flatten axes 1 and 2 of [agent][t][env][...] (runner.py:255-260); works on torch and numpy."""
    s = arr.shape
    return arr.swapaxes(1, 2).reshape(s[0], s[1] * s[2], *s[3:])


def sf0(arr):
    return arr.swapaxes(0, 1).reshape(-1)


class Runner:
    def __init__(self, *, env, models, nsteps, nagent=2, gamma, lam, rho_bar, c_bar, anneal_bound=500, seed=0):
        import torch
        self.torch = torch
        assert nagent == 2
        self.env, self.models = env, models
        self.nenv = env.num_envs
        self.nagent = nagent
        self.nsteps = nsteps
        self.gamma, self.lam, self.rho_bar, self.c_bar = gamma, lam, rho_bar, c_bar
        self.anneal_bound = anneal_bound
        self.device = models[0].device
        self.device_env = bool(getattr(env, 'device_api', False))
        self.D, self.A = models[0].D, models[0].A
        self.obs = torch.zeros((self.nenv, 2, self.D), dtype=torch.float32, device=self.device)
        self.obs.copy_(self._to_dev(env.reset()))
        self.dones = torch.zeros((self.nenv, 2), dtype=torch.uint8, device=self.device)
        self.states = [None, None]
        self._L = _lib.lib()
        self._seed = seed
        self._tick = 0
        self.use_fused = True           # device env: run the step loop through rs_rollout (one library call); False = step by step
        self.tstart = time.time()

    def _to_dev(self, x, dtype=None):
        t = self.torch
        dtype = dtype or t.float32
        if t.is_tensor(x):
            return x.to(device=self.device, dtype=dtype)
        return t.as_tensor(np.asarray(x), device=self.device).to(dtype)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(x):
        return ctypes.c_void_p(x.data_ptr())

    def alpha(self, update):
        # exploration curriculum (runner.py:127-131)
        if update <= self.anneal_bound:
            return float(np.linspace(1, 0, self.anneal_bound)[update - 1])
        return 0.0

    def run(self, update, as_numpy=True, deterministic=False):
        t = self.torch
        T, E, D, A = self.nsteps, self.nenv, self.D, self.A
        dev = self.device
        f32 = dict(dtype=t.float32, device=dev)
        mb_obs = t.empty((2, T, E, D), **f32); mb_actions = t.empty((T, E, 2, A), **f32)
        mb_values = t.empty((2, T, E), **f32); mb_nlp = t.empty((2, T, E), **f32); mb_opp_nlp = t.empty((2, T, E), **f32)
        mb_dones = t.empty((2, T, E), dtype=t.uint8, device=dev)
        mb_shaping = t.empty((2, T, E), **f32); mb_main = t.empty((2, T, E), **f32)
        ep_done = t.empty((T, E), dtype=t.uint8, device=dev); ep_info = t.empty((T, E, 3), **f32)
        m0, m1 = self.models[0].act_model, self.models[1].act_model
        generic = bool(getattr(self.models[1], 'generic', False))      # opponent_mode='fix': a policy_zoo MLP behind the step / action_probability surface
        host_epinfos = []
        if generic:
            assert self.device_env, "a generic opponent needs the device-style env"
            for step in range(T):
                o0, o1 = self.obs[:, 0, :], self.obs[:, 1, :]
                a0, v0, _, nlp0 = m0.step(o0, deterministic=deterministic)                       # runner.py:67
                opp0 = m1.action_probability(o0, given_action=a0)                                # runner.py:85
                a1, _, _, opp1 = m1.step(o1, deterministic=deterministic)                        # runner.py:67 (agent 1)
                v01, nlp1 = m0.value_and_neglogp(o1, given_action=a1)                            # runner.py:89-90
                mb_obs[0, step].copy_(o0); mb_obs[1, step].copy_(o1)
                mb_dones[:, step].copy_(self.dones.t())
                mb_values[0, step].copy_(v0); mb_values[1, step].copy_(v01)
                mb_nlp[0, step].copy_(nlp0); mb_nlp[1, step].copy_(nlp1)
                mb_opp_nlp[0, step].copy_(opp0); mb_opp_nlp[1, step].copy_(opp1)
                mb_actions[step, :, 0].copy_(a0); mb_actions[step, :, 1].copy_(a1)
                obs, rew, done, (info, epi) = self.env.step(mb_actions[step])
                self.obs.copy_(obs); self.dones.copy_(done)
                mb_shaping[:, step].copy_(info[:, :, 6].t()); mb_main[:, step].copy_(info[:, :, 3].t())
                ep_done[step].copy_(done[:, 0]); ep_info[step].copy_(epi)
            return self._finish(update, as_numpy, m0, mb_obs, mb_actions, mb_values, mb_nlp, mb_opp_nlp, mb_dones, mb_shaping, mb_main, ep_done,
                                ep_info, host_epinfos)
        ls0, ls1 = m0.logstd().contiguous(), m1.logstd().contiguous()
        mu00 = t.empty((E, A), **f32); mu10 = t.empty((E, A), **f32); mu11 = t.empty((E, A), **f32); mu01 = t.empty((E, A), **f32)
        v00 = t.empty((E,), **f32); v01 = t.empty((E,), **f32)
        jobs = (_lib.rs_mlp_job * 4)()
        stride = self.obs.stride(0)
        for k, (mm, mu, vv) in enumerate(((m0, mu00, v00), (m1, mu10, None), (m1, mu11, None), (m0, mu01, v01))):
            jobs[k].params = mm.params.data_ptr(); jobs[k].obs_row_stride = stride
            jobs[k].mean = mu.data_ptr(); jobs[k].value = vv.data_ptr() if vv is not None else None
        prec = 1 if m0.precision == 'tf32' else 0
        fused = self.use_fused and self.device_env and getattr(self.env, 'auto_reset', False) and not getattr(self.env, 'mixed', False)
        if fused:
            # the whole step loop in ONE library call (rs_rollout): per step the four policy evaluations, the trajectory writes, the
            # action sampling, the physics step and the reward / episode records -- 5 launches, no host work between them.  The
            # env's own output buffers carry the observation and done flags from step to step.
            env = self.env
            env.d_obs.copy_(self.obs.reshape(env.d_obs.shape)); env.d_done.copy_(self.dones)
            io = _lib.rs_rollout_io()
            scratch = t.empty((4, E, A), **f32)
            for k, v in (('params0', m0.params), ('params1', m1.params), ('obs', env.d_obs), ('rew', env.d_rew), ('done', env.d_done), ('info', env.d_info),
                         ('episode', env.d_epi), ('mb_obs', mb_obs), ('mb_actions', mb_actions), ('mb_values', mb_values), ('mb_nlp', mb_nlp),
                         ('mb_opp_nlp', mb_opp_nlp), ('mb_dones', mb_dones), ('mb_shaping', mb_shaping), ('mb_main', mb_main), ('ep_done', ep_done),
                         ('ep_info', ep_info), ('scratch', scratch)):
                assert v.is_contiguous()
                setattr(io, k, v.data_ptr())
            _lib.check(self._L.rs_rollout(env._h, T, ctypes.byref(io), prec, self._seed, self._tick, 1 if deterministic else 0, self._stream()))
            self._tick += T
            self.obs.copy_(env.d_obs.reshape(self.obs.shape)); self.dones.copy_(env.d_done)
        for step in (range(T) if not fused else ()):
            o0, o1 = self.obs[:, 0, :], self.obs[:, 1, :]
            # the four policy evaluations of runner.py:67-90 in ONE launch (blockIdx.y = job):
            #   models[0].step(obs[:,0]) | models[1].action_probability(obs[:,0], a0) | models[1].step(obs[:,1]) |
            #   models[0].value + action_probability(obs[:,1], a1)
            jobs[0].obs = jobs[1].obs = o0.data_ptr(); jobs[2].obs = jobs[3].obs = o1.data_ptr()
            _lib.check(self._L.rs_mlp_forward_multi(jobs, 4, D, A, E, prec, self._stream()))
            mb_obs[0, step].copy_(o0); mb_obs[1, step].copy_(o1)
            mb_dones[:, step].copy_(self.dones.t())
            mb_values[0, step].copy_(v00); mb_values[1, step].copy_(v01)
            act = mb_actions[step]
            _lib.check(self._L.rs_rollout_sample(E, A, self._p(ls0), self._p(ls1), self._p(mu00), self._p(mu10), self._p(mu11), self._p(mu01),
                                                 self._seed, self._tick, 1 if deterministic else 0, self._p(act), self._p(mb_nlp[0, step]),
                                                 self._p(mb_nlp[1, step]), self._p(mb_opp_nlp[0, step]), self._p(mb_opp_nlp[1, step]),
                                                 self._stream()))
            self._tick += 1
            if self.device_env:
                obs, rew, done, (info, epi) = self.env.step(act)
                self.obs.copy_(obs); self.dones.copy_(done)
                mb_shaping[:, step].copy_(info[:, :, 6].t()); mb_main[:, step].copy_(info[:, :, 3].t())
                ep_done[step].copy_(done[:, 0]); ep_info[step].copy_(epi)
            else:   # any gym-style VecEnv with the reference's numpy surface
                obs, rew, done, infos = self.env.step(act.cpu().numpy())
                self.obs.copy_(self._to_dev(obs)); self.dones.copy_(self._to_dev(np.asarray(done), t.uint8))
                if hasattr(infos, 'column'):          # LazyInfos of the host-style B200SumoVecEnv: whole columns, no per-env dicts
                    sh = infos.column('shaping_reward').T; mn = infos.column('main_reward').T
                    fin = infos.finished()
                elif 'shaping_reward' in infos[0][0]:
                    sh = np.array([[infos[e][a]['shaping_reward'] for e in range(E)] for a in range(2)])
                    mn = np.array([[infos[e][a]['main_reward'] for e in range(E)] for a in range(2)])
                    fin = range(E)
                else:
                    sh = np.asarray(rew, dtype=np.float64).T; mn = sh
                    fin = range(E)
                mb_shaping[:, step].copy_(self._to_dev(sh)); mb_main[:, step].copy_(self._to_dev(mn))
                ep_done[step].zero_()
                for e in fin:
                    ei = infos[e][0].get('episode')
                    if ei:
                        host_epinfos.append(ei)
        return self._finish(update, as_numpy, m0, mb_obs, mb_actions, mb_values, mb_nlp, mb_opp_nlp, mb_dones, mb_shaping, mb_main, ep_done,
                            ep_info, host_epinfos)

    def _finish(self, update, as_numpy, m0, mb_obs, mb_actions, mb_values, mb_nlp, mb_opp_nlp, mb_dones, mb_shaping, mb_main, ep_done, ep_info,
                host_epinfos):
        """Bootstrap values, reward curriculum + IS ratios + V-trace, episode records and the sf01 flattening (runner.py:127-252)."""
        t = self.torch
        T, E = self.nsteps, self.nenv
        last_values = t.stack([m0.forward(self.obs[:, a, :], want_mean=False)[1] for a in range(2)])     # runner.py:184
        out = self.postprocess(update, mb_shaping, mb_main, mb_values, mb_nlp, mb_opp_nlp, mb_dones, last_values, self.dones)
        rewards, returns, ratios = out
        if self.device_env:
            flags = ep_done.cpu().numpy().astype(bool); info_h = ep_info.cpu().numpy()
            tt = round(time.time() - self.tstart, 6)
            sel = info_h[flags].astype(np.float64)         # (step, env) order, as the reference appends them (runner.py:95-98)
            epinfos = [{'r': r, 'dr': dr, 'l': l, 't': tt}
                       for r, dr, l in zip(np.round(sel[:, 0], 6).tolist(), np.round(sel[:, 1], 6).tolist(), sel[:, 2].astype(np.int64).tolist())]
        else:
            epinfos = host_epinfos
        act_a = mb_actions.permute(2, 0, 1, 3)                      # [2][T][E][A]
        res = dict(obs=sf01(mb_obs), returns=sf01(returns), dones=sf01(mb_dones), actions=sf01(act_a), values=sf01(mb_values),
                   neglogpacs=sf01(mb_nlp), rewards=sf01(rewards), opponent_neglogpacs=sf01(mb_opp_nlp),
                   # quirk kept from the reference: sf01 is applied to the 3-D opponent arrays too (runner.py:251)
                   opponent_obs=mb_obs[1].transpose(1, 2).reshape(T, -1), opponent_actions=act_a[1].transpose(1, 2).reshape(T, -1),
                   states=None, epinfos=epinfos, off_policy_ratio=ratios[0].t().reshape(-1), off_env_ratio=ratios[1].t().reshape(-1),
                   ratio=ratios[2].t().reshape(-1))
        if not as_numpy:
            return res
        order = ['obs', 'returns', 'dones', 'actions', 'values', 'neglogpacs', 'rewards', 'opponent_neglogpacs', 'opponent_obs',
                 'opponent_actions', 'states', 'epinfos', 'off_policy_ratio', 'off_env_ratio', 'ratio']
        conv = lambda k, v: (v.cpu().numpy().astype(bool) if k == 'dones' else v.cpu().numpy()) if t.is_tensor(v) else v
        return tuple(conv(k, res[k]) for k in order)

    def postprocess(self, update, shaping, main, values, nlp, opp_nlp, dones, last_values, last_dones):
        """Reward curriculum, IS ratios and V-trace returns (runner.py:127-200) on [2][T][E] device tensors."""
        t = self.torch
        T, E = values.shape[1], values.shape[2]
        alpha = self.alpha(update)
        rewards = (alpha * shaping.double() + (1.0 - alpha) * main.double()).float().contiguous()      # float64 then float32 (runner.py:132-155)
        returns = t.empty_like(rewards)
        ratios = t.empty((3, T, E), dtype=t.float32, device=self.device)
        values = values.contiguous(); nlp = nlp.contiguous(); opp_nlp = opp_nlp.contiguous(); dones = dones.contiguous()
        last_values = last_values.contiguous(); last_dones = last_dones.contiguous()
        _lib.check(self._L.rs_vtrace(T, E, float(self.gamma), float(self.lam), float(self.rho_bar), float(self.c_bar), self._p(rewards),
                                     self._p(values), self._p(dones), self._p(nlp), self._p(opp_nlp), self._p(last_values), self._p(last_dones),
                                     self._p(returns), self._p(ratios), self._stream()))
        return rewards, returns, ratios
