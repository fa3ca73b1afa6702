#!/usr/bin/env python
"""bench.py -- env-steps/s (physics + obs + reward + done + auto-reset) of the batched sumo arena.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E_PER_GPU]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one VecEnv.step over E Ant-vs-Ant env pairs per GPU (BASELINE.json configs[1], E = 4096):
5 substeps x RK4 = 20 forward-dynamics evaluations per pair, fused obs / reward / done / auto-reset.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_PAIR_STEP = 1530          # Ant-vs-Ant, fp32 storage
ENV_ID = 'RoboSumo-Ant-vs-Ant-v0'
METRIC = 'env-steps/sec (physics+obs+reward)'


def read_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  Uses NVML in-process
    (a sample every 20 ms; `nvidia-smi` takes longer to answer than a short timed region lasts) and falls back to the
    `nvidia-smi --query-gpu` recipe when the NVML bindings are missing."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self.stop_flag = False
        self.nvml = None
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception as ex:
            self.nvml = None
            self.err = repr(ex)[:160]

    def _one_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append((mhz, self.max_mhz, bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                             bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)))

    def _one_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(',')]
        if len(parts) >= 6:
            self.samples.append((float(parts[0]), float(parts[1])) + tuple(p.lower().startswith('active') for p in parts[2:6]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._one_nvml()
                else:
                    self._one_smi()
            except Exception as ex:
                self.err = repr(ex)[:160]
            time.sleep(0.02 if self.nvml is not None else 0.1)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable'], 'error': self.err}
        sm = sorted(s[0] for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons, 'samples': len(sm),
                'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


CPU_FLAGS = None


def cpu_port_rate(seconds, envs, threads):
    """Oracle port (oracle/physics_oracle.c) stepping `envs` independent pairs with `threads` host threads for
    about `seconds`; the env-level reward/obs arithmetic is negligible beside the physics and is not included.
    The library is compiled on THIS machine with -O3 -march=native before it is timed (oracle/Makefile `native`)."""
    global CPU_FLAGS
    import numpy as np
    from oracle import physics as op
    from oracle.physics import OracleModel, load_model_json
    if CPU_FLAGS is None:
        try:
            CPU_FLAGS = op.use_native_build()
        except AssertionError:                     # already loaded (smoke / tests in the same process): report what is loaded
            CPU_FLAGS = op.NATIVE_FLAGS if op._NATIVE else op.PORTABLE_FLAGS
    om = OracleModel(load_model_json('ant_ant'))
    rng = np.random.RandomState(0)
    q = np.tile(om.qpos0, (envs, 1)); v = np.zeros((envs, om.nv)); w = np.zeros((envs, om.nv))
    phi = rng.uniform(0, 2 * np.pi, envs)
    for a in range(2):
        q[:, 15 * a] = 1.15 * np.cos(phi + a * np.pi); q[:, 15 * a + 1] = 1.15 * np.sin(phi + a * np.pi); q[:, 15 * a + 2] = 1.25
    q += rng.uniform(-.1, .1, q.shape); v += 0.1 * rng.randn(*v.shape)
    for _ in range(3):                                       # warm-up: let the ants land
        om.step_batch(q, v, rng.randn(envs, om.nu), 5, w, threads)
    n, t0 = 0, time.perf_counter()
    while True:
        om.step_batch(q, v, rng.randn(envs, om.nu), 5, w, threads)
        n += envs
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return n / el, n, el


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path.  MuJoCo 2.1 / mujoco-py cannot be installed here (closed binary,
    no network), so this times the oracle port of the same physics on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    envs = 16 * cores
    per_step_s = 0.5          # bounded sample per step so that the default K / W finish within ~2 minutes
    vals = []
    for i in range(args.warmup + args.steps):
        rate, n, el = cpu_port_rate(per_step_s, envs, cores)
        if i >= args.warmup:
            vals.append((rate, n, el))
    tot_n = sum(v[1] for v in vals); tot_t = sum(v[2] for v in vals)
    value = tot_n / tot_t
    # the same CPU path driven the way the reference drives it: one spawned process per env, pickled pipe messages, Python env
    # code around the physics call (subproc_vec_env.py:6-116, sumo.py / agents.py restated in oracle/env_oracle.py)
    subproc = None
    try:
        import numpy as np
        from oracle.subproc_oracle import SubprocOracleVecEnv
        nw = min(cores, 16)
        venv = SubprocOracleVecEnv('ant_ant', nw, seed=1)
        venv.reset()
        rng = np.random.RandomState(0)
        for _ in range(5):
            venv.step(rng.randn(nw, 2, 8))
        n_sp, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 6.0:
            venv.step(rng.randn(nw, 2, 8)); n_sp += nw
        el = time.perf_counter() - t0
        venv.close()
        subproc = {'value': n_sp / el, 'unit': 'env-steps/s', 'workers': nw, 'envs': nw,
                   'sample': '%d env-steps in %.1f s, one process per env pair, pickled pipes' % (n_sp, el)}
    except Exception as ex:          # never let the secondary number take the reference arm down
        subproc = {'unavailable': repr(ex)[:200]}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / len(vals), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'RoboSumo-Ant-vs-Ant-v0 physics step, CPU oracle port on host cores',
                   'envs_per_step': envs, 'frame_skip': 5, 'integrator': 'RK4'},
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'compiler_flags': 'gcc ' + str(CPU_FLAGS),
                         'sample': '%d env pairs x ~%.1f s per step, %d steps, N(0,1) actions' % (envs, per_step_s, args.steps)},
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'subproc_protocol': subproc,
    }
    print(json.dumps(line), flush=True)


def measure_learner(args, E, local, rank, world, dev):
    """Secondary metrics of BASELINE.json: rollout steps/s (5 policy evaluations + physics per step, runner.py:62-104) and the
    PPO2 update s/iter (V-trace + noptepochs x nminibatches minibatch steps on E*T samples per GPU, all-reduce when N > 1)."""
    import numpy as np
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    from robosumo_selfplay_b200.model import PPOModel
    from robosumo_selfplay_b200.runner import Runner
    from robosumo_selfplay_b200.dist import Comm, EpochSchedule, EpochPermutations
    from robosumo_selfplay_b200 import _lib
    T, nmb, nep = args.nsteps, 32, 6
    comm = Comm() if world > 1 else None
    np.random.seed(0)
    env = B200SumoVecEnv(ENV_ID, num_envs=E, seed=99 + rank, device=local, device_api=True)
    models = [PPOModel(ob_dim=121, ac_dim=8, device=local, comm=comm), PPOModel(ob_dim=121, ac_dim=8, trainable=False, device=local)]
    if comm is not None:
        comm.broadcast(models[0].params, 0)
    runner = Runner(env=env, models=models, nsteps=T, gamma=0.995, lam=1.0, rho_bar=10.0, c_bar=1.0, anneal_bound=1000)
    runner.nsteps = 8
    runner.run(1, as_numpy=False)                       # warm-up
    runner.nsteps = T
    roll_times = []
    R = None
    for rep in range(3):                               # median of three rollouts: the first one pays the cudaMalloc of the ~1 GB of
        R = None                                       # trajectory buffers; afterwards the caching allocator reuses them, as in training
        torch.cuda.synchronize()
        l0 = _lib.lib().rs_launch_count()
        t0 = time.perf_counter()
        R = runner.run(1, as_numpy=False)
        torch.cuda.synchronize()
        roll_times.append(time.perf_counter() - t0)
        roll_launches = _lib.lib().rs_launch_count() - l0
    roll_s = sorted(roll_times)[1]
    data = {k: R[k][0].contiguous() for k in ('obs', 'returns', 'actions', 'values', 'neglogpacs')}
    N_local, N = E * T, E * T * world
    nbt = N // nmb
    lo, hi = rank * N_local, (rank + 1) * N_local
    model = models[0]
    sched = EpochSchedule(dev, N, nbt, lo, hi, comm)
    split = (nbt, lo, hi) if comm is not None else None      # data-parallel: the helper thread hands out this rank's cut of the permutation (dist.host_split)
    l0 = _lib.lib().rs_launch_count()
    def one_update(perms=None):
        # the reference's per-epoch np.random.shuffle, replayed bit-exactly on the host (one epoch ahead on a helper thread); the
        # permutation goes to the device once per epoch, where the minibatches are split by rank and their advantage moments computed
        for inds in (perms if perms is not None else EpochPermutations(N, nep, dtype=np.int32, split=split)):
            for mb, n_loc, gn, sums in sched.load(inds, data['returns'], data['values']):
                model.train_indexed(1e-3, 0.2, data['obs'], data['returns'], data['actions'], data['values'], data['neglogpacs'], None, mb,
                                    global_n=gn, adv_sums=sums)
    one_update()                                        # warm-up
    torch.cuda.synchronize()
    upd_launches = _lib.lib().rs_launch_count() - l0
    if comm is not None:
        comm.barrier()
    t0 = time.perf_counter()
    one_update()
    torch.cuda.synchronize()
    upd_s = time.perf_counter() - t0
    # the same update the way alg_ppo.learn runs it: the six permutations depend on the generator stream only and are drawn on the
    # helper thread WHILE the rollout runs, so inside the training loop the update does not wait for them
    perms = EpochPermutations(N, nep, ahead=nep, dtype=np.int32, split=split)
    runner.run(1, as_numpy=False)
    torch.cuda.synchronize()
    if comm is not None:
        comm.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    one_update(perms)
    e1.record()
    torch.cuda.synchronize()
    upd_loop_s = time.perf_counter() - t0
    upd_gpu_s = e0.elapsed_time(e1) / 1e3
    model.check_peer()
    allreduce = 'none (1 GPU)' if comm is None else ('rs_peer_allreduce: one kernel over NVLink peer memory on the learner stream' if model._peer_h not in (None, False) else 'torch.distributed all_reduce (NCCL)')
    tt = torch.tensor([roll_s, upd_s, upd_loop_s, upd_gpu_s], dtype=torch.float64, device=dev)
    if comm is not None:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)          # max over ranks
    roll_s, upd_s, upd_loop_s, upd_gpu_s = [float(x) for x in tt]
    flops = 114.6e3 * N * nep
    env.close()
    return {'rollout': {'value': E * world * T / roll_s, 'unit': 'env-steps/s', 'T': T, 'launches_per_step': roll_launches / T,
                        'rollout_seconds': roll_times, 'note': 'Runner.run -> rs_rollout: T steps behind one library call; per step 1 MLP launch (4 policy evaluations, tcgen05 tf32), 1 sampling launch, 1 physics launch and 2 trajectory-write launches, no host work in between; median of 3 rollouts'},
            'ppo_update': {'value': upd_loop_s, 'unit': 's/iter', 'samples': N, 'nminibatches': nmb, 'noptepochs': nep, 'minibatch': nbt,
                           'higher_is_better': False, 'dtype': 'tf32 GEMMs (tcgen05) + f32', 'achieved_tflops': flops / upd_loop_s / 1e12,
                           'achieved_gbs': 536.0 * N * nep / upd_loop_s / 1e9, 'launches_per_minibatch': upd_launches / float(nmb * nep),
                           'gpu_seconds': upd_gpu_s, 'gradient_allreduce': allreduce,
                           'standalone': {'value': upd_s, 'unit': 's/iter', 'note': 'same update started cold: includes the host-side replay of the six legacy NumPy shuffles over all %d GLOBAL indices (bit-exact schedule, one epoch ahead of the GPU) on the critical path' % N},
                           'note': 'as alg_ppo.learn runs it (max over ranks): V-trace is part of the rollout; the six permutations are drawn by the helper thread during the preceding rollout; per epoch one H2D of the int32 permutation, one device-side split + advantage-moment launch (one all-reduce per epoch when N > 1), per minibatch 3 launches and ONE gradient all-reduce (N > 1: a 4th launch, the peer-memory all-reduce kernel)'}}


def measure_config1_ref_shape(local, dev):
    """BASELINE.json configs[0]: the reference's own shape -- 8 envs, nsteps = 2048 (learn()'s default) / 8192 (defaults.py:9),
    32 minibatches x 6 epochs (defaults.py:8-26).  Rollout env-steps/s and update s/iter on ONE GPU."""
    import numpy as np
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    from robosumo_selfplay_b200.model import PPOModel
    from robosumo_selfplay_b200.runner import Runner
    from robosumo_selfplay_b200.dist import EpochPermutations, EpochSchedule
    E, T, nmb, nep = 8, 2048, 32, 6
    np.random.seed(1)
    env = B200SumoVecEnv(ENV_ID, num_envs=E, seed=5, device=local, device_api=True)
    models = [PPOModel(ob_dim=121, ac_dim=8, device=local), PPOModel(ob_dim=121, ac_dim=8, trainable=False, device=local)]
    runner = Runner(env=env, models=models, nsteps=64, gamma=0.995, lam=1.0, rho_bar=10.0, c_bar=1.0, anneal_bound=1000)
    runner.run(1, as_numpy=False)
    runner.nsteps = T
    torch.cuda.synchronize(); t0 = time.perf_counter()
    R = runner.run(1, as_numpy=False)
    torch.cuda.synchronize(); roll_s = time.perf_counter() - t0
    out = {'envs': E, 'rollout': {'T': T, 'value': E * T / roll_s, 'unit': 'env-steps/s', 'seconds': roll_s,
                                  'note': '8 pairs occupy 8 warps of 8 SMs: the step is bound by the dependent-instruction latency of one pair (~1 ms), not by throughput'}}
    for Tu, rep in ((2048, 1), (8192, 4)):
        data = {k: R[k][0].repeat(*([rep] + [1] * (R[k][0].dim() - 1))).contiguous() for k in ('obs', 'returns', 'actions', 'values', 'neglogpacs')}
        N = E * Tu
        sched = EpochSchedule(dev, N, N // nmb)
        def upd():
            for inds in EpochPermutations(N, nep, dtype=np.int32):
                for mb, n_loc, gn, sums in sched.load(inds, data['returns'], data['values']):
                    models[0].train_indexed(1e-3, 0.2, data['obs'], data['returns'], data['actions'], data['values'], data['neglogpacs'], None, mb, global_n=gn, adv_sums=sums)
        upd(); torch.cuda.synchronize(); t0 = time.perf_counter(); upd(); torch.cuda.synchronize()
        out['ppo_update_T%d' % Tu] = {'value': time.perf_counter() - t0, 'unit': 's/iter', 'samples': N, 'minibatch': N // nmb, 'nminibatches': nmb, 'noptepochs': nep,
                                      'data': 'rollout of T=2048' + ('' if rep == 1 else ' tiled x%d to the T=8192 sample count' % rep)}
    env.close()
    return out


def measure_config3_morphologies(local, dev, E):
    """BASELINE.json configs[2]: Bug-vs-Bug and Spider-vs-Spider (more legs / contacts, intra-agent collisions), E pairs."""
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    out = {}
    for name, A, bytes_ in (('Bug', 12, 2042), ('Spider', 16, 2554)):
        env = B200SumoVecEnv('RoboSumo-%s-vs-%s-v0' % (name, name), num_envs=E, seed=3, device=local, device_api=True)
        env.reset()
        g = torch.Generator(device=dev); g.manual_seed(5)
        acts = [torch.randn(E, 2, A, device=dev, generator=g) for _ in range(8)]
        for t in range(80):
            env.step(acts[t % 8])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for t in range(30):
            env.step(acts[t % 8])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        bits, n = env.check_status(strict=False)
        out[name.lower()] = {'envs': E, 'ms_per_step': ms, 'value': E / ms * 1e3, 'unit': 'env-steps/s', 'steps': 30, 'status_bits': bits,
                             'roofline_hbm_frac': bytes_ * E / (ms / 1e3) / 1e9 / read_peaks()[0]}
        env.close()
    return out


def measure_config5_eval(local, dev):
    """BASELINE.json configs[4]: the evaluation loop of eval_robosumo_against_fix.py:198-243 -- 32 envs (adjust_z = -0.5), the
    learner deterministic against a policy_zoo MLP (ant, tanh 64-64 with observation filter).  The reference asset
    agent-params-v3.npy does not travel to the GPU box, so the zoo parameters are synthetic of the same layout (24 645 floats)."""
    import numpy as np
    import torch
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    from robosumo_selfplay_b200.model import PPOModel
    from robosumo_selfplay_b200.policy_zoo import ZooMLPPolicy, evaluate_against_fixed
    rng = np.random.RandomState(0)
    n = 3 + 2 * 120 + 1 + 2 * (120 * 64 + 64 + 64 * 64 + 64) + 64 + 1 + 64 * 8 + 8 + 8
    flat = (rng.randn(n) * 0.1).astype(np.float32)
    flat[0:3] = [5.0, 80.0, 10.0]; flat[3:123] = rng.randn(120); flat[123:243] = 20 + rng.rand(120) * 30; flat[243] = 10.0
    np.random.seed(2)
    env = B200SumoVecEnv(ENV_ID, num_envs=32, seed=9, device=local, device_api=True, adjust_z=-0.5)
    model = PPOModel(ob_dim=121, ac_dim=8, trainable=False, device=local)
    zoo = ZooMLPPolicy(flat, 120, 8, device=local)
    evaluate_against_fixed(env, model, zoo, rounds=8, max_steps=50)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = evaluate_against_fixed(env, model, zoo, rounds=10 ** 9, max_steps=600)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    env.close()
    return {'envs': 32, 'steps': res['steps'], 'rounds': res['rounds'], 'value': 32 * res['steps'] / el, 'unit': 'env-steps/s', 'rounds_per_s': res['rounds'] / el,
            'win': res['win'], 'draw': res['draw'], 'lose': res['lose'], 'opponent': 'policy_zoo MLP layout, synthetic parameters',
            'note': 'per step: 2 MLP launches + 1 physics launch + the host-side win/draw/lose tally of the script (one D2H sync per step)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='env pairs per GPU')
    ap.add_argument('--settle', type=int, default=300, help='untimed steps before warm-up: episode phases are mixed (steady state) after ~100 steps')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--no-learner', action='store_true', help='skip the rollout / PPO2-update secondary measurements')
    ap.add_argument('--nsteps', type=int, default=128, help='T of the PPO2-update measurement (config 2: T in {128, 2048})')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from robosumo_selfplay_b200 import _lib
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    assert args.warmup >= 3, "W >= 3 warm-up steps"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    E = args.envs
    L = _lib.lib()
    env = B200SumoVecEnv(ENV_ID, num_envs=E, seed=42 + rank, device=local, device_api=True)
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    pool = [torch.randn(E, 2, 8, device=dev, generator=g) for _ in range(16)]      # N(0,1) actions (untrained policy)
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for t in range(args.settle + args.warmup):
        env.step(pool[t % 16])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.rs_launch_count()
    stream = torch.cuda.current_stream()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for t in range(args.steps):
        if flush is not None:
            flush.fill_(t & 255)                     # evict L2 between timed steps (untimed)
        ev[t][0].record(stream)
        env.step(pool[t % 16])
        ev[t][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = L.rs_launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_t.item()) / args.steps
    clocks = sampler.summary() if rank == 0 else None
    q, v, step, status = env.get_state()
    ncon_note = int((status & 2).sum().item())

    # ---- e2e: the public host API (numpy in / numpy out), copies inside the timed region ----
    henv = B200SumoVecEnv(ENV_ID, num_envs=E, seed=4242 + rank, device=local, device_api=False)
    henv.reset()
    hact = [np.random.RandomState(7 + i).randn(E, 2, 8).astype(np.float32) for i in range(8)]
    for t in range(args.settle // 2 + args.warmup):
        henv.step_async(hact[t % 8]); henv.step_wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n_e2e = max(20, args.steps // 4)
    a32 = hact[0].reshape(E, 16)
    t0 = time.perf_counter()
    for t in range(n_e2e):
        a = hact[t % 8].reshape(E, 16)
        _lib.check(L.rs_step_host(henv._h, henv._np(a), henv._np(henv.h_obs), henv._np(henv.h_rew), henv._np(henv.h_done),
                                  henv._np(henv.h_info), henv._np(henv.h_epi), 1))
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = E * world * n_e2e / float(e2e_t.item())
    # the same through the PYTHON drop-in a Runner / eval script calls: B200SumoVecEnv.step(numpy actions) -> (obs, rews, dones, infos)
    # with float64 obs / rewards as SubprocVecEnv returns them and the lazily materialised infos
    t0 = time.perf_counter()
    for t in range(n_e2e):
        o_, r_, d_, infos_ = henv.step(hact[t % 8])
        _ = infos_[t % E][0]['shaping_reward']                # a consumer touching one entry
    e2e_py_s = time.perf_counter() - t0
    e2e_py_t = torch.tensor([e2e_py_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_py_t, op=dist.ReduceOp.MAX)
    e2e_py_value = E * world * n_e2e / float(e2e_py_t.item())
    h2d = a32.nbytes
    d2h = henv.h_obs.nbytes + henv.h_rew.nbytes + henv.h_done.nbytes + henv.h_info.nbytes + henv.h_epi.nbytes
    learner = None
    config4 = None
    extra_configs = None
    if not args.no_learner:
        # BASELINE.json configs[3]: 65 536 env pairs in total, sharded over the ranks (several waves of blocks per SM: the
        # per-block slowest-warp tail of a single wave averages out)
        del henv
        E4 = 65536 // world
        env4 = B200SumoVecEnv(ENV_ID, num_envs=E4, seed=777 + rank, device=local, device_api=True)
        env4.reset()
        pool4 = [torch.randn(E4, 2, 8, device=dev, generator=g) for _ in range(4)]
        for t in range(60):
            env4.step(pool4[t % 4])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(20):
            env4.step(pool4[t % 4])
        e1.record(stream)
        torch.cuda.synchronize()
        ms4 = torch.tensor([e0.elapsed_time(e1) / 20], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        config4 = {'envs_total': E4 * world, 'envs_per_gpu': E4, 'ms_per_step': float(ms4.item()),
                   'value': E4 * world / (float(ms4.item()) / 1e3), 'unit': 'env-steps/s', 'steps': 20}
        env4.close(); del env4, pool4
        learner = measure_learner(args, E, local, rank, world, dev)
        if rank == 0 and world == 1:
            extra_configs = {'config1_ref_shape': measure_config1_ref_shape(local, dev), 'config3_morphologies': measure_config3_morphologies(local, dev, E),
                             'config5_eval': measure_config5_eval(local, dev)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = E * world / (ms_per_step / 1e3)
    peak, peak_src = read_peaks()
    achieved = ALGO_BYTES_PER_PAIR_STEP * E / (ms_per_step / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'k_step_traffic.json')          # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    fp32 = None
    if os.path.exists(tp):
        try:
            fpp = float(json.load(open(tp)).get('fp32_flop_per_pair_step'))
            sm_clock = (clocks or {}).get('sm_mhz') or 1965.0
            peak_tf = 148 * 128 * 2 * sm_clock * 1e6 / 1e12
            ach_tf = fpp * E * world / (ms_per_step / 1e3) / 1e12
            fp32 = {'achieved': ach_tf, 'peak': peak_tf * world, 'unit': 'TFLOP/s', 'frac': ach_tf / (peak_tf * world), 'flop_per_unit': fpp,
                    'note': 'useful (predicated-on) FP32 FLOP per pair-step from the committed ncu capture (profiles/k_step_traffic.json) against '
                            '148 SMs x 128 lanes x 2 x SM clock: the kernel is bound by dependent-instruction latency and the slowest warp of a block, '
                            'with 12.5 of 32 lanes active on average'}
        except Exception:
            fp32 = None
    cores = os.cpu_count() or 1
    cpu_rate, cpu_n, cpu_el = cpu_port_rate(args.cpu_seconds, 16 * cores, cores)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'RoboSumo-Ant-vs-Ant-v0, %d env pairs per GPU, physics(5x RK4)+obs+reward+done+auto-reset' % E,
                   'envs_per_gpu': E, 'frame_skip': 5, 'integrator': 'RK4', 'solver': 'Newton (primal), pyramidal cone',
                   'actions': 'N(0,1) i.i.d.', 'settle_steps': args.settle,
                   'l2': 'flushed between timed steps (256 MiB write, untimed)' if flush is not None else 'not flushed',
                   'timing': 'CUDA events per step on the launch stream, summed; max over ranks'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_unit': ALGO_BYTES_PER_PAIR_STEP,
                     'fp32': fp32,
                     'note': 'state stays in shared memory for all 20 forward evaluations; the kernel is FP32-latency/'
                             'issue bound by construction, so the HBM fraction is low (DESIGN.md)'},
        'cpu_baseline': {'value': cpu_rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'compiler_flags': 'gcc ' + str(CPU_FLAGS),
                         'sample': '%d env-steps (%d pairs, N(0,1) actions) in %.1f s on %d threads' % (cpu_n, 16 * cores, cpu_el, cores)},
        'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': n_e2e, 'api': 'rs_step_host (B200SumoVecEnv host style: numpy actions in through pinned staging, obs/rew/done/info/episode copied straight into the page-locked numpy result buffers)',
                'python_dropin': {'value': e2e_py_value, 'unit': 'env-steps/s', 'api': 'B200SumoVecEnv.step(numpy) -> float64 obs / rewards, bool dones, lazily materialised infos (what the reference Runner / eval scripts call)'}},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'contact_full_envs': ncon_note,
    }
    if learner is not None:
        line['learner'] = learner
    if config4 is not None:
        line['config4_65536_pairs'] = config4
    if extra_configs is not None:
        line.update(extra_configs)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
