"""Developer bench: step time of the physics kernel against the number of env pairs (one wave ... many waves per SM).

    [RS_B200_LIB=build/variants/librs_<v>.so] python tools/bench_batch_sizes.py [E ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv

sizes = [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192, 16384, 65536]
for E in sizes:
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=1, device_api=True)
    env.reset(); torch.manual_seed(0)
    acts = [torch.randn(E, 2, 8, device='cuda') for _ in range(4)]
    for t in range(100):
        env.step(acts[t % 4])
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(30):
        env.step(acts[t % 4])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 30
    print('E %6d ms/step %.3f env-steps/s %.0f' % (E, ms, E / ms * 1e3), flush=True)
    env.close()
