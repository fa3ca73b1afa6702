"""Developer bench: Ant-vs-Ant throughput versus the number of env pairs (one wave = 4144 pairs on 148 SMs).

    python tools/bench_batch_sizes.py        # needs a B200 and the built library
"""
import sys, torch, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
for E in (2048, 4096, 4144, 8192, 8288, 16384, 32768, 65536):
    env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=1,device_api=True)
    env.reset(); torch.manual_seed(0)
    acts=[torch.randn(E,2,8,device='cuda') for _ in range(4)]
    for t in range(45): env.step(acts[t%4])
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(20): env.step(acts[t%4])
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/20
    print('E %6d ms/step %.3f env-steps/s %.0f'%(E,ms,E/ms*1e3), flush=True)
    env.close()
