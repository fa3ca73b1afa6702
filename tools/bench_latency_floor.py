"""Developer bench: step time with 1, 2, 4, 8, 16 warps (env pairs) per SM -- the dependent-instruction latency floor of the physics kernel.\n\n    python tools/bench_latency_floor.py\n"""
import sys, torch
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
for E in (148, 296, 592, 1184, 2368):
    env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=1,device_api=True)
    env.reset(); torch.manual_seed(0)
    acts=[torch.randn(E,2,8,device='cuda') for _ in range(4)]
    for t in range(100): env.step(acts[t%4])
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(100): env.step(acts[t%4])
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/100
    print('E %5d (%2d warps/SM) ms/step %.3f'%(E,(E+147)//148,ms), flush=True)
    env.close()
