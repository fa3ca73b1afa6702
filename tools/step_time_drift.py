"""Step time and solver statistics over 1500 steps: the mix of episode phases is stationary after ~100 steps.

    python tools/step_time_drift.py        # needs a B200 and the built library
"""
import sys, torch
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
E=4096
env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=42,device_api=True)
env.reset(); torch.manual_seed(0)
acts=[torch.randn(E,2,8,device='cuda') for _ in range(16)]
ev=[torch.cuda.Event(enable_timing=True) for _ in range(32)]
ndone=0
for blk in range(30):
    ev[blk].record()
    dn=0
    for t in range(50):
        o,r,d,i=env.step(acts[t%16]); dn+=int(d[:,0].sum()) if t%10==0 else 0
    dg=env.diagnostics().float()
    torch.cuda.synchronize()
    print('steps %4d-%4d'%(blk*50,blk*50+49),'it/eval %.2f ncon/eval %.2f coupled %.3f done/step~%d'%(dg[:,0].mean().item()/20,dg[:,2].mean().item()/20,dg[:,1].mean().item()/20,dn/5), flush=True)
ev[30].record(); torch.cuda.synchronize()
for blk in range(30): print(blk*50, '%.3f ms/step'%(ev[blk].elapsed_time(ev[blk+1])/50 if blk<30 else 0))
