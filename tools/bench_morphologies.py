"""Developer bench: ms/step, Newton iterations, contacts and solver flags for Ant/Bug/Spider at E=4096 (device style).

    python tools/bench_morphologies.py        # needs a B200 and the built library
"""
import sys, torch, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
for name,A in (('Ant',8),('Bug',12),('Spider',16)):
    E=4096
    env=B200SumoVecEnv('RoboSumo-%s-vs-%s-v0'%(name,name),num_envs=E,seed=1,device_api=True)
    env.reset(); torch.manual_seed(0)
    acts=[torch.randn(E,2,A,device='cuda') for _ in range(8)]
    for t in range(45): env.step(acts[t%8])
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(30): env.step(acts[t%8])
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/30
    d=env.diagnostics().float(); q,v,step,status=env.get_state()
    print(name,'ms/step %.3f'%ms,'env-steps/s %.0f'%(E/ms*1e3),'newton/eval %.2f'%(d[:,0].mean().item()/20),'ncon/eval %.2f'%(d[:,2].mean().item()/20),'coupled %.3f'%(d[:,1].mean().item()/20),'status flags',int((status&7).max()), 'contact_full', int(((status&2)>0).sum()))
    env.close()
