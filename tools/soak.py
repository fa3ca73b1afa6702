"""Soak run: thousands of steps with occasional 3-sigma actions; counts non-finite outputs and solver status flags.

    python tools/soak.py        # needs a B200 and the built library
"""
import sys, torch, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
for name,A,steps in (('Ant',8,4000),('Bug',12,1200),('Spider',16,800)):
    E=4096
    env=B200SumoVecEnv('RoboSumo-%s-vs-%s-v0'%(name,name),num_envs=E,seed=123,device_api=True)
    env.reset(); g=torch.Generator(device='cuda'); g.manual_seed(1)
    bad=0; flags=torch.zeros(4,dtype=torch.long); ndone=0; wins=torch.zeros(3,dtype=torch.long)
    t0=time.time()
    for t in range(steps):
        scale=1.0 if t%3 else 3.0           # occasional very large actions
        obs,rew,done,(info,epi)=env.step(scale*torch.randn(E,2,A,device='cuda',generator=g))
        if t%50==49:
            q,v,step,status=env.get_state()
            bad+=int((~torch.isfinite(obs)).sum())+int((~torch.isfinite(rew)).sum())
            for b in range(3): flags[b]+=int(((status>>b)&1).sum())
            ndone+=int(done[:,0].sum())
    print(name,'steps',steps,'non-finite outputs',bad,'status NaN/contact_full/maxit (sampled env-steps)',flags[:3].tolist(),'max |qvel| %.1f'%float(v.abs().max()),'%.1f s'%(time.time()-t0), flush=True)
    env.close()
