"""Developer bench: PPO2 update wall time split into host shuffles, index upload and GPU time.

    python tools/update_breakdown.py        # needs a B200 and the built library
"""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.model import PPOModel
from robosumo_selfplay_b200.dist import EpochPermutations, legacy_shuffle
D,A=121,8; N=524288; nbt=16384
np.random.seed(0)
m=PPOModel(ob_dim=D,ac_dim=A)
g=torch.Generator(device='cuda'); g.manual_seed(0)
obs=torch.randn(N,D,device='cuda',generator=g); act=torch.randn(N,A,device='cuda',generator=g)*0.5
ret=torch.randn(N,device='cuda',generator=g); val=torch.randn(N,device='cuda',generator=g); old=8+torch.randn(N,device='cuda',generator=g)
def one_update(perms):
    for inds in perms:
        di=torch.as_tensor(inds.astype(np.int32),device='cuda')
        for s0 in range(0,N,nbt):
            m.train_indexed(1e-3,0.2,obs,ret,act,val,old,None,di[s0:s0+nbt],global_n=nbt)
one_update(EpochPermutations(N,6)); torch.cuda.synchronize()
t0=time.perf_counter(); one_update(EpochPermutations(N,6)); torch.cuda.synchronize(); print('update wall %.1f ms'%((time.perf_counter()-t0)*1e3))
a=np.arange(N); t0=time.perf_counter()
for _ in range(6): legacy_shuffle(a)
print('6 shuffles alone %.1f ms'%((time.perf_counter()-t0)*1e3))
pre=list(EpochPermutations(N,6))
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0=time.perf_counter(); s.record(); one_update(pre); e.record(); torch.cuda.synchronize()
print('with precomputed permutations: wall %.1f ms, gpu %.1f ms'%((time.perf_counter()-t0)*1e3, s.elapsed_time(e)))
t0=time.perf_counter()
for inds in pre: di=torch.as_tensor(inds.astype(np.int32),device='cuda')
torch.cuda.synchronize(); print('astype + H2D of 6 index arrays %.1f ms'%((time.perf_counter()-t0)*1e3))
