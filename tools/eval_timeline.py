"""Developer tool: per-warp timeline of the physics kernel (where does an evaluation spend its cycles, who is the slowest warp of a block).

Needs a library built with the instrumentation compiled in:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DRS_DEV_ANT_ONLY -DRS_EXPERIMENT_CLOCK \
         -o build/variants/librs_clk.so robosumo_selfplay_b200/csrc/rs_api.cu
    RS_B200_LIB=build/variants/librs_clk.so python tools/eval_timeline.py
Numbers from this tool are quoted in DESIGN.md section 4; the product library is built without the instrumentation."""
import torch, sys, numpy as np, ctypes
sys.path.insert(0,'/root/repo')
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
from robosumo_selfplay_b200 import _lib
E=4096
env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=42,device_api=True)
env.reset(); torch.manual_seed(0)
for t in range(50): env.step(torch.randn(E,2,8,device='cuda'))
torch.cuda.synchronize()
L=_lib.load() if hasattr(_lib,'load') else env._L
buf=np.zeros(4200*128,np.int64); rc=L.rs_debug_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes)); assert rc==0
T=buf.reshape(4200,128)[:E]
nb=E//28
beg=T[:nb*28,0:60:3].reshape(nb,28,20).astype(np.float64); sb=T[:nb*28,1:60:3].reshape(nb,28,20).astype(np.float64); end=T[:nb*28,2:60:3].reshape(nb,28,20).astype(np.float64)
k0=T[:nb*28,60].reshape(nb,28).astype(np.float64); k1=T[:nb*28,61].reshape(nb,28).astype(np.float64)
dur=end-beg
print('kernel span per block (first begin -> last end of simulate): p50 %.0f max %.0f'%(np.median((k1.max(1)-k0.min(1))),(k1.max(1)-k0.min(1)).max()))
print('begin skew within block per eval (max-min): mean %.0f'%((beg.max(1)-beg.min(1)).mean()))
print('eval duration per warp: mean %.0f ; block max per eval: mean %.0f ; ratio max/mean %.2f'%(dur.mean(), dur.max(1).mean(), dur.max(1).mean()/dur.mean()))
print('pre-solve part mean %.0f, solve part mean %.0f ; block-max of pre-solve %.0f ; block-max solve %.0f'%((sb-beg).mean(),(end-sb).mean(),(sb-beg).max(1).mean(),(end-sb).max(1).mean()))
gap=beg[:,:,1:].min(1)-end[:,:,:-1].max(1)
print('gap between slowest end of eval k and first begin of eval k+1: mean %.0f'%gap.mean())
tot=(k1.max(1)-k0.min(1)); i=np.argmax(tot); j=np.argsort(tot)[nb//2]
for name,b in (('slowest',i),('median',j)):
    print(name,'block',b,'span %.0f'%tot[b],'per-eval block-max dur:',(dur[b].max(0)/1e3).round(0).tolist())
    print('      per-eval mean dur:',(dur[b].mean(0)/1e3).round(0).tolist())

info=T[:nb*28,64:84].reshape(nb,28,20); nit=info&255; cpl=(info>>8)&255; ncon=(info>>16)&255
sol=end-sb
am=sol.argmax(1)   # [nb,20] index of slowest warp
pick=lambda A: np.take_along_axis(A,am[:,None,:],1)[:,0,:]
print('slowest-solve warp per (block,eval): iters mean %.2f (all warps %.2f) ; coupled frac %.2f (all %.3f) ; ncon %.2f (all %.2f)'%(pick(nit).mean(),nit.mean(),(pick(cpl)&1).mean(),(cpl&1).mean(),pick(ncon).mean(),ncon.mean()))
for it in range(1,7):
    m=(nit==it)&((cpl&1)==0); mc=(nit==it)&((cpl&1)==1)
    print('iters=%d: uncoupled n=%6d solve mean %.0f | coupled n=%5d solve mean %.0f'%(it,m.sum(),sol[m].mean() if m.any() else 0,mc.sum(),sol[mc].mean() if mc.any() else 0))
ps=pick(sol); pc=pick(cpl)&1; pn=pick(nit)
print('block-max solve: mean %.0f ; when slowest is coupled (%.0f%%): %.0f ; uncoupled: %.0f'%(ps.mean(),100*pc.mean(),ps[pc==1].mean(),ps[pc==0].mean()))
print('histogram of iterations of the slowest warp:',np.bincount(pn.ravel(),minlength=8)[:10].tolist())

acc=T[:,90:96].astype(np.float64); its=nit.reshape(-1,20).sum(1) if False else None
tot_it=(T[:nb*28,64:84]&255).sum(1).astype(np.float64); A=acc[:nb*28]
names=['jt_forces+grad','build_H','chol_solve','twists+rows+matvec (incl. residual pass)','checks+linesearch+update','(outside / first-pass setup)']
print('cycles per Newton iteration (sum over step / iterations), all warps:')
for i,n in enumerate(names): print('  %-44s %.0f'%(n,A[:,i].sum()/tot_it.sum()))
