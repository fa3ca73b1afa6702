"""Developer tool: how much of a block's time is the wait for its slowest warp?  A trip-level model of k_step's block schedule
(simulate_trips in rs_core.h) fed with the per-evaluation record of a real launch and the lone-warp costs of tools/lone_warp_costs.py;
what-if variants (costs flat in the contact count, coupled = uncoupled, cheaper phases) rank the optimisations before they are written.

    RS_B200_LIB=build/variants/librs_clk.so python tools/trip_model.py record     # on the GPU: writes gpurun_out/evals_4088.npz
    python tools/trip_model.py                                                    # anywhere: the model on that record
"""
import sys
if len(sys.argv) > 1 and sys.argv[1] == "record":
    import ctypes, sys, numpy as np, torch
    sys.path.insert(0,'/root/repo')
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    from robosumo_selfplay_b200 import _lib
    E=4088
    env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=42,device_api=True)
    env.reset(); torch.manual_seed(0)
    for t in range(300): env.step(torch.randn(E,2,8,device='cuda'))
    L=_lib.lib(); L.rs_debug_read.restype=ctypes.c_int; L.rs_debug_read.argtypes=[ctypes.c_void_p, ctypes.c_size_t]
    out=[]; span=[]
    for rep in range(8):
        env.step(torch.randn(E,2,8,device='cuda')); torch.cuda.synchronize()
        buf=np.zeros(4200*128,np.int64); assert L.rs_debug_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes))==0
        T=buf.reshape(4200,128)[:E]
        out.append(T[:,64:84].copy()); span.append((T[:,61].reshape(146,28).max(1)-T[:,60].reshape(146,28).min(1)).copy())
    np.savez_compressed('/root/repo/gpurun_out/evals_4088.npz', info=np.stack(out), span=np.stack(span))
    print('ok', np.stack(span).mean(), np.stack(span).max(1).mean())
    sys.exit(0)
import numpy as np
z=np.load('/root/repo/gpurun_out/evals_4088.npz'); info=z['info']; span=z['span']
nit=(info&255); cpl=(info>>8)&1; ncon=(info>>16)&255       # [S,E,20]
S,E,_=nit.shape; B=E//28
def pre_cost(nc, flat=False): return 12.5e3 + (0*nc if flat else 1.7e3*nc)
def it_cost(nc, cp, flat=False, nocpl=False):
    base = 10.3e3 + (0*nc if flat else (4.6e3*(nc>=1) + 2.4e3*np.maximum(nc-1,0)))
    return base + (0 if nocpl else 11e3*cp)
def sim(flat=False, nocpl=False, pre_scale=1.0, it_scale=1.0, contention=None):
    tot=[]; 
    for s in range(S):
        bt=[]
        for b in range(B):
            n_=nit[s,b*28:(b+1)*28]; c_=cpl[s,b*28:(b+1)*28]; k_=ncon[s,b*28:(b+1)*28]
            ev=np.zeros(28,int); it=np.zeros(28,int); fresh=np.ones(28,bool); done=np.zeros(28,bool)
            T=0.0
            while not done.all():
                act=~done
                f=act&fresh
                idx=np.arange(28)
                if f.any():
                    pc=pre_cost(k_[idx[f],ev[f]],flat)*pre_scale; m=pc.max(); nf=f.sum()
                    T+= m*(contention(nf) if contention else 1)
                    fresh[f]=False; it[f]=0
                ic=it_cost(k_[idx[act],ev[act]],c_[idx[act],ev[act]],flat,nocpl)*it_scale
                T+= ic.max()*(contention(act.sum()) if contention else 1)
                it[act]+=1
                fin=act&(it>=n_[idx,np.minimum(ev,19)])
                ev[fin]+=1; fresh[fin]=True
                done|= ev>=20
            bt.append(T)
        tot.append((np.mean(bt),np.max(bt)))
    tot=np.array(tot); return tot[:,0].mean(), tot[:,1].mean()
cont=lambda n: 1+0.55*(n-1)/27.0
print('measured span (clk build): mean %.0f max %.0f'%(span.mean(), span.max(1).mean()))
for name,kw in (('baseline',{}),('flat in ncon',dict(flat=True)),('coupled=uncoupled',dict(nocpl=True)),('both',dict(flat=True,nocpl=True)),('pre x0.7',dict(pre_scale=0.7)),('iter x0.7',dict(it_scale=0.7))):
    a=sim(**kw); b=sim(contention=cont,**kw)
    print('%-20s no-contention: block mean %.0f max %.0f | with contention: mean %.0f max %.0f'%(name,a[0],a[1],b[0],b[1]))
print('--- what-if: cheaper per-contact costs')
def mk(pre_pc, it_first, it_next, cpl_extra):
    global pre_cost, it_cost
    def pre_cost(nc, flat=False): return 12.5e3 + pre_pc*nc
    def it_cost(nc, cp, flat=False, nocpl=False): return 10.3e3 + it_first*(nc>=1) + it_next*np.maximum(nc-1,0) + cpl_extra*cp
for name,args in (('baseline',(1.7e3,4.6e3,2.4e3,11e3)),('build_H parallel: it 3.0k/0.9k',(1.7e3,3.0e3,0.9e3,11e3)),('+ pre 1.0k',(1.0e3,3.0e3,0.9e3,11e3)),('+ coupled +6k',(1.0e3,3.0e3,0.9e3,6e3)),('+ coupled +3k',(1.0e3,3.0e3,0.9e3,3e3))):
    mk(*args); b=sim(contention=cont)
    print('%-36s with contention: mean %.0f max %.0f'%(name,b[0],b[1]))
