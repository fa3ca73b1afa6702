"""Developer tool: cycles of an evaluation start and of a Newton iteration for a warp that has an SM to itself (E = 148), by contact
count, iteration count and coupling -- the per-phase costs behind tools/trip_model.py.  Needs the instrumented library (tools/eval_timeline.py).

    RS_B200_LIB=build/variants/librs_clk.so python tools/lone_warp_costs.py
"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
from robosumo_selfplay_b200 import _lib
E=148
env=B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0',num_envs=E,seed=42,device_api=True)
env.reset(); torch.manual_seed(0)
for t in range(100): env.step(torch.randn(E,2,8,device='cuda'))
L=_lib.lib(); L.rs_debug_read.restype=ctypes.c_int; L.rs_debug_read.argtypes=[ctypes.c_void_p, ctypes.c_size_t]
pre=[];sol=[];nit=[];cpl=[];ncon=[]
for rep in range(60):
    env.step(torch.randn(E,2,8,device='cuda')); torch.cuda.synchronize()
    buf=np.zeros(4200*128,np.int64); assert L.rs_debug_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes))==0
    T=buf.reshape(4200,128)[:E]
    beg=T[:,0:60:3].astype(np.float64); sb=T[:,1:60:3].astype(np.float64); end=T[:,2:60:3].astype(np.float64)
    info=T[:,64:84]
    pre.append(sb-beg); sol.append(end-sb); nit.append(info&255); cpl.append((info>>8)&1); ncon.append((info>>16)&255)
pre=np.concatenate(pre).ravel(); sol=np.concatenate(sol).ravel(); nit=np.concatenate(nit).ravel(); cpl=np.concatenate(cpl).ravel(); ncon=np.concatenate(ncon).ravel()
print('evals',len(pre),'pre mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f'%(pre.mean(),*np.percentile(pre,[50,90,99]),pre.max()))
print('pre by ncon:',', '.join('%d:%.0f(n=%d,p90 %.0f)'%(k,pre[ncon==k].mean(),(ncon==k).sum(),np.percentile(pre[ncon==k],90)) for k in range(10) if (ncon==k).sum()>5))
for lab,m in (('uncoupled',cpl==0),('coupled',cpl==1)):
    for it in (1,2,3):
        mm=m&(nit==it)
        print(lab,'it',it,'solve by ncon:',', '.join('%d:%.0f(n=%d)'%(k,sol[mm&(ncon==k)].mean(),(mm&(ncon==k)).sum()) for k in range(12) if (mm&(ncon==k)).sum()>3))
