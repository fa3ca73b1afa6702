"""Developer analysis (CPU only): how well can the active state of the joint-limit rows be predicted from earlier evaluations?

Compiles the kernel source as a host emulation with the RS_SOLVE_TRACE hook of rs_core.h recording (sign, residual at the solution)
of every limit row after each forward evaluation, steps a dozen Ant-vs-Ant episodes with N(0,1) actions and scores three
predictors.  Result quoted in DESIGN.md section 3.1: the previous evaluation predicts a row with 95.5 % accuracy, the same RK stage
of the previous substep with 86 %, a linear trend of the residual with 91 %; with 5.6 rows inside their limit region per evaluation
that still leaves one wrong row in 17 % of the evaluations.

    python tools/limit_row_predictors.py
"""
import os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '..', 'build'); os.makedirs(BUILD, exist_ok=True)
CPP = r'''#include <string.h>
#include <stdlib.h>
#include <stdio.h>
static float g_trace[4000000]; static long g_n = 0;
#define RS_SOLVE_TRACE(s, it) { if (g_n + 40 < 4000000) { g_trace[g_n++] = (float)(it); g_trace[g_n++] = (float)(s).ncon; for (int j = 0; j < 16; j++) { g_trace[g_n++] = (s).lsgn[j]; g_trace[g_n++] = (s).ljar[j]; } } }
#include "../robosumo_selfplay_b200/csrc/rs_env.h"
using namespace rs;
extern "C" long trace_get(float* out) { memcpy(out, g_trace, sizeof(float) * g_n); long n = g_n; g_n = 0; return n; }
extern "C" int trace_step(const rs_agent_model* am, float h, int max_newton, float* q, float* v, float* warm, const float* ctrl, int nsub) {
    typedef Slab<4, 4> S;
    static S* s = 0;
    if (!s) s = (S*)calloc(1, sizeof(S));
    Ctx<4, 4> c; c.s = s; c.am = am; c.h = h; c.max_newton = max_newton;
    memcpy(s->q, q, sizeof(float) * S::NQ); memcpy(s->v, v, sizeof(float) * S::NV); memcpy(s->x, warm, sizeof(float) * S::NV);
    for (int a = 0; a < 2; a++) for (int k = 0; k < 8; k++) { int u = a * 8 + k; float x = ctrl[u]; x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x); s->act[u] = am[a].gear * x; }
    simulate(c, nsub);
    memcpy(q, s->q, sizeof(float) * S::NQ); memcpy(v, s->v, sizeof(float) * S::NV); memcpy(warm, s->x, sizeof(float) * S::NV);
    return s->status;
}
'''
open(os.path.join(BUILD, 'emu_trace.cpp'), 'w').write(CPP)
subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', HERE, '-o', os.path.join(BUILD, 'libemu_trace.so'), os.path.join(BUILD, 'emu_trace.cpp')])
import ctypes, numpy as np, sys
sys.path.insert(0, os.path.join(HERE, '..'))
from robosumo_selfplay_b200.morphology import PairSpec
L=ctypes.CDLL(os.path.join(BUILD, 'libemu_trace.so')); L.trace_get.restype=ctypes.c_long
P=lambda a: a.ctypes.data_as(ctypes.c_void_p)
ps=PairSpec('ant','ant'); pack=ps.pack()
rows=[]
rng=np.random.RandomState(0)
for ep in range(12):
    q=ps.qpos0().astype(np.float64); phi=rng.uniform(0,2*np.pi)
    for a in range(2):
        o=a*15; q[o]=1.15*np.cos(phi+a*np.pi); q[o+1]=1.15*np.sin(phi+a*np.pi); q[o+2]=1.25
    q+=rng.uniform(-.1,.1,30); v=0.1*rng.randn(28)
    for o in (3,18): q[o:o+4]/=np.linalg.norm(q[o:o+4])
    qf=q.astype(np.float32); vf=v.astype(np.float32); wf=np.zeros(28,np.float32)
    for t in range(45):
        ctrl=rng.randn(16).astype(np.float32)
        L.trace_step(pack, ctypes.c_float(0.01), 16, P(qf),P(vf),P(wf),P(ctrl),5)
        buf=np.zeros(4000000,np.float32); n=L.trace_get(P(buf)); d=buf[:n].reshape(-1,34)
        if t>=8: rows.append(d)       # 20 evals per step
        if qf[2]<0.29 or qf[17]<0.29 or abs(qf[:2]).max()>2.1 or abs(qf[15:17]).max()>2.1: break

D=np.concatenate(rows)
it=D[:,0]; sg=D[:,2::2]; jar=D[:,3::2]
loaded=(sg!=0)&(jar<0)
N=len(D)
print('evaluations traced', N, 'mean Newton iterations %.2f' % it.mean())
def evalp(name, pred, valid):
    m=valid&(sg!=0)
    print('%-40s rows %6d  row accuracy %.4f  evaluations with a wrong row %.3f'%(name,m.sum(),(pred[m]==loaded[m]).mean(),((pred!=loaded)&m).any(1).mean()))
prev_ok=np.zeros_like(loaded); prev_ok[1:]=(sg[1:]==sg[:-1])&(sg[:-1]!=0)
pred_prev=np.zeros_like(loaded); pred_prev[1:]=loaded[:-1]
evalp('previous evaluation (what the kernel uses)',pred_prev,prev_ok)
p4_ok=np.zeros_like(loaded); p4_ok[4:]=(sg[4:]==sg[:-4])&(sg[:-4]!=0)
pred4=np.zeros_like(loaded); pred4[4:]=loaded[:-4]
evalp('same RK stage of the previous substep',pred4,p4_ok)
t_ok=np.zeros_like(loaded); t_ok[2:]=prev_ok[2:]&prev_ok[1:-1]
jp=np.zeros_like(jar); jp[2:]=2*jar[1:-1]-jar[:-2]
evalp('linear trend of the residual (2 evaluations)',jp<0,t_ok)
new=(sg!=0)&~prev_ok
print('rows entering their limit region: %d, loaded at the solution in %.3f of them'%(new.sum(),loaded[new].mean()))
print('rows inside their limit region per evaluation %.2f, of which loaded %.3f'%((sg!=0).sum(1).mean(),loaded[sg!=0].mean()))
