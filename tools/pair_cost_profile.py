"""Developer tool: how unevenly is the cost of one env step spread over the env pairs and over the blocks of k_step?

Needs the instrumented library (see tools/eval_timeline.py for the build line):
    RS_B200_LIB=build/variants/librs_clk.so python tools/pair_cost_profile.py [E] [steps]
Prints, at steady state: the distribution of a pair's own busy cycles (sum of its 20 evaluation durations, barrier waits
excluded), of its Newton iterations and coupled evaluations, the span of every block against the busiest pair it holds, and
what the kernel would cost if blocks were as long as their busiest pair (no lockstep waiting) or as their mean pair."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
from robosumo_selfplay_b200 import _lib

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
assert E <= 4200
env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=42, device_api=True)
env.reset(); torch.manual_seed(0)
for t in range(100):
    env.step(torch.randn(E, 2, 8, device='cuda'))
L = _lib.lib()
L.rs_debug_read.restype = ctypes.c_int
L.rs_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
wpb = 28
nb = E // wpb
rows = []
for rep in range(NS):
    env.step(torch.randn(E, 2, 8, device='cuda'))
    torch.cuda.synchronize()
    buf = np.zeros(4200 * 128, np.int64)
    assert L.rs_debug_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes)) == 0
    T = buf.reshape(4200, 128)[:nb * wpb]
    beg = T[:, 0:60:3].astype(np.float64); sb = T[:, 1:60:3].astype(np.float64); end = T[:, 2:60:3].astype(np.float64)
    k0 = T[:, 60].astype(np.float64); k1 = T[:, 61].astype(np.float64)
    info = T[:, 64:84]; nit = info & 255; cpl = (info >> 8) & 1; ncon = (info >> 16) & 255
    busy = (end - beg).sum(1)                      # own cycles of a pair, waits excluded
    pre = (sb - beg).sum(1); sol = (end - sb).sum(1)
    span = (k1.reshape(nb, wpb).max(1) - k0.reshape(nb, wpb).min(1))
    bmax = busy.reshape(nb, wpb).max(1); bmean = busy.reshape(nb, wpb).mean(1)
    lock = (end - beg).reshape(nb, wpb, 20).max(1).sum(1)      # sum over evaluations of the slowest warp's duration
    ncp = (cpl.sum(1) > 0).reshape(nb, wpb).sum(1)
    q = lambda a, p: float(np.percentile(a, p))
    print('step %d: pair busy cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f | pre-solve mean %.0f solve mean %.0f' %
          (rep, busy.mean(), q(busy, 50), q(busy, 90), q(busy, 99), busy.max(), pre.mean(), sol.mean()))
    print('   iterations per pair-step mean %.1f p90 %.0f p99 %.0f max %d | pairs with a coupled evaluation %.1f%% (coupled evals %.2f%%) | contacts/eval %.1f' %
          (nit.sum(1).mean(), q(nit.sum(1), 90), q(nit.sum(1), 99), nit.sum(1).max(), 100 * (cpl.sum(1) > 0).mean(), 100 * cpl.mean(), ncon.mean()))
    print('   block span mean %.0f p50 %.0f max %.0f | busiest pair of a block mean %.0f max %.0f | mean pair of a block %.0f | lockstep sum (sum_eval max_warp) mean %.0f max %.0f' %
          (span.mean(), q(span, 50), span.max(), bmax.mean(), bmax.max(), bmean.mean(), lock.mean(), lock.max()))
    i = int(np.argmax(span))
    print('   slowest block %d: span %.0f, busiest pair %.0f, coupled pairs %d, iterations of its pairs %s' %
          (i, span[i], bmax[i], ncp[i], np.sort(nit.sum(1).reshape(nb, wpb)[i])[-6:].tolist()))
    print('   span vs coupled pairs in block: ' + ', '.join('%d:%0.0f(n=%d)' % (k, span[ncp == k].mean(), (ncp == k).sum()) for k in range(0, 6) if (ncp == k).any()))
    # cost of a coupled / uncoupled iteration and of the pre-solve part, per evaluation
    e_sol = (end - sb); e_pre = (sb - beg)
    for lab, m in (('uncoupled', cpl == 0), ('coupled', cpl == 1)):
        if m.any():
            its = nit[m]
            print('   %s evals: n %d, solve cycles by iterations: ' % (lab, m.sum()) +
                  ', '.join('%d it: %.0f (n=%d)' % (k, e_sol[m][its == k].mean(), (its == k).sum()) for k in range(1, 7) if (its == k).any()) +
                  ' | pre-solve %.0f' % e_pre[m].mean())
    for its_k in (1, 2, 3):
        m = (cpl == 0) & (nit == its_k)
        print('   uncoupled evals with %d iteration(s), solve cycles by contact count: ' % its_k +
              ', '.join('%d:%.0f(n=%d)' % (k, e_sol[m & (ncon == k)].mean(), (m & (ncon == k)).sum()) for k in range(0, 13) if (m & (ncon == k)).sum() > 3))
    m = (cpl == 1) & (nit == 1)
    print('   coupled evals with 1 iteration, solve cycles by contact count: ' +
          ', '.join('%d:%.0f(n=%d)' % (k, e_sol[m & (ncon == k)].mean(), (m & (ncon == k)).sum()) for k in range(0, 16) if (m & (ncon == k)).sum() > 1))
    print('   pre-solve cycles by contact count: ' + ', '.join('%d:%.0f' % (k, e_pre[ncon == k].mean()) for k in range(0, 13) if (ncon == k).sum() > 3))
    print('   contact count histogram (all evals): ' + str(np.bincount(ncon.ravel(), minlength=12)[:16].tolist()))
    rows.append((busy.mean(), span.max()))
print('mean over steps: pair busy %.0f cycles, slowest block span %.0f cycles' % (np.mean([r[0] for r in rows]), np.mean([r[1] for r in rows])))

# ---- cost of one Newton iteration by section, uncoupled vs coupled (least squares over the pairs of many steps) ----
if len(sys.argv) > 3:
    NREG = int(sys.argv[3])
    X, Y = [], []
    for rep in range(NREG):
        env.step(torch.randn(E, 2, 8, device='cuda'))
        torch.cuda.synchronize()
        buf = np.zeros(4200 * 128, np.int64)
        assert L.rs_debug_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes)) == 0
        T = buf.reshape(4200, 128)[:E]
        info = T[:, 64:84]; nit = info & 255; cpl = (info >> 8) & 1
        X.append(np.stack([(nit * (cpl == 0)).sum(1), (nit * (cpl == 1)).sum(1), np.full(E, 20)], 1)); Y.append(T[:, 90:96])
    X = np.concatenate(X).astype(np.float64); Y = np.concatenate(Y).astype(np.float64)
    coef = np.linalg.lstsq(X, Y, rcond=None)[0]
    names = ['jt_forces+grad', 'build_H', 'linear solve', 'twists+rows+matvec', 'checks+linesearch+update', 'outside (pre-solve, RK)']
    print('cycles per Newton iteration by section (regression over %d pair-steps, %d with coupled iterations):' % (len(X), int((X[:, 1] > 0).sum())))
    for i, n in enumerate(names):
        print('  %-28s uncoupled it %8.0f | coupled it %8.0f | per evaluation %8.0f' % (n, coef[0, i], coef[1, i], coef[2, i]))
