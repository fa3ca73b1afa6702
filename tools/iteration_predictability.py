"""Developer tool: can a pair's cost in the next env step be predicted from its last one (cost-sorted blocks)?

    python tools/iteration_predictability.py record        # on the GPU: 60 consecutive steps of rs_get_diag -> gpurun_out/diag_seq_4096.npz
    python tools/iteration_predictability.py               # anywhere: correlations, a linear predictor, and what three block
                                                           # assignments (as launched / dealt by last step's count / dealt by the TRUE
                                                           # count) would do to the slowest block in the trip model of tools/trip_model.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
NPZ = os.path.join(ROOT, 'gpurun_out', 'diag_seq_4096.npz')
if len(sys.argv) > 1 and sys.argv[1] == 'record':
    import torch
    sys.path.insert(0, ROOT)
    from robosumo_selfplay_b200.vec_env import B200SumoVecEnv
    E = 4096
    env = B200SumoVecEnv('RoboSumo-Ant-vs-Ant-v0', num_envs=E, seed=42, device_api=True)
    env.reset(); torch.manual_seed(0)
    acc, dn = [], []
    for t in range(360):
        o, r, d, i = env.step(torch.randn(E, 2, 8, device='cuda'))
        if t >= 300:
            acc.append(env.diagnostics().cpu().numpy().copy()); dn.append(d[:, 0].cpu().numpy().copy())
    np.savez_compressed(NPZ, diag=np.stack(acc), done=np.stack(dn))
    sys.exit(0)

z = np.load(NPZ); D = z['diag'].astype(float)
it, cp, nc = D[:, :, 0], D[:, :, 1], D[:, :, 2]
print('Newton iterations per pair-step: mean %.1f sd %.1f, slowest pair of a launch %.1f' % (it.mean(), it.std(), it.max(1).mean()))
print('correlation of a pair\'s count with its count in the previous step %.3f (contacts -> iterations %.3f, coupled evaluations -> iterations %.3f)' % (
    np.mean([np.corrcoef(it[t], it[t + 1])[0, 1] for t in range(len(it) - 1)]), np.mean([np.corrcoef(nc[t], it[t + 1])[0, 1] for t in range(len(it) - 1)]),
    np.mean([np.corrcoef(cp[t], it[t + 1])[0, 1] for t in range(len(it) - 1)])))
X = np.stack([it[:-1].ravel(), nc[:-1].ravel(), cp[:-1].ravel(), np.ones(it[:-1].size)], 1); y = it[1:].ravel()
w = np.linalg.lstsq(X, y, rcond=None)[0]
print('linear predictor (iterations, contacts, coupled, 1) -> R^2 %.3f' % (1 - ((X @ w - y) ** 2).sum() / ((y - y.mean()) ** 2).sum()))
E, B, W = 4096, 147, 28
cost = lambda n: 25.7e3 + (77e3 - 25.7e3) * (n - 1) / 27.0        # cycles of a trip with n active warps (lone ... full block)
def block_time(iters):
    return sum(cost((iters >= t).sum()) for t in range(1, int(iters.max()) + 1))
def deal(pred):       # heaviest first, dealt to the blocks in snake order: every block gets the same mix
    blocks = [[] for _ in range(B)]
    for j, e in enumerate(np.argsort(-pred, kind='stable')):
        rd, pos = divmod(j, B)
        blocks[pos if rd % 2 == 0 else B - 1 - pos].append(e)
    return blocks
res = {'as launched': [], 'dealt by the previous step\'s count': [], 'dealt by the linear predictor': [], 'dealt by the TRUE count (oracle)': []}
for t in range(1, len(it)):
    v = it[t]
    cand = {'as launched': [list(range(b * W, min((b + 1) * W, E))) for b in range(B)], 'dealt by the previous step\'s count': deal(it[t - 1]),
            'dealt by the linear predictor': deal(X[(t - 1) * E:t * E] @ w), 'dealt by the TRUE count (oracle)': deal(v)}
    for k, blocks in cand.items():
        bt = np.array([block_time(v[np.array(b)]) for b in blocks])
        res[k].append((bt.mean(), bt.max()))
for k, v in res.items():
    v = np.array(v); print('%-40s block mean %.0f cycles, slowest block %.0f' % (k, v[:, 0].mean(), v[:, 1].mean()))
