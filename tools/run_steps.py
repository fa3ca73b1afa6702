"""Developer driver for ncu captures: E Ant-vs-Ant pairs, N env steps with N(0,1) actions (python tools/run_steps.py [E] [N] [env-id])."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from robosumo_selfplay_b200.vec_env import B200SumoVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 120
env_id = sys.argv[3] if len(sys.argv) > 3 else 'RoboSumo-Ant-vs-Ant-v0'
env = B200SumoVecEnv(env_id, num_envs=E, seed=42, device_api=True)
env.reset(); torch.manual_seed(0)
A = env.action_space[0].shape[0]
for t in range(N):
    env.step(torch.randn(E, 2, A, device='cuda'))
torch.cuda.synchronize()
print('ok', E, N)
