"""Golden for the policy_zoo opponent (row N1): deterministic actions / vpred of the reference's ant/mlp/agent-params-v3.npy on a
fixed observation batch, computed with the float64 restatement of policy.py (oracle/ppo_oracle.py::zoo_mlp_act).  The parameter
file itself is a reference asset and is NOT copied; only its checksum, size and these outputs are stored.

    python tests/golden/make_zoo_golden.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle.ppo_oracle import zoo_mlp_act  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/robosumo/robosumo/policy_zoo/assets/ant/mlp/agent-params-v3.npy'

if __name__ == '__main__':
    flat = np.load(SRC)
    rng = np.random.RandomState(0)
    obs = rng.randn(64, 120) * 0.7
    obs[:, 2] += 0.8
    act, v = zoo_mlp_act(flat, obs, 120, 8)
    np.savez_compressed(os.path.join(HERE, 'zoo_ant_v3.npz'), obs=obs, act=act, vpred=v, size=flat.size,
                        sha256=hashlib.sha256(flat.tobytes()).hexdigest())
    print(flat.size, act.shape, abs(act).max(), v[:3])
