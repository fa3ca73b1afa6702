"""Shared test helpers: seeded states in the reset distribution (sumo.py:232-253) and friends."""
import numpy as np


def reset_like_state(om, rng, spread=1.15, z=1.25):
    nq, nv = om.nq, om.nv
    q = om.qpos0.copy()
    phi = rng.uniform(0, 2 * np.pi)
    for a in range(2):
        o = a * (nq // 2)
        q[o] = spread * np.cos(phi + a * np.pi)
        q[o + 1] = spread * np.sin(phi + a * np.pi)
        q[o + 2] = z
    q = q + rng.uniform(-0.1, 0.1, nq)
    v = 0.1 * rng.randn(nv)
    om.normalize_qpos(q)
    return q, v


def settled_states(om, rng, n, steps=30, action_scale=0.5, spread=1.15):
    """n states reached by stepping the oracle from reset-like states (contacts exist)."""
    qs, vs = [], []
    for _ in range(n):
        q, v = reset_like_state(om, rng, spread=spread)
        w = np.zeros(om.nv)
        for _ in range(steps):
            om.step(q, v, action_scale * rng.randn(om.nu), 5, w)
        qs.append(q.copy())
        vs.append(v.copy())
    return np.array(qs), np.array(vs)
