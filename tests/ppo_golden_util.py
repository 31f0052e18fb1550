"""Helpers shared by tests/golden/make_golden.py and the PPO parity tests."""
import numpy as np


def seeded_weights(shapes, seed=0):
    """Platform-independent initial ActorCritic weights: N(0, 1/fan_in) from numpy's PCG64 (std = 1 -> ones)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in shapes.items():
        if name == "std":
            out[name] = np.ones(shape, dtype=np.float32)
        elif len(shape) == 2:
            out[name] = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(np.float32)
        else:
            out[name] = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    return out


def sample_tensor(a, stride=97):
    """Compact fingerprint of a tensor: every `stride`-th element, plus sum and sum of squares (float64)."""
    f = np.asarray(a, dtype=np.float32).ravel()
    return np.concatenate([f[::stride].astype(np.float64), [f.astype(np.float64).sum(), (f.astype(np.float64) ** 2).sum()]])
