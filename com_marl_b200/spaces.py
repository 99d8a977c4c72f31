"""Minimal space / spec objects with the attributes the reference's policy and sampler read
(akro.Discrete.n, akro.Box.flat_dim/low/high/shape, EnvSpec.observation_space/action_space —
garage/envs/env_spec.py, garage/envs/base.py:27-60)."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.flat_dim = int(n)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.flat_dim = int(np.prod(self.shape))
        self.dtype = dtype

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box{self.shape}"


class EnvSpec:
    def __init__(self, observation_space, action_space):
        self.observation_space = observation_space
        self.action_space = action_space
