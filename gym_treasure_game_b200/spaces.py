"""Minimal stand-ins for ``gym.spaces.Discrete`` / ``Box`` (gym is optional)."""
from __future__ import annotations

import numpy as np


class Discrete:
    def __init__(self, n, seed=None):
        self.n = int(n)
        self._rng = np.random.default_rng(seed)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = dtype

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s)" % (self.low, self.high, self.shape)
