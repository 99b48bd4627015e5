"""Minimal ``gym.spaces`` stand-ins (gym is not a dependency of this package).

Only what the PVDER-v0 contract uses: ``Discrete(5)`` and ``Box(-10, 10, (11,), float32)``
(reference gym_PVDER/envs/PVDER_env.py:50-51), with ``contains``/``sample``/``n``/``shape``."""
import numpy as np


class Discrete:
    def __init__(self, n, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        if isinstance(x, (bool, np.bool_)):
            return False
        if isinstance(x, (int, np.integer)):
            return 0 <= int(x) < self.n
        if isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            return 0 <= int(x) < self.n
        return False

    __contains__ = contains

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self.n


class Box:
    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        if x.shape != self.shape or not np.can_cast(x.dtype, self.dtype, casting="same_kind"):
            return False
        return bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
