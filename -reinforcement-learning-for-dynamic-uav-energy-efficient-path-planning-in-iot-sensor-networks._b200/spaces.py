"""Observation / action space objects.

The reference declares `spaces.Discrete(5)` and `spaces.Box(-1, 1, (3+fps*N,), float32)`
(uav_env.py:347-355; the padded wrapper widens the Box to +-inf, dqn.py:252-254).  When
gymnasium is installed its classes are used so SB3 accepts them; otherwise these minimal
stand-ins expose the attributes callers read (`n`, `shape`, `dtype`, `low`, `high`, `sample`).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium.spaces import Box, Discrete  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)
            self._rng = np.random.default_rng()

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s)" % (self.shape, self.dtype)
