"""Minimal stand-in for `gymnasium` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference
environment under /root/reference can be imported by oracle/ref_harness.py to
generate golden vectors.  Only what `src/environment/uav_env.py:6-7,240,288,
347-355,407` touches is provided.  `Env.reset(seed)` mirrors what gymnasium
1.2.2 does: `np_random = Generator(PCG64(SeedSequence(seed)))`.
"""
import numpy as np
from . import spaces  # noqa: F401


class _TappedGenerator:
    """Forwards to a numpy Generator; `uniform` calls are reported to a tap so the
    harness can record the reference's initial-buffer-fill draws (uav_env.py:410)."""

    def __init__(self, gen, tap):
        self._gen = gen
        self._tap = tap

    def uniform(self, *a, **k):
        v = self._gen.uniform(*a, **k)
        if self._tap is not None:
            self._tap(v)
        return v

    def __getattr__(self, name):
        return getattr(self._gen, name)


class Env:
    metadata = {}
    np_random = None
    _uniform_tap = None  # set by oracle/ref_harness.py

    def __init__(self, *a, **k):
        pass

    def reset(self, *, seed=None, options=None):
        if seed is not None or self.np_random is None:
            gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
            self.np_random = _TappedGenerator(gen, type(self)._uniform_tap)
        return None

    def close(self):
        pass
