"""gymnasium when it is installed, otherwise the few names the env classes need (the build image
and the GPU boxes have no gymnasium; the reference pins 1.0.0, README.md:8)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register
    HAVE_GYMNASIUM = True
    Env = gym.Env
    VectorEnv = gym.vector.VectorEnv
except Exception:
    HAVE_GYMNASIUM = False

    class Env:
        metadata: dict = {}

        def close(self):
            pass

    class VectorEnv:
        pass

    class _Space:
        def __init__(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class _Discrete(_Space):
        def __init__(self, n, seed=None):
            super().__init__(seed)
            self.n = int(n)

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class _Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            super().__init__(seed)
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = tuple(shape) if shape is not None else np.shape(np.asarray(low))

    class _Dict(_Space, dict):
        def __init__(self, spaces_, seed=None):
            _Space.__init__(self, seed)
            dict.__init__(self, spaces_)
            self.spaces = self

    class spaces:  # noqa: N801 - mirrors the gymnasium module name
        Discrete, Box, Dict = _Discrete, _Box, _Dict

    _REGISTRY = {}

    def register(id, entry_point, **kwargs):  # noqa: A002 - gymnasium's signature
        _REGISTRY[id] = (entry_point, kwargs)

    def make(id, **kwargs):  # noqa: A002
        """Tiny gymnasium.make for the ids registered here."""
        import importlib
        entry, kw = _REGISTRY[id]
        mod, cls = entry.split(":")
        return getattr(importlib.import_module(mod), cls)(**{**kw, **kwargs})
