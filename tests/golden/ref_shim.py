"""Import shim for the *unmodified* reference at /root/reference (build container only).

Test infrastructure.  The reference imports `gymnasium`, `pygame` and `matplotlib`, none of
which exist in this image; this module installs minimal stand-ins in `sys.modules` so the
reference's env / generator / metric / agent code can be imported read-only and driven to
produce golden vectors (see make_golden.py).  Nothing here is used by the product path and
nothing here runs on the GPU box (/root/reference does not exist there).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MAZE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gymnasium_env"))


class _Permissive(types.ModuleType):
    """Module whose every (non-dunder) attribute is a callable no-op returning another no-op."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __add__(self, other):
        return self

    __radd__ = __add__


def _install_gymnasium():
    import numpy as np

    gym = types.ModuleType("gymnasium")

    class Env:
        metadata = {}

        def reset(self, seed=None, options=None):
            return None

        def close(self):
            pass

    class Discrete:
        def __init__(self, n):
            self.n = int(n)

        def sample(self):
            return int(np.random.randint(self.n))

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(dict):
        def __init__(self, spaces=None, **kw):
            super().__init__(spaces or {}, **kw)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete, spaces.Box, spaces.Dict = Discrete, Box, Dict
    gym.Env, gym.spaces = Env, spaces

    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")
    registry = {}

    def register(id, entry_point=None, **kw):
        registry[id] = entry_point

    registration.register, registration.registry = register, registry
    envs.registration = registration
    gym.envs = envs
    gym.register = register

    wrappers = types.ModuleType("gymnasium.wrappers")

    class RecordEpisodeStatistics:
        def __init__(self, env, buffer_length=100):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return getattr(self.env, name)

    wrappers.RecordEpisodeStatistics = RecordEpisodeStatistics
    gym.wrappers = wrappers

    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", spaces), ("gymnasium.envs", envs),
                      ("gymnasium.envs.registration", registration), ("gymnasium.wrappers", wrappers)):
        sys.modules[name] = mod


def _install_pygame():
    import numpy as np

    pg = _Permissive("pygame")
    surfarray = _Permissive("pygame.surfarray")
    surfarray.array3d = lambda *_a, **_k: np.zeros((1, 1, 3), dtype=np.uint8)
    pg.surfarray = surfarray
    pg.QUIT = -1
    event = _Permissive("pygame.event")
    event.get = lambda *_a, **_k: []
    pg.event = event
    sys.modules["pygame"] = pg
    sys.modules["pygame.surfarray"] = surfarray
    sys.modules["pygame.event"] = event


def _install_matplotlib():
    mpl = _Permissive("matplotlib")
    plt = _Permissive("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_installed = False


def install():
    """Make `import gymnasium_env...`, `import lib...`, `import agents...` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name, fn in (("gymnasium", _install_gymnasium), ("pygame", _install_pygame),
                     ("matplotlib", _install_matplotlib)):
        try:
            __import__(name)
        except Exception:
            fn()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference_file(relpath: str, modname: str):
    """Import one reference source file by path (site-packages ships an unrelated `agents`
    package that shadows the reference's namespace package of the same name)."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
