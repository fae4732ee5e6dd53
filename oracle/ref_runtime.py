"""Run the UNMODIFIED reference env on the host cores (bench infrastructure; see oracle/__init__.py).

baseline/_ref holds a plain copy of the reference's Python tree (tools/install_reference.py; git-ignored, shipped to
the GPU box with the snapshot -- the reference has nothing to pip-install).  gymnasium / pygame / matplotlib are absent
from this image, so the stand-ins of tests/golden/ref_shim.py are installed first: pygame is stubbed to no-ops, which
makes these timings an UPPER bound on the reference's speed.  Nothing here is imported by the
product package.
"""
from __future__ import annotations

import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "gymnasium_env", "envs", "simple_maze_env.py"))


def _shim():
    os.environ["MAZE_REFERENCE_ROOT"] = REF_DIR
    spec = importlib.util.spec_from_file_location("_maze_ref_shim", os.path.join(ROOT, "tests", "golden", "ref_shim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_reference_env(shape=(81, 81), algorithm="r-prim", seed=0):
    """gymnasium_env.envs.simple_maze_env.SimpleMazeEnv(shape) of the reference: its own generator (best of six by
    McClendon difficulty), A* per step.  Call in a fresh worker process: it puts the reference first on sys.path."""
    import random
    shim = _shim()
    for name in [m for m in sys.modules if m == "gymnasium_env" or m.startswith("gymnasium_env.") or m == "lib" or m.startswith("lib.")]:
        del sys.modules[name]
    sys.path = [p for p in sys.path if "maze-solving-agent-gymnasium_b200" not in p]
    shim.install()
    from gymnasium_env.envs.base_maze_env import BaseMazeEnv
    from gymnasium_env.envs.simple_maze_env import SimpleMazeEnv
    assert REF_DIR in SimpleMazeEnv.__module__ or REF_DIR in sys.modules[SimpleMazeEnv.__module__].__file__
    random.seed(seed)
    BaseMazeEnv.ALGORITHM = algorithm
    return SimpleMazeEnv(tuple(shape))
