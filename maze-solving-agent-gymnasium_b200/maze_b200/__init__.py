"""B200-native batched maze environment: host side (Python + PyTorch tensors) over the C ABI of
libmaze_b200.so (hand-written sm_100a CUDA).  There is no CPU fallback: importing `cabi` fails
loudly when the library has not been built, and every compute entry point needs a CUDA device.
"""
from . import cabi, dist  # noqa: F401
from .engine import ALGO_IDS, MazeBatch, MazePool  # noqa: F401
from . import mazeset  # noqa: F401
from .vector_env import MazeVectorEnv  # noqa: F401

__all__ = ["cabi", "dist", "mazeset", "MazePool", "MazeBatch", "MazeVectorEnv", "ALGO_IDS"]
