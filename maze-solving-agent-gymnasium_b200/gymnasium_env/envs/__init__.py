"""Entry points of the registered ids (the reference ships no envs/__init__.py, so its own
registration cannot resolve; this one can)."""
from maze_b200.single_env import (BaseMazeEnv, SimpleEnrichMazeEnv, SimpleEnrichVariableMazeEnv, SimpleMazeEnv,  # noqa: F401
                                  SimpleVariableMazeEnv, ToroidalEnrichMazeEnv, ToroidalEnrichVariableMazeEnv,
                                  ToroidalMazeEnv, ToroidalVariableMazeEnv)
from maze_b200.vector_env import MazeVectorEnv  # noqa: F401
