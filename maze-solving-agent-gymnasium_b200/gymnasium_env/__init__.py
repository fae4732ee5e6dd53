"""Same six environment ids as the reference's gymnasium_env/__init__.py:3-31 (registered with
gymnasium when it is installed, with the package's own tiny registry otherwise)."""
from maze_b200._gym import register

for _id, _cls in (("MazeEnv-v0", "SimpleMazeEnv"), ("MazeEnv-v1", "SimpleEnrichMazeEnv"),
                  ("VariableMazeEnv-v0", "SimpleVariableMazeEnv"), ("VariableMazeEnv-v1", "SimpleEnrichVariableMazeEnv"),
                  ("ToroidalMazeEnv-v0", "ToroidalMazeEnv"), ("ToroidalMazeEnv-v1", "ToroidalEnrichMazeEnv")):
    register(id=f"gymnasium_env/{_id}", entry_point=f"gymnasium_env.envs:{_cls}")
