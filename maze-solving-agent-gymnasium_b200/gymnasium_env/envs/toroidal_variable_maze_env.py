from maze_b200.single_env import ToroidalEnrichVariableMazeEnv, ToroidalVariableMazeEnv  # noqa: F401
