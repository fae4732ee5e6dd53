from maze_b200.single_env import ToroidalEnrichMazeEnv, ToroidalMazeEnv  # noqa: F401
