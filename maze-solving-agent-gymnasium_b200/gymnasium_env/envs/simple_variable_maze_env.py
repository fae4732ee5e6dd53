from maze_b200.single_env import SimpleEnrichVariableMazeEnv, SimpleVariableMazeEnv  # noqa: F401
