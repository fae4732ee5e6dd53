from maze_b200.single_env import SimpleEnrichMazeEnv, SimpleMazeEnv  # noqa: F401
