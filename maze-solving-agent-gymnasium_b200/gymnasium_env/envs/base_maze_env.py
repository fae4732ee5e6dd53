from maze_b200.single_env import BaseMazeEnv  # noqa: F401
