from maze_b200.agents import DQAgent  # noqa: F401
