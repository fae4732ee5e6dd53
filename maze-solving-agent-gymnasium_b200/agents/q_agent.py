from maze_b200.agents import QAgent  # noqa: F401
