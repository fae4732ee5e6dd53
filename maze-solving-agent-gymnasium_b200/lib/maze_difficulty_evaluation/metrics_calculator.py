from maze_b200.lib_api import MetricsCalculator  # noqa: F401
