from maze_b200.lib_api import ComplexityEvaluation, MetricsCalculator, maze_metrics  # noqa: F401
