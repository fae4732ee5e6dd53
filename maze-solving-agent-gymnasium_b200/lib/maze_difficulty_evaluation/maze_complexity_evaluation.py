from maze_b200.lib_api import ComplexityEvaluation  # noqa: F401
