from maze_b200.lib_api import gen_maze, gen_maze_no_border, gen_mazes, generate_collection_of_mazes  # noqa: F401
