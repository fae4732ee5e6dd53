"""Oracle generators: structural validity, exact goal selection vs the reference, and output
distribution vs samples drawn from the reference generators (tests/golden/genstats.npz)."""
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle.generation import ALGORITHMS, check_perfect_maze, gen_maze, gen_maze_no_border, maze_shape_stats, select_goal


def mean_close(a, b, nsig=4.5):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    return abs(a.mean() - b.mean()) <= nsig * se + 1e-9, (a.mean(), b.mean(), se)


def test_goal_selection_matches_reference(golden_metrics):
    z, meta = golden_metrics
    n = 0
    for m in meta:
        if m["algo"] == "literal" or m["no_border"]:
            continue
        grid = z[f"m{m['id']}_grid"].copy()
        grid[grid == 2] = 1
        assert select_goal(grid, tuple(m["start"])) == tuple(m["goal"]), m["id"]
        n += 1
    assert n > 40


@pytest.mark.parametrize("algo", ALGORITHMS)
def test_oracle_generators_make_perfect_mazes(algo):
    rng = random.Random(3)
    for shape in ((5, 5), (11, 11), (21, 21), (41, 41)):
        for _ in range(5):
            start, goal, grid = gen_maze(shape, algo, rng)
            ok, why = check_perfect_maze(grid)
            assert ok, why
            assert grid[goal] == 2 and start != goal
            g1 = grid.copy(); g1[goal] = 1
            assert select_goal(g1, start) == goal
    s, g, grid = gen_maze_no_border((15, 15), algo, rng)
    assert grid.shape == (15, 15) and grid[g] == 2 and grid[s] != 0


@pytest.mark.parametrize("algo", ALGORITHMS)
@pytest.mark.parametrize("shape", [21, 41])
def test_oracle_generator_distribution_matches_reference(algo, shape):
    ref = np.load(f"{GOLDEN}/genstats.npz")[f"{algo}_{shape}"]
    rng = random.Random(1234 + shape)
    rows = []
    for _ in range(300 if shape == 21 else 120):
        start, goal, grid = gen_maze((shape, shape), algo, rng)
        st = maze_shape_stats(grid, start, goal)
        rows.append((st["sol_len"], st["dead_ends"], st["junctions"], start[0], start[1]))
    rows = np.array(rows)
    for col, name in enumerate(("sol_len", "dead_ends", "junctions", "start_r", "start_c")):
        ok, info = mean_close(rows[:, col], ref[:, col])
        assert ok, (algo, shape, name, info)
