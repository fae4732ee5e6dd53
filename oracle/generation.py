"""Maze generation oracle (test infrastructure; see oracle/__init__.py).

Restates lib/maze_generation.py of the reference: the three generators (same sampling
distributions, own RNG stream -- the reference draws from Python's global `random` through set
iteration order, which no other implementation can replay), the deterministic goal selection,
border stripping for toroidal mazes, plus structural validity checks.
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np

from .grid import bfs_dist

ALGORITHMS = ("r-prim", "dfs", "prim&kill")
_CELL_STEPS = ((-2, 0), (2, 0), (0, -2), (0, 2))


def _cell_neighbours(r, c, H, W):
    # lib/maze_generation.py:70-73: 0 <= x+dx < height and 0 <= y+dy < width, on odd coordinates
    return [(r + dr, c + dc) for dr, dc in _CELL_STEPS if 0 <= r + dr < H and 0 <= c + dc < W]


def random_prim(grid, start, rng):
    """lib/maze_generation.py:59-99: uniformly random frontier cell, joined through a uniformly
    random neighbour that is already in the maze."""
    H, W = grid.shape
    grid[start] = 1
    frontier = []
    in_frontier = set()
    for n in _cell_neighbours(*start, H, W):
        frontier.append(n); in_frontier.add(n)
    while frontier:
        k = rng.randrange(len(frontier))
        f = frontier[k]
        frontier[k] = frontier[-1]
        frontier.pop()
        in_frontier.discard(f)
        nbrs = [n for n in _cell_neighbours(*f, H, W) if grid[n] == 1]
        if nbrs:
            n = nbrs[rng.randrange(len(nbrs))]
            grid[f] = 1
            grid[(f[0] + n[0]) // 2, (f[1] + n[1]) // 2] = 1
            for q in _cell_neighbours(*f, H, W):
                if grid[q] == 0 and q not in in_frontier:
                    frontier.append(q); in_frontier.add(q)


def depth_first(grid, start, rng):
    """lib/maze_generation.py:101-128: recursive backtracker; the four directions are reshuffled
    on every visit of the stack top, i.e. the next cell is uniform over the unvisited neighbours."""
    H, W = grid.shape
    grid[start] = 1
    stack = [start]
    while stack:
        r, c = stack[-1]
        cand = [n for n in _cell_neighbours(r, c, H, W) if grid[n] == 0]
        if cand:
            n = cand[rng.randrange(len(cand))]
            grid[(r + n[0]) // 2, (c + n[1]) // 2] = 1
            grid[n] = 1
            stack.append(n)
        else:
            stack.pop()


def prim_and_kill(grid, start, rng):
    """lib/maze_generation.py:130-185: random walk ("kill") through unmarked cells until stuck,
    then restart from a uniformly random marked cell that still has an unmarked neighbour."""
    H, W = grid.shape
    cells = [(r, c) for r in range(1, H, 2) for c in range(1, W, 2)]
    for rc in cells:
        grid[rc] = 1
    marked = np.zeros((H, W), dtype=bool)
    marked[start] = True
    n_unmarked = len(cells) - 1

    def walk(cur):
        nonlocal n_unmarked
        while True:
            cand = [n for n in _cell_neighbours(*cur, H, W) if not marked[n]]
            if not cand:
                return
            n = cand[rng.randrange(len(cand))]
            grid[(cur[0] + n[0]) // 2, (cur[1] + n[1]) // 2] = 1
            cur = n
            marked[cur] = True
            n_unmarked -= 1

    walk(start)
    while n_unmarked:
        elig = [rc for rc in cells if marked[rc] and any(not marked[n] for n in _cell_neighbours(*rc, H, W))]
        walk(elig[rng.randrange(len(elig))])


def select_goal(grid, start):
    """lib/maze_generation.py:187-218: among logical cells != start with exactly one open
    neighbour, the one farthest from start; the first in row-major order wins ties."""
    H, W = grid.shape
    d = bfs_dist(grid, start, toroidal=False)
    best, best_d = None, -1
    for r in range(1, H, 2):
        for c in range(1, W, 2):
            if (r, c) == tuple(start) or grid[r, c] != 1:
                continue
            nb = int(grid[r - 1, c] != 0) + int(grid[r + 1, c] != 0) + int(grid[r, c - 1] != 0) + int(grid[r, c + 1] != 0)
            if nb == 1 and d[r, c] > best_d:
                best, best_d = (r, c), int(d[r, c])
    return best


def gen_maze(shape, algorithm="dfs", rng=None):
    """lib/maze_generation.py:6-35 -> (start, goal, grid uint8 [H, W])."""
    rng = rng or random.Random()
    H, W = int(shape[0]), int(shape[1])
    if H % 2 == 0 or W % 2 == 0:
        raise ValueError("block shape must be odd")
    grid = np.zeros((H, W), dtype=np.uint8)
    start = (rng.randrange(1, H - 1, 2), rng.randrange(1, W - 1, 2))
    {"r-prim": random_prim, "dfs": depth_first, "prim&kill": prim_and_kill}[algorithm](grid, start, rng)
    goal = select_goal(grid, start)
    grid[goal] = 2
    return start, goal, grid


def strip_border(start, goal, grid):
    """lib/maze_generation.py:53-55."""
    return (start[0] - 1, start[1] - 1), (goal[0] - 1, goal[1] - 1), np.ascontiguousarray(grid[1:-1, 1:-1])


def gen_maze_no_border(shape, algorithm="dfs", rng=None):
    start, goal, grid = gen_maze((shape[0] + 2, shape[1] + 2), algorithm, rng)
    return strip_border(start, goal, grid)


# ------------------------------------------------------------------------------------------------
# structural checks

def check_perfect_maze(grid):
    """A bordered block grid is a perfect maze iff: cells (odd, odd) all open, pillars
    (even, even) and the border all wall, the passages form a spanning tree of the N x N cell
    lattice (N^2 - 1 passages, connected => acyclic).  Returns (ok, reason)."""
    g = np.asarray(grid)
    H, W = g.shape
    if H % 2 == 0 or W % 2 == 0:
        return False, "even shape"
    n_r, n_c = (H - 1) // 2, (W - 1) // 2
    if (g[0, :] != 0).any() or (g[-1, :] != 0).any() or (g[:, 0] != 0).any() or (g[:, -1] != 0).any():
        return False, "border not wall"
    if (g[1::2, 1::2] == 0).any():
        return False, "closed logical cell"
    if (g[0::2, 0::2] != 0).any():
        return False, "open pillar"
    passages = int((g[1::2, 2:-1:2] != 0).sum() + (g[2:-1:2, 1::2] != 0).sum())
    if passages != n_r * n_c - 1:
        return False, f"{passages} passages, expected {n_r * n_c - 1}"
    d = bfs_dist((g != 0).astype(np.uint8), (1, 1))
    if (d[1::2, 1::2] < 0).any():
        return False, "not connected"
    if int((g == 2).sum()) > 1:
        return False, "more than one goal"
    return True, "ok"


def maze_shape_stats(grid, start, goal):
    """Cheap distribution fingerprints of a bordered maze: solution length in blocks, number of
    dead-end cells, number of junction cells (>= 3 open neighbours)."""
    g = np.asarray(grid)
    d = bfs_dist(g, goal)
    op = (g != 0).astype(np.int32)
    nb = op[:-2, 1:-1] + op[2:, 1:-1] + op[1:-1, :-2] + op[1:-1, 2:]
    cells = nb[0::2, 0::2]
    return dict(sol_len=int(d[tuple(start)]) + 1, dead_ends=int((cells == 1).sum()), junctions=int((cells >= 3).sum()))
