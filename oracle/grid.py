"""Block-grid primitives of the oracle (test infrastructure; see oracle/__init__.py).

A maze is a block grid `g[r][c]` with 0 = wall, 1 = floor, 2 = goal
(reference: lib/maze_generation.py:16-19,33).  All functions take a numpy uint8 array [H, W].
"""
from __future__ import annotations

import heapq
from collections import deque

import numpy as np

# reference: gymnasium_env/envs/base_maze_env.py:19-24  (0 down, 1 up, 2 right, 3 left)
ACTIONS = ((1, 0), (-1, 0), (0, 1), (0, -1))
# neighbour expansion order used by both A* variants (lib/a_star_algos/a_star.py:59)
_ASTAR_NEIGH = ((-1, 0), (1, 0), (0, -1), (0, 1))


def as_grid(maze) -> np.ndarray:
    g = np.asarray(maze, dtype=np.uint8)
    assert g.ndim == 2
    return g


def _h_plain(a, b):
    # lib/a_star_algos/a_star.py:3-7
    return abs(a[0] - b[0]) + abs(a[1] - b[1])


def _h_torus(a, b, rows, cols):
    # lib/a_star_algos/a_star_tor.py:3-12
    dx = abs(a[0] - b[0])
    dy = abs(a[1] - b[1])
    return min(dx, rows - dx) + min(dy, cols - dy)


def astar_path(grid: np.ndarray, start, goal, max_depth=10**6, toroidal=False):
    """Depth-limited A* that returns the path to the goal or, failing that, to the first popped
    node of maximal g.  Port of lib/a_star_algos/a_star.py:9-82 (plain) and
    lib/a_star_algos/a_star_tor.py:15-88 (toroidal): same open-set ordering (f, (r, c)),
    same neighbour order, same g-limit tests, same best-candidate rule."""
    rows, cols = grid.shape
    start = (int(start[0]), int(start[1]))
    goal = (int(goal[0]), int(goal[1]))
    if toroidal:
        h = lambda p: _h_torus(p, goal, rows, cols)
    else:
        h = lambda p: _h_plain(p, goal)
    g = {start: 0}
    parent = {}
    heap = [(h(start), start)]
    far_node, far_g = start, 0
    while heap:
        _, cur = heapq.heappop(heap)
        gc = g[cur]
        if gc > far_g:
            far_g, far_node = gc, cur
        if cur == goal:
            far_node = cur
            break
        if gc >= max_depth:
            continue
        for dr, dc in _ASTAR_NEIGH:
            if toroidal:
                nb = ((cur[0] + dr) % rows, (cur[1] + dc) % cols)
            else:
                nb = (cur[0] + dr, cur[1] + dc)
                if not (0 <= nb[0] < rows and 0 <= nb[1] < cols):
                    continue
            if grid[nb[0], nb[1]] == 0:
                continue
            cand = gc + 1
            if cand > max_depth:
                continue
            if nb not in g or cand < g[nb]:
                g[nb] = cand
                parent[nb] = cur
                heapq.heappush(heap, (cand + h(nb), nb))
    out = [far_node]
    while out[-1] in parent:
        out.append(parent[out[-1]])
    out.reverse()
    return out


def astar_len(grid, start, goal, max_depth=10**6, toroidal=False) -> int:
    return len(astar_path(grid, start, goal, max_depth, toroidal))


def bfs_dist(grid: np.ndarray, src, toroidal=False) -> np.ndarray:
    """Unweighted shortest-path distance from `src` to every open block (int32, -1 = unreached).
    Equals len(A*(x -> src)) - 1 wherever A* is unbounded, since both heuristics are consistent."""
    rows, cols = grid.shape
    dist = np.full((rows, cols), -1, dtype=np.int32)
    sr, sc = int(src[0]), int(src[1])
    dist[sr, sc] = 0
    q = deque([(sr, sc)])
    while q:
        r, c = q.popleft()
        d = dist[r, c] + 1
        for dr, dc in _ASTAR_NEIGH:
            nr, nc = r + dr, c + dc
            if toroidal:
                nr %= rows
                nc %= cols
            elif not (0 <= nr < rows and 0 <= nc < cols):
                continue
            if grid[nr, nc] != 0 and dist[nr, nc] < 0:
                dist[nr, nc] = d
                q.append((nr, nc))
    return dist


def open_neighbours(grid: np.ndarray, r, c) -> int:
    """Count of non-wall 4-neighbours, no wrap (used by generation / metrics on bordered mazes)."""
    return int((grid[r - 1, c] != 0) + (grid[r + 1, c] != 0) + (grid[r, c - 1] != 0) + (grid[r, c + 1] != 0))


def depth_limit(shape) -> int:
    # base_maze_env.py:244  max_depth = 2 * min(H, W)
    return 2 * min(int(shape[0]), int(shape[1]))


def max_steps_for(shape, sol_len: int) -> int:
    """Episode step budget.  simple_maze_env.py:52-58 + metrics_calculator.py:16,22-26:
    CE = (H-1)*((W-1)//2) - 1 ; factor = len(path)/CE ; ceil(((H-1)*(W-1) - 1) * factor)."""
    import math
    H, W = int(shape[0]), int(shape[1])
    ce = (H - 1) * ((W - 1) // 2) - 1
    factor = sol_len / ce
    return math.ceil((((H - 1) * (W - 1)) - 1) * factor)


def best_dir_code_table(grid: np.ndarray, goal, toroidal=False, dgoal=None) -> np.ndarray:
    """Closed form of BaseMazeEnv._find_best_next_cell (base_maze_env.py:224-262) for every block.

    Returns uint8 [H, W]: the action index 0..3 whose neighbour is the best next cell, or 4 when
    no neighbour is valid (best dir = (0, 0)).  For a neighbour n:
        p(n) = len(A*(n -> goal, max_depth=L)) = min(D_goal[n], L) + 1        (L = 2*min(H,W))
    because a capped search that misses the goal ends on a node with g = L (the maze is connected
    and D_goal[n] > L implies such a node exists).  score = p + 0.15*manhattan(n, goal) (plain
    Manhattan on the torus too, :249-252); first strict minimum over actions 0..3 wins, which is
    order-equivalent to comparing the integers 20*p + 3*manhattan.
    """
    rows, cols = grid.shape
    if dgoal is None:
        dgoal = bfs_dist(grid, goal, toroidal)
    L = depth_limit(grid.shape)
    gr, gc = int(goal[0]), int(goal[1])
    code = np.full((rows, cols), 4, dtype=np.uint8)
    for r in range(rows):
        for c in range(cols):
            best = None
            for a, (dr, dc) in enumerate(ACTIONS):
                nr, nc = r + dr, c + dc
                if toroidal:
                    nr %= rows
                    nc %= cols
                    if grid[nr, nc] == 0:
                        continue
                else:
                    # simple_maze_env.py:68  0 < r < H and 0 < c < W and maze != 0
                    if not (0 < nr < rows and 0 < nc < cols) or grid[nr, nc] == 0:
                        continue
                d = int(dgoal[nr, nc])
                if d < 0:
                    continue  # unreachable neighbour (not produced by the generators)
                p = min(d, L) + 1
                score = 20 * p + 3 * (abs(nr - gr) + abs(nc - gc))
                if best is None or score < best:
                    best = score
                    code[r, c] = a
    return code


def best_dir_vector(code: int, pos, shape, toroidal: bool):
    """`agent - best_next` (base_maze_env.py:122).  On the torus best_next is wrapped
    (toroidal_maze_env.py:79-81) so components of +-(S-1) appear."""
    r, c = int(pos[0]), int(pos[1])
    if code == 4:
        return (0, 0)
    dr, dc = ACTIONS[code]
    nr, nc = r + dr, c + dc
    if toroidal:
        nr %= int(shape[0])
        nc %= int(shape[1])
    return (r - nr, c - nc)
