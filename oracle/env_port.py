"""Env step/reset oracle (test infrastructure; see oracle/__init__.py).

PortEnv        the reference's own algorithm, restated: A* for the reward and for every
               candidate of the "best dir" observation, Python list of visited cells.
               Follows gymnasium_env/envs/base_maze_env.py:116-262, lib/maze_view.py:165-197,
               simple_maze_env.py:38-79, toroidal_maze_env.py:57-98, lib/maze_handler.py:4-162.
ClosedFormEnv  the O(1) restatement the CUDA kernels implement: per-maze byte table
               (open | best-dir code | D_goal mod 4), per-env visit counters, host-built
               float64 penalty LUTs.  Differentially tested against PortEnv and the goldens.

Both return the reference's *swapped* tuple (obs, reward, truncated, terminated, info)
(base_maze_env.py:210).
"""
from __future__ import annotations

import math

import numpy as np

from .grid import (ACTIONS, as_grid, astar_len, bfs_dist, best_dir_code_table, best_dir_vector,
                   depth_limit, max_steps_for)

WINDOW = 15  # simple_maze_env.py:130, toroidal_maze_env.py:165


def revisit_penalty(count: int) -> float:
    # base_maze_env.py:171,194   reward = 0.0 ; reward -= 1 - exp(-0.2 * count)
    return 0.0 - (1 - math.exp(-0.2 * count))


def invalid_penalty(k: int) -> float:
    # base_maze_env.py:171,200   reward = 0.0 ; reward -= 1 - exp(-0.15 * k)
    return 0.0 - (1 - math.exp(-0.15 * k))


def shaping_reward(delta: int) -> float:
    # base_maze_env.py:192   (old_dist - new_dist) * 0.5 - 0.05
    return delta * 0.5 - 0.05


def window_origin(pos, shape, toroidal):
    """Top-left block of the 15x15 crop.  Euclid (maze_handler.py:4-54): centred, clamped into the
    grid (so not centred near edges); the whole maze when H == 15.  Torus (:56-80): centred, wraps."""
    k = WINDOW // 2
    if toroidal:
        return int(pos[0]) - k, int(pos[1]) - k
    H = int(shape[0])
    if H == WINDOW:
        return 0, 0
    r0 = min(max(int(pos[0]) - k, 0), H - WINDOW)
    c0 = min(max(int(pos[1]) - k, 0), H - WINDOW)  # the reference uses len(maze) for both axes
    return r0, c0


def window_tensor(grid, non_visited, pos, toroidal):
    """float32 [3,15,15] = [maze==0, maze==1, non_visited]  (maze_handler.py:82-99)."""
    H, W = grid.shape
    r0, c0 = window_origin(pos, grid.shape, toroidal)
    rows = [(r0 + i) % H if toroidal else r0 + i for i in range(WINDOW)]
    cols = [(c0 + i) % W if toroidal else c0 + i for i in range(WINDOW)]
    sub = grid[np.ix_(rows, cols)]
    nv = non_visited[np.ix_(rows, cols)]
    return np.stack([(sub == 0), (sub == 1), nv != 0]).astype(np.float32)


def direction_mask(grid, pos, toroidal):
    """int32[4] in action order; 1 where the neighbour block is not a wall
    (maze_handler.py:122-141 euclid, :143-162 torus)."""
    H, W = grid.shape
    out = np.ones(4, dtype=np.int32)
    for k, (dr, dc) in enumerate(ACTIONS):
        nr, nc = int(pos[0]) + dr, int(pos[1]) + dc
        if toroidal:
            nr %= H
            nc %= W
        if grid[nr, nc] == 0:
            out[k] = 0
    return out


def back_direction_index(last_move: int, toroidal: bool) -> int:
    """Index that get_mask_direction(probs=True) overwrites with 0.25.
    Euclid (simple_maze_env.py:45-49): previous - current = -ACTIONS[last_move].
    Torus (toroidal_maze_env.py:61-68): the (dx, dy) tuple is built column-first, so the 0.25
    lands on a rotated index: down->left(3), up->right(2), right->up(1), left->down(0)."""
    if not toroidal:
        return (1, 0, 3, 2)[last_move]
    return (3, 2, 1, 0)[last_move]


class _EnvBase:
    def __init__(self, grid, start, goal, toroidal=False, enrich=False):
        self.grid = as_grid(grid)
        self.shape = self.grid.shape
        self.start = (int(start[0]), int(start[1]))
        self.goal = (int(goal[0]), int(goal[1]))
        self.toroidal = bool(toroidal)
        self.enrich = bool(enrich)
        self.max_steps = 0
        self.pos = self.start
        self.steps = 0
        self.consec_invalid = 0
        self.last_move = -1
        self.n_moves = 0

    # -- movement rule: lib/maze_view.py:167-180 (euclid), :184-197 (torus)
    def _try_move(self, action):
        dr, dc = ACTIONS[int(action)]
        H, W = self.shape
        nr, nc = self.pos[0] + dr, self.pos[1] + dc
        if self.toroidal:
            nr %= H
            nc %= W
            ok = self.grid[nr, nc] != 0
        else:
            ok = (0 < nr < H - 1) and (0 < nc < W - 1) and self.grid[nr, nc] != 0
        return ok, (nr, nc)

    def _info(self):
        # base_maze_env.py:124-134
        return {"distance": float(abs(self.pos[0] - self.goal[0]) + abs(self.pos[1] - self.goal[1]))}

    def mask_direction(self, probs=False):
        m = direction_mask(self.grid, self.pos, self.toroidal)
        if probs and self.n_moves > 1:
            m = m.astype(np.float32)
            m[back_direction_index(self.last_move, self.toroidal)] = 0.25
        return m

    def _format_obs(self, best_vec, non_visited):
        agent = np.array(self.pos, dtype=np.int32)
        target = np.array(self.goal, dtype=np.int32)
        best = np.array(best_vec, dtype=np.int64)
        if not self.enrich:
            return {"agent": agent, "target": target, "best dir": best}
        shp = np.array(self.shape)
        return {"agent": agent / shp, "target": target / shp, "best dir": best,
                "window": window_tensor(self.grid, non_visited, self.pos, self.toroidal)}


class PortEnv(_EnvBase):
    """The reference's algorithm, A* and all."""

    def __init__(self, grid, start, goal, toroidal=False, enrich=False):
        super().__init__(grid, start, goal, toroidal, enrich)
        self.set_max_steps()
        self.reset()

    def _path_len(self, src, max_depth=10**6):
        return astar_len(self.grid, src, self.goal, max_depth, self.toroidal)

    def set_max_steps(self):
        self.max_steps = max_steps_for(self.shape, self._path_len(self.start))

    def _valid(self, p):
        H, W = self.shape
        if self.toroidal:
            return self.grid[p[0], p[1]] != 0            # toroidal_maze_env.py:83-87
        return 0 < p[0] < H and 0 < p[1] < W and self.grid[p[0], p[1]] != 0  # simple_maze_env.py:68

    def _best_next(self):
        # base_maze_env.py:224-262
        H, W = self.shape
        best, best_score = self.pos, float("inf")
        for a, (dr, dc) in enumerate(ACTIONS):
            n = (self.pos[0] + dr, self.pos[1] + dc)
            if self.toroidal:
                n = (n[0] % H, n[1] % W)
            if not self._valid(n):
                continue
            plen = self._path_len(n, depth_limit(self.shape))
            if plen:
                score = plen + 0.15 * (abs(n[0] - self.goal[0]) + abs(n[1] - self.goal[1]))
                if score < best_score:
                    best_score, best = score, n
            if n == self.goal:
                return n
        return best

    def _obs(self):
        b = self._best_next()
        return self._format_obs((self.pos[0] - b[0], self.pos[1] - b[1]), self.non_visited)

    def reset(self):
        # base_maze_env.py:136-161
        self.pos = self.start
        self.non_visited = (self.grid != 0).astype(np.int32)
        self.non_visited[self.start] = 0
        obs, info = self._obs(), self._info()
        self.steps = 0
        self.consec_invalid = 0
        self.visited = []
        self.last_move, self.n_moves = -1, 0
        return obs, info

    def step(self, action):
        # base_maze_env.py:163-210
        reward, terminated, truncated = 0.0, False, False
        prev = self.pos
        moved, new = self._try_move(action)
        if moved:
            self.pos = new
            self.consec_invalid = 0
            self.last_move = int(action)
            self.n_moves += 1
            if new not in self.visited:
                self.non_visited[new] = 0
                if new == self.goal:
                    reward, terminated = 1, True
                else:
                    reward = (self._path_len(prev) - self._path_len(new)) * 0.5 - 0.05
            else:
                reward -= 1 - math.exp(-0.2 * self.visited.count(new))
            self.visited.append(new)
        else:
            self.consec_invalid += 1
            reward -= 1 - math.exp(-0.15 * self.consec_invalid)
        obs, info = self._obs(), self._info()
        self.steps += 1
        if self.steps > self.max_steps:
            truncated, reward = True, -1
        return obs, reward, truncated, terminated, info


class MazeTables:
    """Per-maze precomputation of the closed form (what the `maze_fields` kernel produces)."""

    def __init__(self, grid, start, goal, toroidal=False):
        self.grid = as_grid(grid)
        self.start = (int(start[0]), int(start[1]))
        self.goal = (int(goal[0]), int(goal[1]))
        self.toroidal = bool(toroidal)
        self.dgoal = bfs_dist(self.grid, self.goal, self.toroidal)
        self.code = best_dir_code_table(self.grid, self.goal, self.toroidal, self.dgoal)
        self.max_steps = max_steps_for(self.grid.shape, int(self.dgoal[self.start]) + 1)
        # one byte per block: bit0 open | bits1-3 best-dir code | bits4-5 D_goal mod 4
        d4 = (np.where(self.dgoal >= 0, self.dgoal, 0) & 3).astype(np.uint8)
        # (wall blocks carry 0: the step never reads more than their open bit)
        opn = (self.grid != 0).astype(np.uint8)
        self.table = (opn | ((self.code << 1) | (d4 << 4)) * opn).astype(np.uint8)


class ClosedFormEnv(_EnvBase):
    """O(1) step driven by MazeTables; the specification of the CUDA step kernel."""

    REVISIT_LUT = [revisit_penalty(c) for c in range(256)]
    INVALID_LUT = [invalid_penalty(k) for k in range(256)]

    def __init__(self, grid, start, goal, toroidal=False, enrich=False, tables=None):
        super().__init__(grid, start, goal, toroidal, enrich)
        self.t = tables or MazeTables(grid, start, goal, toroidal)
        self.max_steps = self.t.max_steps
        self.reset()

    @property
    def non_visited(self):
        nv = ((self.grid != 0) & (self.count == 0)).astype(np.int32)
        nv[self.start] = 0
        return nv

    def _obs(self):
        code = int(self.t.code[self.pos])
        return self._format_obs(best_dir_vector(code, self.pos, self.shape, self.toroidal), self.non_visited)

    def reset(self):
        self.pos = self.start
        self.count = np.zeros(self.shape, dtype=np.uint8)   # saturating visit counters
        self.steps = 0
        self.consec_invalid = 0
        self.last_move, self.n_moves = -1, 0
        return self._obs(), self._info()

    def step(self, action):
        reward, terminated, truncated = 0.0, False, False
        prev = self.pos
        moved, new = self._try_move(action)
        if moved:
            self.pos = new
            self.consec_invalid = 0
            self.last_move = int(action)
            self.n_moves += 1
            c = int(self.count[new])
            if c == 0:
                if new == self.goal:
                    reward, terminated = 1, True
                else:
                    delta = (int(self.t.table[prev] >> 4) - int(self.t.table[new] >> 4)) & 3
                    delta = {0: 0, 1: 1, 3: -1}[delta]
                    reward = shaping_reward(delta)
            else:
                reward = self.REVISIT_LUT[c]
            if c < 255:
                self.count[new] = c + 1
        else:
            self.consec_invalid = min(self.consec_invalid + 1, 255)
            reward = self.INVALID_LUT[self.consec_invalid]
        obs, info = self._obs(), self._info()
        self.steps += 1
        if self.steps > self.max_steps:
            truncated, reward = True, -1
        return obs, reward, truncated, terminated, info
