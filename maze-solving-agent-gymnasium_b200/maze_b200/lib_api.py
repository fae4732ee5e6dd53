"""The reference's maze-library entry points with their original signatures, computed on the GPU:

  lib/maze_generation.py                 gen_maze :6-35, gen_maze_no_border :37-56,
                                         generate_collection_of_mazes :220-247
  lib/maze_difficulty_evaluation/*       ComplexityEvaluation (difficulty_of_maze :319-329,
                                         complexity_of_maze :310-317), MetricsCalculator
                                         (calculate_L :22-26, calculate_D :71-85, calculate_DE :87-127)

plus batched variants (`gen_mazes`, `maze_metrics`) that return device tensors.  Single calls are
batch-of-1 launches: convenient, not fast -- use the batched forms for throughput.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import cabi
from .engine import ALGO_IDS, MazePool

_DEVICE = "cuda"


def _unpack(v):
    return int(v) & 0xffff, int(v) >> 16


def gen_mazes(num: int, shape, algorithm="dfs", toroidal: bool = False, seed=None, candidates: int = 1, device=None,
              slot_id_base: int = 0):
    """Batched generation: a MazePool of `num` mazes (block grids, step tables, metadata on the device)."""
    if seed is None:
        seed = random.getrandbits(62)
    pool = MazePool(num, tuple(shape), device or _DEVICE)
    pool.generate(shapes=tuple(shape), algorithms=algorithm, toroidal=toroidal, seed=seed, candidates=candidates,
                  slot_id_base=slot_id_base)
    return pool


def gen_maze(shape, algorithm: str = "dfs"):
    """-> (start_point, goal_point, maze) with maze a list of lists, 0 wall / 1 floor / 2 goal."""
    if algorithm not in ALGO_IDS:
        # the reference silently generates nothing and then crashes in find_random_position
        raise ValueError(f"unknown maze generation algorithm {algorithm!r} (expected one of {list(ALGO_IDS)})")
    pool = gen_mazes(1, shape, algorithm)
    meta = pool.meta_host()[0]
    return _unpack(meta[cabi.META_START]), _unpack(meta[cabi.META_GOAL]), pool.grid_host(0).astype(np.int64).tolist()


def gen_maze_no_border(shape, algorithm: str = "dfs"):
    """-> (start_point, goal_point, maze, difficulty): generated at shape + 2, scored, outer ring stripped."""
    if algorithm not in ALGO_IDS:
        raise ValueError(f"unknown maze generation algorithm {algorithm!r} (expected one of {list(ALGO_IDS)})")
    pool = MazePool(1, tuple(shape), _DEVICE)
    diff = torch.zeros(1, dtype=torch.float64, device=pool.device)
    pool.generate(shapes=tuple(shape), algorithms=algorithm, toroidal=True, seed=random.getrandbits(62), difficulty_out=diff)
    meta = pool.meta_host()[0]
    return (_unpack(meta[cabi.META_START]), _unpack(meta[cabi.META_GOAL]), pool.grid_host(0).astype(np.int64).tolist(),
            float(diff.item()))


def generate_collection_of_mazes(shape, num_mazes: int, algorithms=("dfs", "r-prim", "prim&kill")):
    """List of distinct int32 [3, H, W] tensors [wall, tile (== 1), non_visited] (maze_generation.py:220-247)."""
    out, seen = [], set()
    while len(out) < num_mazes:
        need = num_mazes - len(out)
        algos = [random.choice(list(algorithms)) for _ in range(need)]
        pool = MazePool(need, tuple(shape), _DEVICE)
        pool.generate(shapes=tuple(shape), algorithms=algos, seed=random.getrandbits(62))
        H, W = int(shape[0]), int(shape[1])
        grids = pool.grids[:, :H * W].reshape(need, H, W)
        meta = pool.meta_host()
        stack = torch.stack([(grids == 0), (grids == 1), (grids != 0)], dim=1).to(torch.int32)
        for k in range(need):
            sr, sc = _unpack(meta[k, cabi.META_START])
            stack[k, 2, sr, sc] = 0
        host = stack.cpu()
        for k in range(need):
            key = host[k].numpy().tobytes()
            if key not in seen and len(out) < num_mazes:
                seen.add(key)
                out.append(host[k])
    return out


def maze_metrics(mazes, starts, goals, toroidal=False, device=None) -> torch.Tensor:
    """Batched metrics: float64 [n, 8] records (cabi.METRIC_NAMES) for host block grids."""
    pool = MazePool.from_grids([np.asarray(m, dtype=np.uint8) for m in mazes], starts, goals, toroidal, device or _DEVICE)
    return pool.difficulty()


def _goal_of(maze):
    rc = np.argwhere(np.asarray(maze) == 2)
    if len(rc) == 0:
        raise ValueError("maze has no goal block (value 2)")
    return int(rc[0][0]), int(rc[0][1])


class ComplexityEvaluation:
    """McClendon complexity / difficulty of a perfect maze (maze_complexity_evaluation.py:38-329)."""

    def __init__(self, maze, start_pos, goal_pos):
        self.maze, self.start_pos, self.goal_pos = maze, tuple(int(x) for x in start_pos), tuple(int(x) for x in goal_pos)
        rec = maze_metrics([maze], [self.start_pos], [self.goal_pos])[0].cpu().numpy()
        self._difficulty, self._complexity = float(rec[0]), float(rec[1])

    def difficulty_of_maze(self):
        return self._difficulty

    def complexity_of_maze(self):
        return self._complexity


class MetricsCalculator:
    """Kim & Crawfis L / D / DE (metrics_calculator.py:3-173).  The path arguments keep the reference's
    signatures; only their first block (the start) and the maze's goal are needed."""

    def __init__(self, maze, sol_path_length: int):
        self.maze = maze
        self.sol_path_length = sol_path_length
        self.maze_size = (len(maze), len(maze[0]))
        self.goal = _goal_of(maze)
        self.CE = (self.maze_size[0] - 1) * ((self.maze_size[1] - 1) // 2) - 1
        self._rec = {}

    def _record(self, path):
        start = (int(path[0][0]), int(path[0][1]))
        if start not in self._rec:
            self._rec[start] = maze_metrics([self.maze], [start], [self.goal])[0].cpu().numpy()
        return self._rec[start]

    def calculate_L(self, path):
        return len(path) / self.CE

    def calculate_D(self, sol_path):
        return float(self._record(sol_path)[4])

    def calculate_DE(self, sol_path):
        return float(self._record(sol_path)[3])
