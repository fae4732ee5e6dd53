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
        from .mazeset import collection_tensor
        stack = collection_tensor(pool)     # maze_collection_encode kernel
        host = stack.cpu()
        for k in range(need):
            key = host[k].numpy().tobytes()
            if key not in seen and len(out) < num_mazes:
                seen.add(key)
                out.append(host[k])
    return out


def maze_metrics(mazes, starts, goals, toroidal=False, device=None, extended: bool = False):
    """Batched metrics: float64 [n, 8] records (cabi.METRIC_NAMES) for host block grids; with
    extended=True also the [n, 20] records of cabi.METRIC_EXT_NAMES (MazePool.difficulty)."""
    pool = MazePool.from_grids([np.asarray(m, dtype=np.uint8) for m in mazes], starts, goals, toroidal, device or _DEVICE)
    return pool.difficulty(extended=extended)


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
    """Kim & Crawfis metrics (metrics_calculator.py:3-255).  The path arguments keep the reference's
    signatures and mean the solution path: only its first block (the start) and the maze's goal are
    needed, the kernel walks the tree itself.  `type` is "AC", "FDE" or "BDE"."""

    _TYPES = {"AC": 0, "FDE": 1, "BDE": 2}

    def __init__(self, maze, sol_path_length: int):
        self.maze = maze
        self.sol_path_length = sol_path_length
        self.maze_size = (len(maze), len(maze[0]))
        self.goal = _goal_of(maze)
        self.CE = (self.maze_size[0] - 1) * ((self.maze_size[1] - 1) // 2) - 1
        self._rec, self._ext = {}, {}

    def _record(self, path):
        start = (int(path[0][0]), int(path[0][1]))
        if start not in self._rec:
            rec, ext = maze_metrics([self.maze], [start], [self.goal], extended=True)
            self._rec[start] = rec[0].cpu().numpy()
            self._ext[start] = ext[0].cpu().numpy()
        return self._rec[start]

    def _extended(self, path):
        self._record(path)
        return self._ext[(int(path[0][0]), int(path[0][1]))]

    def _typed(self, path, type, base):
        if type not in self._TYPES:
            return 0     # the reference compares strings: an unknown type matches no dead end
        return float(self._extended(path)[base + self._TYPES[type]])

    def calculate_density(self):
        # needs no path: score from any open cell (the record's density does not depend on the start)
        rc = np.argwhere(np.asarray(self.maze) == 1)
        return float(self._extended([tuple(rc[0])])[0])

    def calculate_T(self, path):
        return float(self._extended(path)[1])

    def calculate_J(self, path):
        return float(self._extended(path)[2])

    def calculate_CR(self, path):
        return float(self._extended(path)[3])

    def calculate_DE_sub(self, path):
        e = self._extended(path)
        return float(e[4]), float(e[5]), float(e[6])

    def calculate_L_DE(self, path):
        return float(self._extended(path)[7])

    def calculate_T_DE(self, path, type):
        return self._typed(path, type, 8)

    def calculate_D_sharp(self, path, type):
        return self._typed(path, type, 11)

    def calculate_L_sharp(self, path, type):
        return self._typed(path, type, 14)

    def calculate_L(self, path):
        return len(path) / self.CE

    def calculate_D(self, sol_path):
        return float(self._record(sol_path)[4])

    def calculate_DE(self, sol_path):
        return float(self._record(sol_path)[3])
