#!/usr/bin/env python
"""maze_generate + maze_difficulty on 32 768 81x81 mazes per generator: target for the ncu capture of the shipped
(384-thread) maze_difficulty_kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb  # noqa: E402

pool = mb.MazePool(32768, (81, 81))
for algo in ("r-prim", "dfs", "prim&kill"):
    pool.generate(algorithms=algo, seed=3)
    for _ in range(2):
        pool.difficulty()
torch.cuda.synchronize()
