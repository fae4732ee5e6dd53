"""Scratch: a few maze_difficulty launches for an ncu capture (r-prim 81x81)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb
pool = mb.MazePool(16384, (81, 81)); pool.generate(algorithms="r-prim", seed=3)
for _ in range(3): pool.difficulty()
torch.cuda.synchronize()
