"""Scratch: a few maze_window launches at steady state for an ncu capture."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
venv = mb.MazeVectorEnv(B, shape=(81, 81), num_mazes=1000, enrich=True, seed=1234, on_win="next", stats=False)
venv.reset()
acts = torch.randint(0, 4, (B,), dtype=torch.uint8, device="cuda")
for _ in range(300): venv.batch.step(acts, venv._mode)
for _ in range(4): venv.batch.compute_window()
torch.cuda.synchronize()
