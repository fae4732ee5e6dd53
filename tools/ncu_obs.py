#!/usr/bin/env python
"""A few -v1 steps with replay push on 262 144 envs (the bench's extras configuration): target for ncu captures of
maze_window_kernel / maze_dqn_push_kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb  # noqa: E402
from maze_b200.dqn import DeviceReplay  # noqa: E402

B = 262144
env = mb.MazeVectorEnv(B, shape=(81, 81), num_mazes=1000, enrich=True, seed=1234, on_win="next", stats=False)
env.reset()
mem = DeviceReplay(env, 1 << 20, seed=1)
mem.observe()
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(40):
    acts = torch.randint(0, 4, (B,), dtype=torch.uint8, device="cuda", generator=g)
    env.step(acts)
    mem.push(acts)
torch.cuda.synchronize()
