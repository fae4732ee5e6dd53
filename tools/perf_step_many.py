"""Scratch: maze_step_many (K-step bursts, chunked for L2) vs maze_step, steady state."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb
M, B = 1000, 4096000
pool = mb.MazePool(M, (81, 81)); pool.generate(algorithms="r-prim", seed=1234)
mode = mb.cabi.STEP_AUTORESET | mb.cabi.STEP_WIN_NEXT
env_maze = (torch.arange(B, device="cuda", dtype=torch.int32) // (B // M)).clamp_(max=M - 1)
for lay in (sys.argv[1] if len(sys.argv) > 1 else "cell,tile").split(","):
    batch = mb.MazeBatch(pool, B, env_maze=env_maze.clone(), visit_layout=lay)
    batch.reset()
    for K in (16, 64):
        acts = torch.randint(0, 4, (K, B), dtype=torch.uint8, device="cuda")
        for trace in (False, True):
            for chunk in (65536, 262144, 1048576, 4096000):
                for _ in range(max(2, 320 // K)): batch.step_many(acts, mode, chunk_envs=chunk)
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                reps = max(2, 256 // K)
                out = None
                ev[0].record()
                for _ in range(reps): out = batch.step_many(acts, mode, trace=trace, chunk_envs=chunk)
                ev[1].record(); torch.cuda.synchronize()
                us = ev[0].elapsed_time(ev[1]) / (reps * K) * 1e3
                print(f"layout={lay} K={K} trace={trace} chunk={chunk}: {us:6.1f} us/step-equivalent  {B/us*1e6:.3e} steps/s", flush=True)
                del out
    del batch
