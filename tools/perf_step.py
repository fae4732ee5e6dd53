"""Scratch: time maze_step on replicated golden mazes (before the generators exist)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb

z = np.load(os.path.join(ROOT, "tests/golden/metrics.npz"))
meta = json.loads(str(z["meta"]))
ms = [m for m in meta if m["shape"] == 81]
M = 1000
grids = [z[f"m{ms[i % len(ms)]['id']}_grid"] for i in range(M)]
starts = [ms[i % len(ms)]["start"] for i in range(M)]
goals = [ms[i % len(ms)]["goal"] for i in range(M)]
t0 = time.time()
pool = mb.MazePool.from_grids(grids, starts, goals, False)
torch.cuda.synchronize(); print("pool", time.time() - t0)
for B in [int(x) for x in (sys.argv[1:] or ["1048576", "4194304"])]:
    env_maze = (torch.arange(B, device="cuda", dtype=torch.int32) // (B // M)).clamp_(max=M - 1)
    batch = mb.MazeBatch(pool, B, env_maze=env_maze)
    batch.reset()
    K = 64
    acts = torch.randint(0, 4, (K, B), dtype=torch.uint8, device="cuda")
    mode = mb.cabi.STEP_AUTORESET
    for t in range(300):
        batch.step(acts[t % K], mode)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 200
    ev[0].record()
    for t in range(n):
        batch.step(acts[t % K], mode)
    ev[1].record(); torch.cuda.synchronize()
    ms_ = ev[0].elapsed_time(ev[1]) / n
    print(f"B={B} {ms_*1e3:.1f} us/step  {B/ms_*1e3:.3e} steps/s  algorithmic {58*B/ms_/1e6:.0f} GB/s")
    del batch
