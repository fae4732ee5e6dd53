"""Generator throughput at configs[2]'s maze: toroidal 81 x 81 blocks, the three generators mixed (maze_generate: warp kernel +
the torus BFS kernel), against the bordered case.  Bounds what regeneration on win can sustain."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))


def main():
    import maze_b200 as mb
    N = 65536
    for topology in ("euclid", "toroidal"):
        for algos in (["r-prim", "dfs", "prim&kill"], "r-prim", "dfs", "prim&kill"):
            pool = mb.MazePool(N, (81, 81))
            a = algos if isinstance(algos, str) else [algos[i % 3] for i in range(N)]
            pool.generate(shapes=(81, 81), algorithms=a, toroidal=(topology == "toroidal"), seed=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ids = torch.arange(N, dtype=torch.int32, device="cuda")
            e0.record()
            for _ in range(3):
                pool.generate(ids=ids, configure=False, seed=1)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"{topology:9s} {str(algos):36s}: {N / ms * 1e3:.3g} mazes/s ({ms:.1f} ms per {N})", flush=True)
            del pool


if __name__ == "__main__":
    main()
