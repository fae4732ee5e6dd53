#!/usr/bin/env python
"""What does the memory system give maze_step's access pattern?  

Times maze_bench_scatter_rmw (csrc/maze_bench.cu) on the bench's own batch -- 4 096 000 envs on 1 000 81 x 81
mazes, a [slot, B] uint16 visit array of 54 GB -- for every combination of {streams only, RMW only, both} x
{uniform, cell-major per env, cell-major per warp}, and the real maze_step kernel at steady state beside them.
Writes gpurun_out/scatter_rmw.json (copy the summary into profiles/).

    python tools/perf_scatter_rmw.py [--envs 4096000] [--iters 200] [--once PATTERN STREAMS]   # --once: one config, for ncu
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb  # noqa: E402
from maze_b200 import cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096000)
    ap.add_argument("--mazes", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--once", nargs=2, type=int, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.envs
    env = mb.MazeVectorEnv(B, shape=(81, 81), algorithms="r-prim", num_mazes=args.mazes, seed=1234, device=dev, stats=False)
    env.reset()
    b = env.batch
    lib, ctx = cabi.lib(), b.ctx
    acts = torch.randint(0, 4, (64, B), device=dev, dtype=torch.uint8)
    arr, n_elems = b.visits, b.visits.numel()
    stream = cabi.current_stream(dev)

    def launch(pattern, rate, it, streams):
        rc = lib.maze_bench_scatter_rmw(ctx.handle, C.byref(b._c), cabi.ptr(acts[it % 64]), cabi.ptr(arr), n_elems, pattern, rate, it, streams, stream)
        ctx.check(rc, "maze_bench_scatter_rmw")

    def timed(fn, iters):
        for i in range(20):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(20 + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3   # microseconds per launch

    if args.once is not None:
        pattern, streams = args.once
        for i in range(30):
            launch(pattern, 410, i, streams)
        torch.cuda.synchronize()
        return
    out = {"envs": B, "array_bytes": n_elems * 2, "rmw_fraction": 410 / 1024}
    # the real kernel at steady state (autoreset on, same maze pool): what the patterns are compared with
    mode = cabi.STEP_AUTORESET
    for i in range(300):
        b.step(acts[i % 64], mode)
    out["maze_step_us"] = timed(lambda i: b.step(acts[i % 64], mode), args.iters)
    state_keep = b.state.clone()
    out["streams_only_us"] = timed(lambda i: launch(0, 0, i, 1), args.iters)
    for pattern, name in ((0, "uniform"), (1, "cell_major_per_env"), (2, "cell_major_per_warp")):
        out[f"rmw_only_{name}_us"] = timed(lambda i: launch(pattern, 410, i, 0), args.iters)
        out[f"streams_plus_rmw_{name}_us"] = timed(lambda i: launch(pattern, 410, i, 1), args.iters)
    for rate in (102, 205, 410, 820, 1024):
        out[f"streams_plus_rmw_cell_major_per_env_rate{rate}_us"] = timed(lambda i: launch(1, rate, i, 1), args.iters)
    b.state.copy_(state_keep)
    n_rmw = B * 410 / 1024
    out["rmw_per_launch"] = n_rmw
    out["atoms_gbs_rmw_only_cell_major_per_env"] = n_rmw * 128 / out["rmw_only_cell_major_per_env_us"] / 1e3
    out["step_vs_pattern"] = out["streams_plus_rmw_cell_major_per_env_us"] / out["maze_step_us"]
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scatter_rmw.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
