"""GEMM timing under three cache conditions, per launch (CUDA events around every launch):
  warm   -- same operands every launch (L2 holds them after the first)
  cold   -- 512 MB written between launches (operands come from HBM)
  fresh  -- operands rewritten by an elementwise kernel right before the launch (what the train step does)
plus the host time of one launch call.  Explains why the in-situ GEMM times of the train step differ from the loop numbers."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))


def main():
    from maze_b200.dqn_net import gemm_bf16
    torch.manual_seed(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for (M, N, K) in ((16384, 1024, 1600), (8192, 1024, 1600), (1024, 1024, 1600), (16384, 512, 1024), (8192, 1568, 1024)):
        for tn in (256, 512):
            A = torch.randn(M, K, device="cuda").bfloat16()
            A2 = torch.randn(M, K, device="cuda").bfloat16()
            B = torch.randn(N, K, device="cuda").bfloat16()
            Cc = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            res = {}
            bias = torch.randn(N, device="cuda")
            for mode in ("warm", "cold", "fresh", "fresh+bias+lrelu"):
                ts = []
                for it in range(12):
                    if mode == "cold":
                        flush.fill_(it)
                    elif mode.startswith("fresh"):
                        flush.fill_(it)
                        A.copy_(A2)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    if mode.endswith("lrelu"):
                        gemm_bf16(A, B, Cc, tile_n=tn, bias=bias, act=1)
                    else:
                        gemm_bf16(A, B, Cc, tile_n=tn)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
                ts.sort()
                res[mode] = ts[len(ts) // 2] * 1e3
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                gemm_bf16(A, B, Cc, tile_n=tn)
            host = (time.perf_counter() - t0) / 200 * 1e6
            torch.cuda.synchronize()
            res["host_call_us"] = host
            out[f"{M}x{N}x{K}_t{tn}"] = res
            print(f"{M}x{N}x{K} tile {tn}: warm {res['warm']:.1f} us, cold {res['cold']:.1f} us, fresh {res['fresh']:.1f} us, fresh with bias + LeakyReLU {res['fresh+bias+lrelu']:.1f} us, host call {host:.1f} us", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_cold.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
