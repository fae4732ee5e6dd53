#!/usr/bin/env python
"""BASELINE.json configs[4]: a DDQN training loop on thousands of parallel 40x40 maze envs per GPU,
sharded over the GPUs of one box with an NCCL gradient all-reduce (DistributedDataParallel).

    python examples/train_ddqn.py --envs 8192 --iters 200
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_ddqn.py --envs 8192 --iters 200

Everything around the network runs in this repo's kernels: env step + -v1 window (maze_step,
maze_window), masked epsilon-greedy (maze_dqn_select), bit-packed replay ring (maze_dqn_push /
maze_dqn_sample), regeneration of won mazes (maze_generate).  The network is the consumer's: the
architecture of the reference's agents/ddqn_agent.py:18-52 (conv 3->32 3x3 pad 1, LeakyReLU,
Dropout .2, MaxPool 2 -> 1568 + 6 -> 1024 -> 512 -> 4), run in bf16 autocast on the tensor cores;
the update is ddqn_agent.py:113-152 (double-Q target, MSE, gradient clamp +-1, AdamW).
Prints one JSON line with env-steps/s and optimiser steps/s.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
import maze_b200 as mb  # noqa: E402
from maze_b200.dqn import DeviceReplay, MaskedEpsilonGreedy  # noqa: E402


class QNet(nn.Module):
    def __init__(self, hidden: int = 1024):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.LeakyReLU(), nn.Dropout(0.2), nn.MaxPool2d(2, 2))
        self.fc = nn.Sequential(nn.Linear(32 * 7 * 7 + 6, hidden), nn.LeakyReLU(), nn.Linear(hidden, hidden // 2), nn.ReLU(),
                                nn.Linear(hidden // 2, 4))

    def forward(self, state):
        vec, window = state
        return self.fc(torch.cat((self.conv(window).flatten(1), vec), dim=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--batch", type=int, default=4096, help="replay batch per GPU per optimiser step")
    ap.add_argument("--shape", type=int, default=81)
    ap.add_argument("--memory", type=int, default=1 << 20)
    args = ap.parse_args()
    rank, local, world = mb.dist.env_from_torchrun() if hasattr(mb, "dist") else (0, 0, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B = args.envs
    env = mb.MazeVectorEnv(B, shape=(args.shape, args.shape), enrich=True, on_win="regenerate", algorithms="r-prim", seed=1,
                           slot_id_base=rank * B, device=device, stats=True, algorithm_schedule=((5, "prim&kill"), (10, "dfs")))
    memory = DeviceReplay(env, args.memory, seed=rank)
    actor = MaskedEpsilonGreedy(env, 0.9, 0.05, 2000, seed=7, env_id_base=rank * B)
    torch.manual_seed(0)
    source, target = QNet().to(device), QNet().to(device)
    target.load_state_dict(source.state_dict())
    net = nn.parallel.DistributedDataParallel(source, device_ids=[local]) if world > 1 else source
    opt = torch.optim.AdamW(source.parameters(), 1e-4)
    gamma = 0.9
    env.reset()
    memory.observe()
    losses = []
    warmup = min(20, args.iters // 4)
    t0 = time.perf_counter()
    for it in range(args.iters + warmup):
        if it == warmup:                 # cuDNN / cuBLAS initialisation and autotuning stay outside the clock
            torch.cuda.synchronize()
            t0, losses = time.perf_counter(), []
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            source.eval()
            q = source(memory.current_state())
            source.train()
        actions = actor.select(q.float())
        env.step(actions)
        memory.push(actions)
        if (it + 1) * B >= args.batch:       # no device sync: at most B transitions are pushed per step
            state, action, reward, next_state = memory.sample(args.batch)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                qsa = net(state).float().gather(1, action.unsqueeze(1))
                with torch.no_grad():
                    best = source(next_state).float().argmax(1, keepdim=True)
                    nxt = target(next_state).float().gather(1, best).squeeze(1)
                loss = F.mse_loss(qsa, (nxt * gamma + reward).unsqueeze(1))
            opt.zero_grad(set_to_none=True)
            loss.backward()                      # DDP: NCCL all-reduce of the 2.1 M-parameter gradient
            for p in source.parameters():
                p.grad.clamp_(-1, 1)
            opt.step()
            losses.append(loss.detach())
        if it % 50 == 49:                    # update_target()
            target.load_state_dict(source.state_dict())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = env.episode_statistics(reduce=world > 1)
    if rank == 0:
        print(json.dumps({"env_steps_per_s": world * B * args.iters / dt, "optimizer_steps_per_s": len(losses) / dt,
                          "samples_per_s": world * args.batch * len(losses) / dt, "n_gpus": world, "envs_per_gpu": B,
                          "final_loss": float(torch.stack(losses[-10:]).mean()) if losses else None, "episodes": stats["episodes"],
                          "wins": stats["wins"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
