"""Host time against GPU time of the net's train step: is the step bound by the host's launch path?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from maze_b200.dqn_net import DQNNet
    from net_selftest import make_batch
    n = int(os.environ.get("NET_BATCH", "8192"))
    net = DQNNet("cuda", max_batch=n, seed=6)
    b = make_batch(n, seed=30)
    args = (b["vec"], b["pwin"], b["nvec"], b["pnwin"], b["action"], b["reward"])
    for _ in range(5):
        net.train_step(*args, gamma=0.9, lr=1e-4)
    torch.cuda.synchronize()
    host = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net.backward(*args, 0.9)
        t1 = time.perf_counter()
        net.adamw(1e-4)
        t2 = time.perf_counter()
        host.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6))
    host.sort()
    hb, ha = host[len(host) // 2]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(50):
        net.train_step(*args, gamma=0.9, lr=1e-4)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n = {n}: host time of backward() {hb:.0f} us, adamw() {ha:.0f} us (each after a synchronize: nothing queued)")
    print(f"50 train steps: GPU events {e0.elapsed_time(e1) / 50 * 1e3:.0f} us per step, host issue loop {(t1 - t0) / 50 * 1e6:.0f} us per step, "
          f"host + final sync {(t2 - t0) / 50 * 1e6:.0f} us per step")


if __name__ == "__main__":
    main()
