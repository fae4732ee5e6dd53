#!/usr/bin/env python
"""What does HBM3e give a WRITE-ONLY stream?  (the roofline of maze_window, which writes 2 700 B per env and reads little)

MEASURED_PEAKS.json's hbm_gbs is a copy (read + write bytes counted).  This times torch's fill (a plain grid of 16-byte
stores) and a copy on buffers far larger than L2, best of 10, CUDA events.  Writes gpurun_out/write_stream.json."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
n = 1 << 30   # 4 GiB of float32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) * 1e-3)
    return min(t)


out["fill_write_only_GBps"] = 4 * n / best(lambda: a.fill_(1.0)) / 1e9
out["zero_write_only_GBps"] = 4 * n / best(lambda: a.zero_()) / 1e9
out["copy_read_plus_write_GBps"] = 8 * n / best(lambda: b.copy_(a)) / 1e9
out["read_only_sum_GBps"] = 4 * n / best(lambda: a.sum()) / 1e9
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "write_stream.json"), "w"), indent=1)
