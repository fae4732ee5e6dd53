#!/usr/bin/env python
"""Headline benchmark: env-steps/s on 40x40 (81x81 block) mazes -- BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo, N=1 default
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU
    python bench.py --impl reference ...                           # the reference's CPU path

Workload (per GPU): a pool of 1000 r-prim euclidean 81x81 mazes generated on the device
(Philox seed 1234), B envs = 1000 mazes x A agents, uniform random actions from a device-resident
tape, gymnasium next-step autoreset on.  One "step" = one maze_step launch over all B envs.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "maze-solving-agent-gymnasium_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "env-steps/sec (40x40 mazes, whole job)"
UNIT = "env-steps/s"
SHAPE = (81, 81)
NUM_MAZES = 1000
ALGO_BYTES_PER_STEP = 58   # action 1 + state 16 + tables 7 + outputs 34


def workload_name(envs_per_gpu):
    return (f"configs[1]: {NUM_MAZES} constant-size 40x40 (81x81 block) r-prim euclidean mazes per GPU, "
            f"{envs_per_gpu} envs per GPU, uniform random actions, autoreset")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_step_byte():
    """dram bytes per env-step of maze_step from the committed ncu summary, if there is one."""
    path = os.path.join(ROOT, "profiles", "step_kernel_ncu_summary.json")
    try:
        return json.load(open(path))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def reference_mazes(n, seed=1234):
    """81x81 r-prim mazes for the CPU arms (oracle generator; the device generator is the product's)."""
    import random

    from oracle.generation import gen_maze
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        start, goal, grid = gen_maze(SHAPE, "r-prim", rng)
        out.append(dict(grid=grid, start=start, goal=goal, toroidal=False))
    return out


def run_reference(args, rank):
    """The reference's own CPU implementation of the path (the oracle's A*-per-step port of
    gymnasium_env/envs/base_maze_env.py; the Python reference itself cannot travel to the GPU box),
    one worker process per host core, each 'step' a bounded sample."""
    if rank != 0:
        return None
    from oracle.baseline import PersistentVector
    cores = os.cpu_count() or 1
    mazes = reference_mazes(min(cores, 8))
    vec = PersistentVector(mazes, cores, "port")
    # A 'step' is a fixed slice of wall time in which every worker steps its own env as fast as it can (free running:
    # the best this implementation can do on the box -- stepping the envs in lock step like gymnasium's
    # AsyncVectorEnv would wait for the slowest A* search every time and lands at about 40 % of this rate).  The slice
    # is sized so that warm-up + timed steps take about BUDGET_S in total, whatever K and W are.
    BUDGET_S = float(os.environ.get("MAZE_REF_BUDGET_S", "120"))   # (tests shrink it)
    per = max(0.02, BUDGET_S / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        vec.run_for(per)
    steps, secs = 0, 0.0
    for _ in range(args.steps):
        r = vec.run_for(per)
        steps += r["steps"]; secs += r["seconds"]
    vec.close()
    value = steps / secs
    sample = (f"{args.steps} steps x {per * 1e3:.0f} ms x {cores} free-running worker processes ({steps} env-steps in {secs:.1f} s), each "
              f"worker stepping the oracle port (A* per step) of the reference env on an 81x81 r-prim maze, random actions, reset on done")
    return ({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs_per_gpu), "arm": "CPU, oracle port of the reference algorithm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU, so that the pinned staging buffers of the e2e path
    are allocated on the GPU's NUMA node (8 ranks sharing one node's memory halve the copy rate)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:      # CUDA_VISIBLE_DEVICES may renumber the devices: go through the UUID
            handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
    except Exception as exc:   # not fatal: only the host-copy rate depends on it
        print(f"[bench] no CPU affinity for GPU {index}: {exc}", file=sys.stderr)


def measure_extras(mb, torch, device):
    """Secondary numbers of the same path, one GPU (rank 0), CUDA events: valid mazes generated/s
    (BASELINE.json's second metric), difficulty-metric throughput, the -v1 (window) step and the
    fused Q-learning rollout.  Each is a few hundred milliseconds."""
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    out = {"per_gpu": True}
    M = 131072
    pool = mb.MazePool(M, SHAPE, device)
    gen = {}
    for algo in ("r-prim", "dfs", "prim&kill"):
        seed = [7]

        pool.generate(algorithms=algo, seed=6)        # writes the per-slot configuration once

        def go():
            seed[0] += 1
            pool.generate(seed=seed[0], configure=False)
        gen[algo] = M / timed(go, 3)
    out["mazes_per_s_81x81"] = gen
    seed = [100]

    def go6():
        seed[0] += 1
        pool.generate(ids=torch.arange(16384, device=device, dtype=torch.int32), algorithms="r-prim", seed=seed[0],
                      candidates=6, configure=False)
    pool.generate(algorithms="r-prim", seed=5)
    out["best_of_6_mazes_per_s_81x81"] = 16384 / timed(go6, 1)
    ids = torch.arange(32768, device=device, dtype=torch.int32)
    out["difficulty_mazes_per_s_81x81"] = 32768 / timed(lambda: pool.difficulty(ids), 2)
    del pool
    # -v1 observation: step + 15x15 window (2 700 B written per env-step)
    Bv = 262144
    venv = mb.MazeVectorEnv(Bv, shape=SHAPE, num_mazes=1000, enrich=True, seed=1234, on_win="next", stats=False, device=device)
    venv.reset()
    acts = torch.randint(0, 4, (Bv,), dtype=torch.uint8, device=device)
    t = timed(lambda: venv.step(acts), 100)
    out["v1_window_env_steps_per_s"] = Bv / t
    out["v1_window_write_GBps"] = Bv * 2700 / t / 1e9
    # DQN data path: step + bit-packed observation + replay push (245 B per transition), and batch sampling
    from maze_b200.dqn import DeviceReplay
    memory = DeviceReplay(venv, 1 << 20, seed=1)
    memory.observe()

    def step_push():
        venv.batch.step(acts, venv._mode)
        memory.push(acts)
    out["dqn_step_push_env_steps_per_s"] = Bv / timed(step_push, 100)
    out["dqn_replay_samples_per_s"] = 65536 / timed(lambda: memory.sample(65536), 20)
    del venv, memory
    # fused tabular Q-learning rollout (policy + step + update per env, 64 steps per launch)
    from maze_b200.agents import QAgent
    Bq = 1048576
    qenv = mb.MazeVectorEnv(Bq, shape=SHAPE, num_mazes=1000, seed=1234, on_win="next", stats=True, device=device,
                            visit_layout="tile")
    agent = QAgent(qenv, learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=2000, final_epsilon=0.05,
                   discount_factor=0.7, eta=1e-3, envs_per_agent=Bq, capacity=1 << 24)
    qenv.reset()
    K = 64
    t = timed(lambda: agent.rollout(K), 3)
    agent.core.check_overflow()
    out["q_rollout_env_steps_per_s"] = Bq * K / t
    del qenv, agent
    # open-loop bursts: maze_step_many over a 64-step action tape, every per-step output written (bit-identical
    # to 64 maze_step launches); tiled visit layout.  Reported beside the headline, which is one launch per step.
    Bm = 4096000
    menv = mb.MazeVectorEnv(Bm, shape=SHAPE, num_mazes=1000, seed=1234, on_win="next", stats=False, device=device,
                            visit_layout="tile")
    menv.reset()
    tape = torch.randint(0, 4, (K, Bm), dtype=torch.uint8, device=device)
    for _ in range(5):
        menv.step_many(tape, trace=False)
    t = timed(lambda: menv.step_many(tape, trace=True), 4)
    out["step_many_env_steps_per_s"] = Bm * K / t
    out["step_many"] = {"k_steps": K, "envs": Bm, "us_per_step_equivalent": t / K * 1e6, "trace": "agent, best_dir, reward, terminated, truncated for every step",
                        "algorithmic_GBps": ALGO_BYTES_PER_STEP * Bm * K / t / 1e9, "frac_of_measured_hbm_peak": ALGO_BYTES_PER_STEP * Bm * K / t / 1e9 / measured_peak_gbs()[0]}
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import maze_b200 as mb

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.envs_per_gpu
    env = mb.MazeVectorEnv(num_envs=B, shape=SHAPE, topology="euclid", algorithms="r-prim", num_mazes=NUM_MAZES,
                           device=device, seed=1234, slot_id_base=rank * NUM_MAZES, autoreset=True, on_win="next",
                           stats=False)
    gen = torch.Generator(device=device)
    gen.manual_seed(99 + rank)
    TAPE = 16
    tape = torch.randint(0, 4, (TAPE, B), dtype=torch.uint8, device=device, generator=gen)
    env.reset()
    sampler = ClockSampler(local_rank)   # samples from the warm-up on: same kernel, same load
    sampler.start()
    for t in range(args.warmup):
        env.step(tape[t % TAPE])
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(args.steps):
        env.step(tape[t % TAPE])
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    value = world * B * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # end to end: host action buffers in, every output copied back to host, per step
    host_tape = tape[:4].cpu().numpy()
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for t in range(2):
        env.step_host(host_tape[t % 4])
    barrier()
    torch.cuda.synchronize()
    env.reset_host_counters()      # d2h_bytes_per_step() = bytes actually copied in the timed region / steps
    t0 = time.perf_counter()
    for t in range(e2e_steps):
        obs, rew, term, trunc, _ = env.step_host(host_tape[t % 4])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * e2e_steps / e2e_s

    extra = None
    if rank == 0 and not args.no_extras:
        extra = measure_extras(mb, torch, device)
    barrier()

    # independent check that the timed kernel did the work: episode bookkeeping must be moving
    st = env.batch.state_host()
    assert st["steps"].max() > 0 and int(rew.shape[0]) == B

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALGO_BYTES_PER_STEP * B / (ms_per_step * 1e-3) / 1e9
        prof = ncu_traffic_per_step_byte()
        traffic = None
        if prof and prof.get("dram_bytes_per_env_step") is not None:
            traffic = prof["dram_bytes_per_env_step"] * B   # ncu --set full capture of a steady-state launch
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle.baseline import time_env_steps
            cores = os.cpu_count() or 1
            # same mazes as the GPU run: copy a few pool grids back and hand them to the CPU port
            meta = env.pool.meta_host()
            mazes = []
            for m in range(min(cores, 8)):
                mazes.append(dict(grid=env.pool.grid_host(m), start=(int(meta[m, 2]) & 0xffff, int(meta[m, 2]) >> 16),
                                  goal=(int(meta[m, 3]) & 0xffff, int(meta[m, 3]) >> 16), toroidal=False))
            r = time_env_steps(mazes, args.cpu_seconds, cores, "port")
            r2 = time_env_steps(mazes, min(3.0, args.cpu_seconds), cores, "closed")
            cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{r['steps']} env-steps in {r['seconds']:.1f} s: {cores} processes, each one oracle-port env "
                             f"(A* per step, the reference's algorithm) on one of the GPU run's own 81x81 mazes, random actions",
                   "optimised_cpu_closed_form_value": r2["value"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32 state+obs, f64 reward", "data": "synthetic",
            "config": {"workload": workload_name(B), "envs_per_gpu": B, "mazes_per_gpu": NUM_MAZES,
                       "l2": "inputs larger than L2: per-step working set (~100 B x envs = 400 MB) and the 13 KB/env visit arrays exceed the 126 MB L2; no flush needed",
                       "parallelism": f"env-index sharding over {world} GPU(s), no per-step collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_STEP, "kernel": "maze_step_kernel"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step(),
                    "d2h_bytes_per_step": env.d2h_bytes_per_step(), "steps": e2e_steps,
                    "note": "every output of every env is copied to pinned host memory each step (agent, best dir, reward, "
                            "terminated, truncated); the target array is copied again only on steps whose launch rewrote it "
                            "(an env restarting after a win), which the kernel reports through maze_env_batch.target_dirty"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "extra": extra,
        }
    if world > 1:
        dist.destroy_process_group()
    return line


class JsonOnlyStdout:
    """The driver parses stdout as ONE JSON line: everything else a library prints there (NCCL's
    version banner, for one) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self._real, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=500)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=4096 * NUM_MAZES)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    with JsonOnlyStdout() as out:
        if args.impl == "reference":
            line = run_reference(args, rank)
        else:
            line = run_ours(args, rank, local_rank, world)
        if line is not None:
            out.emit(json.dumps(line))


if __name__ == "__main__":
    main()
