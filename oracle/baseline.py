"""CPU baseline timing (test / bench infrastructure; see oracle/__init__.py).

Times the oracle's PORT of the reference env (A* per step, exactly the reference's algorithm and
language) as a vector of independent envs, one worker process per host core -- the analogue of
gymnasium.vector.AsyncVectorEnv over the reference's env classes (gymnasium itself is not
installed in this image).  Also times the closed-form NumPy/Python restatement ("optimised CPU").
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _worker(args):
    kind, grid, start, goal, toroidal, seconds, seed = args
    from .env_port import ClosedFormEnv, PortEnv
    env = (PortEnv if kind == "port" else ClosedFormEnv)(grid, start, goal, toroidal)
    rng = np.random.default_rng(seed)
    env.reset()
    n = 0
    t0 = time.perf_counter()
    deadline = t0 + seconds
    while True:
        acts = rng.integers(0, 4, 16)
        for a in acts:
            _, _, trunc, term, _ = env.step(int(a))
            n += 1
            if trunc or term:
                env.reset()
        if time.perf_counter() >= deadline:
            break
    return n, time.perf_counter() - t0


def time_env_steps(mazes, seconds=10.0, workers=None, kind="port"):
    """mazes: list of dict(grid, start, goal, toroidal).  Returns dict(value steps/s, cores, ...)."""
    workers = workers or os.cpu_count() or 1
    jobs = []
    for w in range(workers):
        m = mazes[w % len(mazes)]
        jobs.append((kind, np.asarray(m["grid"]), tuple(m["start"]), tuple(m["goal"]), bool(m["toroidal"]), seconds, 1000 + w))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if workers == 1:
        res = [_worker(jobs[0])]
    else:
        with ctx.Pool(workers) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    longest = max(r[1] for r in res)
    return dict(value=steps / longest, steps=steps, seconds=longest, wall=wall, cores=workers, kind=kind)
