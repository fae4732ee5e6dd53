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


def _persistent_worker(conn, kind, grid, start, goal, toroidal, seed):
    if kind == "reference":   # the unmodified reference env from baseline/_ref: generates its own 81 x 81 maze (best of six)
        from .ref_runtime import make_reference_env
        env = make_reference_env(np.asarray(grid).shape, "r-prim", seed)
    else:
        from .env_port import ClosedFormEnv, PortEnv
        env = (PortEnv if kind == "port" else ClosedFormEnv)(grid, start, goal, toroidal)
    rng = np.random.default_rng(seed)
    env.reset()
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd is None:
            break
        kind_, arg = cmd
        t0 = time.perf_counter()
        n = 0
        if kind_ == "count":          # exactly `arg` transitions
            for a in rng.integers(0, 4, arg):
                _, _, trunc, term, _ = env.step(int(a))
                if trunc or term:
                    env.reset()
            n = int(arg)
        else:                         # "time": as many transitions as fit into `arg` seconds
            deadline = t0 + arg
            while True:
                _, _, trunc, term, _ = env.step(int(rng.integers(0, 4)))
                n += 1
                if trunc or term:
                    env.reset()
                if time.perf_counter() >= deadline:
                    break
        conn.send((n, time.perf_counter() - t0))


class PersistentVector:
    """One long-lived worker process per core, each owning one env, so that a bench 'step' can be a small
    bounded sample without paying a process spawn each time."""

    def __init__(self, mazes, workers=None, kind="port"):
        self.workers = workers or os.cpu_count() or 1
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for w in range(self.workers):
            m = mazes[w % len(mazes)]
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_persistent_worker, daemon=True,
                            args=(child, kind, np.asarray(m["grid"]), tuple(m["start"]), tuple(m["goal"]), bool(m["toroidal"]), 1000 + w))
            p.start()
            self.conns.append(parent)
            self.procs.append(p)
        for c in self.conns:   # constructors done (the reference's takes ~12 s at 81 x 81: six mazes generated and scored)
            assert c.recv() == "ready"

    def step(self, n):
        """Lock step (the AsyncVectorEnv shape): every env advances by exactly n transitions."""
        return self._run(("count", int(n)))

    def run_for(self, seconds):
        """Free running: every worker steps its env for `seconds` of wall time (no waiting for the slowest env)."""
        return self._run(("time", float(seconds)))

    def _run(self, cmd):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(cmd)
        res = [c.recv() for c in self.conns]
        return dict(steps=sum(r[0] for r in res), seconds=time.perf_counter() - t0, slowest_worker_seconds=max(r[1] for r in res))

    def close(self):
        for c in self.conns:
            c.send(None)
        for p in self.procs:
            p.join(timeout=5)
