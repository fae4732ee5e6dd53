"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference is imported read-only through ref_shim.py (gymnasium / pygame / matplotlib are
stubbed, nothing else).  Every array stored here is an output of reference code:

  steps.npz     mazes (block grid, start, goal, max_steps) + action tapes + the reference's
                (obs, reward, truncated, terminated, info['distance'], direction mask) per step,
                for -v0 and -v1 (window) envs, euclidean and toroidal, all three generators
  bestdir.npz   BaseMazeEnv._find_best_next_cell evaluated on every open block of some mazes
  metrics.npz   gen_maze outputs with ComplexityEvaluation difficulty/complexity and
                MetricsCalculator L / DE / D, plus hallway/branch structure sizes
  qagent.npz    QAgent / DQAgent table contents after scripted transition sequences
  metrics_ext.npz  the MetricsCalculator methods nobody calls (density, T, J, CR, AC/FDE/BDE, L_DE,
                T_DE / D_sharp / L_sharp per dead-end type) on the mazes of metrics.npz
"""
from __future__ import annotations

import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from gymnasium_env.envs.base_maze_env import BaseMazeEnv  # noqa: E402
from gymnasium_env.envs.simple_maze_env import SimpleEnrichMazeEnv, SimpleMazeEnv  # noqa: E402
from gymnasium_env.envs.toroidal_maze_env import ToroidalEnrichMazeEnv, ToroidalMazeEnv  # noqa: E402
from lib.a_star_algos.a_star import astar_limited_partial  # noqa: E402
from lib.maze_difficulty_evaluation.maze_complexity_evaluation import ComplexityEvaluation  # noqa: E402
from lib.maze_difficulty_evaluation.metrics_calculator import MetricsCalculator  # noqa: E402
from lib.maze_generation import gen_maze, gen_maze_no_border  # noqa: E402

ACTIONS = ((1, 0), (-1, 0), (0, 1), (0, -1))

# literal 15x15 maze of /root/reference/testing_Mccledon.py:4-20 (test data, start (1,1), goal (13,1))
LITERAL_15 = None  # filled by _read_literal()


def _read_literal():
    """Parse the commented-out literal out of the reference file instead of transcribing it."""
    import ast
    import re
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "testing_Mccledon.py")).read()
    rows = re.findall(r"#?\s*(\[[0-2,\s]+\])\s*,?", src)
    rows = [ast.literal_eval(r) for r in rows]
    rows = [r for r in rows if len(r) == 15]
    assert len(rows) >= 15, len(rows)
    return np.array(rows[:15], dtype=np.uint8)


def action_towards(best_dir, shape, toroidal):
    """Action index whose move realises obs['best dir'] (= agent - next)."""
    br, bc = int(best_dir[0]), int(best_dir[1])
    if toroidal:
        S0, S1 = shape
        if abs(br) > 1:
            br = -1 if br > 0 else 1
        if abs(bc) > 1:
            bc = -1 if bc > 0 else 1
    for a, (dr, dc) in enumerate(ACTIONS):
        if (-dr, -dc) == (br, bc):
            return a
    return None


def run_trace(env, toroidal, n_steps, follow_p, rng, enrich):
    """Drive a reference env; record everything it returns."""
    rec = {k: [] for k in ("agent", "target", "best", "dist", "mask", "action", "reward", "trunc", "term")}
    if enrich:
        rec["window"] = []

    def push_obs(obs, info):
        if enrich:
            rec["agent"].append(np.asarray(obs["agent"], dtype=np.float64))
            rec["target"].append(np.asarray(obs["target"], dtype=np.float64))
            rec["window"].append(obs["window"].numpy().astype(np.uint8))
        else:
            rec["agent"].append(np.asarray(obs["agent"], dtype=np.int64))
            rec["target"].append(np.asarray(obs["target"], dtype=np.int64))
        rec["best"].append(np.asarray(obs["best dir"], dtype=np.int64))
        rec["dist"].append(float(info["distance"]))
        rec["mask"].append(np.asarray(env.get_mask_direction(probs=True), dtype=np.float32))

    obs, info = env.reset()
    push_obs(obs, info)
    shape = env.maze_shape
    for _ in range(n_steps):
        a = None
        if rng.random() < follow_p:
            a = action_towards(obs["best dir"], shape, toroidal)
        if a is None:
            a = rng.randrange(4)
        obs, reward, truncated, terminated, info = env.step(a)
        rec["action"].append(a)
        rec["reward"].append(float(reward))
        rec["trunc"].append(bool(truncated))
        rec["term"].append(bool(terminated))
        push_obs(obs, info)
    return {k: np.array(v) for k, v in rec.items()}


def make_steps(out_path, specs=None, seed0=1000, budget81=90):
    if specs is None:
        specs = _default_step_specs()
    _make_steps(out_path, specs, seed0, budget81)


def _default_step_specs():
    specs = []
    for topo in ("euclid", "torus"):
        for algo in ("r-prim", "dfs", "prim&kill"):
            for shape in (21, 41):
                specs.append((topo, algo, shape, False))
    specs += [("euclid", "r-prim", 81, False), ("torus", "dfs", 81, False),
              ("euclid", "prim&kill", 15, False), ("torus", "r-prim", 15, False),
              ("euclid", "r-prim", 21, True), ("euclid", "dfs", 41, True), ("euclid", "prim&kill", 15, True),
              ("torus", "r-prim", 21, True), ("torus", "prim&kill", 41, True)]
    return specs


def _make_steps(out_path, specs, seed0, budget81):
    arrays, meta = {}, []
    for i, (topo, algo, shape, enrich) in enumerate(specs):
        random.seed(seed0 + i)
        np.random.seed(seed0 + i)
        BaseMazeEnv.ALGORITHM = algo
        toroidal = topo == "torus"
        cls = {(False, False): SimpleMazeEnv, (False, True): SimpleEnrichMazeEnv,
               (True, False): ToroidalMazeEnv, (True, True): ToroidalEnrichMazeEnv}[(toroidal, enrich)]
        t0 = time.time()
        env = cls((shape, shape))
        rng = random.Random(77 + i)
        tapes = []
        budget = budget81 if shape == 81 else 260
        for j, (follow_p, n) in enumerate(((0.0, budget), (0.75, budget), (1.0, min(budget, env.max_steps_taken + 6)))):
            tr = run_trace(env, toroidal, n, follow_p, rng, enrich)
            for k, v in tr.items():
                arrays[f"m{i}_t{j}_{k}"] = v
            tapes.append(j)
        arrays[f"m{i}_grid"] = np.array(env.maze_map, dtype=np.uint8)
        meta.append(dict(id=i, topology=topo, algo=algo, shape=shape, enrich=enrich,
                         start=[int(x) for x in env._start_pos], goal=[int(x) for x in env._target_location],
                         max_steps=int(env.max_steps_taken), tapes=tapes))
        print(f"steps m{i} {topo} {algo} {shape} enrich={enrich} max_steps={env.max_steps_taken} {time.time()-t0:.1f}s", flush=True)
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(out_path, **arrays)


def make_bestdir(out_path, specs=None, seed0=2000):
    arrays, meta = {}, []
    if specs is None:
        specs = [("euclid", a, s) for a in ("r-prim", "dfs", "prim&kill") for s in (15, 21, 41)]
        specs += [("torus", a, s) for a in ("r-prim", "dfs", "prim&kill") for s in (15, 21, 41)]
        specs += [("euclid", "dfs", 61), ("torus", "dfs", 61)]
    for i, (topo, algo, shape) in enumerate(specs):
        random.seed(seed0 + i)
        BaseMazeEnv.ALGORITHM = algo
        toroidal = topo == "torus"
        env = (ToroidalMazeEnv if toroidal else SimpleMazeEnv)((shape, shape))
        grid = np.array(env.maze_map, dtype=np.uint8)
        nxt = np.full((shape, shape, 2), -1, dtype=np.int32)
        t0 = time.time()
        for r in range(shape):
            for c in range(shape):
                if grid[r, c] != 0:
                    pos = np.array((r, c), dtype=np.int32)
                    b = env._find_best_next_cell(pos)
                    nxt[r, c] = (int(b[0]), int(b[1]))
        arrays[f"m{i}_grid"] = grid
        arrays[f"m{i}_next"] = nxt
        meta.append(dict(id=i, topology=topo, algo=algo, shape=shape,
                         start=[int(x) for x in env._start_pos], goal=[int(x) for x in env._target_location],
                         max_steps=int(env.max_steps_taken)))
        print(f"bestdir m{i} {topo} {algo} {shape} {time.time()-t0:.1f}s", flush=True)
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(out_path, **arrays)


def metric_row(maze, start, goal):
    ce = ComplexityEvaluation(maze, start, goal)
    sol = astar_limited_partial(maze, start, goal)
    mc = MetricsCalculator(maze, len(sol))
    hall_c = [ce.complexity_of_hallway(h) for h in sorted(ce.hallways)]
    return dict(difficulty=ce.difficulty_of_maze(), complexity=ce.complexity_of_maze(),
                L=mc.calculate_L(sol), DE=mc.calculate_DE(sol), D=mc.calculate_D(sol),
                sol_len=len(sol), n_hallways=len(ce.hallways), n_branches=len(ce.branches),
                hall_sum=float(sum(hall_c)))


def make_metrics(out_path):
    arrays, meta = {}, []
    lit = _read_literal()
    i = 0
    arrays[f"m{i}_grid"] = lit
    row = metric_row(lit.tolist(), (1, 1), (13, 1))
    row.update(id=i, algo="literal", shape=15, start=[1, 1], goal=[13, 1], no_border=False)
    meta.append(row)
    i += 1
    for algo in ("r-prim", "dfs", "prim&kill"):
        for shape, n in ((11, 4), (21, 6), (41, 5), (61, 2), (81, 1)):
            for k in range(n):
                random.seed(3000 + 17 * i)
                start, goal, maze = gen_maze((shape, shape), algo)
                t0 = time.time()
                row = metric_row(maze, start, goal)
                row.update(id=i, algo=algo, shape=shape, start=list(start), goal=list(goal), no_border=False)
                arrays[f"m{i}_grid"] = np.array(maze, dtype=np.uint8)
                meta.append(row)
                print(f"metrics m{i} {algo} {shape} diff={row['difficulty']:.4f} {time.time()-t0:.1f}s", flush=True)
                i += 1
    # gen_maze_no_border: difficulty is evaluated on the bordered maze before stripping
    for algo in ("r-prim", "dfs", "prim&kill"):
        for shape in (15, 21):
            random.seed(3000 + 17 * i)
            start, goal, maze, difficulty = gen_maze_no_border((shape, shape), algo)
            arrays[f"m{i}_grid"] = np.array(maze, dtype=np.uint8)
            meta.append(dict(id=i, algo=algo, shape=shape, start=list(start), goal=list(goal), no_border=True,
                             difficulty=difficulty))
            i += 1
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(out_path, **arrays)


def make_qagent(out_path):
    """Scripted transitions through the reference QAgent / DQAgent (agents/q_agent.py, dq_agent.py).
    The agents draw from numpy's global RNG; we record the draws by patching np.random.random and
    the action-space sampler so that the oracle/device replay consumes the same numbers."""
    DQAgent = ref_shim.load_reference_file("agents/dq_agent.py", "_ref_dq_agent").DQAgent
    QAgent = ref_shim.load_reference_file("agents/q_agent.py", "_ref_q_agent").QAgent

    random.seed(4242)
    np.random.seed(4242)
    BaseMazeEnv.ALGORITHM = "r-prim"
    env = SimpleMazeEnv((15, 15))
    out = {}
    for name, cls in (("q", QAgent), ("dq", DQAgent)):
        np.random.seed(99)
        draws_u, draws_a = [], []
        real_random = np.random.random

        def rec_random(*a, **k):
            v = real_random(*a, **k)
            draws_u.append(float(v))
            return v

        real_sample = env.action_space.sample

        def rec_sample():
            v = real_sample()
            draws_a.append(int(v))
            return v

        np.random.random = rec_random
        env.action_space.sample = rec_sample
        try:
            kw = dict(env=env, learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=150, final_epsilon=0.05,
                      discount_factor=0.7, eta=1e-2)
            agent = cls(**kw)
            log = {k: [] for k in ("agent", "target", "best", "action", "reward", "term", "trunc", "gamma")}
            for ep in range(12):
                obs, _ = env.reset()
                done, cum, prev = False, 0.0, 0.0
                while not done:
                    a = agent.get_action(obs)
                    nobs, r, trunc, term, _ = env.step(a)
                    log["agent"].append(np.array(obs["agent"])); log["target"].append(np.array(obs["target"]))
                    log["best"].append(np.array(obs["best dir"])); log["action"].append(a)
                    log["reward"].append(float(r)); log["term"].append(bool(term)); log["trunc"].append(bool(trunc))
                    log["gamma"].append(float(agent.discount_factor))
                    agent.update(obs, a, r, term, nobs)
                    cum += r
                    done = term or trunc
                    obs = nobs
                agent.update_hyperparameter(cum > prev)
            for k, v in log.items():
                out[f"{name}_{k}"] = np.array(v)
            out[f"{name}_final_obs"] = np.concatenate([obs["agent"], obs["target"], obs["best dir"]]).astype(np.int64)
            out[f"{name}_u"] = np.array(draws_u)
            out[f"{name}_a"] = np.array(draws_a, dtype=np.int64)
            tables = [agent.q_values] if name == "q" else [agent.q_a_values, agent.q_b_values]
            for ti, tab in enumerate(tables):
                keys, vals = [], []
                for k, v in tab.items():
                    keys.append(k)
                    vals.append(np.array(v, dtype=np.float64))
                out[f"{name}_tab{ti}_keys"] = np.array(keys)
                out[f"{name}_tab{ti}_vals"] = np.array(vals)
            out[f"{name}_steps_done"] = np.array(agent.steps_done)
            out[f"{name}_gamma_final"] = np.array(agent.discount_factor)
        finally:
            np.random.random = real_random
            env.action_space.sample = real_sample
    out["grid"] = np.array(env.maze_map, dtype=np.uint8)
    out["start"] = np.array(env._start_pos)
    out["goal"] = np.array(env._target_location)
    out["max_steps"] = np.array(env.max_steps_taken)
    np.savez_compressed(out_path, **out)


def make_genstats(out_path):
    """Samples of the reference generators' output distribution (fingerprints only)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.generation import maze_shape_stats
    out = {}
    for algo in ("r-prim", "dfs", "prim&kill"):
        for shape, n in ((21, 400), (41, 300), (81, 60)):
            random.seed(hash((algo, shape)) % 100000 + 5)
            random.seed(5000 + shape + 1000 * ("r-prim", "dfs", "prim&kill").index(algo))
            rows = []
            t0 = time.time()
            for _ in range(n):
                start, goal, maze = gen_maze((shape, shape), algo)
                st = maze_shape_stats(np.array(maze, dtype=np.uint8), start, goal)
                rows.append((st["sol_len"], st["dead_ends"], st["junctions"], start[0], start[1]))
            out[f"{algo}_{shape}"] = np.array(rows, dtype=np.int32)
            print(f"genstats {algo} {shape} n={n} mean={np.mean(rows, axis=0)} {time.time()-t0:.0f}s", flush=True)
    np.savez_compressed(out_path, **out)


def _metric_table_job(args):
    algo, seed = args
    random.seed(seed)
    start, goal, maze = gen_maze((81, 81), algo)
    r = metric_row(maze, start, goal)
    return (r["difficulty"], r["complexity"], r["L"], r["DE"], r["D"])


def make_metric_table_1000(out_path):
    """Round 2: the README table at the README's own sample size (1000 mazes of (81, 81) per generator)."""
    make_metric_table(out_path, n=1000, seed0=170000, procs=max(1, (os.cpu_count() or 2) - 2))


def make_steps81(out_path):
    """Round 2: reference traces at the headline size for the generator x topology pairs steps.npz lacks."""
    specs = [("euclid", "prim&kill", 81, False), ("euclid", "dfs", 81, False), ("torus", "r-prim", 81, False),
             ("torus", "prim&kill", 81, False), ("euclid", "r-prim", 81, True)]
    make_steps(out_path, specs=specs, seed0=1500, budget81=160)


def make_bestdir81(out_path):
    """Round 2: _find_best_next_cell on every open block of 81 x 81 dfs mazes (goal farther than the A* depth
    limit L = 162 from most blocks: the D > L branch of the closed form) and one r-prim maze per topology."""
    specs = [("euclid", "dfs", 81), ("torus", "dfs", 81), ("euclid", "r-prim", 81), ("torus", "prim&kill", 81)]
    make_bestdir(out_path, specs=specs, seed0=2500)


def make_metric_table(out_path, n=120, seed0=70000, procs=None):
    """The README table (generation_algos_metrics_evaluations.py:31-43) as per-maze samples: n mazes
    of (81, 81) per generator from the unmodified reference, columns difficulty, complexity, L, DE, D."""
    import multiprocessing as mp
    out = {}
    with mp.get_context("fork").Pool(procs or os.cpu_count() or 1) as pool:
        for k, algo in enumerate(("r-prim", "prim&kill", "dfs")):
            t0 = time.time()
            rows = pool.map(_metric_table_job, [(algo, seed0 + 1000 * k + i) for i in range(n)], chunksize=4)
            out[algo] = np.array(rows, dtype=np.float64)
            print(f"metric_table {algo} n={n} mean={out[algo].mean(axis=0)} max_difficulty={out[algo][:, 0].max():.2f} {time.time()-t0:.0f}s", flush=True)
    np.savez_compressed(out_path, **out)


EXT_TYPES = ("AC", "FDE", "BDE")


def make_metrics_ext(out_path):
    """The Kim-Crawfis metrics the reference defines but never calls (metrics_calculator.py:18-69,
    175-244): density, T, J, CR, the AC / FDE / BDE split of DE, L_DE and the per-type T_DE /
    D_sharp / L_sharp sums, computed by the unmodified MetricsCalculator on the bordered mazes of
    metrics.npz (shape <= 61)."""
    z = np.load(os.path.join(HERE, "metrics.npz"))
    meta = json.loads(str(z["meta"]))
    rows = []
    for m in meta:
        if m["no_border"] or m["shape"] > 61:
            continue
        t0 = time.time()
        maze = z[f"m{m['id']}_grid"].astype(int).tolist()
        sol = astar_limited_partial(maze, tuple(m["start"]), tuple(m["goal"]))
        mc = MetricsCalculator(maze, len(sol))
        ac, fde, bde = mc.calculate_DE_sub(sol)
        row = dict(id=m["id"], density=mc.calculate_density(), T=mc.calculate_T(sol), J=mc.calculate_J(sol),
                   CR=mc.calculate_CR(sol), AC=ac, FDE=fde, BDE=bde, L_DE=mc.calculate_L_DE(sol),
                   T_DE=[mc.calculate_T_DE(sol, t) for t in EXT_TYPES],
                   D_sharp=[mc.calculate_D_sharp(sol, t) for t in EXT_TYPES],
                   L_sharp=[mc.calculate_L_sharp(sol, t) for t in EXT_TYPES])
        rows.append(row)
        print(f"metrics_ext m{m['id']} {m['algo']} {m['shape']} {time.time()-t0:.1f}s", flush=True)
    np.savez_compressed(out_path, meta=np.array(json.dumps(rows)))


if __name__ == "__main__":
    which = sys.argv[1:] or ["steps", "bestdir", "metrics", "qagent", "genstats", "metric_table"]
    for w in which:
        t0 = time.time()
        {"steps": make_steps, "bestdir": make_bestdir, "metrics": make_metrics, "qagent": make_qagent,
         "genstats": make_genstats, "metric_table": make_metric_table, "metrics_ext": make_metrics_ext,
         "metric_table_1000": make_metric_table_1000, "steps81": make_steps81, "bestdir81": make_bestdir81}[w](
            os.path.join(HERE, f"{w}.npz"))
        print(f"== {w}.npz written in {time.time()-t0:.0f}s", flush=True)
