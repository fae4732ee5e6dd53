"""The reference-facing Python surface (gymnasium_env.envs.*, lib.maze_generation,
lib.maze_difficulty_evaluation, MazeVectorEnv) against reference goldens and the oracle."""
import random

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle.env_port import ClosedFormEnv  # noqa: E402
from oracle.generation import check_perfect_maze, select_goal  # noqa: E402
from oracle.metrics import mcclendon  # noqa: E402
from oracle.qlearn import Draws, OracleQAgent, obs_key, run_episodes  # noqa: E402


def _env_for(m, z):
    """A facade env of the right class with the golden maze installed through the env's own
    update_visited_maze() (how the reference's test() replays learned mazes)."""
    from gymnasium_env.envs import SimpleEnrichMazeEnv, SimpleMazeEnv, ToroidalEnrichMazeEnv, ToroidalMazeEnv
    tor = m["topology"] == "torus"
    cls = {(False, False): SimpleMazeEnv, (False, True): SimpleEnrichMazeEnv,
           (True, False): ToroidalMazeEnv, (True, True): ToroidalEnrichMazeEnv}[(tor, bool(m["enrich"]))]
    env = cls((m["shape"], m["shape"]))
    env.mazes = [[tuple(m["start"]), z[f"m{m['id']}_grid"].astype(np.int64).tolist()]]
    env.next = 0
    env.update_visited_maze(remove=False)
    return env


def test_single_env_classes_replay_reference_traces(golden_steps):
    z, meta = golden_steps
    picks = [m for m in meta if m["shape"] <= 41]
    assert any(m["enrich"] for m in picks) and any(m["topology"] == "torus" for m in picks)
    for m in picks:
        env = _env_for(m, z)
        assert env.max_steps_taken == m["max_steps"]
        assert tuple(env._target_location) == tuple(m["goal"]) and env._start_pos == tuple(m["start"])
        for j in m["tapes"][:2]:
            pre = f"m{m['id']}_t{j}_"
            obs, info = env.reset()
            acts = z[pre + "action"]
            for t in range(len(acts) + 1):
                if m["enrich"]:
                    assert obs["agent"].dtype == np.float64 and obs["window"].dtype == torch.float32
                    np.testing.assert_array_equal(obs["agent"], z[pre + "agent"][t])
                    np.testing.assert_array_equal(obs["target"], z[pre + "target"][t])
                    np.testing.assert_array_equal(obs["window"].numpy(), z[pre + "window"][t].astype(np.float32))
                else:
                    assert obs["agent"].dtype == np.int32 and obs["target"].dtype == np.int32 and obs["best dir"].dtype == np.int64
                    np.testing.assert_array_equal(obs["agent"], z[pre + "agent"][t])
                    np.testing.assert_array_equal(obs["target"], z[pre + "target"][t])
                np.testing.assert_array_equal(obs["best dir"], z[pre + "best"][t], err_msg=f"{pre}{t}")
                assert info["distance"] == z[pre + "dist"][t]
                np.testing.assert_array_equal(env.get_mask_direction(probs=True), z[pre + "mask"][t], err_msg=f"{pre}{t}")
                if t == len(acts):
                    break
                obs, reward, truncated, terminated, info = env.step(int(acts[t]))   # the reference's swapped order
                assert float(reward) == float(z[pre + "reward"][t]), (pre, t)
                assert truncated == bool(z[pre + "trunc"][t]) and terminated == bool(z[pre + "term"][t]), (pre, t)


def test_config1_q_learning_through_the_facade_matches_oracle():
    """BASELINE.json configs[0]: one 10x10-cell (21x21 block) r-prim maze, tabular Q-learning driven
    by the trainer loop (off_policy_trainer.py:29-51,76-78) with the reference-semantics agent."""
    from gymnasium_env.envs import SimpleMazeEnv
    random.seed(123)
    env = SimpleMazeEnv((21, 21))
    grid = np.array(env.maze_map, dtype=np.uint8)
    ok, why = check_perfect_maze(grid)
    assert ok, why
    ora_env = ClosedFormEnv(grid, env._start_pos, tuple(env._target_location), False)
    assert env.max_steps_taken == ora_env.max_steps
    rng = np.random.default_rng(0)
    u, a = rng.random(40000), rng.integers(0, 4, 40000)
    kw = dict(learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=250, final_epsilon=0.05, discount_factor=0.7, eta=1e-3)
    agents = [OracleQAgent(draws=Draws(u, a), **kw) for _ in range(2)]
    logs = [run_episodes(e, ag, n_episodes=25) for e, ag in ((env, agents[0]), (ora_env, agents[1]))]
    assert logs[0]["action"] == logs[1]["action"] and logs[0]["term"] == logs[1]["term"] and logs[0]["trunc"] == logs[1]["trunc"]
    np.testing.assert_array_equal(np.array(logs[0]["reward"]).view(np.uint64), np.array(logs[1]["reward"]).view(np.uint64))
    assert set(agents[0].q_a) == set(agents[1].q_a)
    for k, row in agents[0].q_a.items():
        np.testing.assert_array_equal(row.view(np.uint64), agents[1].q_a[k].view(np.uint64))
    assert sum(logs[0]["term"]) > 0


def test_maze_lifecycle_and_class_global_algorithm():
    from gymnasium_env.envs import (BaseMazeEnv, SimpleEnrichVariableMazeEnv, SimpleMazeEnv, SimpleVariableMazeEnv,
                                    ToroidalMazeEnv, ToroidalVariableMazeEnv)
    random.seed(5)
    env = SimpleMazeEnv((21, 21))
    try:
        env.set_algorithm("dfs")
        assert BaseMazeEnv.ALGORITHM == "dfs" and ToroidalMazeEnv.ALGORITHM == "dfs" and env.get_algorithm() == "dfs"
        first = env.maze_map
        env.update_maze()
        assert env.maze_map != first and len(env.mazes) == 2 and env.get_maze_shape() == (21, 21)
        grid = np.array(env.maze_map, dtype=np.uint8)
        assert check_perfect_maze(grid)[0]
        g1 = grid.copy(); g1[tuple(env._target_location)] = 1
        assert select_goal(g1, env._start_pos) == tuple(env._target_location)
        d, _ = mcclendon(grid, env._start_pos, tuple(env._target_location))
        assert env.get_maze_difficulty() == pytest.approx(d, rel=1e-9)
        env.update_new_maze((31, 31)) if False else None
    finally:
        env.set_algorithm("r-prim")
    # variable envs start at START_SHAPE and grow by (4, 4) per update_maze until max_shape
    var = SimpleVariableMazeEnv((23, 23))
    assert var.get_maze_shape() == (15, 15) and var.get_max_shape() == (23, 23)
    shapes = []
    for _ in range(4):
        var.update_maze()
        shapes.append(var.get_maze_shape())
        assert np.array(var.maze_map).shape == var.get_maze_shape()
    assert shapes == [(19, 19), (23, 23), (23, 23), (23, 23)] and len(var.mazes) == 3
    var.update_new_maze()
    assert var.get_maze_shape()[0] in range(15, 23, 2)
    var.update_visited_maze(remove=True)
    assert len(var.mazes) == 2
    obs, _ = var.reset()
    assert tuple(obs["agent"]) == var._start_pos
    tv = ToroidalVariableMazeEnv((33, 33))
    assert tv.get_maze_shape() == (29, 29)
    tv.update_maze()
    assert tv.get_maze_shape() == (33, 33) and np.array(tv.maze_map).shape == (33, 33)
    ev = SimpleEnrichVariableMazeEnv((19, 19))
    obs, _ = ev.reset()
    assert tuple(obs["window"].shape) == (3, 15, 15) and obs["agent"].dtype == np.float64
    with pytest.raises(ValueError):
        SimpleMazeEnv((20, 20))


def test_best_of_six_lowers_difficulty():
    """generate_maze keeps the least difficult of six draws: facade mazes must be easier on
    average than raw gen_maze output."""
    from gymnasium_env.envs import SimpleMazeEnv
    from lib.maze_generation import gen_mazes
    random.seed(1)
    env = SimpleMazeEnv((21, 21))
    kept = []
    for _ in range(40):
        env.update_maze()
        kept.append(env.get_maze_difficulty())
    raw = gen_mazes(400, (21, 21), "r-prim", seed=3).difficulty()[:, 0].cpu().numpy()
    assert np.mean(kept) < np.mean(raw) - 0.5 * np.std(raw)


def test_lib_entry_points(golden_metrics):
    from lib.maze_difficulty_evaluation.maze_complexity_evaluation import ComplexityEvaluation
    from lib.maze_difficulty_evaluation.metrics_calculator import MetricsCalculator
    from lib.maze_generation import gen_maze, gen_maze_no_border, generate_collection_of_mazes
    z, meta = golden_metrics
    lit = z["m0_grid"].tolist()
    ce = ComplexityEvaluation(lit, (1, 1), (13, 1))
    assert ce.difficulty_of_maze() == pytest.approx(9.950639302928026, rel=1e-12)
    assert ce.complexity_of_maze() == pytest.approx(5.681612603202764, rel=1e-12)
    mc = MetricsCalculator(lit, 61)
    path = [(1, 1)] + [(0, 0)] * 60     # only len(path) and path[0] matter
    assert mc.CE == 97 and mc.calculate_L(path) == 0.6288659793814433
    assert mc.calculate_DE(path) == 0.03278688524590164 and mc.calculate_D(path) == 0.03278688524590164
    random.seed(2)
    for algo in ("r-prim", "dfs", "prim&kill"):
        start, goal, maze = gen_maze((21, 21), algo)
        grid = np.array(maze, dtype=np.uint8)
        assert check_perfect_maze(grid)[0] and grid[goal] == 2 and start[0] % 2 == 1
        s, g, mz, diff = gen_maze_no_border((21, 21), algo)
        bordered = np.pad(np.array(mz, dtype=np.uint8), 1)
        assert check_perfect_maze(bordered)[0] and np.array(mz).shape == (21, 21)
        d, _ = mcclendon(bordered, (s[0] + 1, s[1] + 1), (g[0] + 1, g[1] + 1))
        assert diff == pytest.approx(d, rel=1e-9)
    with pytest.raises(ValueError):
        gen_maze((21, 21), "kruskal")
    coll = generate_collection_of_mazes((15, 15), 6)
    assert len(coll) == 6 and all(tuple(t.shape) == (3, 15, 15) and t.dtype == torch.int32 for t in coll)
    for t in coll:
        assert ((t[0] + (t[1] | (t[2] & ~t[1]))) >= 1).all() and int((t[2] - t[1]).abs().sum()) == 2   # goal + start differ


def test_vector_env_enrich_and_registry():
    import gymnasium_env  # noqa: F401  (registers the ids)
    import maze_b200 as mb
    from maze_b200 import _gym
    if not _gym.HAVE_GYMNASIUM:
        env = _gym.make("gymnasium_env/MazeEnv-v1", maze_shape=(15, 15))
        assert type(env).__name__ == "SimpleEnrichMazeEnv"
    venv = mb.MazeVectorEnv(64, shape=(21, 21), num_mazes=8, enrich=True, seed=4, candidates=6)
    obs, info = venv.reset()
    assert tuple(obs["window"].shape) == (64, 3, 15, 15) and obs["agent"].dtype == torch.float64
    acts = torch.randint(0, 4, (64,), dtype=torch.uint8, device="cuda")
    obs, rew, term, trunc, info = venv.step(acts)
    assert rew.shape == (64,) and term.dtype == torch.bool and info["distance"].shape == (64,)
    h_obs, h_rew, h_term, h_trunc, _ = venv.step_host(acts.cpu().numpy())
    assert h_obs["window"].shape == (64, 3, 15, 15) and h_rew.dtype == np.float64
    mask = venv.get_mask_direction(probs=True)
    assert tuple(mask.shape) == (64, 4)


def test_step_host_refreshes_target_only_when_a_maze_changed():
    """step_host() skips the D2H copy of `target` unless the launch rewrote it (maze_env_batch.target_dirty):
    the host arrays must still equal the device arrays after every step, through wins and pool cycling."""
    import maze_b200 as mb
    B = 64
    env = mb.MazeVectorEnv(B, shape=(11, 11), num_mazes=16, algorithms=["r-prim", "dfs"], seed=3, on_win="next", stats=True)
    env.reset()
    rng = np.random.default_rng(0)
    copies_with_target = 0
    for t in range(400):
        best = env.batch.best_dir.cpu().numpy()
        a = np.zeros(B, dtype=np.uint8)
        a[best[:, 0] == 1] = 1; a[best[:, 0] == -1] = 0; a[best[:, 1] == 1] = 3; a[best[:, 1] == -1] = 2
        rnd = rng.random(B) < 0.1
        a[rnd] = rng.integers(0, 4, int(rnd.sum()))
        before = env._h_bytes
        obs, rew, term, trunc, _ = env.step_host(a)
        copies_with_target += (env._h_bytes - before) > B * 26 + 4
        assert np.array_equal(obs["target"], env.batch.target.cpu().numpy()), t
        assert np.array_equal(obs["agent"], env.batch.agent.cpu().numpy())
        assert np.array_equal(rew, env.batch.reward.cpu().numpy())
    wins = env.episode_statistics()["wins"]
    assert wins > 0 and 0 < copies_with_target < 400     # targets did change, and most steps skipped the copy
    assert env.d2h_bytes_per_step() < B * 34


@pytest.mark.parametrize("enrich", [False, True])
def test_vector_env_declares_spaces_that_match_its_observations(enrich):
    """The north star asks for a gymnasium VectorEnv: single_observation_space / observation_space / action spaces exist
    (gymnasium's classes when it is installed, the package's stand-ins otherwise) and describe what reset / step produce."""
    import maze_b200 as mb
    env = mb.MazeVectorEnv(64, shape=(21, 21), num_mazes=4, enrich=enrich, seed=1)
    obs, _ = env.reset()
    assert env.single_observation_space is not None and env.observation_space is not None
    assert set(env.single_observation_space.keys()) == set(obs.keys())
    for k, v in obs.items():
        single, batched = env.single_observation_space[k], env.observation_space[k]
        assert tuple(batched.shape) == tuple(v.shape) == (64,) + tuple(single.shape), k
        assert np.dtype(single.dtype) == np.dtype(str(v.dtype).replace("torch.", "")), (k, single.dtype, v.dtype)
        lo, hi = float(v.min()), float(v.max())
        assert lo >= float(np.min(single.low)) and hi <= float(np.max(single.high)), k
    a = env.action_space.sample()
    assert len(a) == 64 and env.single_action_space.n == 4
    obs, rew, term, trunc, _ = env.step(a)
    assert rew.shape == (64,) and term.dtype == torch.bool and trunc.dtype == torch.bool
