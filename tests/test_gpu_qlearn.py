"""Device Q-learning / double Q-learning (maze_q_act / maze_q_update / maze_q_rollout) through the
C ABI: bit-exact against the tables of the unmodified reference agents (qagent.npz, replaying the
recorded numpy draws), against the oracle on many independent agents, and behavioural checks of
the Philox-driven path."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle.env_port import ClosedFormEnv  # noqa: E402
from oracle.grid import best_dir_vector  # noqa: E402
from oracle.qlearn import Draws, OracleQAgent, obs_key, parse_reference_key  # noqa: E402

KW = dict(learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=150, final_epsilon=0.05, discount_factor=0.7, eta=1e-2)


def _device_table(agent, which, shapes, tors):
    """{(agent_id, r, c, tr, tc, dr, dc): row} with the best-dir code turned back into the vector."""
    out = {}
    for (aid, r, c, tr, tc, code), row in agent.core.table_host(which).items():
        dr, dc = best_dir_vector(code, (r, c), shapes[aid], tors[aid])
        out[(aid, r, c, tr, tc, int(dr), int(dc))] = row
    return out


def _golden_setup(name, z):
    import maze_b200 as mb
    from maze_b200.agents import DQAgent, QAgent
    pool = mb.MazePool.from_grids([z["grid"]], [tuple(z["start"])], [tuple(z["goal"])], [False])
    batch = mb.MazeBatch(pool, 1, stats=True)
    agent = (DQAgent if name == "dq" else QAgent)(batch, **KW)
    agent.core.attach_tapes(z[f"{name}_u"].reshape(-1, 1), z[f"{name}_a"].reshape(-1, 1))
    batch.reset()
    return mb, batch, agent


def _check_golden(name, z, batch, agent):
    T = len(z[f"{name}_action"])
    assert int(agent.core.steps_done.cpu()[0]) == int(z[f"{name}_steps_done"])
    assert float(agent.core.gamma.cpu()[0]) == float(z[f"{name}_gamma_final"])
    pos = agent.core._tapes[2].cpu().numpy()
    assert pos[0, 0] == len(z[f"{name}_u"]) and pos[1, 0] == len(z[f"{name}_a"])
    shape = z["grid"].shape
    for ti, which in enumerate(["a"] if name == "q" else ["a", "b"]):
        dev = _device_table(agent, which, {0: shape}, {0: False})
        ref = {(0,) + parse_reference_key(str(k)): v for k, v in zip(z[f"{name}_tab{ti}_keys"], z[f"{name}_tab{ti}_vals"])}
        assert set(ref) <= set(dev)
        if name == "q":
            assert set(ref) == set(dev)
        for k, row in dev.items():
            want = ref.get(k, np.zeros(4))
            np.testing.assert_array_equal(row.view(np.uint64), want.view(np.uint64), err_msg=f"{which} {k}")
    st = batch.stats.cpu().numpy()
    return T, st


@pytest.mark.parametrize("name", ["q", "dq"])
def test_unfused_replay_matches_reference_tables(golden_qagent, name):
    z, _ = golden_qagent
    mb, batch, agent = _golden_setup(name, z)
    T = len(z[f"{name}_action"])
    k = 0
    for it in range(T + 11):   # 12 episodes: 11 autoreset steps in between
        acts = agent.get_action()
        pending = bool(batch.state_host()["flags"][0] & mb.cabi.ST_NEEDS_RESET)
        batch.step(acts, mode=mb.cabi.STEP_AUTORESET)
        agent.update()
        if not pending:
            assert int(acts.cpu()[0]) == int(z[f"{name}_action"][k]), (it, k)
            assert float(batch.reward.cpu()[0]) == float(z[f"{name}_reward"][k])
            k += 1
    assert k == T
    _check_golden(name, z, batch, agent)


@pytest.mark.parametrize("name", ["q", "dq"])
@pytest.mark.parametrize("chunks", [1, 7])
def test_fused_rollout_matches_reference_tables(golden_qagent, name, chunks):
    z, _ = golden_qagent
    mb, batch, agent = _golden_setup(name, z)
    total = len(z[f"{name}_action"]) + 11
    done = 0
    for c in range(chunks):
        k = total // chunks if c < chunks - 1 else total - done
        agent.rollout(k, mode=mb.cabi.STEP_AUTORESET)
        done += k
    agent.core.check_overflow()
    T, st = _check_golden(name, z, batch, agent)
    assert st[0] == 11 + int(bool(z[f"{name}_term"][-1] or z[f"{name}_trunc"][-1])) and st[1] == int(z[f"{name}_term"].sum())


@pytest.mark.parametrize("double_q", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_independent_agents_match_oracle(double_q, fused):
    """16 envs on different mazes (euclid + torus), one agent each, numpy-drawn tapes."""
    import maze_b200 as mb
    from maze_b200.agents import DQAgent, QAgent
    z, meta = load_golden("bestdir")
    rows = [m for m in meta if m["shape"] <= 21][:8]
    mazes = [dict(grid=z[f"m{m['id']}_grid"], start=tuple(m["start"]), goal=tuple(m["goal"]), toroidal=m["topology"] == "torus") for m in rows]
    pool = mb.MazePool.from_grids([m["grid"] for m in mazes], [m["start"] for m in mazes], [m["goal"] for m in mazes],
                                  [m["toroidal"] for m in mazes])
    B, K = 2 * len(mazes), 400
    env_maze = np.arange(B) % len(mazes)
    batch = mb.MazeBatch(pool, B, env_maze=torch.from_numpy(env_maze.astype(np.int32)).cuda())
    agent = (DQAgent if double_q else QAgent)(batch, **KW)
    rng = np.random.default_rng(17)
    u = rng.random((3 * K, B))
    a = rng.integers(0, 4, (2 * K, B)).astype(np.uint8)
    agent.core.attach_tapes(u, a)
    batch.reset()
    if fused:
        agent.rollout(K // 2, mode=mb.cabi.STEP_AUTORESET)
        agent.rollout(K - K // 2, mode=mb.cabi.STEP_AUTORESET)
    else:
        for _ in range(K):
            batch.step(agent.get_action(), mode=mb.cabi.STEP_AUTORESET)
            agent.update()
    agent.core.check_overflow()
    shapes = {e: mazes[env_maze[e]]["grid"].shape for e in range(B)}
    tors = {e: mazes[env_maze[e]]["toroidal"] for e in range(B)}
    dev_a = _device_table(agent, "a", shapes, tors)
    dev_b = _device_table(agent, "b", shapes, tors) if double_q else None
    gam, sd = agent.core.gamma.cpu().numpy(), agent.core.steps_done.cpu().numpy()
    wins = 0
    for e in range(B):
        mz = mazes[env_maze[e]]
        env = ClosedFormEnv(mz["grid"], mz["start"], mz["goal"], mz["toroidal"])
        ora = OracleQAgent(draws=Draws(u[:, e], a[:, e]), double_q=double_q, **KW)
        obs, _ = env.reset()
        pending, cum = False, 0
        for _ in range(K):
            if pending:
                obs, _ = env.reset()
                pending = False
                continue
            act = ora.get_action(obs_key(obs))
            nobs, r, trunc, term, _ = env.step(act)
            ora.update(obs_key(obs), act, r, term, obs_key(nobs))
            cum += r
            if term or trunc:
                ora.update_hyperparameter(cum > 0)
                cum, pending = 0, True
                wins += int(term)
            obs = nobs
        assert sd[e] == ora.steps_done and gam[e] == ora.discount_factor, e
        for dev, ref in ((dev_a, ora.q_a), (dev_b, ora.q_b)):
            if dev is None:
                continue
            mine = {k[1:]: v for k, v in dev.items() if k[0] == e}
            assert set(ref) <= set(mine)
            for k, row in mine.items():
                np.testing.assert_array_equal(row.view(np.uint64), ref.get(k, np.zeros(4)).view(np.uint64), err_msg=f"env {e} {k}")
    assert wins > 0


def test_shared_agent_learns_with_philox_draws():
    """One learner fed by 2048 envs of a 15x15 maze: the win rate of late episodes must beat the
    early ones; two runs with the same seed and independent agents are bit-identical."""
    import maze_b200 as mb
    from maze_b200.agents import QAgent
    z, _ = load_golden("qagent")
    pool = mb.MazePool.from_grids([z["grid"]], [tuple(z["start"])], [tuple(z["goal"])], [False])
    B = 2048
    rates = {}
    for lr in (0.1, 0.0):   # lr = 0 never learns: epsilon-greedy on an all-zero table
        batch = mb.MazeBatch(pool, B, stats=True)
        agent = QAgent(batch, envs_per_agent=B, seed=5, **dict(KW, learning_rate=lr))
        batch.reset()
        agent.rollout(1650)
        mid = batch.stats.cpu().numpy().copy()
        agent.rollout(300)
        late = batch.stats.cpu().numpy() - mid
        agent.core.check_overflow()
        assert late[0] > 1000
        rates[lr] = late[1] / late[0]
    # the reference itself wins 6 of its 12 episodes on this maze (47-step budget).  The shared learner's write races make
    # same-seed runs differ: 12 runs on a B200 gave 0.19 .. 0.66 (tools/probe_shared_learner_rates.py), lr = 0 never wins
    assert rates[0.0] < 0.02 and rates[0.1] > 0.08, rates

    tables = []
    for _ in range(2):
        b2 = mb.MazeBatch(pool, 64)
        ag = QAgent(b2, envs_per_agent=1, seed=9, **KW)
        b2.reset()
        ag.rollout(500)
        t = ag.core.table_host("a")
        tables.append({k: v.tobytes() for k, v in t.items()})
    assert tables[0] == tables[1] and len(tables[0]) > 64


def test_shared_learner_runs_differ_within_bounds():
    """With one learner behind many envs, concurrent writers of a table row race and one wins (DESIGN.md section 5,
    deviations): this is not the reference's one-env-per-agent semantics, and runs with the same seed are not bit-identical.
    What the race may and may not do: every run must still learn -- late win rate far above the never-learning baseline, which
    is 0 on this maze -- while the rates themselves spread widely (twelve runs on a B200: 0.19 .. 0.66, median 0.5,
    tools/probe_shared_learner_rates.py; tables: tools/probe_shared_learner.py).  With one env per agent there is no race and
    runs are bit-identical (previous test)."""
    import maze_b200 as mb
    from maze_b200.agents import QAgent
    z, _ = load_golden("qagent")
    pool = mb.MazePool.from_grids([z["grid"]], [tuple(z["start"])], [tuple(z["goal"])], [False])
    B = 2048

    def late_rate(lr):
        batch = mb.MazeBatch(pool, B, stats=True)
        agent = QAgent(batch, envs_per_agent=B, seed=5, **dict(KW, learning_rate=lr))
        batch.reset()
        agent.rollout(1650)
        mid = batch.stats.cpu().numpy().copy()
        agent.rollout(300)
        late = batch.stats.cpu().numpy() - mid
        return late[1] / late[0]

    base = late_rate(0.0)
    rates = [late_rate(0.1) for _ in range(3)]
    assert base < 0.02 and min(rates) > 0.08, (base, rates)      # loose on purpose: the lowest of twelve measured runs was 0.19
    assert max(rates) <= 0.9, rates
