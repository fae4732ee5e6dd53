"""End-to-end scenarios of BASELINE.json configs[2..3] on the device: toroidal mixed-generator mazes
regenerated on every win, and the variable-size curriculum with double Q-learning.  Every maze that
appears must be a valid spanning tree, and the env must stay step-exact against the oracle on the
regenerated mazes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.env_port import ClosedFormEnv  # noqa: E402
from oracle.generation import check_perfect_maze  # noqa: E402


def _meta_rc(v):
    return int(v) & 0xffff, int(v) >> 16


def test_config3_toroidal_mixed_generators_regenerate_on_win():
    import maze_b200 as mb
    B, S = 96, 21
    env = mb.MazeVectorEnv(B, shape=(S, S), topology="toroidal", algorithms=["r-prim", "dfs", "prim&kill"], seed=11,
                           on_win="regenerate", stats=True)
    obs, _ = env.reset()
    pool = env.pool
    oracles = {}

    def oracle_for(e):
        meta = pool.meta_host()[e]
        key = (e, int(meta[mb.cabi.META_SPARE]))
        if key not in oracles:
            grid = pool.grid_host(e).copy()
            assert check_perfect_maze(np.pad(grid, 1))[0], key
            oracles[key] = ClosedFormEnv(grid, _meta_rc(meta[2]), _meta_rc(meta[3]), True)
        return oracles[key]

    cur = [oracle_for(e) for e in range(B)]
    pending = np.zeros(B, bool)
    rng = np.random.default_rng(0)
    regenerated = 0
    for t in range(300):
        best = obs["best dir"].cpu().numpy()
        greedy = np.zeros(B, dtype=np.uint8)
        for a, (dr, dc) in enumerate(((1, 0), (-1, 0), (0, 1), (0, -1))):   # wrap-aware: best dir is agent - next
            hit = (np.sign(best[:, 0]) * (np.abs(best[:, 0]) == 1) == -dr) & (np.sign(best[:, 1]) * (np.abs(best[:, 1]) == 1) == -dc)
            wrap = ((best[:, 0] == dr * (S - 1)) & (dr != 0) & (best[:, 1] == 0)) | ((best[:, 1] == dc * (S - 1)) & (dc != 0) & (best[:, 0] == 0))
            greedy[hit | wrap] = a
        acts = np.where(rng.random(B) < 0.8, greedy, rng.integers(0, 4, B)).astype(np.uint8)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(acts).cuda())
        ag, rw, te, tr = obs["agent"].cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for e in range(B):
            if pending[e]:
                cur[e] = oracle_for(e)          # after a win the slot holds a new maze
                o, _ = cur[e].reset()
                pending[e] = False
                assert rw[e] == 0.0 and not te[e] and not tr[e]
            else:
                o, r, otr, ote, _ = cur[e].step(int(acts[e]))
                assert float(r) == rw[e] and bool(ote) == bool(te[e]) and bool(otr) == bool(tr[e]), (t, e)
                pending[e] = bool(ote or otr)
                regenerated += int(ote)
            assert tuple(o["agent"]) == tuple(ag[e]), (t, e)
    env.drain_regeneration()   # the winners of the last step (their slots are regenerated at the start of the next one)
    stats = env.episode_statistics()
    assert regenerated > 20 and stats["wins"] == regenerated
    gen_counts = pool.meta_host()[:, mb.cabi.META_SPARE]
    assert gen_counts.sum() == B + regenerated


def test_curriculum_kernel_walks_shapes_and_generators():
    """maze_curriculum on every slot, six times: +(4, 4) per win capped at the pool shape
    (simple_variable_maze_env.py:93-112), generator switched at the win thresholds
    (off_policy_trainer.py:302-310), each regenerated maze valid."""
    import maze_b200 as mb
    M = 24
    pool = mb.MazePool(M, (31, 31))
    pool.generate(shapes=(15, 15), algorithms="r-prim", seed=2)
    wins = torch.zeros(M, dtype=torch.int32, device="cuda")
    ids = torch.arange(M, dtype=torch.int32, device="cuda")
    count = torch.tensor([M - 4], dtype=torch.int32, device="cuda")   # the last four slots never win
    for k in range(1, 7):
        pool.curriculum(ids, count, wins, grow=4, schedule=((2, "prim&kill"), (4, "dfs")))
        pool.generate(ids=ids, count_dev=count, configure=False, seed=2)
        meta = pool.meta_host()
        want_shape = min(31, 15 + 4 * k)
        want_algo = mb.cabi.ALGO_DFS if k >= 4 else mb.cabi.ALGO_PRIMKILL if k >= 2 else mb.cabi.ALGO_RPRIM
        for m in range(M):
            won = m < M - 4
            assert meta[m, 0] == meta[m, 1] == (want_shape if won else 15)
            assert (meta[m, mb.cabi.META_FLAGS] >> 8) & 0xff == (want_algo if won else mb.cabi.ALGO_RPRIM)
            assert meta[m, mb.cabi.META_SPARE] == (1 + k if won else 1)
            grid = pool.grid_host(m)
            assert grid.shape == (meta[m, 0], meta[m, 1]) and check_perfect_maze(grid)[0]
    np.testing.assert_array_equal(wins.cpu().numpy(), [6] * (M - 4) + [0] * 4)


def test_config4_variable_size_curriculum_with_double_q():
    """Mazes start at 15x15 blocks, grow by (4, 4) per win up to 31x31 and switch generator after
    2 and 4 wins; a device DQAgent (one learner per env) learns off-policy from a behaviour policy
    that mostly follows the 'best dir' hint.  (Beyond the A* depth limit 2*min(H, W) the hint is
    only a Manhattan heuristic -- in the reference too -- so long prim&kill / dfs mazes are not
    always solved by following it; the curriculum invariants must hold wherever each env got to.)"""
    import maze_b200 as mb
    from maze_b200.agents import DQAgent
    B = 256
    env = mb.MazeVectorEnv(B, shape=(31, 31), start_shape=(15, 15), grow=4, algorithms="r-prim", seed=3, on_win="regenerate",
                           algorithm_schedule=((2, "prim&kill"), (4, "dfs")), stats=True)
    agent = DQAgent(env, learning_rate=0.2, initial_epsilon=0.9, epsilon_decay=300, final_epsilon=0.05,
                    discount_factor=0.8, eta=1e-3, envs_per_agent=1, seed=1)
    obs, _ = env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(800):
        acts = agent.get_action()
        bd = obs["best dir"]   # = agent - next
        follow = torch.where(bd[:, 0] == -1, 0, torch.where(bd[:, 0] == 1, 1, torch.where(bd[:, 1] == -1, 2, 3))).to(torch.uint8)
        use = torch.rand(B, device="cuda", generator=gen) < 0.85
        acts = torch.where(use, follow, acts)
        agent.core.last_action.copy_(acts)
        obs, _, _, _, _ = env.step(acts)
        agent.update()
    agent.core.check_overflow()
    env.drain_regeneration()
    wins = env.wins.cpu().numpy()
    meta = env.pool.meta_host()
    assert wins.max() >= 2 and wins.sum() > B and env.episode_statistics()["wins"] == wins.sum()
    for e in range(B):
        H, W = int(meta[e, 0]), int(meta[e, 1])
        assert H == W == min(31, 15 + 4 * int(wins[e]))
        algo = (int(meta[e, mb.cabi.META_FLAGS]) >> 8) & 0xff
        assert algo == (mb.cabi.ALGO_DFS if wins[e] >= 4 else mb.cabi.ALGO_PRIMKILL if wins[e] >= 2 else mb.cabi.ALGO_RPRIM)
        if e % 8 == 0:
            grid = env.pool.grid_host(e)
            assert grid.shape == (H, W) and check_perfect_maze(grid)[0]
    assert len(agent.core.table_host("a")) > B


@pytest.mark.parametrize("on_win", ["next", "regenerate"])
def test_checkpoint_resume_is_bit_identical(tmp_path, on_win):
    """state_dict -> torch.save -> fresh env -> load_state_dict: the continued run equals the uninterrupted one
    (positions, rewards, flags, regenerated mazes, statistics), also with regeneration on win."""
    import maze_b200 as mb
    kw = dict(shape=(21, 21), algorithms=["r-prim", "dfs", "prim&kill"], seed=11, on_win=on_win, stats=True)
    B = 96
    if on_win == "next":
        kw["num_mazes"] = 12
    env = mb.MazeVectorEnv(B, **kw)
    env.reset()
    rng = np.random.default_rng(2)

    def actions(e):
        # follow the best direction most of the time so that episodes are won and mazes change
        best = e.batch.best_dir.cpu().numpy()
        a = np.full(B, 0, dtype=np.uint8)
        a[best[:, 0] == 1] = 1      # best dir = agent - next: (1, 0) means next is the row above -> action up
        a[best[:, 0] == -1] = 0
        a[best[:, 1] == 1] = 3
        a[best[:, 1] == -1] = 2
        rnd = rng.random(B) < 0.2
        a[rnd] = rng.integers(0, 4, int(rnd.sum()))
        return torch.from_numpy(a).cuda()

    for _ in range(120):
        env.step(actions(env))
    path = str(tmp_path / "env.pt")
    torch.save(env.state_dict(), path)
    rng_state = rng.bit_generator.state
    trace = []
    for _ in range(150):
        obs, rew, term, trunc, _ = env.step(actions(env))
        trace.append((obs["agent"].clone(), obs["target"].clone(), obs["best dir"].clone(), rew.clone(), term.clone(), trunc.clone()))
    assert int(env.batch.stats[1].item()) > 0          # some wins, so mazes did change
    final_meta = env.pool.meta.clone()

    other = mb.MazeVectorEnv(B, **kw)
    other.reset()
    other.load_state_dict(torch.load(path))
    rng.bit_generator.state = rng_state
    for t in range(150):
        obs, rew, term, trunc, _ = other.step(actions(other))
        got = (obs["agent"], obs["target"], obs["best dir"], rew, term, trunc)
        for a, b in zip(got, trace[t]):
            assert torch.equal(a, b), t
    assert torch.equal(other.pool.meta, final_meta)
    a, b = other.episode_statistics(), env.episode_statistics()
    assert {k: v for k, v in a.items() if k != "return_sum"} == {k: v for k, v in b.items() if k != "return_sum"}
    assert a["return_sum"] == pytest.approx(b["return_sum"], rel=1e-12)   # float atomics: summation order varies


@pytest.mark.parametrize("topology", ["euclid", "toroidal"])
def test_terminal_observation_and_replay_see_the_maze_the_episode_was_played_on(topology):
    """Regenerate-on-win with the enriched observation and the device replay ring: the observation returned by the
    winning step, the window pushed as next_state and the re-staged state all come from the OLD maze (the reference
    takes next_obs from env.step() before update_maze(), off_policy_trainer.py:160-171); the new maze appears with the
    autoreset of the following step."""
    import maze_b200 as mb
    from maze_b200.dqn import DeviceReplay, unpack_windows
    B, S = 64, 15
    tor = topology == "toroidal"
    env = mb.MazeVectorEnv(B, shape=(S, S), topology=topology, algorithms=["r-prim", "dfs", "prim&kill"], seed=21, on_win="regenerate",
                           enrich=True, stats=True)
    mem = DeviceReplay(env, 1 << 14, seed=1)
    obs, _ = env.reset()
    mem.observe()
    pool = env.pool

    def oracle_for(e):
        meta = pool.meta_host()[e]
        return ClosedFormEnv(pool.grid_host(e).copy(), _meta_rc(meta[2]), _meta_rc(meta[3]), tor, enrich=True)

    cur = [oracle_for(e) for e in range(B)]
    for o in cur:
        o.reset()
    pending = np.zeros(B, bool)
    rng = np.random.default_rng(4)
    wins = 0
    for t in range(260):
        best = env.batch.best_dir.cpu().numpy()
        greedy = np.zeros(B, dtype=np.uint8)
        for a, (dr, dc) in enumerate(((1, 0), (-1, 0), (0, 1), (0, -1))):
            hit = (np.sign(best[:, 0]) * (np.abs(best[:, 0]) == 1) == -dr) & (np.sign(best[:, 1]) * (np.abs(best[:, 1]) == 1) == -dc)
            wrap = ((best[:, 0] == dr * (S - 1)) & (dr != 0) & (best[:, 1] == 0)) | ((best[:, 1] == dc * (S - 1)) & (dc != 0) & (best[:, 0] == 0))
            greedy[hit | wrap] = a
        acts = np.where(rng.random(B) < 0.85, greedy, rng.integers(0, 4, B)).astype(np.uint8)
        acts_d = torch.from_numpy(acts).cuda()
        obs, rew, term, trunc, _ = env.step(acts_d)
        mem.push(acts_d)
        win, an, tn = obs["window"].cpu().numpy(), obs["agent"].cpu().numpy(), obs["target"].cpu().numpy()
        staged = unpack_windows(mem.stage_win).cpu().numpy()
        svec = mem.stage_vec.cpu().numpy()
        te = term.cpu().numpy()
        for e in range(B):
            if pending[e]:
                cur[e] = oracle_for(e)
                o, _ = cur[e].reset()
                pending[e] = False
            else:
                o, r, otr, ote, _ = cur[e].step(int(acts[e]))
                assert bool(ote) == bool(te[e]), (t, e)
                pending[e] = bool(ote or otr)
                wins += int(ote and not otr)       # a goal reached on the truncating step pays -1, not 1 (base_maze_env.py:205-208)
            np.testing.assert_array_equal(win[e], o["window"], err_msg=f"window env {e} step {t} (terminal: {bool(te[e])})")
            np.testing.assert_array_equal(an[e].view(np.uint64), np.asarray(o["agent"], np.float64).view(np.uint64))
            np.testing.assert_array_equal(tn[e].view(np.uint64), np.asarray(o["target"], np.float64).view(np.uint64))
            np.testing.assert_array_equal(staged[e], o["window"], err_msg=f"staged window env {e} step {t}")
            np.testing.assert_array_equal(svec[e, :4], np.concatenate([o["agent"], o["target"]]).astype(np.float32))
    assert wins > 15
    # every goal transition in the ring ends ON its own episode's goal: next_state agent == next_state target
    n = len(mem)
    goal = (mem.reward[:n] == 1.0)
    assert int(goal.sum()) == wins
    nv = mem.next_vec[:n][goal]
    assert torch.equal(nv[:, 0:2], nv[:, 2:4])
