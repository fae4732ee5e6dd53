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


# ---- configs[2] and configs[3] at their stated sizes (BASELINE.json) ---------------------------
class _SampledOracle:
    """Replays a sample of envs of a regenerate-on-win MazeVectorEnv through the closed-form oracle: rebuilt from the
    pool's grid whenever the env restarts after a win (the slot then holds the regenerated maze)."""

    def __init__(self, env, ids, toroidal):
        self.env, self.ids, self.tor = env, [int(i) for i in ids], toroidal
        self.cur = [self._make(e) for e in self.ids]
        for o in self.cur:
            o.reset()
        self.pending = [False] * len(self.ids)
        self.won = [False] * len(self.ids)
        self.wins = self.steps = self.rebuilt = 0

    def _make(self, e):
        meta = self.env.pool.meta[e].cpu().numpy()
        grid = self.env.pool.grid_host(e).copy()
        assert check_perfect_maze(np.pad(grid, 1) if self.tor else grid)[0], e
        return ClosedFormEnv(grid, _meta_rc(meta[2]), _meta_rc(meta[3]), self.tor)

    def check(self, t, acts, agent, best, reward, term, trunc):
        for k, e in enumerate(self.ids):
            if self.pending[k]:
                if self.won[k]:
                    self.cur[k] = self._make(e)
                    self.rebuilt += 1
                o, _ = self.cur[k].reset()
                self.pending[k] = False
                assert reward[e] == 0.0 and not term[e] and not trunc[e], (t, e)
            else:
                o, r, otr, ote, _ = self.cur[k].step(int(acts[e]))
                assert np.float64(r).view(np.uint64) == np.float64(reward[e]).view(np.uint64), (t, e, r, reward[e])
                assert bool(ote) == bool(term[e]) and bool(otr) == bool(trunc[e]), (t, e)
                self.pending[k], self.won[k] = bool(ote or otr), bool(ote)
                self.wins += int(ote)
                self.steps += 1
            assert tuple(o["agent"]) == tuple(agent[e]) and tuple(o["best dir"]) == tuple(best[e]), (t, e)


def _greedy(best, S_of_env, rng, p):
    """Follow obs['best dir'] (= agent - next; +-(S - 1) components are wrapped moves) with probability p."""
    B = best.shape[0]
    g = np.zeros(B, dtype=np.uint8)
    r, c = best[:, 0], best[:, 1]
    g[(r == -1) | (r == S_of_env - 1)] = 0
    g[(r == 1) | (r == -(S_of_env - 1))] = 1
    g[(c == -1) | (c == S_of_env - 1)] = 2
    g[(c == 1) | (c == -(S_of_env - 1))] = 3
    return np.where(rng.random(B) < p, g, rng.integers(0, 4, B)).astype(np.uint8)


def test_config3_at_size_toroidal_81_mixed_generators_regenerate_on_win():
    """configs[2] at its stated size: toroidal 40x40 = 81 x 81 blocks (the shape the reference's examples pass to both
    topologies, test_q_toroid.py:17; 41 logical lines), r-prim / dfs / prim&kill mixed per slot, 4096 envs, every win
    regenerates the env's maze; 64 sampled envs are replayed through the oracle, all envs are checked by invariants."""
    import maze_b200 as mb
    B, S = 4096, 81
    env = mb.MazeVectorEnv(B, shape=(S, S), topology="toroidal", algorithms=["r-prim", "dfs", "prim&kill"], seed=31, on_win="regenerate", stats=True)
    obs, _ = env.reset()
    rng = np.random.default_rng(1)
    ora = _SampledOracle(env, rng.choice(B, 64, replace=False), True)
    for t in range(600):
        acts = _greedy(obs["best dir"].cpu().numpy(), S, rng, 0.8)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(acts).cuda())
        ora.check(t, acts, obs["agent"].cpu().numpy(), obs["best dir"].cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy())
    env.drain_regeneration()
    stats = env.episode_statistics()
    meta = env.pool.meta_host()
    assert ora.rebuilt >= 10 and ora.steps > 30000 and stats["wins"] > 10 * B // 64
    assert meta[:, mb.cabi.META_SPARE].sum() == B + stats["wins"]               # one generation per slot + one per win
    assert (meta[:, 0] == S).all() and (meta[:, 1] == S).all() and (meta[:, mb.cabi.META_FLAGS] & 1).all()
    algos = (meta[:, mb.cabi.META_FLAGS] >> 8) & 0xff
    assert set(np.unique(algos)) == {0, 1, 2}
    st = env.batch.state_host()
    tab = env.pool.table[torch.arange(B, device="cuda"), torch.from_numpy(st["r"] * S + st["c"]).cuda()]
    assert bool((tab & 1).all())                                                # nobody stands in a wall of its (possibly new) maze


def test_config4_at_size_curriculum_21_to_129_with_double_q():
    """configs[3] at its stated size: mazes between 10x10 cells (21 blocks) and 64x64 cells (129 blocks) growing by (4, 4)
    blocks per win (simple_variable_maze_env.py:93-112), generator switched after 5 and 10 wins
    (off_policy_trainer.py:302-310), device DQAgent learning off-policy; 16 sampled envs replayed through the oracle on
    every maze they meet, shape / generator invariants for all.  The envs start spread over the curriculum (21, 61, 101,
    121 blocks): from 21 alone the climb stalls around 41 blocks within a test's budget, because past the A* depth limit
    the 'best dir' hint is only a Manhattan heuristic (in the reference too) and prim&kill mazes stop being solved by
    following it -- the large shapes must be reached to be tested."""
    import maze_b200 as mb
    from maze_b200.agents import DQAgent
    B = 512
    starts = [(21, 21), (61, 61), (101, 101), (121, 121)]
    env = mb.MazeVectorEnv(B, shape=(129, 129), start_shape=starts, grow=4, algorithms="r-prim", seed=5, on_win="regenerate",
                           algorithm_schedule=((5, "prim&kill"), (10, "dfs")), stats=True)
    agent = DQAgent(env, learning_rate=0.2, initial_epsilon=0.9, epsilon_decay=300, final_epsilon=0.05, discount_factor=0.8, eta=1e-3,
                    envs_per_agent=1, seed=1, capacity=1 << 22)
    obs, _ = env.reset()
    rng = np.random.default_rng(2)
    ora = _SampledOracle(env, rng.choice(B, 16, replace=False), False)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(3000):
        acts = agent.get_action()
        bd = obs["best dir"]
        follow = torch.where(bd[:, 0] == -1, 0, torch.where(bd[:, 0] == 1, 1, torch.where(bd[:, 1] == -1, 2, 3))).to(torch.uint8)
        acts = torch.where(torch.rand(B, device="cuda", generator=gen) < 0.9, follow, acts)
        agent.core.last_action.copy_(acts)
        obs, rew, term, trunc, _ = env.step(acts)
        agent.update()
        ora.check(t, acts.cpu().numpy(), obs["agent"].cpu().numpy(), obs["best dir"].cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(),
                  trunc.cpu().numpy())
    agent.core.check_overflow()
    env.drain_regeneration()
    wins = env.wins.cpu().numpy()
    meta = env.pool.meta_host()
    assert env.episode_statistics()["wins"] == wins.sum() and ora.rebuilt > 50
    assert meta[:, 0].max() == 129 and (meta[:, 0] == 129).sum() >= 20    # the 64 x 64-cell end of the curriculum is reached and played
    for e in range(B):
        assert meta[e, 0] == meta[e, 1] == min(129, starts[e % 4][0] + 4 * int(wins[e]))
        algo = (int(meta[e, mb.cabi.META_FLAGS]) >> 8) & 0xff
        assert algo == (mb.cabi.ALGO_DFS if wins[e] >= 10 else mb.cabi.ALGO_PRIMKILL if wins[e] >= 5 else mb.cabi.ALGO_RPRIM)
