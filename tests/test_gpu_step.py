"""CUDA env step / reset / fields through the C ABI vs (a) the golden vectors recorded from the
unmodified reference and (b) the closed-form oracle on seeded inputs.  Bit-exact: integer obs,
float64 reward compared as uint64 bit patterns."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402


def _engine():
    import maze_b200
    return maze_b200


def _bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def _collect_golden_mazes():
    out = []
    z, meta = load_golden("metrics")
    for m in meta:
        out.append(dict(grid=z[f"m{m['id']}_grid"], start=m["start"], goal=m["goal"], toroidal=bool(m["no_border"])))
    for name in ("bestdir", "bestdir81"):
        z, meta = load_golden(name)
        for m in meta:
            out.append(dict(grid=z[f"m{m['id']}_grid"], start=m["start"], goal=m["goal"], toroidal=m["topology"] == "torus"))
    return out


def test_fields_match_oracle_tables():
    from oracle.env_port import MazeTables
    mb = _engine()
    mazes = _collect_golden_mazes()
    pool = mb.MazePool.from_grids([m["grid"] for m in mazes], [m["start"] for m in mazes],
                                  [m["goal"] for m in mazes], [m["toroidal"] for m in mazes])
    meta = pool.meta_host()
    for i, m in enumerate(mazes):
        t = MazeTables(m["grid"], m["start"], m["goal"], m["toroidal"])
        np.testing.assert_array_equal(pool.table_host(i), t.table, err_msg=f"maze {i}")
        assert meta[i, mb.cabi.META_MAX_STEPS] == t.max_steps
        assert meta[i, mb.cabi.META_SOL_LEN] == int(t.dgoal[tuple(m["start"])]) + 1


@pytest.mark.parametrize("name", ["bestdir", "bestdir81"])
def test_fields_best_dir_matches_reference(name):
    """The step table's best-next code against the reference's _find_best_next_cell on every open block
    (bestdir81.npz: 81 x 81 mazes, dfs included, i.e. the beyond-the-depth-limit branch at the headline size)."""
    mb = _engine()
    z, meta = load_golden(name)
    pool = mb.MazePool.from_grids([z[f"m{m['id']}_grid"] for m in meta], [m["start"] for m in meta],
                                  [m["goal"] for m in meta], [m["topology"] == "torus" for m in meta])
    dr = np.array([1, -1, 0, 0, 0, 0, 0, 0]), np.array([0, 0, 1, -1, 0, 0, 0, 0])
    for i, m in enumerate(meta):
        grid, nxt = z[f"m{m['id']}_grid"], z[f"m{m['id']}_next"]
        S = m["shape"]
        code = (pool.table_host(i).reshape(-1)[:S * S].reshape(S, S) >> mb.cabi.TAB_CODE_SHIFT) & 7
        rr, cc = np.nonzero(grid)
        k = code[rr, cc]
        nr, nc = rr + dr[0][k], cc + dr[1][k]
        if m["topology"] == "torus":
            nr, nc = nr % S, nc % S
        np.testing.assert_array_equal(np.stack([nr, nc], 1), nxt[rr, cc], err_msg=str(m))


@pytest.mark.parametrize("name", ["bestdir", "bestdir81"])
def test_fields_max_steps_match_reference(name):
    mb = _engine()
    z, meta = load_golden(name)
    pool = mb.MazePool.from_grids([z[f"m{m['id']}_grid"] for m in meta], [m["start"] for m in meta],
                                  [m["goal"] for m in meta], [m["topology"] == "torus" for m in meta])
    got = pool.meta_host()[:, mb.cabi.META_MAX_STEPS]
    np.testing.assert_array_equal(got, [m["max_steps"] for m in meta])


@pytest.mark.parametrize("name", ["steps", "steps81"])
def test_step_matches_reference_traces(name):
    """Replay the reference's action tapes (no autoreset, stepping continues past done exactly as
    the reference env allows) and compare every output of every step.  steps81.npz (round 2): 81 x 81 traces for
    prim&kill, dfs (euclid) and r-prim, prim&kill (torus)."""
    mb = _engine()
    z, meta = load_golden(name)
    meta = [m for m in meta if not m["enrich"]]
    pool = mb.MazePool.from_grids([z[f"m{m['id']}_grid"] for m in meta], [m["start"] for m in meta],
                                  [m["goal"] for m in meta], [m["topology"] == "torus" for m in meta])
    np.testing.assert_array_equal(pool.meta_host()[:, mb.cabi.META_MAX_STEPS], [m["max_steps"] for m in meta])
    pairs = [(k, m, j) for k, m in enumerate(meta) for j in m["tapes"]]
    env_maze = torch.tensor([k for k, _, _ in pairs], dtype=torch.int32, device="cuda")
    batch = mb.MazeBatch(pool, len(pairs), env_maze=env_maze)
    tapes = [z[f"m{m['id']}_t{j}_action"] for _, m, j in pairs]
    T = max(len(t) for t in tapes)
    acts = np.zeros((T, len(pairs)), dtype=np.uint8)
    for e, t in enumerate(tapes):
        acts[:len(t), e] = t
    acts_d = torch.from_numpy(acts).cuda()

    def check(t):
        ag, tg, bd = batch.agent.cpu().numpy(), batch.target.cpu().numpy(), batch.best_dir.cpu().numpy()
        rw, te, tr = batch.reward.cpu().numpy(), batch.terminated.cpu().numpy(), batch.truncated.cpu().numpy()
        for e, (_, m, j) in enumerate(pairs):
            pre = f"m{m['id']}_t{j}_"
            if t > len(z[pre + "action"]):
                continue
            np.testing.assert_array_equal(ag[e], z[pre + "agent"][t], err_msg=f"{pre} step {t}")
            np.testing.assert_array_equal(tg[e], z[pre + "target"][t])
            np.testing.assert_array_equal(bd[e], z[pre + "best"][t], err_msg=f"{pre} step {t}")
            if t > 0:
                assert _bits(rw[e]) == _bits(z[pre + "reward"][t - 1]), (pre, t, rw[e], z[pre + "reward"][t - 1])
                assert bool(te[e]) == bool(z[pre + "term"][t - 1]) and bool(tr[e]) == bool(z[pre + "trunc"][t - 1]), (pre, t)

    batch.reset()
    check(0)
    for t in range(T):
        batch.step(acts_d[t], mode=0)
        check(t + 1)


@pytest.mark.parametrize("visit_layout", ["cell", "tile"])
@pytest.mark.parametrize("win_next", [False, True])
def test_step_autoreset_matches_oracle(win_next, visit_layout):
    """Many envs over a mixed pool (euclid + torus, 11..81 blocks), autoreset on, 70 % greedy
    policy so that goals are reached; compared against the closed-form oracle every step."""
    from oracle.vector import OracleVector
    mb = _engine()
    mazes = _collect_golden_mazes()
    M = len(mazes)
    pool = mb.MazePool.from_grids([m["grid"] for m in mazes], [m["start"] for m in mazes],
                                  [m["goal"] for m in mazes], [m["toroidal"] for m in mazes])
    B = 3 * M
    rng = np.random.default_rng(5)
    env_maze = np.arange(B) % M
    batch = mb.MazeBatch(pool, B, env_maze=torch.from_numpy(env_maze.astype(np.int32)).cuda(), stats=True, pool_stride=7,
                          visit_layout=visit_layout)
    ora = OracleVector(mazes, env_maze, autoreset=True, win_next=win_next, pool_stride=7)
    mode = mb.cabi.STEP_AUTORESET | (mb.cabi.STEP_WIN_NEXT if win_next else 0)
    batch.reset()
    ref = ora.reset()
    episodes = wins = 0
    for t in range(400):
        best = ref["best"]
        greedy = np.zeros(B, dtype=np.int64)
        for a, (dr, dc) in enumerate(((1, 0), (-1, 0), (0, 1), (0, -1))):
            hit = (np.sign(best[:, 0]) * (np.abs(best[:, 0]) == 1) == -dr) & (np.sign(best[:, 1]) * (np.abs(best[:, 1]) == 1) == -dc)
            greedy[hit] = a
        acts = np.where(rng.random(B) < 0.7, greedy, rng.integers(0, 4, B)).astype(np.uint8)
        batch.step(torch.from_numpy(acts).cuda(), mode=mode)
        ref = ora.step(acts)
        np.testing.assert_array_equal(batch.agent.cpu().numpy(), ref["agent"], err_msg=f"step {t}")
        np.testing.assert_array_equal(batch.target.cpu().numpy(), ref["target"], err_msg=f"step {t}")
        np.testing.assert_array_equal(batch.best_dir.cpu().numpy(), ref["best"], err_msg=f"step {t}")
        np.testing.assert_array_equal(_bits(batch.reward.cpu().numpy()), _bits(ref["reward"]), err_msg=f"step {t}")
        np.testing.assert_array_equal(batch.terminated.cpu().numpy().astype(bool), ref["term"])
        np.testing.assert_array_equal(batch.truncated.cpu().numpy().astype(bool), ref["trunc"])
        episodes += int((ref["term"] | ref["trunc"]).sum())
        wins += int(ref["term"].sum())
    st = batch.stats.cpu().numpy()
    assert st[0] == episodes and st[1] == wins and wins > 0
    np.testing.assert_array_equal(batch.env_maze.cpu().numpy(), ora.env_maze)


def test_epoch_wraparound_clears_visits():
    """More than 255 episodes on one env forces the visit-epoch wrap; results must not change."""
    from oracle.vector import OracleVector
    mb = _engine()
    z, meta = load_golden("metrics")
    m = next(x for x in meta if x["shape"] == 11)
    maze = dict(grid=z[f"m{m['id']}_grid"], start=m["start"], goal=m["goal"], toroidal=False)
    pool = mb.MazePool.from_grids([maze["grid"]], [maze["start"]], [maze["goal"]], [False])
    B = 40
    batch = mb.MazeBatch(pool, B)
    ora = OracleVector([maze], [0] * B, autoreset=True)
    rng = np.random.default_rng(11)
    batch.reset(); ora.reset()
    n_resets = np.zeros(B, dtype=int)
    t = 0
    while n_resets.min() < 300 and t < 40000:
        acts = rng.integers(0, 4, B).astype(np.uint8)
        batch.step(torch.from_numpy(acts).cuda(), mode=mb.cabi.STEP_AUTORESET)
        ref = ora.step(acts)
        n_resets += (ref["term"] | ref["trunc"])
        if t % 50 == 0 or n_resets.min() in (254, 255, 256, 257):
            np.testing.assert_array_equal(batch.agent.cpu().numpy(), ref["agent"], err_msg=f"step {t}")
            np.testing.assert_array_equal(_bits(batch.reward.cpu().numpy()), _bits(ref["reward"]), err_msg=f"step {t}")
        t += 1
    assert n_resets.min() >= 300
    assert batch.state_host()["epoch"].max() <= 255


def test_argument_errors_are_reported():
    mb = _engine()
    with pytest.raises(ValueError):
        mb.MazePool(2, (20, 21))
    with pytest.raises(mb.cabi.MazeError):
        mb.MazePool(2, (21, 21), device="cpu")


@pytest.mark.parametrize("visit_layout", ["cell", "tile"])
@pytest.mark.parametrize("chunk", [0, 100])
def test_step_many_equals_repeated_steps(visit_layout, chunk):
    """maze_step_many over a K-step action tape == K maze_step calls: per-step trace, final state,
    visit counters, episode statistics (autoreset + pool cycling on)."""
    mb = _engine()
    mazes = _collect_golden_mazes()
    M = len(mazes)
    pool = mb.MazePool.from_grids([m["grid"] for m in mazes], [m["start"] for m in mazes],
                                  [m["goal"] for m in mazes], [m["toroidal"] for m in mazes])
    B, K = 5 * M + 3, 96
    env_maze = torch.from_numpy((np.arange(B) % M).astype(np.int32)).cuda()
    mode = mb.cabi.STEP_AUTORESET | mb.cabi.STEP_WIN_NEXT
    rng = np.random.default_rng(3)
    acts = torch.from_numpy(rng.integers(0, 4, (3 * K, B)).astype(np.uint8)).cuda()
    one = mb.MazeBatch(pool, B, env_maze=env_maze.clone(), stats=True, pool_stride=5, visit_layout=visit_layout, visit_bits=True)
    many = mb.MazeBatch(pool, B, env_maze=env_maze.clone(), stats=True, pool_stride=5, visit_layout=visit_layout, visit_bits=True)
    one.reset(); many.reset()
    for burst in range(3):
        tape = acts[burst * K:(burst + 1) * K]
        ref = {k: [] for k in ("agent", "best_dir", "reward", "terminated", "truncated")}
        for t in range(K):
            one.step(tape[t], mode)
            for k in ref:
                ref[k].append(getattr(one, k).clone())
        tr = many.step_many(tape, mode, trace=True, chunk_envs=chunk)
        for k in ref:
            assert torch.equal(tr[k], torch.stack(ref[k])), (burst, k)
        for k in ("state", "env_maze", "agent", "target", "best_dir", "reward", "terminated", "truncated", "visits", "ep_return", "stats", "visit_bits"):
            assert torch.equal(getattr(many, k), getattr(one, k)), (burst, k)
        assert abs(float(many.stats_return.item()) - float(one.stats_return.item())) < 1e-6
    assert int(one.stats[0].item()) > 0
    assert many.step_many(acts[:4], mode) is None          # no trace: only the final outputs
