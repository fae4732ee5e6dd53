"""configs[1] at full size -- 1000 device-generated 40x40 (81x81 block) r-prim mazes, 1 024 agents
each = 1 048 576 envs -- checked through size-independent properties, plus an oracle replay of a
sampled subset ("64 of the mazes exported as block grids, same action
tapes replayed through the oracle").  Actions: 70 % best-dir following, 30 % uniform, from a
seeded device tape, autoreset on, winners move on to the next maze of the pool."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

M, A, T = 1000, 1024, 420
ACTIONS = ((1, 0), (-1, 0), (0, 1), (0, -1))


def _run(pool, mb, env_slice, seed=7, record=None):
    """Step the envs env_slice of the B = M * A batch for T steps; returns final tensors (+ trace of `record` envs)."""
    lo, hi = env_slice
    B = hi - lo
    env_maze = (torch.arange(lo, hi, device="cuda", dtype=torch.int32) // A)
    batch = mb.MazeBatch(pool, B, env_maze=env_maze, stats=True, pool_stride=1)
    mode = mb.cabi.STEP_AUTORESET | mb.cabi.STEP_WIN_NEXT
    batch.reset()
    trace = []
    g = torch.Generator(device="cuda")
    for t in range(T):
        g.manual_seed(seed * 100003 + t)                      # the same tape whatever the slicing
        u = torch.rand(M * A, device="cuda", generator=g)[lo:hi]
        rnd = (u * 4096).to(torch.int64) % 4
        bd = batch.best_dir
        follow = torch.where(bd[:, 0] == -1, 0, torch.where(bd[:, 0] == 1, 1, torch.where(bd[:, 1] == -1, 2, 3)))
        acts = torch.where(u < 0.7, follow, rnd).to(torch.uint8)
        batch.step(acts, mode)
        if record is not None:
            idx = record
            trace.append((acts[idx].cpu().numpy(), batch.agent[idx].cpu().numpy(), batch.reward[idx].cpu().numpy(),
                          batch.terminated[idx].cpu().numpy(), batch.truncated[idx].cpu().numpy(), batch.best_dir[idx].cpu().numpy(),
                          batch.env_maze[idx].cpu().numpy()))
    torch.cuda.synchronize()
    return batch, trace


def test_full_size_properties_and_sampled_oracle_replay():
    import maze_b200 as mb
    from oracle.env_port import ClosedFormEnv
    from oracle.generation import check_perfect_maze
    pool = mb.MazePool(M, (81, 81))
    pool.generate(algorithms="r-prim", seed=1234)
    meta = pool.meta_host()
    B = M * A
    sample = torch.from_numpy(np.random.default_rng(0).choice(B, 64, replace=False)).cuda()
    batch, trace = _run(pool, mb, (0, B), record=sample)

    # ---- properties over all 1 048 576 envs
    st = batch.state_host()
    r, c = torch.from_numpy(st["r"]).cuda(), torch.from_numpy(st["c"]).cuda()
    em = batch.env_maze.long()
    W = pool.meta[em, mb.cabi.META_W].long()
    tab = pool.table[em, r * W + c]
    assert bool((tab & 1).all()), "an agent stands on a wall"
    np.testing.assert_array_equal(batch.agent.cpu().numpy(), np.stack([st["r"], st["c"]], 1))
    goal = pool.meta[em, mb.cabi.META_GOAL]
    assert torch.equal(batch.target[:, 0], goal & 0xffff) and torch.equal(batch.target[:, 1], goal >> 16)
    assert torch.equal(torch.from_numpy(st["tab"]).cuda().to(torch.uint8), tab), "cached step-table byte is stale"
    max_steps = pool.meta[em, mb.cabi.META_MAX_STEPS].cpu().numpy()
    assert (st["steps"] <= max_steps + 1).all()
    rew, term, trunc = batch.reward.cpu().numpy(), batch.terminated.cpu().numpy().astype(bool), batch.truncated.cpu().numpy().astype(bool)
    assert (rew[trunc] == -1.0).all() and (rew[term & ~trunc] == 1.0).all()
    on_goal = (batch.agent == batch.target).all(1).cpu().numpy()
    assert (on_goal[term]).all()
    lut = set(mb.cabi.reward_lut(0).tolist()) | set(mb.cabi.reward_lut(1).tolist()) | set(mb.cabi.reward_lut(2)[:3].tolist()) | {1.0, -1.0, 0.0}
    assert set(np.unique(rew).tolist()) <= lut
    s = batch.stats.cpu().numpy()
    assert s[0] == s[1] + s[2] and s[1] > 10000 and s[2] > 0      # episodes = wins + truncations; both kinds happened

    # ---- the same run again is bit-identical; so is a run split in two halves (sharding by env index)
    again, _ = _run(pool, mb, (0, B))
    for name in ("state", "agent", "best_dir", "reward", "terminated", "truncated", "env_maze"):
        assert torch.equal(getattr(again, name), getattr(batch, name)), name
    half = B // 2
    for lo, hi in ((0, half), (half, B)):
        part, _ = _run(pool, mb, (lo, hi))
        for name in ("state", "agent", "reward", "terminated", "truncated", "env_maze"):
            assert torch.equal(getattr(part, name), getattr(batch, name)[lo:hi]), (name, lo)
    del again, part

    # ---- oracle replay of the 64 sampled envs over their own action tapes
    envs, mazes_of = {}, {}
    idx = sample.cpu().numpy()

    def oracle_env(m):
        if m not in envs:
            grid = pool.grid_host(m)
            assert check_perfect_maze(grid)[0]
            envs[m] = (grid, (int(meta[m, 2]) & 0xffff, int(meta[m, 2]) >> 16), (int(meta[m, 3]) & 0xffff, int(meta[m, 3]) >> 16))
        return ClosedFormEnv(*envs[m], False)

    cur = [oracle_env(int(e) // A) for e in idx]
    for o in cur:
        o.reset()
    pending = [False] * len(idx)
    won = [False] * len(idx)
    maze_now = [int(e) // A for e in idx]
    for t, (acts, ag, rw, te, tr, bd, em_t) in enumerate(trace):
        for k in range(len(idx)):
            if pending[k]:
                if won[k]:
                    maze_now[k] = (maze_now[k] + 1) % M
                    cur[k] = oracle_env(maze_now[k])
                o, _ = cur[k].reset()
                pending[k] = False
                assert rw[k] == 0.0 and not te[k] and not tr[k]
            else:
                o, r_, otr, ote, _ = cur[k].step(int(acts[k]))
                assert np.float64(r_).view(np.uint64) == rw[k].view(np.uint64), (t, k)
                assert bool(ote) == bool(te[k]) and bool(otr) == bool(tr[k]), (t, k)
                pending[k], won[k] = bool(ote or otr), bool(ote)
            assert tuple(o["agent"]) == tuple(ag[k]) and tuple(o["best dir"]) == tuple(bd[k]), (t, k)
            assert em_t[k] == maze_now[k]
