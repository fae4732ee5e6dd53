"""Regeneration one maze ahead (maze_regen_swap / _prepare / _publish): the shadow pool must install exactly the mazes the
in-place regeneration draws -- whatever the side stream's timing -- because every path draws M(seed, slot, generation count)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_gpu_scenarios import _greedy  # noqa: E402


def _make(ahead, B, S, topology, seed=5, **kw):
    import maze_b200 as mb
    kw.setdefault("algorithms", ["r-prim", "dfs", "prim&kill"])
    return mb.MazeVectorEnv(B, shape=(S, S), topology=topology, seed=seed, on_win="regenerate", stats=True, regenerate_ahead=ahead, **kw)


def _run(env, S, steps, seed, hook=None):
    rng = np.random.default_rng(seed)
    obs, _ = env.reset()
    trace = []
    for t in range(steps):
        acts = _greedy(obs["best dir"].cpu().numpy(), S, rng, 0.9)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(acts).cuda())
        trace.append((obs["agent"].clone(), obs["target"].clone(), obs["best dir"].clone(), rew.clone(), term.clone(), trunc.clone()))
        if hook:
            hook(t, env)
    env.drain_regeneration()
    torch.cuda.synchronize()
    return trace


def _same(a, b):
    assert len(a) == len(b)
    for t, (x, y) in enumerate(zip(a, b)):
        for k, (u, v) in enumerate(zip(x, y)):
            assert torch.equal(u, v), (t, k)


@pytest.mark.parametrize("topology,S,depth", [("euclid", 9, 1), ("toroidal", 9, 2), ("toroidal", 21, 2), ("toroidal", 9, 3)])
def test_shadow_ring_installs_the_mazes_the_in_place_path_draws(topology, S, depth):
    B, steps = 2048, 500
    ref_env, env = _make(0, B, S, topology), _make(depth, B, S, topology)
    assert env.regenerate_depth == depth
    assert env.regenerate_ahead and not ref_env.regenerate_ahead
    _same(_run(ref_env, S, steps, 3), _run(env, S, steps, 3))
    assert torch.equal(ref_env.pool.grids, env.pool.grids) and torch.equal(ref_env.pool.table, env.pool.table)
    assert torch.equal(ref_env.pool.meta, env.pool.meta)
    fast, slow, jobs = env.regeneration_statistics()
    wins = int(env.episode_statistics()["wins"])
    assert fast + slow == wins and wins > B and jobs >= 1
    assert fast > 0


@pytest.mark.parametrize("topology,depth", [("euclid", 2), ("toroidal", 3)])
def test_shadow_ring_follows_the_curriculum(topology, depth):
    """Variable-size envs: +2 blocks per win up to the slot's maximum, generator switched after 2 and 4 wins
    (simple_variable_maze_env.py:93-112, off_policy_trainer.py:302-310).  The ring entries are drawn with the shape and
    generator the curriculum will have reached at their generation count -- same mazes, win counts and records as in place."""
    B, S, steps = 1024, 21, 700
    kw = dict(start_shape=[(9, 9), (11, 11), (13, 13)], grow=2, algorithms="r-prim", algorithm_schedule=((2, "prim&kill"), (4, "dfs")))
    ref_env, env = _make(0, B, S, topology, **kw), _make(depth, B, S, topology, **kw)
    assert env.regenerate_ahead and env.regenerate_depth == depth

    def run(e):
        rng = np.random.default_rng(8)
        obs, _ = e.reset()
        trace = []
        for t in range(steps):
            shape = e.pool.meta[e.batch.env_maze.long(), 0].cpu().numpy()   # the hint wraps at the slot's own size on the torus
            acts = _greedy(obs["best dir"].cpu().numpy(), shape, rng, 0.9)
            obs, rew, term, trunc, _ = e.step(torch.from_numpy(acts).cuda())
            trace.append((obs["agent"].clone(), obs["target"].clone(), obs["best dir"].clone(), rew.clone(), term.clone(), trunc.clone()))
        e.drain_regeneration()
        torch.cuda.synchronize()
        return trace

    _same(run(ref_env), run(env))
    assert torch.equal(ref_env.pool.meta, env.pool.meta) and torch.equal(ref_env.wins, env.wins)
    assert torch.equal(ref_env.pool.grids, env.pool.grids) and torch.equal(ref_env.pool.table, env.pool.table)
    meta = env.pool.meta_host()
    assert meta[:, 0].max() > 13 and len(set(((meta[:, 5] >> 8) & 0xff).tolist())) >= 2   # shapes grew, generators switched
    fast, slow, _ = env.regeneration_statistics()
    assert fast > 0 and fast + slow == int(env.wins.sum().item())


def test_slots_that_win_again_before_their_refill_are_drawn_in_place():
    """No refill job is ever launched (the 'previous job finished' query always says no): the first two wins of a slot take
    the ring entries drawn at construction, every later one fails the ready test and goes through the in-place path.  Same mazes."""
    B, S, steps = 1024, 9, 400
    ref_env, env = _make(0, B, S, "toroidal"), _make(2, B, S, "toroidal")

    class Never:
        def query(self):
            return False

    def hook(t, e):
        if t == 0:
            e._ahead.side_done = Never()

    _same(_run(ref_env, S, steps, 11), _run(env, S, steps, 11, hook))
    assert torch.equal(ref_env.pool.grids, env.pool.grids) and torch.equal(ref_env.pool.meta, env.pool.meta)
    fast, slow, jobs = env.regeneration_statistics()
    assert slow > 0 and fast > B and fast <= 2 * B


def test_refills_keep_up_at_81_blocks():
    """configs[2]'s maze size, default depth: with the side stream refilling, (nearly) every win takes the fast path."""
    B, S, steps = 4096, 81, 300
    env = _make(None, B, S, "toroidal", seed=9)
    assert env.regenerate_depth == 3
    _run(env, S, steps, 5)
    fast, slow, jobs = env.regeneration_statistics()
    wins = int(env.episode_statistics()["wins"])
    assert fast + slow == wins and wins > 0
    assert slow <= max(2, wins // 10), (fast, slow, jobs)   # timing-dependent by nature: the bound is loose, the usual value is 0
    meta = env.pool.meta_host()
    import maze_b200 as mb
    assert meta[:, mb.cabi.META_SPARE].sum() == B + wins


def test_checkpoint_rebuilds_the_shadow_pool(tmp_path):
    B, S = 1024, 9
    ref_env = _make(0, B, S, "toroidal")
    full = _run(ref_env, S, 300, 21)
    env = _make(2, B, S, "toroidal")
    rng = np.random.default_rng(21)
    obs, _ = env.reset()
    for t in range(150):
        acts = _greedy(obs["best dir"].cpu().numpy(), S, rng, 0.9)
        obs, *_ = env.step(torch.from_numpy(acts).cuda())
    torch.save(env.state_dict(), tmp_path / "env.pt")
    env2 = _make(2, B, S, "toroidal")
    env2.load_state_dict(torch.load(tmp_path / "env.pt"))
    obs = {"best dir": env2.batch.best_dir}
    for t in range(150, 300):
        acts = _greedy(obs["best dir"].cpu().numpy(), S, rng, 0.9)
        obs, rew, term, trunc, _ = env2.step(torch.from_numpy(acts).cuda())
        got = (obs["agent"], obs["target"], obs["best dir"], rew, term, trunc)
        for k, (u, v) in enumerate(zip(full[t], got)):
            assert torch.equal(u, v), (t, k)


def test_a_ring_that_does_not_fit_falls_back_to_regeneration_in_place(monkeypatch):
    """The ring is sized against the free device memory when it is first needed: with (pretended) 64 KiB free the env warns,
    regenerates in place and still produces the same steps."""
    B, S = 512, 9
    ref_env, env = _make(0, B, S, "toroidal"), _make(3, B, S, "toroidal")
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (1 << 16, 1 << 30))
    with pytest.warns(UserWarning, match="regeneration in place"):
        got = _run(env, S, 120, 4)
    assert not env.regenerate_ahead and env.regeneration_statistics() == (0, 0, 0)
    _same(_run(ref_env, S, 120, 4), got)
