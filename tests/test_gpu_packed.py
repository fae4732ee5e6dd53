"""The packed wire format of maze_step (MAZE_STEP_PACKED, include/maze_b200.h): one uint32 per env instead of 26 bytes
of wide outputs.  decode(packed) must equal the wide outputs bit for bit -- rewards as uint64 patterns -- on euclidean
and toroidal mazes, with autoreset and pool cycling, at full size (1 M envs), and through step_host_packed()."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _wide(batch):
    return dict(agent=batch.agent.cpu().numpy(), best_dir=batch.best_dir.cpu().numpy(), reward=batch.reward.cpu().numpy(),
                terminated=batch.terminated.cpu().numpy().astype(bool), truncated=batch.truncated.cpu().numpy().astype(bool))


def _compare(dec, wide, what):
    np.testing.assert_array_equal(dec["agent"], wide["agent"], err_msg=what)
    np.testing.assert_array_equal(dec["best_dir"], wide["best_dir"], err_msg=what)
    np.testing.assert_array_equal(dec["reward"].view(np.uint64), wide["reward"].view(np.uint64), err_msg=what)
    np.testing.assert_array_equal(dec["terminated"], wide["terminated"], err_msg=what)
    np.testing.assert_array_equal(dec["truncated"], wide["truncated"], err_msg=what)


@pytest.mark.parametrize("topology,shape,B", [("euclid", 21, 4096), ("toroidal", 21, 4096), ("toroidal", 15, 1000), ("euclid", 81, 1 << 20)])
def test_decoded_records_equal_wide_outputs(topology, shape, B):
    import maze_b200 as mb
    env = mb.MazeVectorEnv(B, shape=(shape, shape), topology=topology, algorithms=["r-prim", "dfs", "prim&kill"], num_mazes=min(B, 500),
                           seed=11, on_win="next")
    env.reset()
    b = env.batch
    g = torch.Generator(device="cuda").manual_seed(1)
    meta = env.pool.meta
    seen_kinds = set()
    T = 140 if B <= 4096 else 40
    for t in range(T):
        u = torch.rand(B, device="cuda", generator=g)
        bd = b.best_dir
        follow = torch.where(bd[:, 0] < 0, 0, torch.where(bd[:, 0] > 0, 1, torch.where(bd[:, 1] < 0, 2, 3)))
        if topology == "toroidal":   # un-wrapped +-(S - 1) components point the other way
            big = bd.abs().max(1)[0] > 1
            follow = torch.where(big, follow ^ 1, follow)
        acts = torch.where(u < 0.8, follow, (u * 4096).long() % 4).to(torch.uint8)
        env.step(acts, extra_mode=mb.cabi.STEP_PACKED)          # wide outputs and the record from the same launch
        rec = b.packed.cpu().numpy().view(np.uint32)
        m = meta[b.env_maze.long()]
        dec = mb.cabi.decode_records(rec, m[:, :2].cpu().numpy(), (m[:, mb.cabi.META_FLAGS] & 1).to(torch.uint8).cpu().numpy())
        _compare(dec, _wide(b), f"step {t}")
        seen_kinds |= set(np.unique((rec >> mb.cabi.REC_KIND_SHIFT) & 3).tolist())
    assert seen_kinds >= {0, 1, 2}
    if B <= 4096:   # long enough for goals / truncations / resets (the constant rewards) to occur
        assert seen_kinds == {0, 1, 2, 3} and env.episode_statistics()["wins"] > 0


def test_no_wide_mode_leaves_the_wide_buffers_alone_and_steps_identically():
    import maze_b200 as mb
    envs = [mb.MazeVectorEnv(2048, shape=(21, 21), algorithms="r-prim", num_mazes=64, seed=2, on_win="next") for _ in range(2)]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(80):
        acts = torch.randint(0, 4, (2048,), device="cuda", generator=g).to(torch.uint8)
        envs[0].step(acts)
        before = envs[1].batch.reward.clone()
        envs[1].step(acts, extra_mode=mb.cabi.STEP_PACKED | mb.cabi.STEP_NO_WIDE)
        assert torch.equal(envs[1].batch.reward, before)                       # not written
        assert torch.equal(envs[0].batch.state, envs[1].batch.state), t         # same transition
        dec = mb.cabi.decode_records(envs[1].batch.packed.cpu().numpy())
        _compare(dec, _wide(envs[0].batch), f"step {t}")
    with pytest.raises(mb.cabi.MazeError):
        envs[0].batch.step(acts, mb.cabi.STEP_NO_WIDE)


@pytest.mark.parametrize("chunks", [1, 3])
@pytest.mark.parametrize("topology", ["euclid", "toroidal"])
def test_step_host_packed_equals_step_host(topology, chunks):
    """chunks = 3: the batch is stepped as three env ranges on alternating streams (copy-in / kernel / copy-out overlapped)."""
    import maze_b200 as mb
    kw = dict(shape=(21, 21), topology=topology, algorithms=["r-prim", "dfs"], num_mazes=100, seed=5, on_win="next")
    a, b = mb.MazeVectorEnv(3000, **kw), mb.MazeVectorEnv(3000, **kw)
    a.reset()
    b.reset()
    rng = np.random.default_rng(0)
    for t in range(120):
        acts = rng.integers(0, 4, 3000).astype(np.uint8)
        oa, ra, ta, ua, _ = a.step_host(acts)
        ob, rb, tb, ub, _ = b.step_host_packed(acts, decode=True, chunks=chunks)
        for k in ("agent", "target", "best dir"):
            np.testing.assert_array_equal(oa[k], ob[k], err_msg=f"{k} step {t}")
        np.testing.assert_array_equal(ra.view(np.uint64), rb.view(np.uint64))
        np.testing.assert_array_equal(ta, tb)
        np.testing.assert_array_equal(ua, ub)
    # euclid: 4 B per env and step.  The short toroidal episodes change some env's maze on most steps here, and every such
    # step also refreshes the `target` mirror (8 B per env)
    assert b.d2h_bytes_per_step() < (0.3 if topology == "euclid" else 0.5) * a.d2h_bytes_per_step()
    assert torch.equal(a.batch.state, b.batch.state) and torch.equal(a.batch.visits, b.batch.visits)


def test_statistics_count_steps():
    import maze_b200 as mb
    env = mb.MazeVectorEnv(1000, shape=(15, 15), algorithms="r-prim", num_mazes=10, seed=1)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    resets = 0
    for t in range(50):
        resets += int((env.batch.terminated | env.batch.truncated).sum())     # these envs autoreset on the next step
        env.step(torch.randint(0, 4, (1000,), device="cuda", generator=g).to(torch.uint8))
    s = env.episode_statistics()
    assert s["steps"] == 50 * 1000 - resets
