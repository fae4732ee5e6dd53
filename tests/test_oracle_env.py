"""Oracle (CPU restatement) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from oracle.env_port import ClosedFormEnv, MazeTables, PortEnv
from oracle.grid import ACTIONS, astar_len, bfs_dist, best_dir_vector


def _replay(env_cls, z, m, one_tape_above=10**9):
    i = m["id"]
    toroidal = m["topology"] == "torus"
    env = env_cls(z[f"m{i}_grid"], m["start"], m["goal"], toroidal=toroidal, enrich=m["enrich"])
    assert env.max_steps == m["max_steps"]
    tapes = m["tapes"] if m["shape"] <= one_tape_above else m["tapes"][:1]   # A* per step: keep the large mazes short
    for j in tapes:
        pre = f"m{i}_t{j}_"
        obs, info = env.reset()
        T = len(z[pre + "action"])

        def check(t, obs, info):
            if m["enrich"]:
                assert obs["agent"].dtype == np.float64
                np.testing.assert_array_equal(obs["agent"].view(np.uint64), z[pre + "agent"][t].view(np.uint64))
                np.testing.assert_array_equal(obs["target"].view(np.uint64), z[pre + "target"][t].view(np.uint64))
                np.testing.assert_array_equal(obs["window"].astype(np.uint8), z[pre + "window"][t])
            else:
                np.testing.assert_array_equal(obs["agent"], z[pre + "agent"][t])
                np.testing.assert_array_equal(obs["target"], z[pre + "target"][t])
            np.testing.assert_array_equal(obs["best dir"], z[pre + "best"][t])
            assert info["distance"] == z[pre + "dist"][t]
            np.testing.assert_array_equal(env.mask_direction(probs=True).astype(np.float32), z[pre + "mask"][t])

        check(0, obs, info)
        for t in range(T):
            obs, r, trunc, term, info = env.step(int(z[pre + "action"][t]))
            assert np.float64(r).view(np.uint64) == z[pre + "reward"][t].view(np.uint64), (t, r, z[pre + "reward"][t])
            assert trunc == bool(z[pre + "trunc"][t]) and term == bool(z[pre + "term"][t]), t
            check(t + 1, obs, info)


_Z = {}


def _golden(name):
    from conftest import load_golden
    if name not in _Z:
        _Z[name] = load_golden(name)
    return _Z[name]


def _metas():
    # steps.npz: round 1; steps81.npz: round 2, the headline size for the generator x topology pairs round 1 lacked
    return [dict(m, file=f) for f in ("steps", "steps81") for m in _golden(f)[1]]


def _id(m):
    return f"{m['file']}-{m['topology']}-{m['algo']}-{m['shape']}{'-v1' if m['enrich'] else ''}"


@pytest.mark.parametrize("m", _metas(), ids=_id)
def test_closed_form_env_matches_reference(m):
    _replay(ClosedFormEnv, _golden(m["file"])[0], m)


@pytest.mark.parametrize("m", _metas(), ids=_id)
def test_port_env_matches_reference(m):
    # every shape, 81 x 81 included: this port is what bench.py times as the reference arm / cpu_baseline
    _replay(PortEnv, _golden(m["file"])[0], m, one_tape_above=21)


@pytest.mark.parametrize("name", ["bestdir", "bestdir81"])
def test_best_dir_table_matches_reference(name):
    """bestdir81.npz (round 2): every open block of 81 x 81 mazes, dfs included -- most of a dfs maze lies farther from
    the goal than the A* depth limit L = 162, the `D > L` branch of the closed form."""
    z, meta = _golden(name)
    if name == "bestdir81":
        from oracle.grid import bfs_dist as _bfs
        far = [int((_bfs(z[f"m{m['id']}_grid"], m["goal"], m["topology"] == "torus")[z[f"m{m['id']}_grid"] != 0] > 2 * m["shape"]).sum()) for m in meta]
        assert max(far) > 1000, far   # the far branch really is exercised
    for m in meta:
        i = m["id"]
        grid, nxt = z[f"m{i}_grid"], z[f"m{i}_next"]
        toroidal = m["topology"] == "torus"
        t = MazeTables(grid, m["start"], m["goal"], toroidal)
        assert t.max_steps == m["max_steps"]
        for r, c in zip(*np.nonzero(grid)):
            v = best_dir_vector(int(t.code[r, c]), (r, c), grid.shape, toroidal)
            assert (r - v[0], c - v[1]) == tuple(nxt[r, c]), (m, r, c)


def test_astar_port_lengths_equal_bfs(golden_bestdir):
    z, meta = golden_bestdir
    rng = np.random.default_rng(0)
    for m in meta[:12]:
        grid = z[f"m{m['id']}_grid"]
        toroidal = m["topology"] == "torus"
        d = bfs_dist(grid, m["goal"], toroidal)
        cells = np.argwhere(grid != 0)
        for r, c in cells[rng.choice(len(cells), size=12, replace=False)]:
            assert astar_len(grid, (r, c), m["goal"], toroidal=toroidal) == d[r, c] + 1
            L = 6
            assert astar_len(grid, (r, c), m["goal"], max_depth=L, toroidal=toroidal) == min(d[r, c], L) + 1


def test_penalty_luts_saturate():
    assert ClosedFormEnv.REVISIT_LUT[188] == -1.0 and ClosedFormEnv.REVISIT_LUT[255] == -1.0
    assert ClosedFormEnv.INVALID_LUT[250] == -1.0 and ClosedFormEnv.INVALID_LUT[255] == -1.0
    assert ClosedFormEnv.REVISIT_LUT[187] != -1.0 or ClosedFormEnv.REVISIT_LUT[186] != -1.0
