"""maze_difficulty through the C ABI: McClendon difficulty / complexity and Kim-Crawfis L / DE / D
against values the unmodified reference produced (tests/golden/metrics.npz, incl. the literal
15x15 known answers) and against the oracle on device-generated mazes of every generator."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle.generation import ALGORITHMS  # noqa: E402
from oracle.metrics import kim_crawfis, mcclendon  # noqa: E402

REL = 1e-9   # float sums run in a different order than networkx iteration


def _pool_from_rows(z, rows):
    import maze_b200 as mb
    grids = [z[f"m{m['id']}_grid"] for m in rows]
    return mb, mb.MazePool.from_grids(grids, [m["start"] for m in rows], [m["goal"] for m in rows],
                                      [bool(m["no_border"]) for m in rows])


def test_literal_maze_known_answers():
    z, meta = load_golden("metrics")
    mb, pool = _pool_from_rows(z, meta[:1])
    out = pool.difficulty().cpu().numpy()[0]
    # reference outputs on the literal maze of testing_Mccledon.py:4-20
    assert out[0] == pytest.approx(9.950639302928026, rel=1e-12)
    assert out[1] == pytest.approx(5.681612603202764, rel=1e-12)
    assert out[2] == 0.6288659793814433 and out[3] == 0.03278688524590164 and out[4] == 0.03278688524590164
    assert out[5] == 61


def test_golden_mazes_match_reference():
    z, meta = load_golden("metrics")
    mb, pool = _pool_from_rows(z, meta)
    out = pool.difficulty().cpu().numpy()
    for k, m in enumerate(meta):
        tag = (m["algo"], m["shape"], m["id"])
        assert out[k, 0] == pytest.approx(m["difficulty"], rel=REL), tag
        if m["no_border"]:
            continue   # the reference records only the difficulty for border-less mazes
        assert out[k, 1] == pytest.approx(m["complexity"], rel=REL), tag
        assert out[k, 2] == m["L"] and out[k, 3] == m["DE"] and out[k, 4] == m["D"], tag
        assert out[k, 5] == m["sol_len"], tag


def test_ids_select_slots():
    z, meta = load_golden("metrics")
    mb, pool = _pool_from_rows(z, meta)
    full = pool.difficulty().cpu().numpy()
    ids = [7, 3, 3, 20]
    part = pool.difficulty(ids).cpu().numpy()
    np.testing.assert_array_equal(part, full[ids])


@pytest.mark.parametrize("algo", ALGORITHMS)
@pytest.mark.parametrize("shape,toroidal", [(21, False), (41, False), (81, False), (129, False), (41, True), (79, True)])
def test_generated_mazes_match_oracle(algo, shape, toroidal):
    import maze_b200 as mb
    n = 24 if shape <= 41 else 8
    pool = mb.MazePool(n, (shape, shape))
    pool.generate(algorithms=algo, toroidal=toroidal, seed=1000 + shape)
    out = pool.difficulty().cpu().numpy()
    meta = pool.meta_host()
    for m in range(n):
        grid = pool.grid_host(m)
        start = (int(meta[m, 2]) & 0xffff, int(meta[m, 2]) >> 16)
        goal = (int(meta[m, 3]) & 0xffff, int(meta[m, 3]) >> 16)
        if toroidal:
            grid = np.pad(grid, 1)
            start, goal = (start[0] + 1, start[1] + 1), (goal[0] + 1, goal[1] + 1)
        d, c = mcclendon(grid, start, goal)
        k = kim_crawfis(grid, start, goal)
        assert out[m, 0] == pytest.approx(d, rel=REL), (algo, shape, m)
        assert out[m, 1] == pytest.approx(c, rel=REL), (algo, shape, m)
        assert out[m, 2] == k["L"] and out[m, 3] == k["DE"] and out[m, 4] == k["D"], (algo, shape, m)
        assert out[m, 5] == k["sol_len"] and out[m, 6] == k["dead_end_count"]


def test_unreachable_goal_gives_nan():
    import maze_b200 as mb
    g = np.zeros((7, 7), dtype=np.uint8)
    g[1, 1:4] = 1
    g[5, 5] = 2
    pool = mb.MazePool.from_grids([g], [(1, 1)], [(5, 5)], False)
    out = pool.difficulty().cpu().numpy()[0]
    assert np.isnan(out[:5]).all()


@pytest.mark.parametrize("algo,toroidal", [("r-prim", False), ("dfs", False), ("prim&kill", True)])
def test_best_of_k_selection(algo, toroidal):
    """BaseMazeEnv.generate_maze (base_maze_env.py:78-97): keep the least difficult of 1 + 5 draws,
    strict <.  Candidate c of a slot is the same maze for every k > c, so the kept difficulty is a
    running minimum over k and the kept maze only changes when the minimum does."""
    import maze_b200 as mb
    n, shape = 48, (21, 21)
    prev_d, prev_grids = None, None
    for k in range(1, 7):
        pool = mb.MazePool(n, shape)
        dout = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        pool.generate(algorithms=algo, toroidal=toroidal, seed=77, candidates=k, difficulty_out=dout)
        d = dout.cpu().numpy()
        rescored = pool.difficulty().cpu().numpy()[:, 0]
        # the reported difficulty is the kept maze's (shared-memory float atomics: last-ulp run-to-run noise)
        np.testing.assert_allclose(d, rescored, rtol=1e-12)
        grids = [pool.grid_host(m).copy() for m in range(n)]
        if prev_d is not None:
            assert (d <= prev_d * (1 + 1e-12)).all()
            for m in range(n):
                if abs(d[m] - prev_d[m]) <= 1e-12 * prev_d[m]:
                    np.testing.assert_array_equal(grids[m], prev_grids[m])
        prev_d, prev_grids = d, grids
    # six draws must have found easier mazes for most slots
    pool1 = mb.MazePool(n, shape)
    pool1.generate(algorithms=algo, toroidal=toroidal, seed=77)
    d1 = pool1.difficulty().cpu().numpy()[:, 0]
    assert (prev_d <= d1 * (1 + 1e-12)).all() and (prev_d < d1 * (1 - 1e-9)).mean() > 0.5
