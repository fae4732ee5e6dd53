"""Device generators (maze_generate) through the C ABI: every maze is a valid spanning tree, the
goal is the reference's farthest-leaf choice, the step table equals the oracle's, and the output
distribution matches samples from the reference generators."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402
from oracle.env_port import MazeTables  # noqa: E402
from oracle.generation import ALGORITHMS, check_perfect_maze, maze_shape_stats, select_goal  # noqa: E402
from test_oracle_generation import mean_close  # noqa: E402


def _gen(num, shape, algo, toroidal=False, seed=0, base=0, max_shape=None):
    import maze_b200 as mb
    pool = mb.MazePool(num, max_shape or (shape if isinstance(shape, tuple) else (shape, shape)))
    pool.generate(shapes=shape if isinstance(shape, tuple) else (shape, shape), algorithms=algo, toroidal=toroidal,
                  seed=seed, slot_id_base=base)
    torch.cuda.synchronize()
    return mb, pool


def _unpack(v):
    return int(v) & 0xffff, int(v) >> 16


@pytest.mark.parametrize("algo", ALGORITHMS)
@pytest.mark.parametrize("shape", [5, 11, 21, 41, 81, 129])
def test_generated_mazes_are_valid_and_consistent(algo, shape):
    n = 64 if shape <= 41 else 12
    mb, pool = _gen(n, shape, algo, seed=shape)
    meta = pool.meta_host()
    seen = set()
    for m in range(n):
        grid = pool.grid_host(m)
        ok, why = check_perfect_maze(grid)
        assert ok, (algo, shape, m, why)
        start, goal = _unpack(meta[m, mb.cabi.META_START]), _unpack(meta[m, mb.cabi.META_GOAL])
        assert grid[goal] == 2 and start[0] % 2 == 1 and start[1] % 2 == 1
        g1 = grid.copy(); g1[goal] = 1
        assert select_goal(g1, start) == goal, (algo, shape, m)
        t = MazeTables(grid, start, goal, False)
        np.testing.assert_array_equal(pool.table_host(m), t.table)
        assert meta[m, mb.cabi.META_MAX_STEPS] == t.max_steps
        assert meta[m, mb.cabi.META_SOL_LEN] == int(t.dgoal[start]) + 1
        seen.add(grid.tobytes())
    assert len(seen) >= (n if shape > 5 else 4)


@pytest.mark.parametrize("algo", ALGORITHMS)
def test_generated_toroidal_mazes(algo):
    S = 21
    mb, pool = _gen(32, S, algo, toroidal=True, seed=9)
    meta = pool.meta_host()
    for m in range(32):
        grid = pool.grid_host(m)
        assert grid.shape == (S, S)
        bordered = np.pad(grid, 1)
        ok, why = check_perfect_maze(bordered)
        assert ok, why
        start, goal = _unpack(meta[m, mb.cabi.META_START]), _unpack(meta[m, mb.cabi.META_GOAL])
        b1 = bordered.copy(); b1[b1 == 2] = 1
        assert select_goal(b1, (start[0] + 1, start[1] + 1)) == (goal[0] + 1, goal[1] + 1)
        t = MazeTables(grid, start, goal, True)
        np.testing.assert_array_equal(pool.table_host(m), t.table)
        assert meta[m, mb.cabi.META_MAX_STEPS] == t.max_steps


def test_mixed_shapes_and_algorithms_in_one_launch():
    import maze_b200 as mb
    shapes = [(s, s) for s in (21, 25, 29, 41, 61, 129)] * 4
    algos = [ALGORITHMS[i % 3] for i in range(len(shapes))]
    pool = mb.MazePool(len(shapes), (129, 129))
    pool.generate(shapes=shapes, algorithms=algos, seed=3)
    meta = pool.meta_host()
    for m, (shp, a) in enumerate(zip(shapes, algos)):
        grid = pool.grid_host(m)
        assert grid.shape == shp
        ok, why = check_perfect_maze(grid)
        assert ok, why
        assert (meta[m, mb.cabi.META_FLAGS] >> 8) == mb.ALGO_IDS[a]


def test_generation_is_deterministic_and_shard_invariant():
    mb, a = _gen(64, 21, "r-prim", seed=5)
    _, b = _gen(64, 21, "r-prim", seed=5)
    assert torch.equal(a.grids, b.grids) and torch.equal(a.meta, b.meta)
    _, c = _gen(64, 21, "r-prim", seed=6)
    assert not torch.equal(a.grids, c.grids)
    # slots 32..63 generated as a second "rank" with slot_id_base = 32 reproduce the same mazes
    _, d = _gen(32, 21, "r-prim", seed=5, base=32)
    assert torch.equal(a.grids[32:], d.grids)
    # regenerating a slot gives a new maze (generation count is part of the RNG key)
    before = a.grids[:4].clone()
    a.generate(ids=[0, 1, 2, 3], configure=False, seed=5)
    assert not torch.equal(before, a.grids[:4])
    ok, why = check_perfect_maze(a.grid_host(0))
    assert ok, why


@pytest.mark.parametrize("algo", ALGORITHMS)
@pytest.mark.parametrize("shape", [21, 41, 81])
def test_generator_distribution_matches_reference(algo, shape):
    """Means of (solution length, dead ends, junctions, start row/col) over 1500 device mazes vs the
    reference's own samples; tolerance 4.5 standard errors of the difference of means."""
    ref = np.load(f"{GOLDEN}/genstats.npz")[f"{algo}_{shape}"]
    n = 1500 if shape < 81 else 400
    mb, pool = _gen(n, shape, algo, seed=100 + shape)
    meta = pool.meta_host()
    grids = pool.grids.cpu().numpy()[:, :shape * shape].reshape(n, shape, shape)
    rows = []
    for m in range(n):
        start, goal = _unpack(meta[m, mb.cabi.META_START]), _unpack(meta[m, mb.cabi.META_GOAL])
        st = maze_shape_stats(grids[m], start, goal)
        assert st["sol_len"] == meta[m, mb.cabi.META_SOL_LEN]
        rows.append((st["sol_len"], st["dead_ends"], st["junctions"], start[0], start[1]))
    rows = np.array(rows)
    for col, name in enumerate(("sol_len", "dead_ends", "junctions", "start_r", "start_c")):
        ok, info = mean_close(rows[:, col], ref[:, col])
        assert ok, (algo, shape, name, info)


@pytest.mark.parametrize("algo", ALGORITHMS)
def test_largest_toroidal_shape(algo):
    """127 x 127 border-less mazes are generated at 129 x 129 = 64 x 64 cells, the lattice limit."""
    mb, pool = _gen(6, 127, algo, toroidal=True, seed=3)
    meta = pool.meta_host()
    for m in range(6):
        grid = pool.grid_host(m)
        assert grid.shape == (127, 127)
        ok, why = check_perfect_maze(np.pad(grid, 1))
        assert ok, (algo, m, why)
        t = MazeTables(grid, _unpack(meta[m, mb.cabi.META_START]), _unpack(meta[m, mb.cabi.META_GOAL]), True)
        np.testing.assert_array_equal(pool.table_host(m), t.table)
    with pytest.raises(ValueError):
        mb.MazePool(2, (129, 129)).generate(toroidal=True)


def test_generate_argument_errors():
    import maze_b200 as mb
    pool = mb.MazePool(4, (21, 21))
    with pytest.raises(ValueError):
        pool.generate(algorithms="kruskal")
    with pytest.raises(ValueError):
        pool.generate(shapes=(20, 20))
    with pytest.raises(ValueError):
        pool.generate(shapes=(41, 41))


@pytest.mark.parametrize("algo", ALGORITHMS)
def test_generation_is_invariant_to_sharding(algo):
    """RNG streams are keyed by the GLOBAL slot id: two ranks owning halves of the slot range
    generate exactly the mazes one rank owning all of it would."""
    from maze_b200.dist import shard_range
    import maze_b200 as mb
    M, shape = 40, (21, 21)
    whole = mb.MazePool(M, shape)
    whole.generate(algorithms=algo, seed=31, slot_id_base=0, candidates=2)
    for world in (2, 3):
        for r in range(world):
            sh = shard_range(M, r, world)
            part = mb.MazePool(sh.count, shape)
            part.generate(algorithms=algo, seed=31, slot_id_base=sh.start, candidates=2)
            assert torch.equal(part.grids, whole.grids[sh.start:sh.stop])
            assert torch.equal(part.table, whole.table[sh.start:sh.stop])
            assert torch.equal(part.meta, whole.meta[sh.start:sh.stop])
