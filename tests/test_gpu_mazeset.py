"""maze_pack / maze_unpack / maze_collection_encode through the C ABI against oracle/mazeset.py:
identical packed bytes, lossless round trips, .mzs files exchanged both ways, and the
generate_collection_of_mazes channel encode (lib/maze_generation.py:236-242)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle import mazeset as oms  # noqa: E402


def _golden_pool(bordered_only):
    import maze_b200 as mb
    z, meta = load_golden("metrics")
    rows = [m for m in meta if not (bordered_only and m["no_border"])]
    grids = [z[f"m{m['id']}_grid"] for m in rows]
    pool = mb.MazePool.from_grids(grids, [m["start"] for m in rows], [m["goal"] for m in rows], [bool(m["no_border"]) for m in rows])
    return mb, pool, grids, rows


@pytest.mark.parametrize("fmt", ["bitmap", "walls"])
def test_packed_records_equal_the_oracle_and_round_trip(fmt):
    mb, pool, grids, rows = _golden_pool(bordered_only=(fmt == "walls"))
    packed = mb.mazeset.pack(pool, fmt=fmt)
    host = packed.cpu().numpy()
    for k, g in enumerate(grids):
        rec = oms.pack_bitmap(g) if fmt == "bitmap" else oms.pack_walls(g)
        assert np.array_equal(host[k, :len(rec)], rec), rows[k]["id"]
        assert not host[k, len(rec):].any()
    other = mb.MazePool(pool.num_mazes, pool.max_shape)
    other.grids.fill_(7)                       # unpack must overwrite whole slots
    mb.mazeset.unpack(other, packed, pool.meta, fmt=fmt)
    assert torch.equal(other.grids, pool.grids)
    assert torch.equal(other.table, pool.table)
    assert torch.equal(other.meta, pool.meta)   # max_steps / sol_len recomputed by maze_fields


def test_walls_format_refuses_toroidal_mazes():
    mb, pool, _, _ = _golden_pool(bordered_only=False)
    with pytest.raises(ValueError):
        mb.mazeset.pack(pool, fmt="walls")


@pytest.mark.parametrize("fmt", ["bitmap", "walls"])
def test_files_are_exchanged_with_the_oracle_both_ways(tmp_path, fmt):
    mb, pool, grids, rows = _golden_pool(bordered_only=(fmt == "walls"))
    path = str(tmp_path / "dev.mzs")
    nbytes = mb.mazeset.save(pool, path, fmt=fmt)
    import os
    assert os.path.getsize(path) == nbytes
    back, metas, fmt_id = oms.read_file(path)
    assert fmt_id == mb.mazeset.FORMATS[fmt]
    assert np.array_equal(metas, pool.meta_host())
    for g, b in zip(grids, back):
        want = (np.asarray(g) != 0).astype(np.uint8)
        assert np.array_equal(b != 0, want != 0) and (b == 2).sum() == 1
    # oracle-written file -> device
    path2 = str(tmp_path / "ora.mzs")
    oms.write_file(path2, grids, pool.meta_host(), mb.mazeset.FORMATS[fmt], max_shape=pool.max_shape)
    assert open(path, "rb").read() == open(path2, "rb").read()
    loaded = mb.mazeset.load(path2)
    assert torch.equal(loaded.grids, pool.grids) and torch.equal(loaded.table, pool.table) and torch.equal(loaded.meta, pool.meta)


def test_generated_pool_survives_a_file_round_trip_and_steps_identically(tmp_path):
    import maze_b200 as mb
    pool = mb.MazePool(300, (81, 81))
    pool.generate(algorithms=["r-prim", "dfs", "prim&kill"] * 100, seed=5)
    path = str(tmp_path / "gen.mzs")
    n = mb.mazeset.save(pool, path, fmt="walls")
    assert n == 32 + 300 * (32 + 800)
    loaded = mb.mazeset.load(path)
    assert torch.equal(loaded.grids, pool.grids) and torch.equal(loaded.table, pool.table) and torch.equal(loaded.meta, pool.meta)
    ids = [5, 200, 17]
    sub = mb.mazeset.pack(pool, ids, fmt="bitmap")
    assert torch.equal(sub, mb.mazeset.pack(pool, fmt="bitmap")[ids])


def test_collection_encode_matches_the_reference_formula():
    import maze_b200 as mb
    pool = mb.MazePool(12, (21, 21))
    pool.generate(algorithms=["r-prim", "dfs", "prim&kill"] * 4, seed=9)
    got = mb.mazeset.collection_tensor(pool).cpu().numpy()
    meta = pool.meta_host()
    assert got.dtype == np.int32 and got.shape == (12, 3, 21, 21)
    for m in range(12):
        start = (int(meta[m, 2]) & 0xffff, int(meta[m, 2]) >> 16)
        assert np.array_equal(got[m], oms.collection_tensor(pool.grid_host(m), start))
    part = mb.mazeset.collection_tensor(pool, [3, 1]).cpu().numpy()
    assert np.array_equal(part, got[[3, 1]])
    from lib.maze_generation import generate_collection_of_mazes
    coll = generate_collection_of_mazes((11, 11), 5)
    assert len(coll) == 5 and all(t.shape == (3, 11, 11) and t.dtype == torch.int32 for t in coll)
    assert len({t.numpy().tobytes() for t in coll}) == 5
