"""Packed maze-set formats: the numpy oracle (oracle/mazeset.py) round-trips every reference-generated
maze of tests/golden/metrics.npz through both record formats and through a .mzs file."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import mazeset as oms


def _mazes():
    z, meta = load_golden("metrics")
    return [(z[f"m{m['id']}_grid"], m) for m in meta]


def test_bitmap_round_trip_is_lossless_for_both_topologies():
    for grid, m in _mazes():
        H, W = grid.shape
        rec = oms.pack_bitmap(grid)
        assert len(rec) == (H * W + 7) // 8
        back = oms.unpack_bitmap(rec, H, W, m["goal"])
        want = (grid != 0).astype(np.uint8)
        want[m["goal"][0], m["goal"][1]] = 2
        assert np.array_equal(back, want)


def test_wall_nibbles_round_trip_bordered_mazes():
    n = 0
    for grid, m in _mazes():
        if m["no_border"]:
            continue
        H, W = grid.shape
        rec = oms.pack_walls(grid)
        assert len(rec) == (((H - 1) // 2) * ((W - 1) // 2) + 1) // 2
        assert np.array_equal(oms.unpack_walls(rec, H, W, m["goal"]), grid)
        n += 1
    assert n > 40
    # 40 x 40 logical cells: 800 B of wall nibbles, 821 B of bitmap (representation note)
    assert oms.stride_of((81, 81), oms.WALLS) == 800 and oms.stride_of((81, 81), oms.BITMAP) == 821


@pytest.mark.parametrize("fmt", [oms.BITMAP, oms.WALLS])
def test_file_round_trip(tmp_path, fmt):
    rows = [(g, m) for g, m in _mazes() if not m["no_border"]][:20]
    metas = np.zeros((len(rows), 8), dtype=np.int32)
    for k, (g, m) in enumerate(rows):
        metas[k, :4] = g.shape[0], g.shape[1], m["start"][0] | (m["start"][1] << 16), m["goal"][0] | (m["goal"][1] << 16)
    path = str(tmp_path / "set.mzs")
    oms.write_file(path, [g for g, _ in rows], metas, fmt)
    grids, metas2, fmt2 = oms.read_file(path)
    assert fmt2 == fmt and np.array_equal(metas2, metas)
    for (g, _), back in zip(rows, grids):
        assert np.array_equal(back, g)
