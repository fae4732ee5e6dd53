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


def test_product_loader_rejects_bad_files_before_touching_the_gpu(tmp_path):
    """maze_b200.mazeset.load validates magic / version / format / length on the host (ValueError), and
    packed_stride knows both record sizes; none of this needs a device."""
    import struct
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maze-solving-agent-gymnasium_b200"))
    from maze_b200 import mazeset
    assert mazeset.packed_stride((81, 81), "bitmap") == 821 and mazeset.packed_stride((81, 81), "walls") == 800
    with pytest.raises(ValueError):
        mazeset.packed_stride((81, 81), "rle")
    good_header = struct.pack("<8s6I", b"MAZEB200", 1, 0, 1, 11, 11, 16)
    cases = {
        "magic": struct.pack("<8s6I", b"NOTAMAZE", 1, 0, 1, 11, 11, 16) + bytes(32 + 16),
        "version": struct.pack("<8s6I", b"MAZEB200", 9, 0, 1, 11, 11, 16) + bytes(32 + 16),
        "format": struct.pack("<8s6I", b"MAZEB200", 1, 5, 1, 11, 11, 16) + bytes(32 + 16),
        "length": good_header + bytes(10),
    }
    for name, blob in cases.items():
        path = tmp_path / f"{name}.mzs"
        path.write_bytes(blob)
        with pytest.raises(ValueError):
            mazeset.load(str(path))
