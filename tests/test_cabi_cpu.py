"""CPU-side checks of the boundary: the library loads, exports every symbol the header declares,
and its host-side reward tables equal Python's math.exp expressions bit for bit."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "maze_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(maze_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from maze_b200 import cabi
    lib = cabi.lib()
    names = _declared_symbols()
    assert "maze_step" in names and "maze_fields" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/maze_b200.h but not exported"
        assert n in cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.maze_abi_version() == cabi.ABI_VERSION


def test_reward_luts_equal_reference_expressions():
    from maze_b200 import cabi
    rev, inv, shp = cabi.reward_lut(0), cabi.reward_lut(1), cabi.reward_lut(2)
    for i in range(256):
        assert rev[i] == 0.0 - (1 - math.exp(-0.2 * i))      # base_maze_env.py:194
        assert inv[i] == 0.0 - (1 - math.exp(-0.15 * i))     # base_maze_env.py:200
    assert list(shp[:3]) == [-1 * 0.5 - 0.05, 0 * 0.5 - 0.05, 1 * 0.5 - 0.05]
    assert rev[188] == -1.0 and inv[250] == -1.0              # saturating uint8 counters are exact


def test_struct_layout_matches_header():
    import ctypes as C
    from maze_b200 import cabi
    assert C.sizeof(cabi.MazeEnvBatch) == 16 + 16 * 8 + 16 + 8 + 8 + 8 + 16 + 8
    assert cabi.MazeEnvBatch.meta.offset == 16
    # the library reports the sizes it was compiled with (host-only call, no GPU needed)
    lib = cabi.lib()
    for which, struct in enumerate((cabi.MazeEnvBatch, cabi.MazeQAgent, cabi.MazeReplay, cabi.MazeStepTrace, cabi.MazeDqnNet)):
        assert lib.maze_sizeof(which) == C.sizeof(struct), struct.__name__
    assert lib.maze_sizeof(99) == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    from maze_b200 import MazePool, cabi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((cabi.MazeError, RuntimeError)):
        MazePool(1, (21, 21), device="cuda")
    with pytest.raises(cabi.MazeError):
        MazePool(1, (21, 21), device="cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "maze-solving-agent-gymnasium_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src


def test_packed_record_decoder_runs_on_the_host():
    """maze_step_decode_host is plain C (no GPU): records built here by the documented bit layout decode to the wide arrays
    by the reference's rules -- best dir = agent - (wrapped) neighbour, reward from the LUT the record names."""
    from maze_b200 import cabi
    rng = np.random.default_rng(0)
    n, S = 5000, 21
    r, c = rng.integers(0, S, n), rng.integers(0, S, n)
    code = rng.integers(0, 5, n)
    term, trunc = rng.integers(0, 2, n), rng.integers(0, 2, n)
    kind = rng.integers(0, 4, n)
    index = np.where(kind == 2, rng.integers(0, 3, n), np.where(kind == 3, rng.integers(0, 3, n), rng.integers(0, 256, n)))
    rec = (r | (c << 8) | (code << cabi.REC_CODE_SHIFT) | (term << cabi.REC_TERM_SHIFT) | (trunc << cabi.REC_TRUNC_SHIFT)
           | (kind << cabi.REC_KIND_SHIFT) | (index << cabi.REC_INDEX_SHIFT)).astype(np.uint32)
    tor = rng.integers(0, 2, n).astype(np.uint8)
    shape = np.full((n, 2), S, dtype=np.int32)
    d = cabi.decode_records(rec, shape, tor)
    np.testing.assert_array_equal(d["agent"], np.stack([r, c], 1))
    np.testing.assert_array_equal(d["terminated"], term.astype(bool))
    np.testing.assert_array_equal(d["truncated"], trunc.astype(bool))
    table = cabi.reward_table()
    np.testing.assert_array_equal(d["reward"].view(np.uint64), table[kind, index].view(np.uint64))
    assert table[3, 0] == 0.0 and table[3, 1] == 1.0 and table[3, 2] == -1.0 and table[2, 2] == 1 * 0.5 - 0.05
    dr, dc = np.array([1, -1, 0, 0, 0]), np.array([0, 0, 1, -1, 0])
    nr, nc = r + dr[code], c + dc[code]
    wrap = tor.astype(bool)
    nr, nc = np.where(wrap, nr % S, nr), np.where(wrap, nc % S, nc)
    np.testing.assert_array_equal(d["best_dir"], np.stack([r - nr, c - nc], 1))
    # bordered mazes need neither shapes nor topology
    d2 = cabi.decode_records(rec)
    np.testing.assert_array_equal(d2["best_dir"], np.stack([-dr[code], -dc[code]], 1))
