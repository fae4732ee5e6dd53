"""Packed maze-set formats, numpy restatement (test infrastructure; see oracle/__init__.py).

Independent reader / writer of the `.mzs` files and of the two packed record formats defined in
include/maze_b200.h (MAZE_PACK_BITMAP, MAZE_PACK_WALLS); the reference has no such format (its mazes
are lists of lists, lib/maze_generation.py:17), so parity here is device <-> oracle, both ways, on
reference-generated grids (tests/golden/metrics.npz), plus the reference's own channel encode
(lib/maze_generation.py:236-242) for `collection_tensor`.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC, VERSION = b"MAZEB200", 1
BITMAP, WALLS = 0, 1


def pack_bitmap(grid) -> np.ndarray:
    return np.packbits((np.asarray(grid).reshape(-1) != 0).astype(np.uint8), bitorder="little")


def unpack_bitmap(rec, H, W, goal) -> np.ndarray:
    g = np.unpackbits(np.asarray(rec, dtype=np.uint8), bitorder="little")[:H * W].reshape(H, W).astype(np.uint8)
    if g[goal[0], goal[1]]:
        g[goal[0], goal[1]] = 2
    return g


def pack_walls(grid) -> np.ndarray:
    g = np.asarray(grid)
    H, W = g.shape
    nr, nc = (H - 1) // 2, (W - 1) // 2
    nib = np.zeros(nr * nc + (nr * nc & 1), dtype=np.uint8)
    for i in range(nr):
        for j in range(nc):
            r, c = 2 * i + 1, 2 * j + 1
            nib[i * nc + j] = (g[r - 1, c] == 0) * 1 + (g[r, c + 1] == 0) * 2 + (g[r + 1, c] == 0) * 4 + (g[r, c - 1] == 0) * 8
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


def unpack_walls(rec, H, W, goal) -> np.ndarray:
    rec = np.asarray(rec, dtype=np.uint8)
    nr, nc = (H - 1) // 2, (W - 1) // 2
    g = np.zeros((H, W), dtype=np.uint8)
    for i in range(nr):
        for j in range(nc):
            ci = i * nc + j
            nib = (rec[ci >> 1] >> (4 * (ci & 1))) & 0xf
            r, c = 2 * i + 1, 2 * j + 1
            g[r, c] = 1
            if not nib & 2 and c + 1 < W - 1:
                g[r, c + 1] = 1
            if not nib & 4 and r + 1 < H - 1:
                g[r + 1, c] = 1
    if g[goal[0], goal[1]]:
        g[goal[0], goal[1]] = 2
    return g


def stride_of(max_shape, fmt) -> int:
    H, W = max_shape
    return (H * W + 7) // 8 if fmt == BITMAP else (((H - 1) // 2) * ((W - 1) // 2) + 1) // 2


def write_file(path, grids, metas, fmt=BITMAP, max_shape=None):
    """grids: list of block grids; metas: int32 [n, 8] records (H, W, start, goal, max_steps, flags, sol_len, spare)."""
    metas = np.asarray(metas, dtype="<i4")
    if max_shape is None:
        max_shape = (max(np.asarray(g).shape[0] for g in grids), max(np.asarray(g).shape[1] for g in grids))
    stride = stride_of(max_shape, fmt)
    body = np.zeros((len(grids), stride), dtype=np.uint8)
    for k, g in enumerate(grids):
        rec = pack_bitmap(g) if fmt == BITMAP else pack_walls(g)
        body[k, :len(rec)] = rec
    with open(path, "wb") as f:
        f.write(struct.pack("<8s6I", MAGIC, VERSION, fmt, len(grids), max_shape[0], max_shape[1], stride))
        f.write(metas.tobytes())
        f.write(body.tobytes())


def read_file(path):
    """-> (grids, metas int32 [n, 8], fmt)."""
    raw = open(path, "rb").read()
    magic, version, fmt, count, max_h, max_w, stride = struct.unpack_from("<8s6I", raw, 0)
    assert magic == MAGIC and version == VERSION, (magic, version)
    metas = np.frombuffer(raw, dtype="<i4", count=count * 8, offset=32).reshape(count, 8)
    body = np.frombuffer(raw, dtype=np.uint8, count=count * stride, offset=32 + metas.nbytes).reshape(count, stride)
    grids = []
    for k in range(count):
        H, W, goal = int(metas[k, 0]), int(metas[k, 1]), (int(metas[k, 3]) & 0xffff, int(metas[k, 3]) >> 16)
        grids.append(unpack_bitmap(body[k], H, W, goal) if fmt == BITMAP else unpack_walls(body[k], H, W, goal))
    return grids, metas.astype(np.int32), fmt


def collection_tensor(grid, start) -> np.ndarray:
    """int32 [3, H, W] exactly as generate_collection_of_mazes builds it (lib/maze_generation.py:236-242)."""
    g = np.asarray(grid)
    tile = (g == 1).astype(np.int32)
    wall = (g == 0).astype(np.int32)
    non_visited = (g != 0).astype(np.int32)
    non_visited[start[0], start[1]] = 0
    return np.stack([wall, tile, non_visited])
