"""Packed maze sets: device pack / unpack (csrc/maze_pack.cu) and the `.mzs` file built on them.

The reference keeps mazes as Python lists of lists (lib/maze_generation.py:17) and has no interchange
format; SURVEY.md section 8(f) rank 3 asks for one so that maze sets (the 1000-maze parity sets, golden
fixtures, curriculum pools) can be stored and exchanged.  File layout, little-endian:

    0   8  magic  b"MAZEB200"
    8   4  u32    version (1)
    12  4  u32    format  (0 = bitmap: 1 bit per block, 1 = walls: 4 wall bits per logical cell)
    16  4  u32    count   number of mazes
    20  4  u32    max_h
    24  4  u32    max_w
    28  4  u32    stride  bytes per packed record
    32     i32 [count, 8] meta records (H, W, start, goal, max_steps, flags, sol_len, spare; start / goal = row | col << 16)
    ...    u8  [count, stride] packed records

A 40 x 40 (81 x 81 block) maze takes 821 B as a bitmap and 800 B as wall nibbles, against 6 561 B as a block grid.
`oracle/mazeset.py` reads and writes the same files with numpy only (tests exchange files both ways).
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import cabi
from .engine import MazePool

MAGIC = b"MAZEB200"
VERSION = 1
FORMATS = {"bitmap": cabi.PACK_BITMAP, "walls": cabi.PACK_WALLS}
_HEADER = struct.Struct("<8s6I")


def packed_stride(max_shape, fmt: str) -> int:
    H, W = int(max_shape[0]), int(max_shape[1])
    if fmt == "bitmap":
        return (H * W + 7) // 8
    if fmt == "walls":
        return (((H - 1) // 2) * ((W - 1) // 2) + 1) // 2
    raise ValueError(f"unknown packed format {fmt!r} (expected 'bitmap' or 'walls')")


def pack(pool: MazePool, ids=None, fmt: str = "bitmap") -> torch.Tensor:
    """uint8 [n, stride] device tensor: the packed records of the given slots (all if None)."""
    stride = packed_stride(pool.max_shape, fmt)
    if fmt == "walls" and bool((pool.meta[:, cabi.META_FLAGS] & cabi.FLAG_TOROIDAL).any().item()):
        raise ValueError("the wall-nibble format cannot express a toroidal maze (passage blocks link across the seam); use 'bitmap'")
    ids_t = None if ids is None else torch.as_tensor(list(ids), dtype=torch.int32, device=pool.device)
    n = pool.num_mazes if ids_t is None else ids_t.numel()
    out = torch.empty((n, stride), dtype=torch.uint8, device=pool.device)
    rc = cabi.lib().maze_pack(pool.ctx.handle, cabi.ptr(pool.grids), cabi.ptr(pool.meta), cabi.ptr(ids_t), n, pool.slot,
                              FORMATS[fmt], cabi.ptr(out), stride, cabi.current_stream(pool.device))
    pool.ctx.check(rc, "maze_pack")
    return out


def unpack(pool: MazePool, packed: torch.Tensor, meta: torch.Tensor, ids=None, fmt: str = "bitmap"):
    """Fill slots of `pool` from packed records: writes their meta records, expands the grids on the
    device and rebuilds step tables + step budgets (maze_fields)."""
    packed = packed.to(device=pool.device, dtype=torch.uint8).contiguous()
    meta = meta.to(device=pool.device, dtype=torch.int32).contiguous()
    n = packed.shape[0]
    if meta.shape != (n, cabi.META_WORDS):
        raise ValueError("meta must be int32 [n, 8]")
    ids_l = list(range(n)) if ids is None else list(ids)
    idx = torch.as_tensor(ids_l, dtype=torch.long, device=pool.device)
    hw = (meta[:, cabi.META_H] * meta[:, cabi.META_W]).max().item()
    if hw > pool.slot:
        raise ValueError("a packed maze does not fit the pool's slot")
    pool.meta[idx] = meta
    ids_t = torch.as_tensor(ids_l, dtype=torch.int32, device=pool.device)
    rc = cabi.lib().maze_unpack(pool.ctx.handle, cabi.ptr(packed), packed.shape[1], FORMATS[fmt], cabi.ptr(pool.grids),
                                cabi.ptr(pool.meta), cabi.ptr(ids_t), n, pool.slot, cabi.current_stream(pool.device))
    pool.ctx.check(rc, "maze_unpack")
    pool.compute_fields(ids_l)


def save(pool: MazePool, path: str, fmt: str = "bitmap", ids=None) -> int:
    """Write the pool (or the given slots) as a .mzs file; returns the number of bytes written."""
    packed = pack(pool, ids, fmt).cpu().numpy()
    meta = (pool.meta if ids is None else pool.meta[torch.as_tensor(list(ids), dtype=torch.long, device=pool.device)]).cpu().numpy()
    with open(path, "wb") as f:
        f.write(_HEADER.pack(MAGIC, VERSION, FORMATS[fmt], packed.shape[0], pool.max_shape[0], pool.max_shape[1], packed.shape[1]))
        f.write(np.ascontiguousarray(meta, dtype="<i4").tobytes())
        f.write(packed.tobytes())
    return _HEADER.size + meta.nbytes + packed.nbytes


def load(path: str, device="cuda") -> MazePool:
    """Read a .mzs file into a new MazePool (grids, meta, step tables on the device)."""
    with open(path, "rb") as f:
        raw = f.read()
    magic, version, fmt_id, count, max_h, max_w, stride = _HEADER.unpack_from(raw, 0)
    if magic != MAGIC or version != VERSION:
        raise ValueError(f"{path}: not a maze-set file (magic {magic!r}, version {version})")
    fmt = {v: k for k, v in FORMATS.items()}.get(fmt_id)
    if fmt is None:
        raise ValueError(f"{path}: unknown packed format {fmt_id}")
    need = _HEADER.size + count * (4 * cabi.META_WORDS + stride)
    if len(raw) != need:
        raise ValueError(f"{path}: {len(raw)} bytes, header says {need}")
    meta = np.frombuffer(raw, dtype="<i4", count=count * cabi.META_WORDS, offset=_HEADER.size).reshape(count, cabi.META_WORDS)
    packed = np.frombuffer(raw, dtype=np.uint8, count=count * stride, offset=_HEADER.size + meta.nbytes).reshape(count, stride)
    pool = MazePool(count, (max_h, max_w), device)
    unpack(pool, torch.from_numpy(packed.copy()), torch.from_numpy(meta.astype(np.int32)), fmt=fmt)
    return pool


def collection_tensor(pool: MazePool, ids=None) -> torch.Tensor:
    """int32 [n, 3, H, W] = [wall, tile (== 1), non_visited (start cleared)] of a constant-shape pool
    (generate_collection_of_mazes, lib/maze_generation.py:236-242)."""
    H, W = pool.max_shape
    ids_t = None if ids is None else torch.as_tensor(list(ids), dtype=torch.int32, device=pool.device)
    n = pool.num_mazes if ids_t is None else ids_t.numel()
    out = torch.empty((n, 3, H, W), dtype=torch.int32, device=pool.device)
    rc = cabi.lib().maze_collection_encode(pool.ctx.handle, cabi.ptr(pool.grids), cabi.ptr(pool.meta), cabi.ptr(ids_t), n,
                                           pool.slot, H, W, cabi.ptr(out), cabi.current_stream(pool.device))
    pool.ctx.check(rc, "maze_collection_encode")
    return out
