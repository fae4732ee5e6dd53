// Wall bit-planes of a bordered maze in the registers of one warp, and bit-parallel breadth-first
// search over the cell lattice (shared by the generators and the difficulty metrics).
//
// A maze of N x N logical cells (N <= 64) is two planes: E (bit j of row i = open passage between
// cells (i, j) and (i, j+1)) and S (between (i, j) and (i+1, j)), one 64-bit word per lattice row,
// rows l and l+32 in lane l.  Cell sets use the same layout.  All row / column arguments are
// warp-uniform and all 32 lanes must call.
#pragma once
#include "maze_fields.cuh"

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;

struct RowSets {
    u64 a0, a1;   // rows lane, lane + 32
};
struct Walls {
    RowSets e, s;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ u64 row_get(const RowSets& s, int i) {   // row i, broadcast to every lane
    return __shfl_sync(FULL, (i >> 5) ? s.a1 : s.a0, i & 31);
}
__device__ __forceinline__ void row_or(RowSets& s, int i, u64 bits) {
    if ((i & 31) == lane_id()) {
        if (i >> 5) s.a1 |= bits; else s.a0 |= bits;
    }
}
__device__ __forceinline__ int cell_bit(const RowSets& s, int i, int j) { return (int)((row_get(s, i) >> j) & 1ull); }

// ---- bit-parallel breadth-first search over the cell lattice ----------------------------------

struct Reach {   // cells adjacent to the frontier, by where their frontier neighbour sits
    u64 l0, l1, r0, r1, a0, a1, b0, b1;   // left / right / above / below, rows lane / lane + 32
};

__device__ __forceinline__ void bfs_expand(const Walls& w, u64 f0, u64 f1, Reach& x) {
    const int lane = lane_id();
    x.l0 = (f0 & w.e.a0) << 1;  x.l1 = (f1 & w.e.a1) << 1;
    x.r0 = (f0 >> 1) & w.e.a0;  x.r1 = (f1 >> 1) & w.e.a1;
    const u64 d0 = f0 & w.s.a0, d1 = f1 & w.s.a1;          // frontier cells whose south wall is open
    const u64 u0 = __shfl_up_sync(FULL, d0, 1), u1 = __shfl_up_sync(FULL, d1, 1), wrap_a = __shfl_sync(FULL, d0, 31);
    x.a0 = lane == 0 ? 0ull : u0;
    x.a1 = lane == 0 ? wrap_a : u1;
    const u64 n0 = __shfl_down_sync(FULL, f0, 1), n1 = __shfl_down_sync(FULL, f1, 1), wrap_b = __shfl_sync(FULL, f1, 0);
    x.b0 = (lane == 31 ? wrap_b : n0) & w.s.a0;
    x.b1 = (lane == 31 ? 0ull : n1) & w.s.a1;
}


// Wall planes of the bordered block grid `grid` (0 wall / != 0 open, pitch Wb) -- warp 0 of a CTA.
__device__ __forceinline__ void walls_from_grid(const uint8_t* grid, int Wb, int nr, int nc, Walls& w) {
    const int lane = lane_id();
    w.e.a0 = w.e.a1 = w.s.a0 = w.s.a1 = 0ull;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int i = lane + 32 * sl;
        if (i >= nr) continue;
        const uint8_t* row = grid + (2 * i + 1) * Wb;
        u64 e = 0ull, s = 0ull;
        for (int j = 0; j < nc; ++j) {
            if (j + 1 < nc && row[2 * j + 2] != 0) e |= 1ull << j;
            if (i + 1 < nr && row[Wb + 2 * j + 1] != 0) s |= 1ull << j;
        }
        if (sl) { w.e.a1 = e; w.s.a1 = s; } else { w.e.a0 = e; w.s.a0 = s; }
    }
}

// Block distances from cell (si, sj) written into dist[] (pitch Wb) for every cell reached and for
// the passage block it was reached through; everything else must hold DIST_INF already.  One
// bit-parallel level per cell distance: 2 * level for cells, 2 * level - 1 for the passage.
// (Tried: a corridor fast path that follows a singleton frontier in the shared-memory grid without shuffles. The
// frontier of a dfs maze is rarely exactly one cell -- side branches keep it at two to four -- so it lost 5 %.)
__device__ __forceinline__ void cell_bfs_distances(const Walls& w, int si, int sj, unsigned short* dist, int Wb) {
    const int lane = lane_id();
    u64 f0 = 0ull, f1 = 0ull;
    if ((si & 31) == lane) { if (si >> 5) f1 = 1ull << sj; else f0 = 1ull << sj; }
    u64 v0 = f0, v1 = f1;
    if (lane == 0) dist[(2 * si + 1) * Wb + 2 * sj + 1] = 0;
    // one pass over the newly reached cells of a row; a cell reached from several sides at once (mazes with
    // cycles) keeps the first source in the order left, right, above, below
    auto scatter = [&](u64 bits, u64 l, u64 r, u64 a, int row, int level) {
        while (bits) {
            const int j = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            const int off = ((l >> j) & 1ull) ? -1 : (((r >> j) & 1ull) ? 1 : (((a >> j) & 1ull) ? -Wb : Wb));
            const int b = (2 * row + 1) * Wb + 2 * j + 1;
            dist[b] = (unsigned short)(2 * level);
            dist[b + off] = (unsigned short)(2 * level - 1);
        }
    };
    for (int level = 1;; ++level) {
        Reach x;
        bfs_expand(w, f0, f1, x);
        const u64 n0 = (x.l0 | x.r0 | x.a0 | x.b0) & ~v0, n1 = (x.l1 | x.r1 | x.a1 | x.b1) & ~v1;
        if (!__ballot_sync(FULL, (n0 | n1) != 0ull)) break;
        scatter(n0, x.l0, x.r0, x.a0, lane, level);
        scatter(n1, x.l1, x.r1, x.a1, lane + 32, level);
        v0 |= n0; v1 |= n1; f0 = n0; f1 = n1;
    }
    __syncwarp();
}
