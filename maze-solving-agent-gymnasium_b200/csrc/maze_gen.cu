// Batched maze generation on the device (sm_100a): r-prim, dfs, prim&kill + goal selection +
// border stripping for toroidal mazes + step table, one CTA per maze.
//
// The sequential carving loop runs on warp 0 with the cell sets held as row bitmaps in REGISTERS
// (one 64-bit word per lattice row, rows l and l+32 in lane l), so "pick a uniformly random
// frontier / eligible cell" is popcount + warp scan + ballot + find-nth-set-bit, with no shared
// memory traffic for the sets; the block grid itself lives in shared memory.  The data-parallel
// phases (BFS from start, farthest-leaf goal, BFS from goal, table encode) use the whole CTA.
//
// Reference: lib/maze_generation.py:6-35 (gen_maze), :37-56 (gen_maze_no_border), :59-99 (r-prim),
// :101-128 (dfs), :130-185 (prim&kill), :187-218 (goal = farthest leaf, row-major tie-break).
// RNG: Philox4x32-10 keyed by (seed, global slot id, generation count) -- the reference draws
// from Python's global `random` through set iteration order, which cannot be replayed, so parity
// for generators is structural (spanning tree) + distributional (see tests).
#include "maze_metrics.cuh"

namespace {

constexpr int GEN_THREADS = FIELD_THREADS;
constexpr unsigned FULL = 0xffffffffu;

struct GenParams {
    uint8_t* grids;          // [M, slot] out (may be NULL)
    int32_t* meta;           // [M, 8] in: H, W, FLAGS ; out: START, GOAL, MAX_STEPS, SOL_LEN, SPARE(gen count)
    uint8_t* table;          // [M, slot] out
    const int32_t* ids;      // [n] slots to generate (NULL = 0..n-1)
    const int32_t* count_dev;// optional device-side n
    int n;
    int slot;
    int smem_hw;
    int smem_cells;          // 0 when no metrics are needed
    int candidates;          // best-of-k by McClendon difficulty (base_maze_env.py:78-97); 1 = raw generator
    double* difficulty;      // [n] optional out: difficulty of the maze kept for item k
    unsigned long long seed;
    long long slot_id_base;
};

// ---- row-bitmap helpers (warp-uniform i, j) ------------------------------------------------

struct RowSets {
    unsigned long long a0, a1;   // rows lane, lane + 32
};

__device__ __forceinline__ int select64(unsigned long long w, int k) {   // k-th (0-based) set bit
    const unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
    const int pl = __popc(lo);
    return k < pl ? (int)__fns(lo, 0, k + 1) : 32 + (int)__fns(hi, 0, k - pl + 1);
}

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// uniformly random set bit over the two rows of all lanes; returns (i << 8) | j, or -1 if empty
__device__ __forceinline__ int pick_uniform(unsigned long long s0, unsigned long long s1, Philox& rng) {
    const int lane = threadIdx.x & 31;
    const int c0 = __popcll(s0), c1 = __popcll(s1);
    const int incl = warp_incl_scan(c0 + c1);
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return -1;
    const int k = (int)rng.below((unsigned)total);
    const int owner = __ffs(__ballot_sync(FULL, incl > k)) - 1;
    int packed = 0;
    if (lane == owner) {
        const int kk = k - (incl - c0 - c1);
        packed = kk < c0 ? ((lane << 8) | select64(s0, kk)) : (((lane + 32) << 8) | select64(s1, kk - c0));
    }
    return __shfl_sync(FULL, packed, owner);
}

// 4-bit mask of lattice neighbours of (i, j) whose bit in `s` equals `want`:
// bit0 up (i-1), bit1 down (i+1), bit2 left (j-1), bit3 right (j+1); out-of-lattice never counts
__device__ __forceinline__ unsigned neighbour_mask(const RowSets& s, int i, int j, int nr, int nc, bool want) {
    const int lane = threadIdx.x & 31;
    unsigned my = 0;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int row = lane + 32 * sl;
        const unsigned long long w = sl ? s.a1 : s.a0;
        if (row == i - 1 && (((w >> j) & 1ull) != 0) == want) my |= 1u;
        if (row == i + 1 && row < nr && (((w >> j) & 1ull) != 0) == want) my |= 2u;
        if (row == i) {
            if (j > 0 && (((w >> (j - 1)) & 1ull) != 0) == want) my |= 4u;
            if (j + 1 < nc && (((w >> (j + 1)) & 1ull) != 0) == want) my |= 8u;
        }
    }
    return __reduce_or_sync(FULL, my);
}

__device__ __forceinline__ void set_bit(RowSets& s, int i, int j) {
    const int lane = threadIdx.x & 31;
    if ((i & 31) == lane) {
        if (i >> 5) s.a1 |= 1ull << j; else s.a0 |= 1ull << j;
    }
}

__device__ __forceinline__ void dir_delta(int d, int& di, int& dj) {   // 0 up 1 down 2 left 3 right
    di = (d == 1) - (d == 0);
    dj = (d == 3) - (d == 2);
}

// open cell (i, j) and the wall towards direction d in the block grid
__device__ __forceinline__ void carve(uint8_t* grid, int Wb, int i, int j, int d) {
    if ((threadIdx.x & 31) == 0) {
        int di, dj;
        dir_delta(d, di, dj);
        const int r = 2 * i + 1, c = 2 * j + 1;
        grid[r * Wb + c] = 1;
        grid[(r + di) * Wb + (c + dj)] = 1;
    }
}

// ---- generators (all 32 lanes of warp 0, uniform control flow) -----------------------------

// lib/maze_generation.py:59-99
__device__ void gen_random_prim(uint8_t* grid, int Wb, int nr, int nc, int si, int sj, Philox& rng) {
    const int lane = threadIdx.x & 31;
    const unsigned long long colmask = nc >= 64 ? ~0ull : ((1ull << nc) - 1ull);
    RowSets in = {0ull, 0ull}, fr = {0ull, 0ull};
    auto add_cell = [&](int i, int j) {
        const unsigned long long bit = 1ull << j;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int row = lane + 32 * sl;
            unsigned long long& w_in = sl ? in.a1 : in.a0;
            unsigned long long& w_fr = sl ? fr.a1 : fr.a0;
            if (row >= nr) continue;
            if (row == i) {
                w_in |= bit;
                w_fr &= ~bit;
                w_fr |= ((bit << 1) | (bit >> 1)) & ~w_in & colmask;
            } else if (row == i - 1 || row == i + 1) {
                w_fr |= bit & ~w_in;
            }
        }
    };
    if (lane == 0) grid[(2 * si + 1) * Wb + 2 * sj + 1] = 1;
    add_cell(si, sj);
    for (;;) {
        const int p = pick_uniform(fr.a0, fr.a1, rng);
        if (p < 0) break;
        const int i = p >> 8, j = p & 0xff;
        const unsigned nb = neighbour_mask(in, i, j, nr, nc, true);   // always non-empty for a frontier cell
        const int d = (int)__fns(nb, 0, (int)rng.below((unsigned)__popc(nb)) + 1);
        carve(grid, Wb, i, j, d);
        add_cell(i, j);
    }
}

// lib/maze_generation.py:101-128 (first unvisited neighbour of a fresh shuffle == uniform choice)
__device__ void gen_depth_first(uint8_t* grid, unsigned short* stack, int Wb, int nr, int nc, int si, int sj, Philox& rng) {
    const int lane = threadIdx.x & 31;
    RowSets vis = {0ull, 0ull};
    set_bit(vis, si, sj);
    if (lane == 0) grid[(2 * si + 1) * Wb + 2 * sj + 1] = 1;
    int sp = 0, i = si, j = sj;   // (i, j) is the stack top, kept in registers
    for (;;) {
        const unsigned nb = neighbour_mask(vis, i, j, nr, nc, false);
        if (nb) {
            const int d = (int)__fns(nb, 0, (int)rng.below((unsigned)__popc(nb)) + 1);
            int di, dj;
            dir_delta(d, di, dj);
            if (lane == 0) stack[sp] = (unsigned short)((i << 8) | j);
            ++sp;
            // carve from the new cell back towards the old one
            i += di; j += dj;
            carve(grid, Wb, i, j, d ^ 1);
            set_bit(vis, i, j);
        } else {
            if (sp == 0) break;
            --sp;
            __syncwarp();
            const int t = stack[sp];
            i = t >> 8; j = t & 0xff;
        }
    }
}

// lib/maze_generation.py:130-185
__device__ void gen_prim_and_kill(uint8_t* grid, int Wb, int nr, int nc, int si, int sj, Philox& rng) {
    const int lane = threadIdx.x & 31;
    const unsigned long long colmask = nc >= 64 ? ~0ull : ((1ull << nc) - 1ull);
    for (int t = lane; t < nr * nc; t += 32) grid[(2 * (t / nc) + 1) * Wb + 2 * (t % nc) + 1] = 1;   // :140-142
    // marked rows; bits / rows outside the lattice read as "marked" so they never look eligible
    RowSets mk;
    mk.a0 = lane < nr ? ~colmask : ~0ull;
    mk.a1 = lane + 32 < nr ? ~colmask : ~0ull;
    set_bit(mk, si, sj);
    int unmarked = nr * nc - 1;
    int i = si, j = sj;
    for (;;) {
        // random walk until no unmarked neighbour (:154-185)
        for (;;) {
            const unsigned nb = neighbour_mask(mk, i, j, nr, nc, false);
            if (!nb) break;
            const int d = (int)__fns(nb, 0, (int)rng.below((unsigned)__popc(nb)) + 1);
            int di, dj;
            dir_delta(d, di, dj);
            i += di; j += dj;
            carve(grid, Wb, i, j, d ^ 1);
            set_bit(mk, i, j);
            --unmarked;
        }
        if (unmarked == 0) break;
        // restart from a uniformly random marked cell with an unmarked neighbour (:150-152)
        unsigned long long up0 = __shfl_up_sync(FULL, mk.a0, 1);
        unsigned long long up1 = __shfl_up_sync(FULL, mk.a1, 1);
        const unsigned long long last0 = __shfl_sync(FULL, mk.a0, 31);
        if (lane == 0) { up0 = ~0ull; up1 = last0; }
        unsigned long long dn0 = __shfl_down_sync(FULL, mk.a0, 1);
        unsigned long long dn1 = __shfl_down_sync(FULL, mk.a1, 1);
        const unsigned long long first1 = __shfl_sync(FULL, mk.a1, 0);
        if (lane == 31) { dn0 = first1; dn1 = ~0ull; }
        const unsigned long long e0 = mk.a0 & colmask & (~up0 | ~dn0 | (~mk.a0 << 1) | (~mk.a0 >> 1));
        const unsigned long long e1 = mk.a1 & colmask & (~up1 | ~dn1 | (~mk.a1 << 1) | (~mk.a1 >> 1));
        const int p = pick_uniform(lane < nr ? e0 : 0ull, lane + 32 < nr ? e1 : 0ull, rng);
        if (p < 0) break;   // cannot happen on a connected lattice
        i = p >> 8; j = p & 0xff;
    }
}

template <bool kScored>   // kScored: candidates > 1 or a difficulty output (keeps the raw path lean)
__global__ void __launch_bounds__(GEN_THREADS)
maze_generate_kernel(GenParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_start, s_goal, s_keep_start, s_keep_goal, s_take;
    __shared__ unsigned s_best;
    __shared__ double s_keep_diff;
    __shared__ MazeMetrics s_metrics;
    const int tid = threadIdx.x;
    const int n = p.count_dev ? min(*p.count_dev, p.n) : p.n;
    FieldSmem f = field_smem_carve(smem, p.smem_hw);
    const size_t keep_bytes = ((size_t)p.smem_hw + 15) & ~(size_t)15;
    uint8_t* keep = smem + field_smem_bytes(p.smem_hw);
    MetricsSmem ms = metrics_smem_carve(keep + keep_bytes, p.smem_cells, f.queue);

    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        const int m = p.ids ? p.ids[item] : item;
        int32_t* mm = p.meta + (size_t)m * MAZE_META_WORDS;
        const int H = mm[MAZE_META_H], W = mm[MAZE_META_W];
        const int flags = mm[MAZE_META_FLAGS];
        const bool tor = (flags & MAZE_FLAG_TOROIDAL) != 0;
        const int algo = (flags >> 8) & 0xff;
        const int gen_count = mm[MAZE_META_SPARE];
        const int Hb = tor ? H + 2 : H, Wb = tor ? W + 2 : W;   // :48 gen_maze(shape + 2)
        const int nr = (Hb - 1) / 2, nc = (Wb - 1) / 2;

        for (int cand = 0; cand < p.candidates; ++cand) {
            __syncthreads();   // previous item / candidate fully consumed before smem is reused
            for (int i = tid; i < Hb * Wb; i += GEN_THREADS) f.grid[i] = 0;
            __syncthreads();

            if (tid < 32) {
                Philox rng;
                rng.init(p.seed, (unsigned long long)(p.slot_id_base + m) | ((unsigned long long)(unsigned)gen_count << 40),
                         (unsigned)cand);
                const int si = (int)rng.below((unsigned)nr);   // :21 uniform logical cell
                const int sj = (int)rng.below((unsigned)nc);
                if (algo == MAZE_ALGO_RPRIM) gen_random_prim(f.grid, Wb, nr, nc, si, sj, rng);
                else if (algo == MAZE_ALGO_DFS) gen_depth_first(f.grid, f.queue, Wb, nr, nc, si, sj, rng);
                else gen_prim_and_kill(f.grid, Wb, nr, nc, si, sj, rng);
                if (tid == 0) { s_start = (2 * si + 1) * Wb + 2 * sj + 1; s_best = 0u; }
            }
            __syncthreads();

            // goal = farthest leaf from start, first in row-major order on ties (:187-218)
            const int start_idx = s_start;
            block_bfs(f, Hb, Wb, false, start_idx);
            for (int t = tid; t < nr * nc; t += GEN_THREADS) {
                const int r = 2 * (t / nc) + 1, c = 2 * (t % nc) + 1;
                const int idx = r * Wb + c;
                if (idx == start_idx) continue;
                const int open_nb = (f.grid[idx - Wb] != 0) + (f.grid[idx + Wb] != 0) + (f.grid[idx - 1] != 0) + (f.grid[idx + 1] != 0);
                if (open_nb == 1) atomicMax(&s_best, ((unsigned)f.dist[idx] << 15) | (unsigned)(32767 - idx));
            }
            __syncthreads();
            if (tid == 0) {
                const unsigned key = s_best;
                s_goal = key ? 32767 - (int)(key & 32767u) : start_idx;
                f.grid[s_goal] = 2;   // :33
            }
            __syncthreads();

            if constexpr (kScored) {
                // McClendon difficulty of the bordered maze (base_maze_env.py:86-92; for border-less
                // mazes lib/maze_generation.py:51); f.dist still holds the distances from start
                maze_metrics(f, ms, Hb, Wb, start_idx, s_goal, s_metrics, false);
                if (tid == 0) {
                    s_take = (cand == 0) || (s_metrics.difficulty < s_keep_diff);   // strict <, first wins ties
                    if (s_take) { s_keep_diff = s_metrics.difficulty; s_keep_start = start_idx; s_keep_goal = s_goal; }
                }
                __syncthreads();
                if (p.candidates > 1 && s_take)
                    for (int i = tid; i < Hb * Wb; i += GEN_THREADS) keep[i] = f.grid[i];
            }
        }
        if (kScored && p.candidates > 1) {
            __syncthreads();
            for (int i = tid; i < Hb * Wb; i += GEN_THREADS) f.grid[i] = keep[i];
            if (tid == 0) { s_start = s_keep_start; s_goal = s_keep_goal; }
            __syncthreads();
        }
        const int start_idx = s_start;
        int sr = start_idx / Wb, sc = start_idx % Wb, gr = s_goal / Wb, gc = s_goal % Wb;

        if (tor) {   // :53-55 strip the outer ring
            uint8_t* tmp = reinterpret_cast<uint8_t*>(f.dist);
            for (int i = tid; i < H * W; i += GEN_THREADS) tmp[i] = f.grid[(i / W + 1) * Wb + (i % W) + 1];
            __syncthreads();
            for (int i = tid; i < H * W; i += GEN_THREADS) f.grid[i] = tmp[i];
            __syncthreads();
            sr -= 1; sc -= 1; gr -= 1; gc -= 1;
        }

        block_bfs(f, H, W, tor, gr * W + gc);
        encode_step_table(f, H, W, tor, gr, gc, p.table + (size_t)m * p.slot);
        if (p.grids) {
            uint8_t* g = p.grids + (size_t)m * p.slot;
            for (int i = tid; i < H * W; i += GEN_THREADS) g[i] = f.grid[i];
        }
        if (tid == 0) {
            const int d = f.dist[sr * W + sc];
            const int sol_len = d == DIST_INF ? 0 : d + 1;
            mm[MAZE_META_START] = sr | (sc << 16);
            mm[MAZE_META_GOAL] = gr | (gc << 16);
            mm[MAZE_META_SOL_LEN] = sol_len;
            mm[MAZE_META_MAX_STEPS] = max_steps_budget(H, W, sol_len);
            mm[MAZE_META_SPARE] = gen_count + 1;
            if (kScored && p.difficulty) p.difficulty[item] = s_keep_diff;
        }
    }
}

}  // namespace

extern "C" int maze_generate(maze_ctx* ctx, uint8_t* grids, int32_t* meta, uint8_t* table, const int32_t* ids,
                             const int32_t* count_dev, int n, int slot, int max_h, int max_w,
                             uint64_t seed, int64_t slot_id_base, int candidates, double* difficulty, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!meta || !table) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_generate pointer");
    if (n <= 0 || slot <= 0) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_generate n / slot");
    if (candidates < 1 || candidates > 64) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_generate candidates (1..64)");
    if (max_h < 5 || max_w < 5 || !(max_h & 1) || !(max_w & 1) || max_h + 2 > MAZE_GEN_MAX_DIM || max_w + 2 > MAZE_GEN_MAX_DIM)
        return maze_fail_arg(ctx, MAZE_E_SHAPE, "maze_generate: max shape must be odd, >= 5 and <= MAZE_GEN_MAX_DIM - 2");
    if (max_h * max_w > slot) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_generate: slot smaller than max shape");
    GenParams p;
    p.grids = grids; p.meta = meta; p.table = table; p.ids = ids; p.count_dev = count_dev;
    p.n = n; p.slot = slot;
    p.smem_hw = (max_h + 2) * (max_w + 2);
    const bool scored = candidates > 1 || difficulty != nullptr;
    p.smem_cells = scored ? ((max_h + 1) / 2) * ((max_w + 1) / 2) : 0;
    p.candidates = candidates;
    p.difficulty = difficulty;
    p.seed = seed; p.slot_id_base = slot_id_base;
    size_t smem = field_smem_bytes(p.smem_hw);
    if (scored) smem += (((size_t)p.smem_hw + 15) & ~(size_t)15) + metrics_smem_bytes(p.smem_cells);
    auto kernel = scored ? maze_generate_kernel<true> : maze_generate_kernel<false>;
    MAZE_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MAZE_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, GEN_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int resident = per_sm * (ctx->num_sms > 0 ? ctx->num_sms : 148);
    const int grid = n < resident ? n : resident;
    kernel<<<grid, GEN_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
