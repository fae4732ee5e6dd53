// Enriched (-v1) observation and direction-mask kernels (sm_100a).
//
// maze_window: one warp per env writes the float32 [3, 15, 15] crop around the agent
// (lib/maze_handler.py:4-99) -- 2 700 B of coalesced streaming stores per env; the maze comes from
// the L2-resident step table, the visit counters from the env-major visit array (a window row is
// 30 contiguous bytes).  HBM-write bound.
// maze_direction_mask: one thread per env, four table bytes (lib/maze_handler.py:122-162,
// simple_maze_env.py:41-50, toroidal_maze_env.py:57-70).
#include "maze_env.cuh"

namespace {

constexpr int OBS_THREADS = 256;
constexpr int WIN = MAZE_WINDOW;
constexpr int WIN_CELLS = WIN * WIN;

// A CTA of 8 warps serves 8 consecutive envs.  Each warp gathers its env's 225 window blocks and writes
// one byte (0 / 1) per output float into shared memory, laid out exactly like the output; then all 256
// threads stream the 8 x 2 700 bytes out as 16-byte stores, four shared bytes -> one float4 (8 envs x
// 675 floats start on a 16-byte boundary; one env's 2 700 bytes do not).  ncu of the first version
// (one 4-byte store per value, div / mod per block): 1 080 warp instructions per env, 77 % issue-slot
// busy, table rows missing L2 behind the streamed output -- hence the incremental indexing, the byte
// staging and the evict_last policy on the table loads below.
constexpr int WIN_ENVS = OBS_THREADS / 32;
constexpr int WIN_FLOATS = 3 * WIN_CELLS;

#ifndef MAZE_WIN_MINB
#define MAZE_WIN_MINB 5
#endif
// kTiled: 4 x 4-tiled visit index; kUnit: visit_cell_stride == 1 (env-major array, the -v1 default)
template <bool kTiled, bool kUnit>
__global__ void __launch_bounds__(OBS_THREADS, MAZE_WIN_MINB)
maze_window_kernel(maze_env_batch b, float* __restrict__ window, double* __restrict__ agent_norm,
                   double* __restrict__ target_norm) {
    __shared__ __align__(16) uint8_t s_out[WIN_ENVS * WIN_FLOATS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e0 = blockIdx.x * WIN_ENVS, e = e0 + w;
    uint8_t* so = s_out + w * WIN_FLOATS;
    if (e < b.num_envs) {
        const EnvState st = unpack_state(b.state[e]);
        const int m = b.env_maze[e];
        const int4 m0 = __ldg(reinterpret_cast<const int4*>(b.meta + (size_t)m * MAZE_META_WORDS));
        const int H = m0.x, W = m0.y;
        const bool tor = (__ldg(b.meta + (size_t)m * MAZE_META_WORDS + MAZE_META_FLAGS) & MAZE_FLAG_TOROIDAL) != 0;
        const int start_idx = (m0.z & 0xffff) * W + (m0.z >> 16);
        const int goal_idx = (m0.w & 0xffff) * W + (m0.w >> 16);
        if (H < WIN || W < WIN) {   // no 15 x 15 crop exists (the reference cannot build one either)
            for (int i = lane; i < WIN_FLOATS; i += 32) so[i] = 0;
        } else {
            int r0 = st.r - WIN / 2, c0 = st.c - WIN / 2;
            if (!tor) {   // extract_submaze: clamped, and the reference uses len(maze) for both axes
                r0 = min(max(r0, 0), H - WIN);
                c0 = min(max(c0, 0), W - WIN);   // the reference clamps with len(maze) (maze_handler.py:21-29: square mazes only); W keeps a non-square slot in bounds
            }
            const uint8_t* tab = b.table + (size_t)m * b.slot;
            const uint64_t pol_table = l2_policy<MAZE_TABLE_POLICY>();
            const uint16_t* vbase = b.visits + (size_t)e * b.visit_env_stride;
            const int wt = (W + 3) >> 2;
            // Lane -> window block mapping: lanes 0-14 take the columns of an even window row, lanes 16-30 those of
            // the odd row below it, eight row pairs per lane (lanes 15 / 31 and row 15 are idle: 225 of 256 slots).
            // The column part of every index is then fixed per lane and the row part is a compile-time step, so the
            // loop body is a handful of integer instructions around its two loads (the kernel is issue-bound).
            // Table byte and visit word of all eight blocks are requested before any is used; the visit word is
            // fetched whether or not the block is open (same sectors, one dependent round trip less).
            constexpr int K = (WIN + 1) / 2;
            const int half = lane >> 4, col = min(lane & 15, WIN - 1);
            const bool col_ok = (lane & 15) < WIN;
            int cc = c0 + col;
            if (tor) cc = cc < 0 ? cc + W : (cc >= W ? cc - W : cc);   // extract_submaze_toroid: (position + i - k) % maze_shape
            const int cpart = kTiled ? (((cc >> 2) << 4) | (cc & 3)) : cc;
            int idxs[K], tb[K];
            unsigned vis[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int rr = r0 + min(2 * k + half, WIN - 1);
                if (tor) rr = rr < 0 ? rr + H : (rr >= H ? rr - H : rr);
                idxs[k] = rr * W + cc;
                tb[k] = pol_load_nc<MAZE_TABLE_POLICY>(tab + idxs[k], pol_table);
                const int vi = kTiled ? ((((rr >> 2) * wt) << 4) | ((rr & 3) << 2)) + cpart : idxs[k];
                vis[k] = __ldcs(kUnit ? vbase + vi : vbase + (size_t)vi * b.visit_cell_stride);
            }
            uint8_t* sl = so + half * WIN + col;   // block (row 2 k + half, col) is output index 30 k + half * 15 + col
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (!col_ok || 2 * k + half >= WIN) continue;
                const bool open = (tb[k] & MAZE_TAB_OPEN) != 0;
                // non_visited (base_maze_env.py:148-149,183-184): open, not the start, no visit in this episode
                const bool fresh = open && idxs[k] != start_idx && !((int)(vis[k] >> 8) == st.epoch && (vis[k] & 0xffu) != 0);
                sl[2 * WIN * k] = open ? 0 : 1;                                          // maze == 0
                sl[WIN_CELLS + 2 * WIN * k] = (open && idxs[k] != goal_idx) ? 1 : 0;     // maze == 1
                sl[2 * WIN_CELLS + 2 * WIN * k] = fresh ? 1 : 0;
            }
        }
        if (lane < 2) {
            const int goal = m0.w;
            const double shape = (double)(lane == 0 ? H : W);
            if (agent_norm) agent_norm[(size_t)e * 2 + lane] = __ddiv_rn((double)(lane == 0 ? st.r : st.c), shape);
            if (target_norm) target_norm[(size_t)e * 2 + lane] = __ddiv_rn((double)(lane == 0 ? (goal & 0xffff) : (goal >> 16)), shape);
        }
    }
    __syncthreads();
    const int n_float = min(WIN_ENVS, b.num_envs - e0) * WIN_FLOATS;
    float* out = window + (size_t)e0 * WIN_FLOATS;   // 16-byte aligned: e0 is a multiple of 8
    for (int q = threadIdx.x; 4 * q < n_float; q += OBS_THREADS) {
        const int f = 4 * q;
        const unsigned v = *reinterpret_cast<const unsigned*>(s_out + f);   // four values, one byte each
        if (f + 3 < n_float) {
            __stcs(reinterpret_cast<float4*>(out) + q,
                   make_float4(__uint_as_float((v & 1u) * 0x3f800000u), __uint_as_float((v >> 8 & 1u) * 0x3f800000u),
                               __uint_as_float((v >> 16 & 1u) * 0x3f800000u), __uint_as_float((v >> 24 & 1u) * 0x3f800000u)));
        } else {
            for (int i = f; i < n_float; ++i) __stcs(out + i, (float)s_out[i]);
        }
    }
}

__global__ void __launch_bounds__(OBS_THREADS)
maze_direction_mask_kernel(maze_env_batch b, int probs, float4* __restrict__ mask) {
    const int e = blockIdx.x * OBS_THREADS + threadIdx.x;
    if (e >= b.num_envs) return;
    const EnvState st = unpack_state(b.state[e]);
    const int m = b.env_maze[e];
    const int2 shape = __ldg(reinterpret_cast<const int2*>(b.meta + (size_t)m * MAZE_META_WORDS));
    const int H = shape.x, W = shape.y;
    const bool tor = (__ldg(b.meta + (size_t)m * MAZE_META_WORDS + MAZE_META_FLAGS) & MAZE_FLAG_TOROIDAL) != 0;
    const uint8_t* tab = b.table + (size_t)m * b.slot;
    float v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int dr, dc;
        action_delta(a, dr, dc);
        int nr = st.r + dr, nc = st.c + dc;
        bool open;
        if (tor) {
            nr = nr < 0 ? H - 1 : (nr >= H ? 0 : nr);
            nc = nc < 0 ? W - 1 : (nc >= W ? 0 : nc);
            open = (__ldg(tab + nr * W + nc) & MAZE_TAB_OPEN) != 0;
        } else {   // agents stand on interior blocks of a bordered maze; stay in the slot regardless
            open = nr >= 0 && nr < H && nc >= 0 && nc < W && (__ldg(tab + nr * W + nc) & MAZE_TAB_OPEN) != 0;
        }
        v[a] = open ? 1.0f : 0.0f;
    }
    if (probs && ((st.flags >> MAZE_ST_NMOVES_SHIFT) & 3) >= 2) {   // len(visited_cell) > 1
        const int last = (st.flags >> MAZE_ST_MOVE_SHIFT) & 3;
        // euclid: previous - current = -ACTIONS[last] -> the opposite action.  torus: the tuple is
        // built (dx, dy), i.e. column-first, before the same index lookup (toroidal_maze_env.py:63-68)
        const int back = tor ? 3 - last : (last ^ 1);
        v[back] = 0.25f;
    }
    mask[e] = make_float4(v[0], v[1], v[2], v[3]);
}

// maze_render: the frame MazeViewTemplate keeps on its pygame surface (lib/maze_view.py:12-14,88-104,
// 148-152), as uint8 [n, out_h, out_w, 3] for n selected envs.  One CTA per (env, tile row); a thread
// produces four pixels (12 bytes) at a time.
//   tile (r, c) = 16 x 16 pixels at (16 c, 16 r): CELL_COLORS[maze[r][c]] with a one-pixel outline,
//   (59, 66, 82) as drawn by __draw_maze, (208, 135, 112) once the agent has left the block
//   (_draw_cell, called by move_agent / _reset_agent on the block the agent leaves);
//   agent = 8 x 8 square at offset (4, 4) of its tile; the layers are one pixel smaller than the
//   window (:41-44), so the last pixel row / column stay black, like everything beyond the maze.
// "Left" is derived from the episode's visit counters: start once the episode has a legal move, the
// agent's own block when it is there for at least the second time, any other block with a visit.
__global__ void __launch_bounds__(OBS_THREADS)
maze_render_kernel(maze_env_batch b, const int32_t* __restrict__ env_ids, int out_h, int out_w, uint8_t* __restrict__ out) {
    constexpr int T = MAZE_RENDER_TILE;
    const int k = blockIdx.x, r = blockIdx.y;
    const int e = env_ids ? env_ids[k] : k;
    const EnvState st = unpack_state(b.state[e]);
    const int m = b.env_maze[e];
    const int4 m0 = __ldg(reinterpret_cast<const int4*>(b.meta + (size_t)m * MAZE_META_WORDS));
    const int H = m0.x, W = m0.y;
    const int start_r = m0.z & 0xffff, start_c = m0.z >> 16, goal_r = m0.w & 0xffff, goal_c = m0.w >> 16;
    const uint8_t* tab = b.table + (size_t)m * b.slot;
    const bool has_moved = ((st.flags >> MAZE_ST_NMOVES_SHIFT) & 3) >= 1;
    __shared__ uint8_t s_tile[MAZE_GEN_MAX_DIM];   // per tile of this row: value (0 / 1 / 2) | trail << 2
    for (int c = threadIdx.x; c < out_w / T; c += OBS_THREADS) {
        int code = 0;
        if (r < H && c < W) {
            const bool open = (__ldg(tab + r * W + c) & MAZE_TAB_OPEN) != 0;
            const int v = open ? ((r == goal_r && c == goal_c) ? 2 : 1) : 0;
            bool trail = false;
            if (open) {
                const unsigned vis = *VISIT_AT(b, e, visit_index(b, r, c, W));
                const int cnt = ((int)(vis >> 8) == st.epoch) ? (int)(vis & 0xff) : 0;
                if (r == start_r && c == start_c) trail = has_moved;
                else if (r == st.r && c == st.c) trail = cnt >= 2;
                else trail = cnt >= 1;
            }
            code = v | (trail ? 4 : 0) | 8;
        }
        s_tile[c] = (uint8_t)code;
    }
    __syncthreads();
    const int groups = out_w / 4;   // four pixels per thread and iteration
    for (int p = threadIdx.x; p < T * groups; p += OBS_THREADS) {
        const int ly = p / groups, x0 = (p % groups) * 4, y = r * T + ly;
        if (y >= out_h) break;
        unsigned char px[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x0 + i, c = x / T, lx = x % T;
            const int code = s_tile[c];
            int cr = 0, cg = 0, cb = 0;
            if ((code & 8) && x != W * T - 1 && y != H * T - 1) {
                const int v = code & 3;
                if (lx == 0 || lx == T - 1 || ly == 0 || ly == T - 1) {
                    if (code & 4) { cr = 208; cg = 135; cb = 112; } else { cr = 59; cg = 66; cb = 82; }
                } else if (r == st.r && c == st.c && lx >= T / 4 && lx < T / 4 + T / 2 && ly >= T / 4 && ly < T / 4 + T / 2) {
                    cr = 94; cg = 129; cb = 172;                                     // AGENT_COLOR
                } else if (v == 0) { cr = 46; cg = 52; cb = 64; }                    // CELL_COLORS: wall, floor, goal
                else if (v == 1) { cr = 236; cg = 239; cb = 244; }
                else { cr = 163; cg = 190; cb = 140; }
            }
            px[3 * i] = (unsigned char)cr; px[3 * i + 1] = (unsigned char)cg; px[3 * i + 2] = (unsigned char)cb;
        }
        unsigned* o = reinterpret_cast<unsigned*>(out + (((size_t)k * out_h + y) * out_w + x0) * 3);
#pragma unroll
        for (int w = 0; w < 3; ++w)
            __stcs(o + w, (unsigned)px[4 * w] | ((unsigned)px[4 * w + 1] << 8) | ((unsigned)px[4 * w + 2] << 16) | ((unsigned)px[4 * w + 3] << 24));
    }
}

}  // namespace

extern "C" int maze_render(maze_ctx* ctx, const maze_env_batch* b, const int32_t* env_ids, int n, uint8_t* out,
                           int out_h, int out_w, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!out) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_render: out");
    if (n <= 0 || out_h <= 0 || out_w <= 0 || out_h % MAZE_RENDER_TILE || out_w % MAZE_RENDER_TILE ||
        out_h / MAZE_RENDER_TILE > MAZE_GEN_MAX_DIM || out_w / MAZE_RENDER_TILE > MAZE_GEN_MAX_DIM)
        return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_render: n / out_h / out_w (multiples of MAZE_RENDER_TILE, at most MAZE_GEN_MAX_DIM tiles)");
    if ((uintptr_t)out & 3) return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_render: out must be 4-byte aligned");
    if (env_ids == nullptr && n > b->num_envs) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_render: n exceeds num_envs");
    maze_render_kernel<<<dim3(n, out_h / MAZE_RENDER_TILE), OBS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, env_ids, out_h, out_w, out);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_window(maze_ctx* ctx, const maze_env_batch* b, float* window, double* agent_norm,
                           double* target_norm, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!window) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_window: window");
    if (((uintptr_t)window & 15) || ((uintptr_t)agent_norm & 7) || ((uintptr_t)target_norm & 7))
        return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_window pointer alignment");
    const int per_cta = OBS_THREADS / 32;
    const int grid = (b->num_envs + per_cta - 1) / per_cta;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool unit = b->visit_cell_stride == 1;
    if (b->visit_tiled) {
        if (unit) maze_window_kernel<true, true><<<grid, OBS_THREADS, 0, st>>>(*b, window, agent_norm, target_norm);
        else maze_window_kernel<true, false><<<grid, OBS_THREADS, 0, st>>>(*b, window, agent_norm, target_norm);
    } else {
        if (unit) maze_window_kernel<false, true><<<grid, OBS_THREADS, 0, st>>>(*b, window, agent_norm, target_norm);
        else maze_window_kernel<false, false><<<grid, OBS_THREADS, 0, st>>>(*b, window, agent_norm, target_norm);
    }
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_direction_mask(maze_ctx* ctx, const maze_env_batch* b, int probs, float* mask, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!mask) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_direction_mask: mask");
    if ((uintptr_t)mask & 15) return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_direction_mask: mask must be 16-byte aligned");
    const int grid = (b->num_envs + OBS_THREADS - 1) / OBS_THREADS;
    maze_direction_mask_kernel<<<grid, OBS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        *b, probs, reinterpret_cast<float4*>(mask));
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
