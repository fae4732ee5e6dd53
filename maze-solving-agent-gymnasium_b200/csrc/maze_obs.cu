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

__global__ void __launch_bounds__(OBS_THREADS)
maze_window_kernel(maze_env_batch b, float* __restrict__ window, double* __restrict__ agent_norm,
                   double* __restrict__ target_norm) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (OBS_THREADS / 32) + (threadIdx.x >> 5);
    if (e >= b.num_envs) return;
    const EnvState st = unpack_state(b.state[e]);
    const int m = b.env_maze[e];
    const int4 m0 = __ldg(reinterpret_cast<const int4*>(b.meta + (size_t)m * MAZE_META_WORDS));
    const int H = m0.x, W = m0.y;
    const bool tor = (__ldg(b.meta + (size_t)m * MAZE_META_WORDS + MAZE_META_FLAGS) & MAZE_FLAG_TOROIDAL) != 0;
    const int start_idx = (m0.z & 0xffff) * W + (m0.z >> 16);
    const int goal_idx = (m0.w & 0xffff) * W + (m0.w >> 16);
    float* out = window + (size_t)e * (3 * WIN_CELLS);
    if (H < WIN || W < WIN) {   // no 15 x 15 crop exists (the reference cannot build one either)
        for (int i = lane; i < 3 * WIN_CELLS; i += 32) out[i] = 0.0f;
        return;
    }
    int r0 = st.r - WIN / 2, c0 = st.c - WIN / 2;
    if (!tor) {   // extract_submaze: clamped, and the reference uses len(maze) for both axes
        r0 = min(max(r0, 0), H - WIN);
        c0 = min(max(c0, 0), H - WIN);
    }
    const uint8_t* tab = b.table + (size_t)m * b.slot;
#pragma unroll
    for (int k = 0; k < (WIN_CELLS + 31) / 32; ++k) {
        const int i = k * 32 + lane;
        if (i >= WIN_CELLS) break;
        int rr = r0 + i / WIN, cc = c0 + i % WIN;
        if (tor) {   // extract_submaze_toroid: (position + i - k) % maze_shape
            rr = rr < 0 ? rr + H : (rr >= H ? rr - H : rr);
            cc = cc < 0 ? cc + W : (cc >= W ? cc - W : cc);
        }
        const int idx = rr * W + cc;
        const bool open = (__ldg(tab + idx) & MAZE_TAB_OPEN) != 0;
        bool fresh = false;   // non_visited (base_maze_env.py:148-149,183-184)
        if (open && idx != start_idx) {
            const unsigned v = *VISIT_AT(b, e, visit_index(b, rr, cc, W));
            fresh = !((int)(v >> 8) == st.epoch && (v & 0xffu) != 0);
        }
        __stcs(out + i, open ? 0.0f : 1.0f);                                    // maze == 0
        __stcs(out + WIN_CELLS + i, (open && idx != goal_idx) ? 1.0f : 0.0f);   // maze == 1
        __stcs(out + 2 * WIN_CELLS + i, fresh ? 1.0f : 0.0f);
    }
    if (lane < 2) {
        const int goal = m0.w;
        const double shape = (double)(lane == 0 ? H : W);
        if (agent_norm) agent_norm[(size_t)e * 2 + lane] = __ddiv_rn((double)(lane == 0 ? st.r : st.c), shape);
        if (target_norm) target_norm[(size_t)e * 2 + lane] = __ddiv_rn((double)(lane == 0 ? (goal & 0xffff) : (goal >> 16)), shape);
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

}  // namespace

extern "C" int maze_window(maze_ctx* ctx, const maze_env_batch* b, float* window, double* agent_norm,
                           double* target_norm, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!window) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_window: window");
    if (((uintptr_t)window & 3) || ((uintptr_t)agent_norm & 7) || ((uintptr_t)target_norm & 7))
        return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_window pointer alignment");
    const int per_cta = OBS_THREADS / 32;
    const int grid = (b->num_envs + per_cta - 1) / per_cta;
    maze_window_kernel<<<grid, OBS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, window, agent_norm, target_norm);
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
