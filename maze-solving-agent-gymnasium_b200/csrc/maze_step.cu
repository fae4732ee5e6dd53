// Env step / reset kernels (sm_100a).
//
// One thread per env.  State is one packed 64-bit word per env (coalesced 8-byte load/store);
// the maze is read through its one-byte-per-block step table (L2/L1 resident for a maze pool);
// the only scattered DRAM access is the 2-byte visit counter of the block stepped onto.
// Reference semantics: gymnasium_env/envs/base_maze_env.py:136-210, lib/maze_view.py:165-197.
#include "maze_common.cuh"

namespace {

constexpr int STEP_THREADS = 256;

struct StepLuts {
    const double* revisit;  // [256]
    const double* invalid;  // [256]
    double shaping_same, shaping_closer, shaping_farther;   // D[prev]-D[cur] = 0, +1, -1
};

// Warp-cooperative zeroing of the visit arrays of the lanes in `need` (epoch wrap-around:
// once per 255 episodes per env).  Must be called by all 32 lanes.
__device__ __forceinline__ void warp_clear_visits(unsigned need, uint16_t* my_visits, int slot) {
    const int lane = threadIdx.x & 31;
    while (need) {
        int src = __ffs(need) - 1;
        need &= need - 1;
        unsigned long long base = __shfl_sync(0xffffffffu, (unsigned long long)my_visits, src);
        uint32_t* p = reinterpret_cast<uint32_t*>(base);  // slot is even and rows are 4-byte aligned
        for (int i = lane; i < slot / 2; i += 32) p[i] = 0u;
    }
}

// Episode (re)start: BaseMazeEnv.reset, base_maze_env.py:136-161.  The start block is NOT
// marked visited (:159), only excluded from non_visited (:149).
__device__ __forceinline__ void begin_episode(EnvState& s, int start, int tab_at_start, bool& wrapped) {
    s.r = start & 0xffff;
    s.c = start >> 16;
    s.consec = 0;
    s.flags = 0;
    s.steps = 0;
    s.epoch += 1;
    wrapped = s.epoch > 255;
    if (wrapped) s.epoch = 1;
    s.tab = tab_at_start;
}

template <bool kStats>
__global__ void __launch_bounds__(STEP_THREADS)
maze_step_kernel(maze_env_batch b, const uint8_t* __restrict__ actions, uint32_t mode, StepLuts luts) {
    const int e = blockIdx.x * STEP_THREADS + threadIdx.x;
    const bool valid = e < b.num_envs;
    const int ee = valid ? e : b.num_envs - 1;

    EnvState s = unpack_state(b.state[ee]);
    int m = b.env_maze[ee];
    const int a = actions[ee] & 3;
    uint16_t* my_visits = b.visits + (size_t)ee * b.slot;

    const bool do_reset = valid && (mode & MAZE_STEP_AUTORESET) && (s.flags & MAZE_ST_NEEDS_RESET);
    if (do_reset && (mode & MAZE_STEP_WIN_NEXT) && (s.flags & MAZE_ST_WON)) {
        m += b.pool_stride;
        if (m >= b.num_mazes) m -= b.num_mazes;
        b.env_maze[e] = m;
    }

    const int4* mp = reinterpret_cast<const int4*>(b.meta + (size_t)m * MAZE_META_WORDS);
    const int4 m0 = __ldg(mp);
    const int4 m1 = __ldg(mp + 1);
    const int H = m0.x, W = m0.y, start = m0.z, goal = m0.w;
    const int max_steps = m1.x;
    const bool tor = (m1.y & MAZE_FLAG_TOROIDAL) != 0;
    const uint8_t* __restrict__ tab = b.table + (size_t)m * b.slot;

    double reward = 0.0;
    int term = 0, trunc = 0;
    bool wrapped = false;

    if (do_reset) {
        const int sidx = (start & 0xffff) * W + (start >> 16);
        begin_episode(s, start, __ldg(tab + sidx), wrapped);
    } else if (valid) {
        int dr, dc;
        action_delta(a, dr, dc);
        int nr = s.r + dr, nc = s.c + dc;
        bool inb;
        if (tor) {  // lib/maze_view.py:185-186
            nr = nr < 0 ? H - 1 : (nr >= H ? 0 : nr);
            nc = nc < 0 ? W - 1 : (nc >= W ? 0 : nc);
            inb = true;
        } else {    // lib/maze_view.py:169
            inb = (nr > 0) & (nr < H - 1) & (nc > 0) & (nc < W - 1);
        }
        const int idx = inb ? nr * W + nc : s.r * W + s.c;
        const int tb = __ldg(tab + idx);
        const uint32_t vis = my_visits[idx];
        const bool moved = inb && (tb & MAZE_TAB_OPEN);
        if (moved) {
            const int cnt = ((int)(vis >> 8) == s.epoch) ? (int)(vis & 0xff) : 0;
            if (cnt == 0) {
                if ((nr | (nc << 16)) == goal) {
                    reward = 1.0;   // base_maze_env.py:185-187
                    term = 1;
                } else {            // :189-192, len(path) = D_goal + 1
                    const int dd = ((s.tab >> MAZE_TAB_D4_SHIFT) - (tb >> MAZE_TAB_D4_SHIFT)) & 3;
                    reward = dd == 1 ? luts.shaping_closer : (dd == 3 ? luts.shaping_farther : luts.shaping_same);
                }
            } else {
                reward = __ldg(luts.revisit + cnt);   // :194
            }
            my_visits[idx] = (uint16_t)((s.epoch << 8) | (cnt < 255 ? cnt + 1 : 255));   // :196
            s.r = nr;
            s.c = nc;
            s.tab = tb;
            s.consec = 0;
            int nm = (s.flags >> MAZE_ST_NMOVES_SHIFT) & 3;
            nm = nm < 2 ? nm + 1 : 2;
            s.flags = (a << MAZE_ST_MOVE_SHIFT) | (nm << MAZE_ST_NMOVES_SHIFT);
        } else {
            s.consec = s.consec < 255 ? s.consec + 1 : 255;   // :199-200
            reward = __ldg(luts.invalid + s.consec);
            s.flags &= ~(MAZE_ST_NEEDS_RESET | MAZE_ST_WON);
        }
        s.steps = s.steps < 65535 ? s.steps + 1 : 65535;
        if (s.steps > max_steps) {   // :205-208 (overrides a goal reward on the same step)
            trunc = 1;
            reward = -1.0;
        }
        if (term | trunc) s.flags |= MAZE_ST_NEEDS_RESET | (term ? MAZE_ST_WON : 0);
    }

    // epoch wrap-around: the visit array must really be cleared (rare)
    const unsigned need = __ballot_sync(0xffffffffu, wrapped);
    if (need) warp_clear_visits(need, my_visits, b.slot);

    if (!valid) return;

    b.state[e] = pack_state(s);
    st_cs(reinterpret_cast<int2*>(b.agent) + e, make_int2(s.r, s.c));
    st_cs(reinterpret_cast<int2*>(b.target) + e, make_int2(goal & 0xffff, goal >> 16));
    st_cs(reinterpret_cast<int2*>(b.best_dir) + e,
          best_dir_from_code((s.tab >> MAZE_TAB_CODE_SHIFT) & 7, s.r, s.c, H, W, tor));
    st_cs(b.reward + e, reward);
    b.terminated[e] = (uint8_t)term;
    b.truncated[e] = (uint8_t)trunc;

    if (kStats) {
        if (b.ep_return) {
            double g = do_reset ? 0.0 : b.ep_return[e] + reward;
            b.ep_return[e] = g;
            if ((term | trunc) && b.stats_return) atomicAdd(b.stats_return, g);
        }
        if (b.stats) {
            if (term | trunc) {
                atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 0), 1ull);
                if (term) atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 1), 1ull);
                else atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 2), 1ull);
            }
        }
        if (term && (mode & MAZE_STEP_WIN_QUEUE) && b.queue) {
            int at = atomicAdd(b.queue_count, 1);
            b.queue[at] = m;
        }
    }
}

__global__ void __launch_bounds__(STEP_THREADS)
maze_reset_kernel(maze_env_batch b, const uint8_t* __restrict__ mask) {
    const int e = blockIdx.x * STEP_THREADS + threadIdx.x;
    const bool valid = e < b.num_envs;
    const int ee = valid ? e : b.num_envs - 1;
    const bool sel = valid && (mask == nullptr || mask[ee] != 0);
    uint16_t* my_visits = b.visits + (size_t)ee * b.slot;
    bool wrapped = false;
    EnvState s;
    int H = 0, W = 0, goal = 0;
    bool tor = false;
    if (sel) {
        s = unpack_state(b.state[ee]);
        const int m = b.env_maze[ee];
        const int4* mp = reinterpret_cast<const int4*>(b.meta + (size_t)m * MAZE_META_WORDS);
        const int4 m0 = __ldg(mp);
        const int4 m1 = __ldg(mp + 1);
        H = m0.x; W = m0.y; goal = m0.w;
        tor = (m1.y & MAZE_FLAG_TOROIDAL) != 0;
        const int start = m0.z;
        const int sidx = (start & 0xffff) * W + (start >> 16);
        begin_episode(s, start, __ldg(b.table + (size_t)m * b.slot + sidx), wrapped);
    }
    const unsigned need = __ballot_sync(0xffffffffu, wrapped);
    if (need) warp_clear_visits(need, my_visits, b.slot);
    if (!sel) return;
    b.state[e] = pack_state(s);
    reinterpret_cast<int2*>(b.agent)[e] = make_int2(s.r, s.c);
    reinterpret_cast<int2*>(b.target)[e] = make_int2(goal & 0xffff, goal >> 16);
    reinterpret_cast<int2*>(b.best_dir)[e] =
        best_dir_from_code((s.tab >> MAZE_TAB_CODE_SHIFT) & 7, s.r, s.c, H, W, tor);
    b.reward[e] = 0.0;
    b.terminated[e] = 0;
    b.truncated[e] = 0;
    if (b.ep_return) b.ep_return[e] = 0.0;
}

int check_batch(maze_ctx* ctx, const maze_env_batch* b) {
    if (!b) return maze_fail_arg(ctx, MAZE_E_NULL, "batch");
    if (!b->meta || !b->table || !b->env_maze || !b->state || !b->visits || !b->agent || !b->target ||
        !b->best_dir || !b->reward || !b->terminated || !b->truncated)
        return maze_fail_arg(ctx, MAZE_E_NULL, "batch pointer");
    if (b->num_envs <= 0 || b->num_mazes <= 0 || b->slot <= 0 || (b->slot & 1))
        return maze_fail_arg(ctx, MAZE_E_RANGE, "num_envs / num_mazes / slot (must be even)");
    if (((uintptr_t)b->meta & 15) || ((uintptr_t)b->state & 7) || ((uintptr_t)b->agent & 7) ||
        ((uintptr_t)b->target & 7) || ((uintptr_t)b->best_dir & 7) || ((uintptr_t)b->reward & 7) ||
        ((uintptr_t)b->visits & 3))
        return maze_fail_arg(ctx, MAZE_E_ALIGN, "batch pointer alignment");
    return 0;
}

}  // namespace

extern "C" int maze_step(maze_ctx* ctx, const maze_env_batch* b, const uint8_t* actions, uint32_t mode, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = check_batch(ctx, b)) return rc;
    if (!actions) return maze_fail_arg(ctx, MAZE_E_NULL, "actions");
    StepLuts luts;
    luts.revisit = ctx->d_lut_revisit;
    luts.invalid = ctx->d_lut_invalid;
    luts.shaping_same = ctx->h_shaping[0];
    luts.shaping_closer = ctx->h_shaping[1];
    luts.shaping_farther = ctx->h_shaping[3];
    const int grid = (b->num_envs + STEP_THREADS - 1) / STEP_THREADS;
    const bool stats = b->ep_return || b->stats || ((mode & MAZE_STEP_WIN_QUEUE) && b->queue);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (stats) maze_step_kernel<true><<<grid, STEP_THREADS, 0, st>>>(*b, actions, mode, luts);
    else maze_step_kernel<false><<<grid, STEP_THREADS, 0, st>>>(*b, actions, mode, luts);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_reset(maze_ctx* ctx, const maze_env_batch* b, const uint8_t* mask, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = check_batch(ctx, b)) return rc;
    const int grid = (b->num_envs + STEP_THREADS - 1) / STEP_THREADS;
    maze_reset_kernel<<<grid, STEP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, mask);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
