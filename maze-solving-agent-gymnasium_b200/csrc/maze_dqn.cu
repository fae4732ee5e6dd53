// DQN / DDQN data path on the device (sm_100a): bit-packed observation encode, replay ring,
// uniform sampling with unpack, and the masked epsilon-greedy action selection.
// Reference: lib/replay_memory.py:8-24, agents/ddqn_agent.py:95-108, lib/maze_handler.py:4-162,
// lib/trainers/off_policy_trainer.py:153-171.
//
// A -v1 observation is 6 floats + a 3 x 15 x 15 window of {0, 1}: the window is kept as 24 words
// (one ballot per two window rows and channel), so a transition costs 2 x (24 + 96) + 5 = 245 B of HBM
// instead of 5.5 KB, and a 1 M-transition ring fits in 245 MB.
#include "maze_env.cuh"
#include "maze_replay.cuh"
#include "maze_window.cuh"

namespace {

constexpr int DQN_THREADS = 256;
constexpr int WIN = MAZE_WINDOW;
constexpr int WIN_CELLS = WIN * WIN;
constexpr int WORDS = MAZE_WINDOW_WORDS;
constexpr unsigned FULL = 0xffffffffu;

// Observation of one env by HALF a warp (16 lanes, two envs per warp): lane r < 15 of the half owns window row r --
// it walks the 15 blocks of its row (one table byte and one visit word each; a 4 x 4 visit tile is one 32-byte sector,
// shared by the four lanes whose rows cross it) and assembles the row's three 15-bit masks in registers; lane 2 k then
// takes row 2 k + 1's masks from its neighbour and holds word k of each channel (window rows 2 k in bits 0-14,
// 2 k + 1 in bits 16-30).  Round 1 used a whole warp per env with one ballot per block pair and channel: ~850 warp
// instructions per env (profiles/r01i_dqn_push_details.txt); this form needs about a sixth of that.
// Same rules as maze_window_kernel (maze_obs.cu).
struct PackedObs {
    float vec;          // half-lane < 6: the state vector
    unsigned word[3];   // even half-lane 2 k: word k of channels 0 (wall), 1 (floor), 2 (not visited); k = 0..7
};

// Second half of the encode: the row masks of the 16 lanes -> packed words and the state vector.
__device__ __forceinline__ PackedObs finish_obs16(const EnvState& st, const MazeView& mz, bool valid, int hl, unsigned m0, unsigned m1, unsigned m2) {
    PackedObs o;
    o.vec = 0.f;
    o.word[0] = o.word[1] = o.word[2] = 0u;
    // odd rows move to the even lane above them (all 32 lanes take part in the shuffles)
    const unsigned n0 = __shfl_down_sync(FULL, m0, 1), n1 = __shfl_down_sync(FULL, m1, 1), n2 = __shfl_down_sync(FULL, m2, 1);
    if (!(hl & 1)) {
        const bool has_odd = hl + 1 < WIN;
        o.word[0] = m0 | (has_odd ? n0 << 16 : 0u);
        o.word[1] = m1 | (has_odd ? n1 << 16 : 0u);
        o.word[2] = m2 | (has_odd ? n2 << 16 : 0u);
    }
    if (valid && hl < 6) {
        // float32 of the float64 quotient, like torch.tensor(np.concatenate([...]), dtype=float32); one division per lane
        const int2 bd = best_dir_from_code((st.tab >> MAZE_TAB_CODE_SHIFT) & 7, st.r, st.c, mz.H, mz.W, mz.tor);
        const double num = (double)(hl == 0 ? st.r : (hl == 1 ? st.c : (hl == 2 ? (mz.goal & 0xffff) : (mz.goal >> 16))));
        const double q = __ddiv_rn(num, (double)((hl & 1) ? mz.W : mz.H));
        o.vec = (float)(hl < 4 ? q : (hl == 4 ? (double)bd.x : (double)bd.y));
    }
    return o;
}

__device__ __forceinline__ PackedObs encode_obs16(const maze_env_batch& b, int e, bool valid, int hl) {
    unsigned m0 = 0, m1 = 0, m2 = 0;
    EnvState st{};
    MazeView mz{};
    if (valid) {
        st = unpack_state(b.state[e]);
        mz = load_maze(b, b.env_maze[e]);
        window_row_masks(b, e, st, mz, hl, m0, m1, m2);
    }
    return finish_obs16(st, mz, valid, hl, m0, m1, m2);
}

__device__ __forceinline__ void store_obs16(const PackedObs& o, int hl, float* __restrict__ vec6, uint32_t* __restrict__ words24) {
    if (hl < 6) vec6[hl] = o.vec;
    if (!(hl & 1)) {
        const int k = hl >> 1;
        words24[k] = o.word[0];
        words24[8 + k] = o.word[1];
        words24[16 + k] = o.word[2];
    }
}

__global__ void __launch_bounds__(DQN_THREADS)
maze_dqn_observe_kernel(maze_env_batch b, maze_replay r) {
    const int hl = threadIdx.x & 15;
    const int e = blockIdx.x * (DQN_THREADS / 16) + (threadIdx.x >> 4);
    const bool valid = e < b.num_envs;
    const PackedObs o = encode_obs16(b, e, valid, hl);
    if (valid) store_obs16(o, hl, r.stage_vec + (size_t)e * 6, r.stage_win + (size_t)e * WORDS);
}

// The kernel is bound by the latency of its dependent loads (state -> maze record -> table / bitmap rows -> ring), not by
// bytes or instructions (ncu, profiles/r02i_obs_details.txt: long-scoreboard stalls, 36 % occupancy): throughput = envs in
// flight per SM / chain latency.  So (i) everything that does not depend on the gather is issued before it -- the slot
// claim (one atomicAdd per warp) and the loads of the staged observations travel while the window rows are fetched -- and
// (ii) every half-warp carries TWO envs through the chain together (four per warp): each level of the chain is issued
// for both before either is consumed, which doubles the loads in flight at the same occupancy (1.83e9 -> 3.0e9 env-steps/s
// with the step; four envs per half-warp or more CTAs per SM at fewer registers measured slower: 2.1 - 2.7e9).  This kernel is the
// bordered-maze / visit-bitmap case only (maze_env_batch.flags & MAZE_BATCH_BORDERED: the caller's promise that no maze of the
// batch is toroidal); maze_dqn_push_generic_kernel below handles everything else, one env per half-warp.
constexpr int PUSH_THREADS = 256;
constexpr int PUSH_ENVS = 2;   // envs per 16-lane group

__global__ void __launch_bounds__(PUSH_THREADS, 4)
maze_dqn_push_bordered_kernel(maze_env_batch b, maze_replay r, const uint8_t* __restrict__ actions) {
    const int hl = threadIdx.x & 15, lane = threadIdx.x & 31;
    const int group = blockIdx.x * (PUSH_THREADS / 16) + (threadIdx.x >> 4);
    int e[PUSH_ENVS];
    bool valid[PUSH_ENVS], real[PUSH_ENVS], fast[PUSH_ENVS];
    uint64_t raw_state[PUSH_ENVS];
    int maze_id[PUSH_ENVS];
    // level 1: state and maze id of both envs
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        e[u] = group * PUSH_ENVS + u;
        valid[u] = e[u] < b.num_envs;
        raw_state[u] = valid[u] ? b.state[e[u]] : 0ull;
        maze_id[u] = valid[u] ? b.env_maze[e[u]] : 0;
    }
    EnvState st[PUSH_ENVS];
    MazeView mz[PUSH_ENVS];
    // level 2: maze records; the staged observations (the `state` of the transitions), action and reward
    float sv[PUSH_ENVS];
    uint32_t sw0[PUSH_ENVS], sw1[PUSH_ENVS];
    uint8_t act[PUSH_ENVS];
    float rew[PUSH_ENVS];
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        st[u] = unpack_state(raw_state[u]);
        real[u] = valid[u] && st[u].steps != 0;   // a real transition (steps == 0 only right after a reset)
        mz[u] = MazeView{};
        if (valid[u]) mz[u] = load_maze(b, maze_id[u]);
        sv[u] = 0.f;
        sw0[u] = sw1[u] = 0u;
        act[u] = 0;
        rew[u] = 0.f;
        if (real[u]) {
            if (hl < 6) sv[u] = r.stage_vec[(size_t)e[u] * 6 + hl];
            sw0[u] = r.stage_win[(size_t)e[u] * WORDS + hl];
            if (hl < WORDS - 16) sw1[u] = r.stage_win[(size_t)e[u] * WORDS + 16 + hl];
            if (hl == 0) {
                act[u] = actions[e[u]] & 3;
                rew[u] = (float)b.reward[e[u]];
            }
        }
    }
    // slot claim for the (up to) four real envs of the warp: ballot bits 0 / 16 of each env index
    unsigned real_bits[PUSH_ENVS];
    int n_real = 0;
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        real_bits[u] = __ballot_sync(FULL, real[u] && hl == 0);
        n_real += __popc(real_bits[u]);
    }
    unsigned long long at = 0;
    if (lane == 0 && n_real) at = atomicAdd(r.pushed, (unsigned long long)n_real);
    // level 3: the window rows of both envs
    WindowRowRaw rows[PUSH_ENVS];
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        fast[u] = valid[u] && mz[u].H >= WIN && mz[u].W >= WIN;
        rows[u] = WindowRowRaw{};
        if (fast[u] && hl < WIN) rows[u] = window_row_load_bits(b, e[u], st[u], mz[u], hl);
    }
    at = __shfl_sync(FULL, at, 0);
    PackedObs o[PUSH_ENVS];
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        unsigned m0 = 0, m1 = 0, m2 = 0;
        if (fast[u] && hl < WIN) window_row_fold_bits(rows[u], mz[u], m0, m1, m2);   // mazes below 15 x 15 have no window (zeros), as in window_row_masks
        o[u] = finish_obs16(st[u], mz[u], valid[u], hl, m0, m1, m2);
    }
    // ring order inside the warp: env 0 of the low half, env 0 of the high half, env 1 of the low half, env 1 of the high half
    int before = 0;
#pragma unroll
    for (int u = 0; u < PUSH_ENVS; ++u) {
        if (real[u]) {
            const unsigned long long mine = at + (unsigned long long)(before + ((lane >= 16 && (real_bits[u] & 1u)) ? 1 : 0));
            const size_t slot = (size_t)(mine % (unsigned long long)r.capacity);
            if (hl < 6) r.vec[slot * 6 + hl] = sv[u];
            r.win[slot * WORDS + hl] = sw0[u];
            if (hl < WORDS - 16) r.win[slot * WORDS + 16 + hl] = sw1[u];
            store_obs16(o[u], hl, r.next_vec + slot * 6, r.next_win + slot * WORDS);
            if (hl == 0) {
                r.action[slot] = act[u];
                r.reward[slot] = rew[u];
            }
        }
        before += __popc(real_bits[u]);
        // (the staged observation was read into registers above, before anything re-stages it)
        if (valid[u]) store_obs16(o[u], hl, r.stage_vec + (size_t)e[u] * 6, r.stage_win + (size_t)e[u] * WORDS);
    }
}

// Every other case (torus, no visit bitmap): one env per half-warp, the generic row gather of maze_window.cuh.

__global__ void __launch_bounds__(PUSH_THREADS, 4)
maze_dqn_push_generic_kernel(maze_env_batch b, maze_replay r, const uint8_t* __restrict__ actions) {
    const int hl = threadIdx.x & 15, lane = threadIdx.x & 31;
    const int e = blockIdx.x * (PUSH_THREADS / 16) + (threadIdx.x >> 4);
    const bool valid = e < b.num_envs;
    bool real = false;
    if (valid) real = unpack_state(b.state[e]).steps != 0;   // a real transition (steps == 0 only right after a reset)
    // slot claim for both envs of the warp
    const unsigned real_halves = (__ballot_sync(FULL, real && hl == 0));          // bit 0 / bit 16
    unsigned long long at = 0;
    if (lane == 0 && real_halves) at = atomicAdd(r.pushed, (unsigned long long)__popc(real_halves));
    // staged observation of the env (the `state` of its transition)
    float sv = 0.f;
    uint32_t sw0 = 0, sw1 = 0;
    uint8_t act = 0;
    float rew = 0.f;
    if (real) {
        if (hl < 6) sv = r.stage_vec[(size_t)e * 6 + hl];
        sw0 = r.stage_win[(size_t)e * WORDS + hl];
        if (hl < WORDS - 16) sw1 = r.stage_win[(size_t)e * WORDS + 16 + hl];
        if (hl == 0) {
            act = actions[e] & 3;
            rew = (float)b.reward[e];
        }
    }
    const PackedObs o = encode_obs16(b, e, valid, hl);
    at = __shfl_sync(FULL, at, 0);
    if (real) {
        const unsigned long long mine = at + ((lane >= 16 && (real_halves & 1u)) ? 1ull : 0ull);
        const size_t slot = (size_t)(mine % (unsigned long long)r.capacity);
        if (hl < 6) r.vec[slot * 6 + hl] = sv;
        r.win[slot * WORDS + hl] = sw0;
        if (hl < WORDS - 16) r.win[slot * WORDS + 16 + hl] = sw1;
        store_obs16(o, hl, r.next_vec + slot * 6, r.next_win + slot * WORDS);
        if (hl == 0) {
            r.action[slot] = act;
            r.reward[slot] = rew;
        }
    }
    // (the staged observation was read into registers above, before anything re-stages it)
    if (valid) store_obs16(o, hl, r.stage_vec + (size_t)e * 6, r.stage_win + (size_t)e * WORDS);
}

__device__ __forceinline__ void unpack_window(const uint32_t* __restrict__ words, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const unsigned w = lane < WORDS ? words[lane] : 0u;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned word = __shfl_sync(FULL, w, ch * 8 + k);
            const int row = 2 * k + (lane >> 4), col = lane & 15;   // word k = window rows 2 k (bits 0-14) and 2 k + 1 (bits 16-30)
            if (row < WIN && col < WIN) __stcs(out + ch * WIN_CELLS + row * WIN + col, (float)((word >> lane) & 1u));
        }
    }
}

__global__ void __launch_bounds__(DQN_THREADS)
maze_dqn_sample_kernel(maze_replay r, int n, unsigned long long seed, unsigned long long draw, float* __restrict__ vec,
                       float* __restrict__ win, float* __restrict__ next_vec, float* __restrict__ next_win,
                       int64_t* __restrict__ action, float* __restrict__ reward) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (DQN_THREADS / 32) + (threadIdx.x >> 5);
    if (k >= n) return;
    const unsigned long long pushed = *r.pushed;
    const unsigned long long filled = pushed < (unsigned long long)r.capacity ? pushed : (unsigned long long)r.capacity;
    if (filled == 0) return;
    const size_t slot = replay_slot(r, filled, n, seed, draw, k);
    if (lane < 6) {
        vec[(size_t)k * 6 + lane] = r.vec[slot * 6 + lane];
        next_vec[(size_t)k * 6 + lane] = r.next_vec[slot * 6 + lane];
    }
    if (lane == 0) {
        action[k] = (int64_t)r.action[slot];
        reward[k] = r.reward[slot];
    }
    unpack_window(r.win + slot * WORDS, win + (size_t)k * 3 * WIN_CELLS);
    unpack_window(r.next_win + slot * WORDS, next_win + (size_t)k * 3 * WIN_CELLS);
}

__global__ void __launch_bounds__(DQN_THREADS)
maze_dqn_select_kernel(maze_env_batch b, const float4* __restrict__ q_values, const double* __restrict__ eps_lut, int eps_len,
                       uint32_t* __restrict__ steps_done, unsigned long long seed, long long env_id_base,
                       uint8_t* __restrict__ actions) {
    const int e = blockIdx.x * DQN_THREADS + threadIdx.x;
    if (e >= b.num_envs) return;
    const EnvState st = unpack_state(b.state[e]);
    if (st.flags & MAZE_ST_NEEDS_RESET) { actions[e] = 0; return; }   // the next step is an autoreset
    const uint32_t sd = steps_done[e];
    steps_done[e] = sd + 1;
    const double eps = __ldg(eps_lut + (sd < (uint32_t)eps_len ? (int)sd : eps_len - 1));
    Philox rng;
    rng.init(seed, (uint64_t)(env_id_base + e), sd);
    rng.refill();
    const double u = (double)(((uint64_t)(rng.o0 >> 5) << 26) | (uint64_t)(rng.o1 >> 6)) * (1.0 / 9007199254740992.0);
    int a;
    if (u < eps) {   // ps = mask / mask.sum(); np.random.choice(4, p=ps)  (ddqn_agent.py:102-104)
        const MazeView mz = load_maze(b, b.env_maze[e]);
        float w[4];
        float total = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int dr, dc;
            action_delta(k, dr, dc);
            int nr = st.r + dr, nc = st.c + dc;
            bool open;
            if (mz.tor) {
                nr = nr < 0 ? mz.H - 1 : (nr >= mz.H ? 0 : nr);
                nc = nc < 0 ? mz.W - 1 : (nc >= mz.W ? 0 : nc);
                open = (__ldg(mz.tab + nr * mz.W + nc) & MAZE_TAB_OPEN) != 0;
            } else {
                open = nr >= 0 && nr < mz.H && nc >= 0 && nc < mz.W && (__ldg(mz.tab + nr * mz.W + nc) & MAZE_TAB_OPEN) != 0;
            }
            w[k] = open ? 1.f : 0.f;
        }
        if (((st.flags >> MAZE_ST_NMOVES_SHIFT) & 3) >= 2) {
            const int last = (st.flags >> MAZE_ST_MOVE_SHIFT) & 3;
            w[mz.tor ? 3 - last : (last ^ 1)] = 0.25f;   // see maze_direction_mask
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) total += w[k];
        const float x = (float)(rng.o2 * (1.0 / 4294967296.0)) * total;   // inverse CDF over the four weights
        a = 3;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc += w[k];
            if (x < acc) { a = k; break; }
        }
        if (total == 0.f) a = (int)(rng.o2 >> 30);
        else if (w[a] == 0.f) {   // x fell on the upper edge through rounding: take the last allowed action
            for (int k = 3; k >= 0; --k) if (w[k] > 0.f) { a = k; break; }
        }
    } else {          // source_net(state).max(1)[1]: first maximum
        const float4 q = __ldg(q_values + e);
        a = 0;
        float best = q.x;
        if (q.y > best) { best = q.y; a = 1; }
        if (q.z > best) { best = q.z; a = 2; }
        if (q.w > best) { best = q.w; a = 3; }
    }
    actions[e] = (uint8_t)a;
}

int check_replay(maze_ctx* ctx, const maze_replay* r, bool need_stage) {
    if (!r) return maze_fail_arg(ctx, MAZE_E_NULL, "replay");
    if (!r->pushed || !r->vec || !r->next_vec || !r->win || !r->next_win || !r->action || !r->reward)
        return maze_fail_arg(ctx, MAZE_E_NULL, "replay pointer");
    if (need_stage && (!r->stage_vec || !r->stage_win)) return maze_fail_arg(ctx, MAZE_E_NULL, "replay staging pointer");
    if (r->capacity < 1) return maze_fail_arg(ctx, MAZE_E_RANGE, "replay capacity");
    return 0;
}

}  // namespace

extern "C" int maze_dqn_observe(maze_ctx* ctx, const maze_env_batch* b, const maze_replay* r, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (int rc = check_replay(ctx, r, true)) return rc;
    const int per = DQN_THREADS / 16;
    maze_dqn_observe_kernel<<<(b->num_envs + per - 1) / per, DQN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *r);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_dqn_push(maze_ctx* ctx, const maze_env_batch* b, const maze_replay* r, const uint8_t* actions, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (int rc = check_replay(ctx, r, true)) return rc;
    if (!actions) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_dqn_push: actions");
    if (r->capacity < b->num_envs)   // one launch claims up to num_envs slots: a smaller ring would hand one slot to several warps
        return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_dqn_push: replay capacity must be >= num_envs");
    if ((b->flags & MAZE_BATCH_BORDERED) && b->visit_bits) {
        const int per = PUSH_THREADS / 16 * PUSH_ENVS;
        maze_dqn_push_bordered_kernel<<<(b->num_envs + per - 1) / per, PUSH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *r, actions);
    } else {
        const int per = PUSH_THREADS / 16;
        maze_dqn_push_generic_kernel<<<(b->num_envs + per - 1) / per, PUSH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *r, actions);
    }
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_dqn_sample(maze_ctx* ctx, const maze_replay* r, int n, uint64_t seed, uint64_t draw, float* vec, float* win,
                               float* next_vec, float* next_win, int64_t* action, float* reward, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = check_replay(ctx, r, false)) return rc;
    if (n < 1) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_dqn_sample: n");
    if (!vec || !win || !next_vec || !next_win || !action || !reward) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_dqn_sample output");
    const int per = DQN_THREADS / 32;
    maze_dqn_sample_kernel<<<(n + per - 1) / per, DQN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        *r, n, seed, draw, vec, win, next_vec, next_win, action, reward);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_dqn_select(maze_ctx* ctx, const maze_env_batch* b, const float* q_values, const double* eps_lut, int eps_len,
                               uint32_t* steps_done, uint64_t seed, int64_t env_id_base, uint8_t* actions, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!q_values || !eps_lut || !steps_done || !actions) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_dqn_select pointer");
    if (eps_len < 1) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_dqn_select: eps_len");
    if ((uintptr_t)q_values & 15) return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_dqn_select: q_values must be 16-byte aligned");
    maze_dqn_select_kernel<<<(b->num_envs + DQN_THREADS - 1) / DQN_THREADS, DQN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        *b, reinterpret_cast<const float4*>(q_values), eps_lut, eps_len, steps_done, seed, env_id_base, actions);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
