// Tabular Q-learning / double Q-learning on the device (sm_100a).
//
// Reference: agents/q_agent.py:8-79 (QAgent), agents/dq_agent.py:5-73 (DQAgent), driven by the loop
// of lib/trainers/off_policy_trainer.py:38-51,76-78.  The reference's defaultdict keyed by str(obs)
// becomes an open-addressing hash table in HBM: 8-byte keys, rows of 4 float64 (one 32-byte sector).
// Per env-step the learner touches the row of the next observation (key probe + row read) and
// read-modify-writes one entry of the current row, so the traffic is a handful of random 32-byte
// sectors: HBM-random bound, not coalesced-stream bound.  maze_q_rollout fuses policy + env step +
// update for K steps per launch with the env state in registers (no per-step launch, no obs
// round-trip through HBM).
#include <cmath>
#include "maze_env.cuh"

namespace {

constexpr int Q_THREADS = 128;
constexpr int Q_MAX_PROBE = 1 << 14;

struct QCursor {   // per-env learner state held in registers
    uint32_t slot, steps_done, pos_u, pos_a;
    double ep_return;
};

__device__ __forceinline__ uint64_t q_key(const EnvState& st, int goal, unsigned agent_id) {
    const unsigned code = ((unsigned)st.tab >> MAZE_TAB_CODE_SHIFT) & 7u;   // 'best dir' of the current block
    return (uint64_t)((unsigned)st.r | ((unsigned)st.c << 8) | ((unsigned)(goal & 0xff) << 16) | ((unsigned)((goal >> 16) & 0xff) << 24)) |
           ((uint64_t)code << 32) | ((uint64_t)agent_id << 35);
}

// defaultdict lookup: find the row of `key`, claiming a zero row if it does not exist yet
__device__ __forceinline__ uint32_t q_find_or_insert(const maze_q_agent& ag, uint64_t key) {
    const uint64_t mask = (uint64_t)ag.capacity - 1;
    uint64_t i = (key * 0x9E3779B97F4A7C15ull) >> 20 & mask;
    for (int probe = 0; probe < Q_MAX_PROBE; ++probe) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(ag.keys + i);
        if (k == key) return (uint32_t)i;
        if (k == MAZE_Q_EMPTY) {
            k = atomicCAS(reinterpret_cast<unsigned long long*>(ag.keys + i), MAZE_Q_EMPTY, (unsigned long long)key);
            if (k == MAZE_Q_EMPTY || k == key) return (uint32_t)i;
        }
        i = (i + 1) & mask;
    }
    *ag.overflow = 1;
    return (uint32_t)(key & mask);   // keep running on a (wrong) row; the host checks `overflow`
}

__device__ __forceinline__ int argmax4(const double* row) {   // np.argmax: first maximum
    const double2 lo = *reinterpret_cast<const double2*>(row), hi = *reinterpret_cast<const double2*>(row + 2);
    int a = 0;
    double best = lo.x;
    if (lo.y > best) { best = lo.y; a = 1; }
    if (hi.x > best) { best = hi.x; a = 2; }
    if (hi.y > best) { best = hi.y; a = 3; }
    return a;
}

__device__ __forceinline__ double max4(const double* row) {
    const double2 lo = *reinterpret_cast<const double2*>(row), hi = *reinterpret_cast<const double2*>(row + 2);
    return fmax(fmax(lo.x, lo.y), fmax(hi.x, hi.y));
}

// One get_action call (q_agent.py:44-54 / dq_agent.py:36-47).  `coin` receives an independent
// uniform bit-source for DQAgent.update's table choice when no tape is attached.
__device__ __forceinline__ int q_get_action(const maze_q_agent& ag, int B, int e, QCursor& cur, uint32_t* coin_word) {
    const int li = cur.steps_done < (uint32_t)ag.eps_len ? (int)cur.steps_done : ag.eps_len - 1;
    const double eps = __ldg(ag.eps_lut + li);
    double u;
    uint32_t rnd_action;
    if (ag.u_tape) {
        u = ag.u_tape[(size_t)min(cur.pos_u, (uint32_t)ag.u_len - 1) * B + e];
        ++cur.pos_u;
        rnd_action = 0;
    } else {
        Philox rng;
        rng.init(ag.seed, (uint64_t)(ag.env_id_base + e), cur.steps_done);
        rng.refill();
        u = (double)(((uint64_t)(rng.o0 >> 5) << 26) | (uint64_t)(rng.o1 >> 6)) * (1.0 / 9007199254740992.0);   // 53 bits
        rnd_action = rng.o2 >> 30;
        if (coin_word) *coin_word = rng.o3;
    }
    ++cur.steps_done;
    if (u < eps) {
        if (ag.a_tape) {
            rnd_action = ag.a_tape[(size_t)min(cur.pos_a, (uint32_t)ag.a_len - 1) * B + e] & 3u;
            ++cur.pos_a;
        }
        return (int)rnd_action;
    }
    return argmax4(ag.q_a + (size_t)cur.slot * 4);
}

// QAgent.update (q_agent.py:56-72) or DQAgent.update (dq_agent.py:49-66) for one transition.
// The entry is written with a plain store: with one env per agent that IS the reference update;
// when many envs share an agent, concurrent updates of one entry race and one of them wins
// (summing them with atomics would multiply the learning rate by the number of writers).
__device__ __forceinline__ void q_learn(const maze_q_agent& ag, int B, int e, QCursor& cur, uint32_t next_slot, int action,
                                        double reward, bool terminated, double gamma) {
    double* row_a = ag.q_a + (size_t)cur.slot * 4;
    if (!ag.q_b) {
        const double future = terminated ? 0.0 : max4(ag.q_a + (size_t)next_slot * 4);
        const double td = __dadd_rn(__dadd_rn(reward, __dmul_rn(gamma, future)), -row_a[action]);
        row_a[action] = __dadd_rn(row_a[action], __dmul_rn(ag.lr, td));
        return;
    }
    // double Q: the coin is drawn before the bootstrap action; the bootstrap action is epsilon-greedy
    // on Q_A and advances steps_done; `terminated` is ignored (dq_agent.py:57-64)
    bool update_a;
    uint32_t coin = 0;
    if (ag.u_tape) {
        update_a = ag.u_tape[(size_t)min(cur.pos_u, (uint32_t)ag.u_len - 1) * B + e] < 0.5;
        ++cur.pos_u;
    }
    QCursor at_next = cur;
    at_next.slot = next_slot;
    const int best = q_get_action(ag, B, e, at_next, &coin);
    cur.steps_done = at_next.steps_done; cur.pos_u = at_next.pos_u; cur.pos_a = at_next.pos_a;
    if (!ag.u_tape) update_a = (coin >> 31) == 0;
    double* row_b = ag.q_b + (size_t)cur.slot * 4;
    if (update_a) {
        const double boot = ag.q_b[(size_t)next_slot * 4 + best];
        const double td = __dadd_rn(__dadd_rn(reward, __dmul_rn(gamma, boot)), -row_a[action]);
        row_a[action] = __dadd_rn(row_a[action], __dmul_rn(ag.lr, td));
    } else {
        const double boot = ag.q_a[(size_t)next_slot * 4 + best];
        const double td = __dadd_rn(__dadd_rn(reward, __dmul_rn(gamma, boot)), -row_b[action]);
        row_b[action] = __dadd_rn(row_b[action], __dmul_rn(ag.lr, td));
    }
}

// update_hyperparameter after an episode (q_agent.py:75-79; off_policy_trainer.py:76-78 compares the
// episode return with prev_cum_rew, which that loop resets to 0 every episode)
__device__ __forceinline__ void q_episode_end(const maze_q_agent& ag, unsigned agent_id, QCursor& cur) {
    const double delta = cur.ep_return > 0.0 ? ag.eta : -ag.eta;
    if (ag.envs_per_agent == 1) ag.gamma[agent_id] = __dadd_rn(ag.gamma[agent_id], delta);
    else atomicAdd(ag.gamma + agent_id, delta / ag.envs_per_agent);   // one round of episodes ~ one reference episode
    cur.ep_return = 0.0;
}

__device__ __forceinline__ QCursor load_cursor(const maze_q_agent& ag, int B, int e) {
    QCursor c;
    c.slot = ag.slot[e];
    c.steps_done = ag.steps_done[e];
    c.ep_return = ag.ep_return[e];
    c.pos_u = ag.tape_pos ? ag.tape_pos[e] : 0;
    c.pos_a = ag.tape_pos ? ag.tape_pos[B + e] : 0;
    return c;
}

__device__ __forceinline__ void store_cursor(const maze_q_agent& ag, int B, int e, const QCursor& c) {
    ag.slot[e] = c.slot;
    ag.steps_done[e] = c.steps_done;
    ag.ep_return[e] = c.ep_return;
    if (ag.tape_pos) { ag.tape_pos[e] = c.pos_u; ag.tape_pos[B + e] = c.pos_a; }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Q_THREADS)
maze_q_act_kernel(maze_env_batch b, maze_q_agent ag, uint8_t* __restrict__ actions) {
    const int e = blockIdx.x * Q_THREADS + threadIdx.x;
    if (e >= b.num_envs) return;
    const EnvState st = unpack_state(b.state[e]);
    if (st.flags & MAZE_ST_NEEDS_RESET) {   // the next step is an autoreset: no decision is made
        actions[e] = 0;
        ag.last_action[e] = 0;
        return;
    }
    QCursor cur = load_cursor(ag, b.num_envs, e);
    if (cur.slot == MAZE_Q_NO_SLOT) {
        const int goal = __ldg(b.meta + (size_t)b.env_maze[e] * MAZE_META_WORDS + MAZE_META_GOAL);
        cur.slot = q_find_or_insert(ag, q_key(st, goal, (unsigned)(e / ag.envs_per_agent)));
    }
    const int a = q_get_action(ag, b.num_envs, e, cur, nullptr);
    actions[e] = (uint8_t)a;
    ag.last_action[e] = (uint8_t)a;
    store_cursor(ag, b.num_envs, e, cur);
}

__global__ void __launch_bounds__(Q_THREADS)
maze_q_update_kernel(maze_env_batch b, maze_q_agent ag) {
    const int e = blockIdx.x * Q_THREADS + threadIdx.x;
    if (e >= b.num_envs) return;
    const EnvState st = unpack_state(b.state[e]);
    const unsigned agent_id = (unsigned)(e / ag.envs_per_agent);
    const int goal = __ldg(b.meta + (size_t)b.env_maze[e] * MAZE_META_WORDS + MAZE_META_GOAL);
    QCursor cur = load_cursor(ag, b.num_envs, e);
    const uint32_t next_slot = q_find_or_insert(ag, q_key(st, goal, agent_id));
    if (st.steps == 0 || cur.slot == MAZE_Q_NO_SLOT) {   // the step was a reset: only re-anchor
        cur.slot = next_slot;
        store_cursor(ag, b.num_envs, e, cur);
        return;
    }
    const double reward = b.reward[e];
    const bool term = b.terminated[e] != 0, trunc = b.truncated[e] != 0;
    cur.ep_return = __dadd_rn(cur.ep_return, reward);
    q_learn(ag, b.num_envs, e, cur, next_slot, ag.last_action[e] & 3, reward, term, ag.gamma[agent_id]);
    cur.slot = next_slot;
    if (term || trunc) q_episode_end(ag, agent_id, cur);
    store_cursor(ag, b.num_envs, e, cur);
}

__global__ void __launch_bounds__(Q_THREADS)
maze_q_rollout_kernel(maze_env_batch b, maze_q_agent ag, int k_steps, uint32_t mode, StepLuts luts) {
    const int e = blockIdx.x * Q_THREADS + threadIdx.x;
    if (e >= b.num_envs) return;
    const int B = b.num_envs;
    const unsigned agent_id = (unsigned)(e / ag.envs_per_agent);
    EnvState st = unpack_state(b.state[e]);
    int m = b.env_maze[e];
    MazeView mz = load_maze(b, m);
    QCursor cur = load_cursor(ag, B, e);
    double reward = 0.0;
    int term = 0, trunc = 0;
    unsigned long long n_episodes = 0, n_wins = 0, n_steps = 0;
    double return_sum = 0.0;

    for (int k = 0; k < k_steps; ++k) {
        if ((mode & MAZE_STEP_AUTORESET) && (st.flags & MAZE_ST_NEEDS_RESET)) {
            if ((mode & MAZE_STEP_WIN_NEXT) && (st.flags & MAZE_ST_WON)) {
                m += b.pool_stride;
                if (m >= b.num_mazes) m -= b.num_mazes;
                mz = load_maze(b, m);
            }
            bool wrapped;
            begin_episode(st, mz.start, __ldg(mz.tab + (mz.start & 0xffff) * mz.W + (mz.start >> 16)), wrapped);
            visit_bits_clear(b, e);
            if (wrapped)
                for (int i = 0; i < b.visit_slot; ++i) *VISIT_AT(b, e, i) = 0;
            cur.slot = q_find_or_insert(ag, q_key(st, mz.goal, agent_id));
            reward = 0.0; term = 0; trunc = 0;
            continue;
        }
        if (cur.slot == MAZE_Q_NO_SLOT) cur.slot = q_find_or_insert(ag, q_key(st, mz.goal, agent_id));
        const int a = q_get_action(ag, B, e, cur, nullptr);
        const StepResult r = env_transition(b, e, st, mz, a, luts);
        reward = r.reward; term = r.term; trunc = r.trunc;
        ++n_steps;
        const uint32_t next_slot = q_find_or_insert(ag, q_key(st, mz.goal, agent_id));
        cur.ep_return = __dadd_rn(cur.ep_return, reward);
        q_learn(ag, B, e, cur, next_slot, a, reward, term != 0, ag.gamma[agent_id]);
        cur.slot = next_slot;
        if (term | trunc) {
            ++n_episodes;
            n_wins += term;
            return_sum += cur.ep_return;
            q_episode_end(ag, agent_id, cur);
        }
    }

    b.state[e] = pack_state(st);
    b.env_maze[e] = m;
    store_cursor(ag, B, e, cur);
    reinterpret_cast<int2*>(b.agent)[e] = make_int2(st.r, st.c);
    reinterpret_cast<int2*>(b.target)[e] = make_int2(mz.goal & 0xffff, mz.goal >> 16);
    if (b.target_dirty) *b.target_dirty = 1;
    reinterpret_cast<int2*>(b.best_dir)[e] = best_dir_from_code((st.tab >> MAZE_TAB_CODE_SHIFT) & 7, st.r, st.c, mz.H, mz.W, mz.tor);
    b.reward[e] = reward;
    b.terminated[e] = (uint8_t)term;
    b.truncated[e] = (uint8_t)trunc;
    if (b.ep_return) b.ep_return[e] = (term | trunc) ? 0.0 : cur.ep_return;   // the agent's running return is the env's
    if (b.stats && n_steps) atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 3), n_steps);
    if (b.stats && n_episodes) {
        atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 0), n_episodes);
        if (n_wins) atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 1), n_wins);
        if (n_episodes - n_wins) atomicAdd(reinterpret_cast<unsigned long long*>(b.stats + 2), n_episodes - n_wins);
        if (b.stats_return) atomicAdd(b.stats_return, return_sum);
    }
}

int check_agent(maze_ctx* ctx, const maze_env_batch* b, const maze_q_agent* ag) {
    if (!ag) return maze_fail_arg(ctx, MAZE_E_NULL, "q agent");
    if (!ag->keys || !ag->q_a || !ag->overflow || !ag->eps_lut || !ag->gamma || !ag->slot || !ag->steps_done ||
        !ag->last_action || !ag->ep_return)
        return maze_fail_arg(ctx, MAZE_E_NULL, "q agent pointer");
    if (ag->capacity < 16 || (ag->capacity & (ag->capacity - 1)) || ag->capacity > (1ll << 31))
        return maze_fail_arg(ctx, MAZE_E_RANGE, "q agent capacity must be a power of two in [16, 2^31]");
    if (ag->envs_per_agent < 1 || ag->eps_len < 1)
        return maze_fail_arg(ctx, MAZE_E_RANGE, "q agent envs_per_agent / eps_len");
    if ((b->num_envs + ag->envs_per_agent - 1) / ag->envs_per_agent > (1 << 29))
        return maze_fail_arg(ctx, MAZE_E_RANGE, "q agent: more than 2^29 agents");
    if ((ag->u_tape || ag->a_tape) && (!ag->u_tape || !ag->a_tape || !ag->tape_pos || ag->u_len < 1 || ag->a_len < 1))
        return maze_fail_arg(ctx, MAZE_E_NULL, "q agent replay tapes need u_tape, a_tape, tape_pos and lengths");
    if (((uintptr_t)ag->q_a & 31) || ((uintptr_t)ag->q_b & 31) || ((uintptr_t)ag->keys & 7))
        return maze_fail_arg(ctx, MAZE_E_ALIGN, "q agent tables must be 32-byte aligned");
    return 0;
}

}  // namespace

extern "C" int maze_q_epsilon_lut(double initial_epsilon, double final_epsilon, double decay, double* out, int n) {
    if (!out) return MAZE_E_NULL;
    if (n < 1) return MAZE_E_RANGE;
    for (int i = 0; i < n; ++i)
        out[i] = final_epsilon + (initial_epsilon - final_epsilon) * std::exp(-1. * i / decay);   // q_agent.py:49
    return 0;
}

extern "C" int maze_q_act(maze_ctx* ctx, const maze_env_batch* b, const maze_q_agent* agent, uint8_t* actions, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (int rc = check_agent(ctx, b, agent)) return rc;
    if (!actions) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_q_act: actions");
    const int grid = (b->num_envs + Q_THREADS - 1) / Q_THREADS;
    maze_q_act_kernel<<<grid, Q_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *agent, actions);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_q_update(maze_ctx* ctx, const maze_env_batch* b, const maze_q_agent* agent, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (int rc = check_agent(ctx, b, agent)) return rc;
    const int grid = (b->num_envs + Q_THREADS - 1) / Q_THREADS;
    maze_q_update_kernel<<<grid, Q_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *agent);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_q_rollout(maze_ctx* ctx, const maze_env_batch* b, const maze_q_agent* agent, int k_steps,
                              uint32_t mode, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (int rc = check_agent(ctx, b, agent)) return rc;
    if (k_steps < 1) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_q_rollout: k_steps");
    if (mode & MAZE_STEP_WIN_QUEUE)
        return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_q_rollout: MAZE_STEP_WIN_QUEUE needs a generation launch between steps; use maze_step");
    const int grid = (b->num_envs + Q_THREADS - 1) / Q_THREADS;
    maze_q_rollout_kernel<<<grid, Q_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*b, *agent, k_steps, mode, step_luts(ctx));
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
