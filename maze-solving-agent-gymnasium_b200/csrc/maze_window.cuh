// One window row of one env: the three 15-bit masks of lib/maze_handler.py:82-99 ([maze == 0, maze == 1 (goal
// excluded), non_visited]) for window row `row` (0..14) of the 15 x 15 crop around the agent (extract_submaze :4-54
// clamped on bordered mazes, extract_submaze_toroid :56-80 wrapped on the torus).  Bit j = window column j.
// One lane per row: 15 table bytes and 15 visit words are loaded first, then folded into the masks in registers.
// Shared by the bit-packed replay encode (maze_dqn.cu) and the float window of the -v1 observation (maze_obs.cu).
#pragma once
#include "maze_env.cuh"

// The bitmap path of window_row_masks (bordered mazes with maze_env_batch.visit_bits) in two halves, so that a kernel
// can issue the loads of several envs before it folds any of them (maze_dqn_push_kernel).
struct WindowRowRaw {
    uint32_t tword[5];   // the row's 15 table bytes inside five aligned words
    uint32_t b_lo, b_hi; // the two bitmap words its 15 "visited" bits lie in
    int first, c0;       // table index of window column 0; its maze column
};

// requires: !mz.tor, b.visit_bits, mz.H >= WIN, mz.W >= WIN, row < WIN
__device__ __forceinline__ WindowRowRaw window_row_load_bits(const maze_env_batch& b, int e, const EnvState& st, const MazeView& mz, int row) {
    constexpr int WIN = MAZE_WINDOW;
    const int r0 = min(max(st.r - WIN / 2, 0), mz.H - WIN), c0 = min(max(st.c - WIN / 2, 0), mz.W - WIN);
    const int rr = r0 + row;
    WindowRowRaw w;
    w.first = rr * mz.W + c0;
    w.c0 = c0;
    const int off = w.first & 3;
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(mz.tab + (w.first - off));   // slots are 16-byte aligned and padded
    const uint32_t* brow = b.visit_bits + (size_t)e * b.visit_bits_stride + rr * b.visit_bits_pitch + (c0 >> 5);
#pragma unroll
    for (int k = 0; k < 5; ++k) w.tword[k] = (4 * k - off < WIN) ? __ldg(tw + k) : 0u;
    w.b_lo = brow[0];
    w.b_hi = ((c0 & 31) + WIN > 32) ? brow[1] : 0u;
    return w;
}

__device__ __forceinline__ void window_row_fold_bits(const WindowRowRaw& w, const MazeView& mz, unsigned& m0, unsigned& m1, unsigned& m2) {
    constexpr int WIN = MAZE_WINDOW;
    const int off = w.first & 3;
    const unsigned seen = __funnelshift_r(w.b_lo, w.b_hi, w.c0 & 31) & ((1u << WIN) - 1u);
    unsigned openm = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jt = 4 * k + q - off;               // window column of table byte q of word k
            if (jt >= 0 && jt < WIN) openm |= ((w.tword[k] >> (8 * q)) & MAZE_TAB_OPEN) << jt;
        }
    const unsigned all = (1u << WIN) - 1u;
    const int start_idx = (mz.start & 0xffff) * mz.W + (mz.start >> 16);
    const int goal_idx = (mz.goal & 0xffff) * mz.W + (mz.goal >> 16);
    const int gj = goal_idx - w.first, sj = start_idx - w.first;   // goal / start inside this row of the window?
    const unsigned goal_bit = (gj >= 0 && gj < WIN) ? 1u << gj : 0u, start_bit = (sj >= 0 && sj < WIN) ? 1u << sj : 0u;
    m0 = ~openm & all;
    m1 = openm & ~goal_bit;
    m2 = openm & ~start_bit & ~seen;
}

__device__ __forceinline__ void window_row_masks(const maze_env_batch& b, int e, const EnvState& st, const MazeView& mz, int row,
                                                 unsigned& m0, unsigned& m1, unsigned& m2) {
    constexpr int WIN = MAZE_WINDOW;
    m0 = m1 = m2 = 0u;
    const int H = mz.H, W = mz.W;
    if (H < WIN || W < WIN || row >= WIN) return;   // no 15 x 15 crop exists (the reference cannot build one either)
    const int start_idx = (mz.start & 0xffff) * W + (mz.start >> 16);
    const int goal_idx = (mz.goal & 0xffff) * W + (mz.goal >> 16);
    int r0 = st.r - WIN / 2, c0 = st.c - WIN / 2;
    if (!mz.tor) {
        r0 = min(max(r0, 0), H - WIN);
        c0 = min(max(c0, 0), W - WIN);   // the reference clamps with len(maze) (maze_handler.py:21-29: square mazes only); W keeps a non-square slot in bounds
    }
    int rr = r0 + row;
    if (mz.tor) rr = rr < 0 ? rr + H : (rr >= H ? rr - H : rr);
    const uint8_t* trow = mz.tab + rr * W;
    const uint16_t* vbase = b.visits + (size_t)e * b.visit_env_stride;
    const int wt = (W + 3) >> 2;
    if (b.visit_bits && !mz.tor) {
        // Bitmap path (bordered mazes): the row's "visited" bits are 15 consecutive bits of one bitmap row -- two words -- and
        // its table bytes five aligned 32-bit words: 7 loads per lane, and the 15 lanes of an env read 15 consecutive bitmap
        // rows (180 contiguous bytes at 81 x 81: three 64-byte DRAM atoms where the counter tiles cost up to 25 sectors).
        const WindowRowRaw raw = window_row_load_bits(b, e, st, mz, row);
        window_row_fold_bits(raw, mz, m0, m1, m2);
        return;
    }
    if (!mz.tor && b.visit_tiled && b.visit_cell_stride == 1) {
        // Fast path (bordered mazes, the tiled env-major visit array of the -v1 default): the row's 15 table bytes come
        // as five aligned 32-bit words, its visit words as one 8-byte load per 4 x 4 tile the row crosses -- 10 loads per
        // lane instead of 30, and a third of the L1 sector look-ups, which is what bounded the one-load-per-block form
        // (ncu profiles/r02i_obs_details.txt: 10.5 sectors per request, long-scoreboard stalls).
        const int first = rr * W + c0, off = first & 3;
        const uint32_t* tw = reinterpret_cast<const uint32_t*>(mz.tab + (first - off));   // slots are 16-byte aligned and padded
        const int tc0 = c0 >> 2;
        const unsigned long long* vt = reinterpret_cast<const unsigned long long*>(vbase + ((((rr >> 2) * wt) << 4) | ((rr & 3) << 2)));
        uint32_t tword[5];
        unsigned long long vword[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            tword[k] = (4 * k - off < WIN) ? __ldg(tw + k) : 0u;
            vword[k] = (tc0 + k < wt) ? vt[(size_t)(tc0 + k) * 4] : 0ull;   // tile (rr >> 2, tc0 + k): 16 entries = 4 x 8 bytes further
        }
        unsigned openm = 0, seen = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int jt = 4 * k + q - off;               // window column of table byte q of word k
                if (jt >= 0 && jt < WIN) openm |= ((tword[k] >> (8 * q)) & MAZE_TAB_OPEN) << jt;
                const int jv = 4 * (tc0 + k) + q - c0;        // window column of visit entry q of tile k
                const unsigned v = (unsigned)(vword[k] >> (16 * q)) & 0xffffu;
                if (jv >= 0 && jv < WIN) seen |= (((int)(v >> 8) == st.epoch && (v & 0xffu) != 0) ? 1u : 0u) << jv;
            }
        const unsigned all = (1u << WIN) - 1u;
        const int gj = goal_idx - first, sj = start_idx - first;   // goal / start inside this row of the window?
        const unsigned goal_bit = (gj >= 0 && gj < WIN) ? 1u << gj : 0u, start_bit = (sj >= 0 && sj < WIN) ? 1u << sj : 0u;
        m0 = ~openm & all;
        m1 = openm & ~goal_bit;
        m2 = openm & ~start_bit & ~seen;
        return;
    }
    const int vrow = b.visit_tiled ? ((((rr >> 2) * wt) << 4) | ((rr & 3) << 2)) : rr * W;
    int tb[WIN];
    unsigned vis[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) {   // all loads of the row first
        int cc = c0 + j;
        if (mz.tor) cc = cc < 0 ? cc + W : (cc >= W ? cc - W : cc);
        tb[j] = __ldg(trow + cc);
        if (b.visit_bits) {   // wrapped window on the torus: one bitmap word per block (the word is shared by up to 15 columns: L1)
            const uint32_t wv = b.visit_bits[(size_t)e * b.visit_bits_stride + rr * b.visit_bits_pitch + (cc >> 5)];
            vis[j] = ((wv >> (cc & 31)) & 1u) ? (unsigned)((st.epoch << 8) | 1) : 0u;   // in the counters' terms: seen this episode
        } else {
            const int vi = b.visit_tiled ? vrow + (((cc >> 2) << 4) | (cc & 3)) : vrow + cc;
            vis[j] = vbase[(size_t)vi * b.visit_cell_stride];
        }
    }
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
        int cc = c0 + j;
        if (mz.tor) cc = cc < 0 ? cc + W : (cc >= W ? cc - W : cc);
        const int idx = rr * W + cc;
        const bool open = (tb[j] & MAZE_TAB_OPEN) != 0;
        // non_visited (base_maze_env.py:148-149,183-184): open, not the start, no visit in this episode
        const bool fresh = open && idx != start_idx && !((int)(vis[j] >> 8) == st.epoch && (vis[j] & 0xffu) != 0);
        m0 |= (open ? 0u : 1u) << j;
        m1 |= ((open && idx != goal_idx) ? 1u : 0u) << j;
        m2 |= (fresh ? 1u : 0u) << j;
    }
}
