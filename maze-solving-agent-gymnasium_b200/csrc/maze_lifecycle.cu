// Maze-set lifecycle on the device: what the reference's trainers do on the host after a win.
//   - variable-size envs grow their maze by (4, 4) blocks per win until max_shape
//     (gymnasium_env/envs/simple_variable_maze_env.py:93-112, toroidal_variable_maze_env.py:118-136)
//   - the neural trainer switches the generator after 5 and 10 wins
//     (lib/trainers/off_policy_trainer.py:302-310: r-prim -> prim&kill -> dfs)
// One thread per queued slot rewrites the slot's meta record (H, W, generator id); maze_generate then
// draws the new maze from it.  With the regeneration queue of maze_step this keeps the whole
// "win -> harder maze" loop on the GPU.
#include "maze_common.cuh"

namespace {

__global__ void maze_curriculum_kernel(int32_t* __restrict__ meta, int32_t* __restrict__ wins, const int32_t* __restrict__ ids,
                                       const int32_t* __restrict__ count_dev, int n, int grow, int max_h, int max_w,
                                       int wins_a, int algo_a, int wins_b, int algo_b) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int count = count_dev ? min(*count_dev, n) : n;
    if (k >= count) return;
    const int m = ids ? ids[k] : k;
    int32_t* mm = meta + (size_t)m * MAZE_META_WORDS;
    const int w = wins[m] + 1;
    wins[m] = w;
    if (grow > 0) {   // shape <= max_shape compares tuples lexicographically in the reference; shapes are square there
        const int H = mm[MAZE_META_H] + grow, W = mm[MAZE_META_W] + grow;
        if (H <= max_h && W <= max_w) { mm[MAZE_META_H] = H; mm[MAZE_META_W] = W; }
    }
    int algo = -1;
    if (algo_b >= 0 && w >= wins_b) algo = algo_b;
    else if (algo_a >= 0 && w >= wins_a) algo = algo_a;
    if (algo >= 0) mm[MAZE_META_FLAGS] = (mm[MAZE_META_FLAGS] & ~0xff00) | (algo << 8);
}

}  // namespace

extern "C" int maze_curriculum(maze_ctx* ctx, int32_t* meta, int32_t* wins, const int32_t* ids, const int32_t* count_dev,
                               int n, int grow, int max_h, int max_w, int wins_a, int algo_a, int wins_b, int algo_b,
                               void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!meta || !wins) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_curriculum pointer");
    if (n <= 0) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_curriculum n");
    if (grow < 0 || (grow & 1)) return maze_fail_arg(ctx, MAZE_E_SHAPE, "maze_curriculum: grow must be even and >= 0 (block shapes stay odd)");
    if (algo_a > MAZE_ALGO_PRIMKILL || algo_b > MAZE_ALGO_PRIMKILL) return maze_fail_arg(ctx, MAZE_E_ALGO, "maze_curriculum generator id");
    maze_curriculum_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        meta, wins, ids, count_dev, n, grow, max_h, max_w, wins_a, algo_a, wins_b, algo_b);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Regeneration ahead of time.  Drawing a maze is sequential carving by one warp (0.5 - 1 ms for 81 x 81 blocks, 2.7 ms with
// the torus fields), so a step that waits for that step's regenerations costs milliseconds however few envs won.  The
// generator is a pure function of (seed, slot id, generation count, slot configuration), so the NEXT `depth` mazes of every
// slot can be drawn before they are needed: a shadow ring holds, for slot m and generation count g, the maze M(m, g) in
// entry g % depth; a win copies the entry of the live count over the live slot (13 KB) and queues the slot for a refill
// that runs on a side stream while the envs keep stepping.
//   ready_gen[j][m] == g + 1   <=>   entry j of slot m is complete and holds M(m, g)
// is the only thing the fast path trusts: a slot that wins more than `depth` times before a refill is published fails
// the test and is drawn in place on the stepping stream, exactly as without the ring.  Results therefore never depend on
// timing: every path installs the same maze M(m, count).  (depth 1 is not enough in practice: about 1 % of the episodes
// on 81 x 81 mazes end within the refill latency, a few slots per step of 262 144 envs -- every step took the slow path.)
//   maze_regen_swap     stepping stream: winners of the step -> copy (fast) or slow queue; every winner is appended to the
//                       refill queue of the current batch once (queued_tag)
//   maze_regen_prepare  side stream: for a queued slot with live count L the ring should hold M(m, L) .. M(m, L + depth - 1);
//                       entries that already do are left alone (so a refill never rewrites an entry that the stepping
//                       stream may be copying), the others are un-published, get their generation count and the shape /
//                       generator the curriculum gives that count, and go to the work queue of their ring index
//   maze_regen_publish  side stream, after maze_generate on the ring entries: ready_gen = entry's count (release)
// The live meta record is written after the slot's bytes (fence in between), so a prepare that already sees the new
// count can only start a refill of that entry after the copy has finished reading it.
namespace {

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr int SWAP_THREADS = 128;

// shadow_* are [depth, n, ...] arrays: ring entry j of slot m at index j * n + m
__global__ void __launch_bounds__(SWAP_THREADS)
maze_regen_swap_kernel(uint8_t* __restrict__ grids, uint8_t* __restrict__ table, int32_t* __restrict__ meta, const uint8_t* __restrict__ sh_grids,
                       const uint8_t* __restrict__ sh_table, const int32_t* __restrict__ sh_meta, const int32_t* __restrict__ ready_gen, int depth,
                       const int32_t* __restrict__ queue, const int32_t* __restrict__ queue_count, int n, int slot,
                       int32_t* __restrict__ refill_queue, int32_t* __restrict__ refill_count, int32_t* __restrict__ queued_tag, int batch,
                       int32_t* __restrict__ slow_queue, int32_t* __restrict__ slow_count, int32_t* __restrict__ stats,
                       int32_t* __restrict__ wins) {
    __shared__ int s_entry;
    const int count = min(*queue_count, n);
    for (int k = blockIdx.x; k < count; k += gridDim.x) {
        const int m = queue[k];
        if (threadIdx.x == 0) {
            const int live = meta[(size_t)m * MAZE_META_WORDS + MAZE_META_SPARE];
            const int j = live % depth;
            const bool valid = ld_acquire(ready_gen + (size_t)j * n + m) == live + 1;
            s_entry = valid ? j : -1;
            if (atomicExch(queued_tag + m, batch) != batch) refill_queue[atomicAdd(refill_count, 1)] = m;
            if (!valid) slow_queue[atomicAdd(slow_count, 1)] = m;
            if (stats) atomicAdd(stats + (valid ? 0 : 1), 1);
            if (valid && wins) wins[m] += 1;   // the curriculum's win count (slow slots: maze_curriculum on the slow queue)
        }
        __syncthreads();
        const int j = s_entry;
        if (j >= 0) {
            const size_t e = (size_t)j * n + m;
            const uint4* sg = reinterpret_cast<const uint4*>(sh_grids + e * slot);
            const uint4* stb = reinterpret_cast<const uint4*>(sh_table + e * slot);
            uint4* lg = reinterpret_cast<uint4*>(grids + (size_t)m * slot);
            uint4* lt = reinterpret_cast<uint4*>(table + (size_t)m * slot);
            for (int i = threadIdx.x; i < slot / 16; i += SWAP_THREADS) {
                lg[i] = __ldcg(sg + i);
                lt[i] = __ldcg(stb + i);
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x < MAZE_META_WORDS) meta[(size_t)m * MAZE_META_WORDS + threadIdx.x] = __ldcg(sh_meta + e * MAZE_META_WORDS + threadIdx.x);
        }
        __syncthreads();
    }
}

// The configuration (shape, generator) of generation count g follows from a snapshot of the slot taken while nothing was
// in flight (base_meta / base_wins, at ring construction) and the curriculum's rule applied g - base count + 1 times, in
// closed form: maze_curriculum_kernel grows both sides together while both fit and picks the generator from the win count.
// Reading the LIVE record's shape here instead would race with maze_regen_swap rewriting it word by word.
struct CurriculumRule {
    int grow, max_h, max_w, wins_a, algo_a, wins_b, algo_b;
};

// work_queue [depth, n], work_count [depth]
__global__ void maze_regen_prepare_kernel(const int32_t* __restrict__ meta, const int32_t* __restrict__ base_meta, const int32_t* __restrict__ base_wins,
                                          CurriculumRule rule, int32_t* __restrict__ sh_meta, int32_t* __restrict__ ready_gen, int depth,
                                          const int32_t* __restrict__ refill_queue, const int32_t* __restrict__ refill_count, int n,
                                          int32_t* __restrict__ work_queue, int32_t* __restrict__ work_count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(*refill_count, n)) return;
    const int m = refill_queue[k];
    const int live = *reinterpret_cast<const volatile int32_t*>(meta + (size_t)m * MAZE_META_WORDS + MAZE_META_SPARE);
    const int32_t* bm = base_meta + (size_t)m * MAZE_META_WORDS;
    const int bH = bm[MAZE_META_H], bW = bm[MAZE_META_W], bflags = bm[MAZE_META_FLAGS], bcount = bm[MAZE_META_SPARE];
    const int bwins = base_wins ? base_wins[m] : 0;
    const int room = rule.grow > 0 ? max(0, min((rule.max_h - bH) / rule.grow, (rule.max_w - bW) / rule.grow)) : 0;
    for (int g = live; g < live + depth; ++g) {
        const int j = g % depth;
        int32_t* rg = ready_gen + (size_t)j * n + m;
        if (ld_acquire(rg) == g + 1) continue;   // already holds M(m, g): nothing to draw, and the stepping stream may be reading it
        st_release(rg, 0);
        const int steps = g - bcount + 1;        // curriculum steps between the snapshot and the regeneration that draws M(m, g)
        const int grown = rule.grow * min(max(steps, 0), room);
        const int w = bwins + steps;
        int algo = -1;
        if (rule.algo_b >= 0 && w >= rule.wins_b) algo = rule.algo_b;
        else if (rule.algo_a >= 0 && w >= rule.wins_a) algo = rule.algo_a;
        int32_t* sm = sh_meta + ((size_t)j * n + m) * MAZE_META_WORDS;
        sm[MAZE_META_H] = bH + grown;
        sm[MAZE_META_W] = bW + grown;
        sm[MAZE_META_FLAGS] = algo >= 0 ? ((bflags & ~0xff00) | (algo << 8)) : bflags;
        sm[MAZE_META_SPARE] = g;
        work_queue[(size_t)j * n + atomicAdd(work_count + j, 1)] = m;
    }
}

__global__ void maze_regen_publish_kernel(const int32_t* __restrict__ sh_meta, int32_t* __restrict__ ready_gen, int depth,
                                          const int32_t* __restrict__ work_queue, const int32_t* __restrict__ work_count, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (k >= min(work_count[j], n)) return;
    const int m = work_queue[(size_t)j * n + k];
    __threadfence();
    st_release(ready_gen + (size_t)j * n + m, sh_meta[((size_t)j * n + m) * MAZE_META_WORDS + MAZE_META_SPARE]);
}

}  // namespace

extern "C" int maze_regen_swap(maze_ctx* ctx, uint8_t* grids, uint8_t* table, int32_t* meta, const uint8_t* shadow_grids, const uint8_t* shadow_table,
                               const int32_t* shadow_meta, const int32_t* ready_gen, int depth, const int32_t* queue, const int32_t* queue_count,
                               int n, int slot, int32_t* refill_queue, int32_t* refill_count, int32_t* queued_tag, int batch, int32_t* slow_queue,
                               int32_t* slow_count, int32_t* stats, int32_t* wins, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!grids || !table || !meta || !shadow_grids || !shadow_table || !shadow_meta || !ready_gen || !queue || !queue_count || !refill_queue ||
        !refill_count || !queued_tag || !slow_queue || !slow_count)
        return maze_fail_arg(ctx, MAZE_E_NULL, "maze_regen_swap pointer");
    if (n <= 0 || slot <= 0 || (slot % 16) != 0 || depth < 1 || depth > 8)
        return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_regen_swap: n / slot (a multiple of 16) / depth (1..8)");
    if (((uintptr_t)grids | (uintptr_t)table | (uintptr_t)shadow_grids | (uintptr_t)shadow_table) & 15)
        return maze_fail_arg(ctx, MAZE_E_ALIGN, "maze_regen_swap: 16-byte aligned grids / tables");
    const int sms = ctx->num_sms > 0 ? ctx->num_sms : 148;
    const int grid = n < 4 * sms ? n : 4 * sms;
    maze_regen_swap_kernel<<<grid, SWAP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(grids, table, meta, shadow_grids, shadow_table, shadow_meta, ready_gen,
                                                                                       depth, queue, queue_count, n, slot, refill_queue, refill_count,
                                                                                       queued_tag, batch, slow_queue, slow_count, stats, wins);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_regen_prepare(maze_ctx* ctx, const int32_t* meta, const int32_t* base_meta, const int32_t* base_wins, int grow, int max_h,
                                  int max_w, int wins_a, int algo_a, int wins_b, int algo_b, int32_t* shadow_meta, int32_t* ready_gen, int depth,
                                  const int32_t* refill_queue, const int32_t* refill_count, int n, int32_t* work_queue, int32_t* work_count,
                                  void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!meta || !base_meta || !shadow_meta || !ready_gen || !refill_queue || !refill_count || !work_queue || !work_count)
        return maze_fail_arg(ctx, MAZE_E_NULL, "maze_regen_prepare pointer");
    if (n <= 0 || depth < 1 || depth > 8) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_regen_prepare n / depth");
    if (grow < 0 || (grow & 1)) return maze_fail_arg(ctx, MAZE_E_SHAPE, "maze_regen_prepare: grow must be even and >= 0");
    if (algo_a > MAZE_ALGO_PRIMKILL || algo_b > MAZE_ALGO_PRIMKILL) return maze_fail_arg(ctx, MAZE_E_ALGO, "maze_regen_prepare generator id");
    const CurriculumRule rule{grow, max_h, max_w, wins_a, algo_a, wins_b, algo_b};
    maze_regen_prepare_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(meta, base_meta, base_wins, rule, shadow_meta, ready_gen, depth,
                                                                                            refill_queue, refill_count, n, work_queue, work_count);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_regen_publish(maze_ctx* ctx, const int32_t* shadow_meta, int32_t* ready_gen, int depth, const int32_t* work_queue,
                                  const int32_t* work_count, int n, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!shadow_meta || !ready_gen || !work_queue || !work_count) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_regen_publish pointer");
    if (n <= 0 || depth < 1 || depth > 8) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_regen_publish n / depth");
    maze_regen_publish_kernel<<<dim3((n + 255) / 256, depth), 256, 0, static_cast<cudaStream_t>(stream)>>>(shadow_meta, ready_gen, depth, work_queue,
                                                                                                         work_count, n);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
