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
