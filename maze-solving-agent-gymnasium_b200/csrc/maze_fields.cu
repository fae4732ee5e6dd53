// maze_fields: step table + step budget for mazes whose block grids the caller uploaded.
#include "maze_fields.cuh"

namespace {

__global__ void __launch_bounds__(FIELD_THREADS)
maze_fields_kernel(const uint8_t* __restrict__ grids, int32_t* __restrict__ meta, uint8_t* __restrict__ table,
                   const int32_t* __restrict__ ids, int slot) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int m = ids ? ids[blockIdx.x] : blockIdx.x;
    int32_t* mm = meta + (size_t)m * MAZE_META_WORDS;
    const int H = mm[MAZE_META_H], W = mm[MAZE_META_W];
    const int start = mm[MAZE_META_START], goal = mm[MAZE_META_GOAL];
    const bool tor = (mm[MAZE_META_FLAGS] & MAZE_FLAG_TOROIDAL) != 0;
    const int hw = H * W;
    FieldSmem f = field_smem_carve(smem, hw);
    const uint8_t* g = grids + (size_t)m * slot;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) f.grid[i] = g[i];
    __syncthreads();
    const int gr = goal & 0xffff, gc = goal >> 16;
    block_bfs(f, H, W, tor, gr * W + gc);
    encode_step_table(f, H, W, tor, gr, gc, table + (size_t)m * slot);
    if (threadIdx.x == 0) {
        const int d = f.dist[(start & 0xffff) * W + (start >> 16)];
        const int sol_len = (d == DIST_INF) ? 0 : d + 1;
        mm[MAZE_META_SOL_LEN] = sol_len;
        mm[MAZE_META_MAX_STEPS] = max_steps_budget(H, W, sol_len);
    }
}

}  // namespace

extern "C" int maze_fields(maze_ctx* ctx, const uint8_t* grids, int32_t* meta, uint8_t* table,
                           const int32_t* ids, int n, int slot, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!grids || !meta || !table) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_fields pointer");
    if (n <= 0 || slot <= 0) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_fields n / slot");
    if (slot > MAZE_GEN_MAX_DIM * MAZE_GEN_MAX_DIM + 1)
        return maze_fail_arg(ctx, MAZE_E_SHAPE, "maze_fields: slot exceeds MAZE_GEN_MAX_DIM^2");
    const size_t smem = field_smem_bytes(slot);
    MAZE_CHECK(cudaFuncSetAttribute(maze_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    maze_fields_kernel<<<n, FIELD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(grids, meta, table, ids, slot);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
