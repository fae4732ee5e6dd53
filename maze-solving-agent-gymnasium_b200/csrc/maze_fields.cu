// maze_fields: step table + step budget for mazes whose block grids the caller uploaded.
#include "maze_fields.cuh"

namespace {

__global__ void __launch_bounds__(FIELD_THREADS)
maze_fields_kernel(const uint8_t* __restrict__ grids, int32_t* __restrict__ meta, uint8_t* __restrict__ table,
                   const int32_t* __restrict__ ids, int slot) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int m = ids ? ids[blockIdx.x] : blockIdx.x;
    const int32_t* mm = meta + (size_t)m * MAZE_META_WORDS;
    FieldSmem f = field_smem_carve(smem, mm[MAZE_META_H] * mm[MAZE_META_W]);
    fields_of_slot(f, grids, meta, table, m, slot);
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
