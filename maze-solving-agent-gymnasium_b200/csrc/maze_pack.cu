// Packed maze-set formats (wire / on-disk interchange) and the auto-encoder's channel encode (sm_100a).
//
//   MAZE_PACK_BITMAP  one bit per block, row-major, LSB first: lossless for any block grid and the
//                     only faithful packed form of a toroidal maze (representation
//                     note: across the seam passage blocks link to passage blocks)
//   MAZE_PACK_WALLS   four wall bits per logical cell (N = 1, E = 2, S = 4, W = 8), two cells per byte,
//                     low nibble first: euclidean generator output (every logical cell open)
//   maze_collection_encode   int32 [n, 3, H, W] = [wall, tile (== 1), non_visited] of
//                     generate_collection_of_mazes (lib/maze_generation.py:236-242)
// All three are pure byte shuffles: one thread per packed byte / per 8 blocks, coalesced both ways.
#include "maze_common.cuh"

namespace {

constexpr int PACK_THREADS = 256;

struct PackArgs {
    const int32_t* meta;
    const int32_t* ids;
    int n, slot, stride;
};

__device__ __forceinline__ int slot_of(const PackArgs& a, int k) { return a.ids ? a.ids[k] : k; }

__global__ void __launch_bounds__(PACK_THREADS)
maze_pack_bitmap_kernel(PackArgs a, const uint8_t* __restrict__ grids, uint8_t* __restrict__ packed) {
    const int k = blockIdx.x, t = blockIdx.y * PACK_THREADS + threadIdx.x;
    if (t >= a.stride) return;
    const int m = slot_of(a, k);
    const int hw = a.meta[(size_t)m * MAZE_META_WORDS + MAZE_META_H] * a.meta[(size_t)m * MAZE_META_WORDS + MAZE_META_W];
    unsigned bits = 0;
    if (8 * t < hw) {
        // slot is a multiple of 16 and 8 t + 8 <= roundup8(hw) <= slot: one aligned 8-byte load
        const uint64_t v = *reinterpret_cast<const uint64_t*>(grids + (size_t)m * a.slot + 8 * t);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (8 * t + i < hw && ((v >> (8 * i)) & 0xff) != 0) bits |= 1u << i;
    }
    packed[(size_t)k * a.stride + t] = (uint8_t)bits;
}

__global__ void __launch_bounds__(PACK_THREADS)
maze_unpack_bitmap_kernel(PackArgs a, const uint8_t* __restrict__ packed, uint8_t* __restrict__ grids) {
    const int k = blockIdx.x, t = blockIdx.y * PACK_THREADS + threadIdx.x;
    if (8 * t >= a.slot) return;
    const int m = slot_of(a, k);
    const int32_t* mm = a.meta + (size_t)m * MAZE_META_WORDS;
    const int W = mm[MAZE_META_W], hw = mm[MAZE_META_H] * W;
    const int goal = (mm[MAZE_META_GOAL] & 0xffff) * W + (mm[MAZE_META_GOAL] >> 16);
    const unsigned bits = (8 * t < hw && t < a.stride) ? packed[(size_t)k * a.stride + t] : 0u;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = 8 * t + i;
        if (idx < hw && (bits >> i & 1u)) v |= (uint64_t)(idx == goal ? 2 : 1) << (8 * i);   // goal block carries 2 (lib/maze_generation.py:33)
    }
    *reinterpret_cast<uint64_t*>(grids + (size_t)m * a.slot + 8 * t) = v;   // also zeroes the slot's tail
}

// one thread per byte = two logical cells
__global__ void __launch_bounds__(PACK_THREADS)
maze_pack_walls_kernel(PackArgs a, const uint8_t* __restrict__ grids, uint8_t* __restrict__ packed) {
    const int k = blockIdx.x, t = blockIdx.y * PACK_THREADS + threadIdx.x;
    if (t >= a.stride) return;
    const int m = slot_of(a, k);
    const int32_t* mm = a.meta + (size_t)m * MAZE_META_WORDS;
    const int H = mm[MAZE_META_H], W = mm[MAZE_META_W], nr = (H - 1) / 2, nc = (W - 1) / 2;
    const uint8_t* g = grids + (size_t)m * a.slot;
    unsigned out = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ci = 2 * t + half;
        if (ci >= nr * nc) continue;
        const int b = (2 * (ci / nc) + 1) * W + 2 * (ci % nc) + 1;
        const unsigned nib = (g[b - W] == 0 ? 1u : 0u) | (g[b + 1] == 0 ? 2u : 0u) | (g[b + W] == 0 ? 4u : 0u) | (g[b - 1] == 0 ? 8u : 0u);
        out |= nib << (4 * half);
    }
    packed[(size_t)k * a.stride + t] = (uint8_t)out;
}

// one thread per block of the grid: cells open, pillars / border closed, a passage block open iff the
// cell on its north / west side has no S / E wall
__global__ void __launch_bounds__(PACK_THREADS)
maze_unpack_walls_kernel(PackArgs a, const uint8_t* __restrict__ packed, uint8_t* __restrict__ grids) {
    const int k = blockIdx.x, idx = blockIdx.y * PACK_THREADS + threadIdx.x;
    if (idx >= a.slot) return;
    const int m = slot_of(a, k);
    const int32_t* mm = a.meta + (size_t)m * MAZE_META_WORDS;
    const int H = mm[MAZE_META_H], W = mm[MAZE_META_W], nc = (W - 1) / 2;
    const int goal = (mm[MAZE_META_GOAL] & 0xffff) * W + (mm[MAZE_META_GOAL] >> 16);
    int v = 0;
    if (idx < H * W) {
        const int r = idx / W, c = idx % W;
        auto nibble = [&](int i, int j) {
            const int ci = i * nc + j;
            return (packed[(size_t)k * a.stride + (ci >> 1)] >> (4 * (ci & 1))) & 0xf;
        };
        if (r > 0 && r < H - 1 && c > 0 && c < W - 1) {
            if ((r & 1) && (c & 1)) v = 1;
            else if ((r & 1) && !(c & 1)) v = (nibble((r - 1) >> 1, (c - 2) >> 1) & 2) ? 0 : 1;   // east passage of the cell on the left
            else if (!(r & 1) && (c & 1)) v = (nibble((r - 2) >> 1, (c - 1) >> 1) & 4) ? 0 : 1;   // south passage of the cell above
        }
        if (v && idx == goal) v = 2;
    }
    grids[(size_t)m * a.slot + idx] = (uint8_t)v;
}

__global__ void __launch_bounds__(PACK_THREADS)
maze_collection_encode_kernel(PackArgs a, const uint8_t* __restrict__ grids, int hw, int32_t* __restrict__ out) {
    const int k = blockIdx.x, idx = blockIdx.y * PACK_THREADS + threadIdx.x;
    if (idx >= hw) return;
    const int m = slot_of(a, k);
    const int32_t* mm = a.meta + (size_t)m * MAZE_META_WORDS;
    const int start = (mm[MAZE_META_START] & 0xffff) * mm[MAZE_META_W] + (mm[MAZE_META_START] >> 16);
    const int g = grids[(size_t)m * a.slot + idx];
    int32_t* o = out + (size_t)k * 3 * hw + idx;
    __stcs(o, g == 0 ? 1 : 0);                            // wall_mask   :237
    __stcs(o + hw, g == 1 ? 1 : 0);                       // tile_mask   :236 (the goal, 2, is neither)
    __stcs(o + 2 * hw, (g != 0 && idx != start) ? 1 : 0);   // non_visited :238-239
}

int check_pack(maze_ctx* ctx, const void* grids, const void* meta, const void* packed, int n, int slot, int format, int stride,
               const char* who) {
    if (!grids || !meta || !packed) return maze_fail_arg(ctx, MAZE_E_NULL, who);
    if (n <= 0 || slot <= 0 || (slot & 15) || stride <= 0) return maze_fail_arg(ctx, MAZE_E_RANGE, who);
    if (format != MAZE_PACK_BITMAP && format != MAZE_PACK_WALLS) return maze_fail_arg(ctx, MAZE_E_RANGE, who);
    if ((uintptr_t)grids & 15) return maze_fail_arg(ctx, MAZE_E_ALIGN, who);
    return 0;
}

}  // namespace

extern "C" int maze_pack(maze_ctx* ctx, const uint8_t* grids, const int32_t* meta, const int32_t* ids, int n, int slot,
                         int format, uint8_t* packed, int packed_stride, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = check_pack(ctx, grids, meta, packed, n, slot, format, packed_stride, "maze_pack")) return rc;
    const PackArgs a = {meta, ids, n, slot, packed_stride};
    const dim3 grid(n, (packed_stride + PACK_THREADS - 1) / PACK_THREADS);   // maze index on x: n may exceed 65 535
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (format == MAZE_PACK_BITMAP) {
        if (packed_stride > slot / 8) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_pack: bitmap stride exceeds slot / 8");
        maze_pack_bitmap_kernel<<<grid, PACK_THREADS, 0, st>>>(a, grids, packed);
    } else {
        maze_pack_walls_kernel<<<grid, PACK_THREADS, 0, st>>>(a, grids, packed);
    }
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_unpack(maze_ctx* ctx, const uint8_t* packed, int packed_stride, int format, uint8_t* grids,
                           const int32_t* meta, const int32_t* ids, int n, int slot, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = check_pack(ctx, grids, meta, packed, n, slot, format, packed_stride, "maze_unpack")) return rc;
    const PackArgs a = {meta, ids, n, slot, packed_stride};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (format == MAZE_PACK_BITMAP)
        maze_unpack_bitmap_kernel<<<dim3(n, (slot / 8 + PACK_THREADS - 1) / PACK_THREADS), PACK_THREADS, 0, st>>>(a, packed, grids);
    else
        maze_unpack_walls_kernel<<<dim3(n, (slot + PACK_THREADS - 1) / PACK_THREADS), PACK_THREADS, 0, st>>>(a, packed, grids);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int maze_collection_encode(maze_ctx* ctx, const uint8_t* grids, const int32_t* meta, const int32_t* ids, int n,
                                      int slot, int h, int w, int32_t* out, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (!grids || !meta || !out) return maze_fail_arg(ctx, MAZE_E_NULL, "maze_collection_encode pointer");
    if (n <= 0 || slot <= 0 || h <= 0 || w <= 0 || h * w > slot) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_collection_encode n / slot / shape");
    const PackArgs a = {meta, ids, n, slot, 0};
    maze_collection_encode_kernel<<<dim3(n, (h * w + PACK_THREADS - 1) / PACK_THREADS), PACK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, grids, h * w, out);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
