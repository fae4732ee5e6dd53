// Micro-benchmark of the access pattern that bounds maze_step: per launch, every env streams
// its coalesced words (13 B read, 34 + 8 B written = the step kernel's 47 B of streams plus the 8-byte state store)
// and a fraction of the envs do ONE 2-byte read-modify-write at a scattered address of a large array (the visit
// counter of the block stepped onto).  No maze logic: what remains is what HBM3e + L2 give this pattern.
//   pattern 0  uniform random element of the array
//   pattern 1  cell-major [slot, B] addressing: element = cell * B + env with a random cell per env and launch
//              (neighbouring envs land in different 64-byte atoms, like decorrelated agents)
//   pattern 2  as 1, but the 32 envs of a warp share the cell (agents that still walk together: one atom per warp)
#include "maze_common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256)
bench_scatter_kernel(const uint64_t* __restrict__ state_in, uint64_t* __restrict__ state_out, const int32_t* __restrict__ env_maze,
                     const uint8_t* __restrict__ actions, int2* __restrict__ agent, int2* __restrict__ best_dir, double* __restrict__ reward,
                     uint8_t* __restrict__ term, uint8_t* __restrict__ trunc, uint16_t* __restrict__ arr, long long n_elems, int B, int slot,
                     int pattern, uint32_t rmw_per_1024, uint32_t launch, int streams) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= B) return;
    uint64_t s = 0;
    int m = 0, a = 0;
    if (streams) {
        s = state_in[e];
        m = env_maze[e];
        a = __ldcs(actions + e);
    }
    const uint32_t h = mix32((uint32_t)e * 0x9E3779B9u + launch * 0x85EBCA6Bu);
    uint32_t v = 0;
    if ((h & 1023u) < rmw_per_1024) {
        long long idx;
        if (pattern == 0) {
            idx = (long long)(((unsigned long long)mix32(h + 1) << 32 | mix32(h + 2)) % (unsigned long long)n_elems);
        } else {
            const uint32_t hc = pattern == 2 ? mix32(((uint32_t)e >> 5) * 0x9E3779B9u + launch) : mix32(h + 3);
            idx = (long long)(hc % (uint32_t)slot) * B + e;
        }
        v = arr[idx];
        arr[idx] = (uint16_t)(v + 1 + (uint32_t)(s & 1));
    }
    if (streams) {
        state_out[e] = s + v + (uint64_t)(m + a);
        __stcs(agent + e, make_int2((int)(s & 0xff), (int)((s >> 8) & 0xff)));
        __stcs(best_dir + e, make_int2(a, m));
        __stcs(reward + e, (double)v);
        __stcs(term + e, (uint8_t)(v & 1));
        __stcs(trunc + e, (uint8_t)0);
    }
}

}  // namespace

// arr: uint16 [n_elems] (n_elems >= slot * B for patterns 1, 2).  rmw_per_1024: envs out of 1024 that touch the array
// per launch (the step kernel: ~410).  streams != 0 adds the coalesced reads / writes of maze_step.
extern "C" int maze_bench_scatter_rmw(maze_ctx* ctx, const maze_env_batch* b, const uint8_t* actions, uint16_t* arr, int64_t n_elems, int pattern,
                                      int rmw_per_1024, uint32_t launch, int streams, void* stream) {
    if (!ctx) return MAZE_E_NULL;
    if (int rc = maze_check_batch(ctx, b)) return rc;
    if (!arr || !actions || n_elems < 1 || pattern < 0 || pattern > 2) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_bench_scatter_rmw arguments");
    if (pattern != 0 && n_elems < (int64_t)b->slot * b->num_envs) return maze_fail_arg(ctx, MAZE_E_RANGE, "maze_bench_scatter_rmw: array smaller than slot * B");
    bench_scatter_kernel<<<(b->num_envs + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        b->state, b->state, b->env_maze, actions, reinterpret_cast<int2*>(b->agent), reinterpret_cast<int2*>(b->best_dir), b->reward, b->terminated,
        b->truncated, arr, (long long)n_elems, b->num_envs, b->slot, pattern, (uint32_t)rmw_per_1024, launch, streams);
    MAZE_CHECK(cudaGetLastError());
    return 0;
}
