// Which ring slot does draw number `draw` give to batch position k?  Shared by maze_dqn_sample and
// maze_dqn_sample_packed so that the two return the same transitions.
//   with replacement:     uniform in [0, filled), Philox keyed by (seed, draw, k)
//   without replacement:  position k of a pseudo-random permutation of [0, filled) -- a 4-round Feistel network over the
//                         next even power of two, keyed by Philox(seed, draw), cycle-walked into range -- i.e. the first
//                         n images are n distinct slots, which is random.sample(memory, n) (lib/replay_memory.py:20-21)
#pragma once
#include "maze_common.cuh"

__device__ __forceinline__ uint32_t replay_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ size_t replay_slot(const maze_replay& r, unsigned long long filled, int n, unsigned long long seed,
                                              unsigned long long draw, int k) {
    Philox rng;
    if (!r.without_replacement || (unsigned long long)n > filled || filled > 0x7fffffffull) {
        rng.init(seed, draw, (uint32_t)k);
        rng.refill();
        const unsigned long long u = ((unsigned long long)rng.o0 << 32) | rng.o1;
        return (size_t)__umul64hi(u, filled);   // uniform in [0, filled)
    }
    rng.init(seed, draw, 0xffffffffu);          // one key block per draw, the same for every k
    rng.refill();
    const uint32_t key[4] = {rng.o0, rng.o1, rng.o2, rng.o3};
    int half = 1;
    while ((1ull << (2 * half)) < filled) ++half;
    const uint32_t mask = (1u << half) - 1u;
    uint32_t x = (uint32_t)k;
    do {   // cycle walking: the domain is at most 4 x filled, so a few iterations on average
        uint32_t L = x >> half, R = x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t F = replay_mix(R ^ key[round]) & mask;
            const uint32_t nl = R;
            R = L ^ F;
            L = nl;
        }
        x = (L << half) | R;
    } while (x >= filled);
    return (size_t)x;
}
