// rng.cuh -- counter-based RNG for the path kernels.
//
// Takes the place of the reference's rand_double() (reference include/util/rand_util.h:85-117:
// a thread_local 32-bit LCG whose stream depends on OpenMP scheduling).  Here every random
// number is a pure function of (seed, pixel, sample, bounce), so an image does not depend on
// how pixel-samples are mapped to threads, blocks or GPUs.  Generator: Philox4x32-10
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), one call per ray
// segment = four 32-bit words.
#pragma once
#include <cstdint>

namespace b200rt {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// 24 high bits -> float in [0, 1)
__host__ __device__ inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace b200rt
