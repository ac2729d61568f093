// ce2e_rng.h -- counter-based random numbers for the on-device environment reset.
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11):
// a keyed bijection of a 128-bit counter, ten rounds of two 32x32->64 multiplies.  No state: the
// draws of environment e in its k-th episode are a pure function of (seed, e, k, draw index), so a
// reset is reproducible whatever the batch size, launch geometry or number of GPUs, and the NumPy
// restatement in oracle/ regenerates them bit for bit.  Host + device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CE2E_HD __host__ __device__ __forceinline__
#else
#define CE2E_HD inline
#endif

namespace ce2e {

struct Philox4 {
    uint32_t v[4];
};

CE2E_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// 24 random bits -> fp32 in [0, 1), exactly representable (so every later product has one rounding)
CE2E_HD float u01_24(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

}  // namespace ce2e
