// Philox4x32-10 on the device.  Same counter/key layout as oracle/philox.py (the test-side
// restatement); tests compare the two bit for bit through lm_pair_uniforms.
//
// Replaces the reference's np.random.rand() (interactions.py:20) / numpy.random.uniform
// (particle_advecter.py:241-242) with an order-independent, per-pair / per-particle stream.
#pragma once
#include <stdint.h>

namespace lm {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two 32-bit words -> double in [0,1) with 53 random bits
__device__ __forceinline__ double u53(uint32_t x0, uint32_t x1)
{
    return ((double)(x0 >> 5) * 67108864.0 + (double)(x1 >> 6)) * (1.0 / 9007199254740992.0);
}

// per-pair uniform: counter = (i, j, step_lo, step_hi), key = seed;  requires i < j
__device__ __forceinline__ double pair_uniform(uint32_t i, uint32_t j, uint32_t step_lo, uint32_t step_hi,
                                               uint32_t seed_lo, uint32_t seed_hi)
{
    uint32_t x[4];
    philox4x32_10(i, j, step_lo, step_hi, seed_lo, seed_hi, x);
    return u53(x[0], x[1]);
}

// per-particle uniforms (diffusion kick): counter = (id, 0xD1FF0000|stream, step_lo, step_hi)
__device__ __forceinline__ void particle_uniforms(uint32_t pid, uint32_t stream, uint32_t step_lo, uint32_t step_hi,
                                                  uint32_t seed_lo, uint32_t seed_hi, double &ua, double &ub)
{
    uint32_t x[4];
    philox4x32_10(pid, 0xD1FF0000u | (stream & 0xFFFFu), step_lo, step_hi, seed_lo, seed_hi, x);
    ua = u53(x[0], x[1]);
    ub = u53(x[2], x[3]);
}

}  // namespace lm
