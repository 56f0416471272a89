// Philox counter-based generators (Salmon et al., SC'11) on the device.  Same counter/key layout as oracle/philox.py
// (the test-side restatement); tests compare the two bit for bit through lm_pair_uniforms.
//
// Per-pair stream (the reference's np.random.rand() in interactions.py:20, made order-independent):
//     key(seed, step) = word 0 of Philox4x32-10(counter = (step_lo, step_hi, 'RPS1', 0), key = seed)    once per step, host
//     (x0, x1)        = Philox2x32-10(counter = (i, j), key = key(seed, step))                          i < j particle ids
//     u               = ((x0 >> 5) * 2^26 + (x1 >> 6)) / 2^53
// Philox2x32 yields exactly the 64 bits one draw needs with one 32 x 32 -> 64 multiplication per round -- half the
// instructions of Philox4x32 (the draw is the single most expensive operation of the interaction step: ~45 of them per
// differing pair).  Per-particle uniforms (diffusion kick) keep Philox4x32: two draws per call.
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

__host__ __device__ __forceinline__ void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k, uint32_t &x0, uint32_t &x1)
{
    constexpr uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p = (unsigned long long)M * c0;
        const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ k ^ c1; c1 = lo;
        k += W;
    }
    x0 = c0; x1 = c1;
}

// the 53-bit integer m of the pair's draw, u = m * 2^-53
__host__ __device__ __forceinline__ unsigned long long pair_draw_m(uint32_t i, uint32_t j, uint32_t key)
{
    uint32_t x0, x1;
    philox2x32_10(i, j, key, x0, x1);
    return ((unsigned long long)(x0 >> 5) << 26) | (unsigned long long)(x1 >> 6);
}

// key of the per-pair stream of one step (host side: csrc/api.cu computes it once per call)
inline uint32_t pair_stream_key(uint64_t seed, uint64_t step)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = 0x52505331u /* 'RPS1' */, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += W0; k1 += W1;
    }
    return c0;
}

// two 32-bit words -> double in [0,1) with 53 random bits
__device__ __forceinline__ double u53(uint32_t x0, uint32_t x1)
{
    return ((double)(x0 >> 5) * 67108864.0 + (double)(x1 >> 6)) * (1.0 / 9007199254740992.0);
}

// per-pair uniform: Philox2x32-10, counter = (i, j), key = pair_stream_key(seed, step);  requires i < j
__device__ __forceinline__ double pair_uniform(uint32_t i, uint32_t j, uint32_t key)
{
    return (double)pair_draw_m(i, j, key) * (1.0 / 9007199254740992.0);
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
