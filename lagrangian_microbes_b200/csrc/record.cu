// Lossless delta packing of the per-step position record ("on-GPU quantise/delta-pack").
//
// The reference stores float32 lon / lat of every microbe after every step (particle_advecter.py:233-235,
// interaction_simulator.py:108-110): 8 B per microbe-step of positions + 1 B of species, and at 10^9..10^10
// microbe-steps/s that record is what the PCIe link carries (the end-to-end number).  A microbe moves
// ~0.01 degrees per step, i.e. a few hundred float32 ulps at these magnitudes, so the record of step k is sent as the
// DIFFERENCE to the record of step k-1 in units of ulps:
//
//     key(x)  = the float32 bit pattern mapped to an unsigned integer that is monotone in x
//               (negative: ~bits, else bits | 0x80000000) -- a bijection on all 2^32 patterns, NaNs included
//     d       = key(cur) - key(prev)                          (exact, as a 64-bit integer)
//     |d| <= 32767  ->  int16 d;   else  ->  int16 -32768 and one escape entry {2 i + coordinate, raw bits of cur}
//
// 4 B instead of 8 B per microbe-step, bit-exact after decoding (io.py::unpack_delta_record adds the deltas to the
// previous record's keys and applies the escapes).  HBM-bound: R 16 + W 4 B per microbe; four microbes per thread
// with 16-byte loads and 8-byte stores when the arrays are aligned, a scalar path otherwise and for the tail.
#include <algorithm>
#include <thread>
#include <vector>

#include "lm_internal.cuh"

namespace lm {

struct alignas(8) Delta4 { int16_t v[4]; };

__device__ __forceinline__ uint32_t mono_key(float x)
{
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one coordinate of one microbe; escapes are appended in arbitrary order (the decoder does not depend on it)
__device__ __forceinline__ int16_t delta_or_escape(float prev, float cur, long long slot, uint2 *esc, long long esc_cap,
                                                   unsigned int *esc_count)
{
    const long long d = (long long)mono_key(cur) - (long long)mono_key(prev);
    if (d >= -32767 && d <= 32767) return (int16_t)d;
    const unsigned int k = atomicAdd(esc_count, 1u);          // counts every escape, stored or not: the caller sees an overflow
    if ((long long)k < esc_cap) esc[k] = make_uint2((unsigned int)slot, __float_as_uint(cur));
    return (int16_t)-32768;
}

template <bool VEC>
__global__ void __launch_bounds__(256) record_delta_pack_kernel(const float *__restrict__ prev_lon, const float *__restrict__ prev_lat,
                                                                const float *__restrict__ lon, const float *__restrict__ lat,
                                                                long long n, int16_t *__restrict__ dlon, int16_t *__restrict__ dlat,
                                                                uint2 *esc, long long esc_cap, unsigned int *esc_count)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (VEC) {
        const long long i = 4 * t;
        if (i + 3 < n) {
            const float4 pl = __ldg(reinterpret_cast<const float4 *>(prev_lon + i)), pa = __ldg(reinterpret_cast<const float4 *>(prev_lat + i));
            const float4 cl = __ldg(reinterpret_cast<const float4 *>(lon + i)), ca = __ldg(reinterpret_cast<const float4 *>(lat + i));
            Delta4 ol, oa;
            ol.v[0] = delta_or_escape(pl.x, cl.x, 2 * i + 0, esc, esc_cap, esc_count);
            ol.v[1] = delta_or_escape(pl.y, cl.y, 2 * i + 2, esc, esc_cap, esc_count);
            ol.v[2] = delta_or_escape(pl.z, cl.z, 2 * i + 4, esc, esc_cap, esc_count);
            ol.v[3] = delta_or_escape(pl.w, cl.w, 2 * i + 6, esc, esc_cap, esc_count);
            oa.v[0] = delta_or_escape(pa.x, ca.x, 2 * i + 1, esc, esc_cap, esc_count);
            oa.v[1] = delta_or_escape(pa.y, ca.y, 2 * i + 3, esc, esc_cap, esc_count);
            oa.v[2] = delta_or_escape(pa.z, ca.z, 2 * i + 5, esc, esc_cap, esc_count);
            oa.v[3] = delta_or_escape(pa.w, ca.w, 2 * i + 7, esc, esc_cap, esc_count);
            *reinterpret_cast<Delta4 *>(dlon + i) = ol;
            *reinterpret_cast<Delta4 *>(dlat + i) = oa;
        } else {
            for (long long j = i; j < n; ++j) {                // the last, partial group of four
                dlon[j] = delta_or_escape(__ldg(prev_lon + j), __ldg(lon + j), 2 * j, esc, esc_cap, esc_count);
                dlat[j] = delta_or_escape(__ldg(prev_lat + j), __ldg(lat + j), 2 * j + 1, esc, esc_cap, esc_count);
            }
        }
    } else if (t < n) {
        dlon[t] = delta_or_escape(__ldg(prev_lon + t), __ldg(lon + t), 2 * t, esc, esc_cap, esc_count);
        dlat[t] = delta_or_escape(__ldg(prev_lat + t), __ldg(lat + t), 2 * t + 1, esc, esc_cap, esc_count);
    }
}

static bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

cudaError_t launch_record_delta_pack(const float *prev_lon, const float *prev_lat, const float *lon, const float *lat, int64_t n,
                                     int16_t *dlon, int16_t *dlat, uint32_t *esc, int64_t esc_cap, uint32_t *esc_count,
                                     cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(esc_count, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess || n <= 0) return e;
    uint2 *esc2 = reinterpret_cast<uint2 *>(esc);
    const bool vec = aligned_to(prev_lon, 16) && aligned_to(prev_lat, 16) && aligned_to(lon, 16) && aligned_to(lat, 16) &&
                     aligned_to(dlon, 8) && aligned_to(dlat, 8);
    if (vec) {
        const long long threads = (n + 3) / 4;
        record_delta_pack_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(prev_lon, prev_lat, lon, lat, (long long)n, dlon, dlat,
                                                                                   esc2, (long long)esc_cap, esc_count);
    } else {
        record_delta_pack_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(prev_lon, prev_lat, lon, lat, (long long)n, dlon, dlat,
                                                                              esc2, (long long)esc_cap, esc_count);
    }
    return cudaGetLastError();
}


// ---- host decoder (plain C++ on HOST arrays; no device work) ---------------------------------------------------
// branch-free forms of mono_key and its inverse (the loop below vectorises)
static inline uint32_t host_key(uint32_t b) { return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u); }
static inline uint32_t host_unkey(uint32_t k) { return k ^ (~(uint32_t)((int32_t)k >> 31) | 0x80000000u); }

static int64_t unpack_range(const uint32_t *__restrict__ prev, const int16_t *__restrict__ d, uint32_t *out, int64_t a, int64_t b)
{
    int64_t marked = 0;
    if (out == prev) {                                         // decoding in place
        for (int64_t i = a; i < b; ++i) {
            const int32_t v = d[i], is_esc = v == -32768;
            marked += is_esc;
            out[i] = host_unkey(host_key(out[i]) + (uint32_t)(v & (is_esc - 1)));
        }
        return marked;
    }
    uint32_t *__restrict__ o = out;
    for (int64_t i = a; i < b; ++i) {
        const int32_t v = d[i], is_esc = v == -32768;
        marked += is_esc;
        o[i] = host_unkey(host_key(prev[i]) + (uint32_t)(v & (is_esc - 1)));       // exact: the true key is in range
    }
    return marked;
}

// -> number of escape markers met, or -1 if an escape entry points outside the arrays
int64_t record_delta_unpack_host(const float *prev_lon, const float *prev_lat, const int16_t *dlon, const int16_t *dlat,
                                 const uint32_t *esc, int64_t n_esc, int64_t n, float *lon_out, float *lat_out, int n_threads)
{
    const uint32_t *pl = reinterpret_cast<const uint32_t *>(prev_lon), *pa = reinterpret_cast<const uint32_t *>(prev_lat);
    uint32_t *ol = reinterpret_cast<uint32_t *>(lon_out), *oa = reinterpret_cast<uint32_t *>(lat_out);
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n / 65536));          // a thread is not worth less than 64k microbes
    std::vector<int64_t> marked((size_t)T, 0);
    auto work = [&](int t) {
        const int64_t a = n * t / T, b = n * (t + 1) / T;
        marked[(size_t)t] = unpack_range(pl, dlon, ol, a, b) + unpack_range(pa, dlat, oa, a, b);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (int64_t m : marked) total += m;
    for (int64_t k = 0; k < n_esc; ++k) {
        const uint32_t slot = esc[2 * k];
        if ((int64_t)(slot >> 1) >= n) return -1;
        ((slot & 1u) ? oa : ol)[slot >> 1] = esc[2 * k + 1];
    }
    return total;
}

}  // namespace lm
