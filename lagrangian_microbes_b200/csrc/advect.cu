// RK4 advection + diffusion kick.
//
// Replaces  pset.execute(parcels.AdvectionRK4, runtime=dt, dt=dt, ...)   (particle_advecter.py:222-223)
// and       p.lat += uniform(-1,1)*sqrt(6|dt|Kh); p.lon += ...            (particle_advecter.py:240-242)
//
// The arithmetic follows the float32-faithful statement of parcels 2.0.0beta2's JIT path in
// DESIGN.md §4.1 (the CPU restatement the tests compare against is oracle/rk4.py): float32
// particle state and grid, float32 xsi/eta widened to double, double bilinear sums rounded to
// float32, float32 time interpolation, double unit conversion with the SAMPLE point's latitude,
// double stage updates rounded to float32 (stage 3 pure float32).  Every operation is written
// with an explicit-rounding intrinsic so nvcc can not contract a*b+c into an FMA: the x86-64
// code Parcels compiles has none.
//
// One thread per particle; the state is kept in (cell, id) order by the binning stage, so the
// 32 particles of a warp sit within a few 0.01-degree cells and their 16 corner reads per stage
// collapse to a handful of broadcast L1 hits: the kernel is bound by its fp64 arithmetic
// (4 x cos + 8 x div per particle), not by HBM (16 B per particle).
#include <algorithm>

#include "lm_internal.cuh"
#include "philox.cuh"

namespace lm {

struct StageDev {
    int ti[4];
    int interp[4];
    float frac[4];
    // float32 RK4 on the interleaved field: time interval ti2 and which of its two levels: 0 the earlier, 1 the later, 2 both (interpolate)
    int ti2[4];
    int pick[4];
};

// Index search on one grid axis (parcels.h::search_indices_rectilinear restated): the cell i with
// vals[i] <= x <= vals[i+1], found from a uniform-spacing guess and fixed up -- the result is that of Parcels'
// linear search for any ascending axis.  Returns the two bracketing axis values, which the caller needs anyway,
// so a regular axis costs two independent loads and no dependent ones.  v_first / v_last are the axis ends
// (kernel parameters).  Returns -1 out of bounds (also catches NaN).
__device__ __forceinline__ int search_axis(const float *__restrict__ vals, int n, float x, float v_first, float v_last,
                                           float inv_d, float &lo, float &hi)
{
    if (!(x >= v_first) || !(x <= v_last)) return -1;
    int i = (int)((x - v_first) * inv_d);
    i = max(0, min(i, n - 2));
    lo = __ldg(vals + i);
    hi = __ldg(vals + i + 1);
    while (i < n - 2 && x > hi) { ++i; lo = hi; hi = __ldg(vals + i + 1); }
    while (i > 0 && x < lo) { --i; hi = lo; lo = __ldg(vals + i); }
    return i;
}

// cos(x) for |x| <= pi/2 (latitudes): fdlibm's __kernel_cos / __kernel_sin polynomials (error < 1 ulp, as is
// the libm the CPU restatement links), without the generic argument reduction of the CUDA library routine.
__device__ __noinline__ double cos_generic(double x) { return cos(x); }

__device__ __forceinline__ double cos_lat(double x)
{
    const double ax = fabs(x);
    if (!(ax <= 1.5707963267948966)) return cos_generic(x);       // not a latitude: the library routine
    if (ax <= 0.78539816339744830962) {
        const double z = x * x;
        const double r = z * (4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 +
                         z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11)))));
        return 1.0 - (0.5 * z - z * r);
    }
    // cos(x) = sin(pi/2 - |x|), pi/2 in two parts
    const double y = (1.57079632679489655800e+00 - ax) + 6.12323399573676603587e-17;
    const double z = y * y;
    const double r = 8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 + z * (2.75573137070700676789e-06 +
                     z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10)));
    return y + y * z * (-1.66666666666666324348e-01 + z * r);
}

// Sample converted (u, v) [deg/s] at the float32 point (x, y).  Returns false when out of bounds.
__device__ __forceinline__ bool sample_uv(const FieldDev &f, float x, float y, int ti, int interp, float frac,
                                          float &u, float &v)
{
    float lx0, lx1, ly0, ly1;
    const int xi = search_axis(f.lon, f.X, x, f.lon0, f.lon1, f.inv_dx, lx0, lx1);
    const int yi = search_axis(f.lat, f.Y, y, f.lat0, f.lat1, f.inv_dy, ly0, ly1);
    if (xi < 0 || yi < 0) return false;
    const double xsi = (double)__fdiv_rn(__fsub_rn(x, lx0), __fsub_rn(lx1, lx0));
    const double eta = (double)__fdiv_rn(__fsub_rn(y, ly0), __fsub_rn(ly1, ly0));
    const double omx = __dsub_rn(1.0, xsi), ome = __dsub_rn(1.0, eta);
    const double w00 = __dmul_rn(omx, ome), w01 = __dmul_rn(xsi, ome);
    const double w11 = __dmul_rn(xsi, eta), w10 = __dmul_rn(omx, eta);
    const size_t slab = (size_t)f.Y * f.X;
    const size_t off = (size_t)ti * slab + (size_t)yi * f.X + xi;

    auto bilinear = [&](const float *__restrict__ d) -> float {
        const float d00 = __ldg(d), d01 = __ldg(d + 1), d10 = __ldg(d + f.X), d11 = __ldg(d + f.X + 1);
        double t = __dmul_rn(w00, (double)d00);
        t = __dadd_rn(t, __dmul_rn(w01, (double)d01));
        t = __dadd_rn(t, __dmul_rn(w11, (double)d11));
        t = __dadd_rn(t, __dmul_rn(w10, (double)d10));
        return __double2float_rn(t);
    };

    float uu = bilinear(f.U + off), vv = bilinear(f.V + off);
    if (interp) {
        const float u1 = bilinear(f.U + off + slab), v1 = bilinear(f.V + off + slab);
        uu = __fadd_rn(uu, __fmul_rn(__fsub_rn(u1, uu), frac));
        vv = __fadd_rn(vv, __fmul_rn(__fsub_rn(v1, vv), frac));
    }
    // u *= 1.0 / (1852. * 60. * cos(y * M_PI / 180));   v *= 1.0 / (1852. * 60.)
    const double ang = __ddiv_rn(__dmul_rn((double)y, 3.14159265358979323846), 180.0);
    const double cu = __ddiv_rn(1.0, __dmul_rn(111120.0, cos_lat(ang)));
    constexpr double cv = 1.0 / 111120.0;
    u = __double2float_rn(__dmul_rn((double)uu, cu));
    v = __double2float_rn(__dmul_rn((double)vv, cv));
    return true;
}

__global__ void __launch_bounds__(256) advect_rk4_kernel(FieldDev f, float *__restrict__ lon, float *__restrict__ lat,
                                                         int n, StageDev st, float dt, Counters *ctr)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float x = lon[p], y = lat[p];
    const double xd = (double)x, yd = (double)y, dtd = (double)dt;
    float u1, v1, u2, v2, u3, v3, u4, v4;
    bool ok = sample_uv(f, x, y, st.ti[0], st.interp[0], st.frac[0], u1, v1);
    if (ok) {
        const float x1 = __double2float_rn(__dadd_rn(xd, __dmul_rn(__dmul_rn((double)u1, .5), dtd)));
        const float y1 = __double2float_rn(__dadd_rn(yd, __dmul_rn(__dmul_rn((double)v1, .5), dtd)));
        ok = sample_uv(f, x1, y1, st.ti[1], st.interp[1], st.frac[1], u2, v2);
    }
    if (ok) {
        const float x2 = __double2float_rn(__dadd_rn(xd, __dmul_rn(__dmul_rn((double)u2, .5), dtd)));
        const float y2 = __double2float_rn(__dadd_rn(yd, __dmul_rn(__dmul_rn((double)v2, .5), dtd)));
        ok = sample_uv(f, x2, y2, st.ti[2], st.interp[2], st.frac[2], u3, v3);
    }
    if (ok) {
        const float x3 = __fadd_rn(x, __fmul_rn(u3, dt));     // no double literal on this line in the reference
        const float y3 = __fadd_rn(y, __fmul_rn(v3, dt));
        ok = sample_uv(f, x3, y3, st.ti[3], st.interp[3], st.frac[3], u4, v4);
    }
    if (!ok) {   // Parcels raises OutOfBoundsError; we leave the particle where it is and count it
        atomicAdd(&ctr->n_oob, 1ull);
        return;
    }
    const float su = __fadd_rn(__fadd_rn(__fadd_rn(u1, __fmul_rn(2.f, u2)), __fmul_rn(2.f, u3)), u4);
    const float sv = __fadd_rn(__fadd_rn(__fadd_rn(v1, __fmul_rn(2.f, v2)), __fmul_rn(2.f, v3)), v4);
    lon[p] = __double2float_rn(__dadd_rn(xd, __dmul_rn(__ddiv_rn((double)su, 6.0), dtd)));
    lat[p] = __double2float_rn(__dadd_rn(yd, __dmul_rn(__ddiv_rn((double)sv, 6.0), dtd)));
}

// ---- LM_OPT_ADVECT_MODE = 1: the same RK4 step in float32 arithmetic -------------------------------------------
// north_star's bar for positions is 1e-6 relative to the reference RK4; the particle state is float32 (half an ulp at
// 200 degrees is 3.8e-8 relative), a step moves a microbe ~0.01 degrees, so a float32 evaluation of the DISPLACEMENT
// (relative error ~1e-6 of 0.01 degrees = 1e-8 degrees) is three orders of magnitude below the rounding of the stored
// position itself.  What the bit-faithful kernel above pays for -- 10 fp64 divisions, 4 fp64 cosines and 64 fp64
// bilinear terms per microbe, all on explicit-rounding intrinsics -- buys bit-identity with an x86 evaluation order,
// not accuracy.  Here: FMA bilinear sums, MUFU reciprocal / cosine (|error| of __cosf on [-pi/2, pi/2] is 2^-21.4
// absolute), the same index search, the same out-of-bounds policy.  tests/test_gpu_advect_fast.py holds it to 1e-6
// relative against the float64 restatement, step by step from identical inputs and over config 1's 24 steps.
__device__ __forceinline__ bool sample_uv_fast(const FieldDev &f, float x, float y, int ti, int interp, float frac,
                                               float &u, float &v)
{
    float lx0, lx1, ly0, ly1;
    const int xi = search_axis(f.lon, f.X, x, f.lon0, f.lon1, f.inv_dx, lx0, lx1);
    const int yi = search_axis(f.lat, f.Y, y, f.lat0, f.lat1, f.inv_dy, ly0, ly1);
    if (xi < 0 || yi < 0) return false;
    const float xsi = __fdividef(x - lx0, lx1 - lx0), eta = __fdividef(y - ly0, ly1 - ly0);
    const float omx = 1.f - xsi, ome = 1.f - eta;
    const float w00 = omx * ome, w01 = xsi * ome, w11 = xsi * eta, w10 = omx * eta;
    const size_t slab = (size_t)f.Y * f.X;
    const size_t off = (size_t)ti * slab + (size_t)yi * f.X + xi;
    const float *__restrict__ pu = f.U + off, *__restrict__ pv = f.V + off;
    // all loads of the sample first (16 with time interpolation): one exposed latency instead of four
    const float u00 = __ldg(pu), u01 = __ldg(pu + 1), u10 = __ldg(pu + f.X), u11 = __ldg(pu + f.X + 1);
    const float v00 = __ldg(pv), v01 = __ldg(pv + 1), v10 = __ldg(pv + f.X), v11 = __ldg(pv + f.X + 1);
    float uu = fmaf(w10, u10, fmaf(w11, u11, fmaf(w01, u01, w00 * u00)));
    float vv = fmaf(w10, v10, fmaf(w11, v11, fmaf(w01, v01, w00 * v00)));
    if (interp) {
        const float *__restrict__ qu = pu + slab, *__restrict__ qv = pv + slab;
        const float a00 = __ldg(qu), a01 = __ldg(qu + 1), a10 = __ldg(qu + f.X), a11 = __ldg(qu + f.X + 1);
        const float b00 = __ldg(qv), b01 = __ldg(qv + 1), b10 = __ldg(qv + f.X), b11 = __ldg(qv + f.X + 1);
        const float u1 = fmaf(w10, a10, fmaf(w11, a11, fmaf(w01, a01, w00 * a00)));
        const float v1 = fmaf(w10, b10, fmaf(w11, b11, fmaf(w01, b01, w00 * b00)));
        uu = fmaf(u1 - uu, frac, uu);
        vv = fmaf(v1 - vv, frac, vv);
    }
    constexpr float inv_m_per_deg = (float)(1.0 / 111120.0);
    u = uu * __fdividef(inv_m_per_deg, __cosf(y * 0.017453292519943295f));
    v = vv * inv_m_per_deg;
    return true;
}

// the same sample from the interleaved copy: (u_t, v_t, u_t+1, v_t+1) of one grid point and time interval in ONE 16-byte load
__device__ __forceinline__ bool sample_uv4(const FieldDev &f, float x, float y, int ti2, int pick, float frac, float &u, float &v)
{
    float lx0, lx1, ly0, ly1;
    const int xi = search_axis(f.lon, f.X, x, f.lon0, f.lon1, f.inv_dx, lx0, lx1);
    const int yi = search_axis(f.lat, f.Y, y, f.lat0, f.lat1, f.inv_dy, ly0, ly1);
    if (xi < 0 || yi < 0) return false;
    const float xsi = __fdividef(x - lx0, lx1 - lx0), eta = __fdividef(y - ly0, ly1 - ly0);
    const float omx = 1.f - xsi, ome = 1.f - eta;
    const float w00 = omx * ome, w01 = xsi * ome, w11 = xsi * eta, w10 = omx * eta;
    const float4 *__restrict__ p = f.UV4 + ((size_t)ti2 * f.Y + yi) * f.X + xi;
    const float4 c00 = __ldg(p), c01 = __ldg(p + 1), c10 = __ldg(p + f.X), c11 = __ldg(p + f.X + 1);
    const float u0 = fmaf(w10, c10.x, fmaf(w11, c11.x, fmaf(w01, c01.x, w00 * c00.x)));
    const float v0 = fmaf(w10, c10.y, fmaf(w11, c11.y, fmaf(w01, c01.y, w00 * c00.y)));
    const float u1 = fmaf(w10, c10.z, fmaf(w11, c11.z, fmaf(w01, c01.z, w00 * c00.z)));
    const float v1 = fmaf(w10, c10.w, fmaf(w11, c11.w, fmaf(w01, c01.w, w00 * c00.w)));
    // the same arithmetic as sample_uv_fast on the same values: the two layouts give identical positions
    const float uu = pick == 2 ? fmaf(u1 - u0, frac, u0) : (pick == 1 ? u1 : u0);
    const float vv = pick == 2 ? fmaf(v1 - v0, frac, v0) : (pick == 1 ? v1 : v0);
    constexpr float inv_m_per_deg = (float)(1.0 / 111120.0);
    u = uu * __fdividef(inv_m_per_deg, __cosf(y * 0.017453292519943295f));
    v = vv * inv_m_per_deg;
    return true;
}

__global__ void __launch_bounds__(256) advect_rk4_uv4_kernel(FieldDev f, float *__restrict__ lon, float *__restrict__ lat,
                                                             int n, StageDev st, float dt, Counters *ctr)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float x = lon[p], y = lat[p];
    const float h = 0.5f * dt;
    float u1, v1, u2, v2, u3, v3, u4, v4;
    bool ok = sample_uv4(f, x, y, st.ti2[0], st.pick[0], st.frac[0], u1, v1);
    if (ok) ok = sample_uv4(f, fmaf(u1, h, x), fmaf(v1, h, y), st.ti2[1], st.pick[1], st.frac[1], u2, v2);
    if (ok) ok = sample_uv4(f, fmaf(u2, h, x), fmaf(v2, h, y), st.ti2[2], st.pick[2], st.frac[2], u3, v3);
    if (ok) ok = sample_uv4(f, fmaf(u3, dt, x), fmaf(v3, dt, y), st.ti2[3], st.pick[3], st.frac[3], u4, v4);
    if (!ok) {
        atomicAdd(&ctr->n_oob, 1ull);
        return;
    }
    const float sixth = dt * (1.f / 6.f);
    lon[p] = fmaf(u1 + 2.f * (u2 + u3) + u4, sixth, x);
    lat[p] = fmaf(v1 + 2.f * (v2 + v3) + v4, sixth, y);
}

__global__ void __launch_bounds__(256) interleave_field_kernel(const float *__restrict__ U, const float *__restrict__ V,
                                                               float4 *__restrict__ out, size_t slab, size_t total)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
        out[k] = make_float4(__ldg(U + k), __ldg(V + k), __ldg(U + k + slab), __ldg(V + k + slab));
}

cudaError_t launch_interleave_field(const FieldDev &f, float4 *uv4, cudaStream_t s, int64_t *launches)
{
    const size_t slab = (size_t)f.Y * f.X, total = (size_t)(f.T - 1) * slab;
    if (total == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)std::min<size_t>((total + 255) / 256, (size_t)kNumSMs * 16);
    interleave_field_kernel<<<grid, 256, 0, s>>>(f.U, f.V, uv4, slab, total);
    ++*launches;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) advect_rk4_fast_kernel(FieldDev f, float *__restrict__ lon, float *__restrict__ lat,
                                                              int n, StageDev st, float dt, Counters *ctr)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float x = lon[p], y = lat[p];
    const float h = 0.5f * dt;
    float u1, v1, u2, v2, u3, v3, u4, v4;
    bool ok = sample_uv_fast(f, x, y, st.ti[0], st.interp[0], st.frac[0], u1, v1);
    if (ok) ok = sample_uv_fast(f, fmaf(u1, h, x), fmaf(v1, h, y), st.ti[1], st.interp[1], st.frac[1], u2, v2);
    if (ok) ok = sample_uv_fast(f, fmaf(u2, h, x), fmaf(v2, h, y), st.ti[2], st.interp[2], st.frac[2], u3, v3);
    if (ok) ok = sample_uv_fast(f, fmaf(u3, dt, x), fmaf(v3, dt, y), st.ti[3], st.interp[3], st.frac[3], u4, v4);
    if (!ok) {   // same policy as the bit-faithful kernel: the particle stays where it is and is counted
        atomicAdd(&ctr->n_oob, 1ull);
        return;
    }
    const float sixth = dt * (1.f / 6.f);
    lon[p] = fmaf(u1 + 2.f * (u2 + u3) + u4, sixth, x);
    lat[p] = fmaf(v1 + 2.f * (v2 + v3) + v4, sixth, y);
}

cudaError_t launch_advect(const FieldDev &f, float *lon, float *lat, int n, const lm_stage_times &st, float dt,
                          Counters *ctr, cudaStream_t s, int64_t *launches, int mode)
{
    if (n <= 0) return cudaSuccess;
    StageDev sd;
    for (int k = 0; k < 4; ++k) {
        sd.ti[k] = st.ti[k];
        sd.interp[k] = st.interp[k];
        sd.frac[k] = st.frac[k];
        // interleaved layout: interval ti (both levels when interpolating, else its earlier level), or -- at the last
        // level, which starts no interval -- the later level of the interval before
        const bool last = !st.interp[k] && st.ti[k] >= f.T - 1;
        sd.ti2[k] = last ? st.ti[k] - 1 : st.ti[k];
        sd.pick[k] = st.interp[k] ? 2 : (last ? 1 : 0);
    }
    const int block = 256;
    if (mode == 1 && f.UV4) advect_rk4_uv4_kernel<<<(n + block - 1) / block, block, 0, s>>>(f, lon, lat, n, sd, dt, ctr);
    else if (mode == 1) advect_rk4_fast_kernel<<<(n + block - 1) / block, block, 0, s>>>(f, lon, lat, n, sd, dt, ctr);
    else advect_rk4_kernel<<<(n + block - 1) / block, block, 0, s>>>(f, lon, lat, n, sd, dt, ctr);
    ++*launches;
    return cudaGetLastError();
}

// p.lat += uniform(-1, 1) * amp; p.lon += uniform(-1, 1) * amp   (lat drawn first; numpy's
// uniform(-1,1) = -1 + 2*u).  float32 attribute + Python float -> double sum, stored as float32.
__global__ void __launch_bounds__(256) diffuse_kernel(float *__restrict__ lon, float *__restrict__ lat,
                                                      const int32_t *__restrict__ ids, int n, double amp,
                                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo,
                                                      uint32_t step_hi)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t pid = ids ? (uint32_t)ids[p] : (uint32_t)p;
    double ua, ub;
    particle_uniforms(pid, 0u, step_lo, step_hi, seed_lo, seed_hi, ua, ub);
    const double ka = __dmul_rn(__dadd_rn(-1.0, __dmul_rn(2.0, ua)), amp);
    const double kb = __dmul_rn(__dadd_rn(-1.0, __dmul_rn(2.0, ub)), amp);
    lat[p] = __double2float_rn(__dadd_rn((double)lat[p], ka));
    lon[p] = __double2float_rn(__dadd_rn((double)lon[p], kb));
}

cudaError_t launch_diffuse(float *lon, float *lat, const int32_t *ids, int n, double amp, uint64_t seed, uint64_t step,
                           cudaStream_t s, int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    const int block = 256;
    diffuse_kernel<<<(n + block - 1) / block, block, 0, s>>>(lon, lat, ids, n, amp, (uint32_t)seed,
                                                              (uint32_t)(seed >> 32), (uint32_t)step,
                                                              (uint32_t)(step >> 32));
    ++*launches;
    return cudaGetLastError();
}

}  // namespace lm
