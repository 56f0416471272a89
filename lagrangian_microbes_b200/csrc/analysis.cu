// On-device analysis reductions on a snapshot of the microbes (the callers
// downstream of the per-timestep path).
//
//  pair_distance_hist_kernel   the reference's pair-distance histogram
//      sandbox/pairwise_distance_histogram_distributed.jl:33-44 (haversine_distance32), :54-64
//      (pairwise_distance_histogram_1point: for i, for j > i: bin = round(Int8, 10 log10(max(1, d))); hist[bin] += 1)
//      and its unfinished CUDA version, sandbox/pairwise_distance_histogram_gpu.jl:14-43, which stops at
//      "dist_hist[bin] += 1  # Cannot work of course!!!".  All N(N-1)/2 pairs of one point set, compute-bound.
//
//      Design: a persistent grid (two CTAs per SM) walks tiles of 256 anchors x 1024 partners of the upper triangle;
//      the partners (lat, lon, cos lat) are staged once per tile in shared memory and read as broadcasts, each thread
//      keeps its anchor in registers.  The float32 haversine argument a = sin^2(dlat/2) + cos cos sin^2(dlon/2) is
//      evaluated in the reference's operation order; bin = round(10 log10(max(1, 2 R asin(min(1, sqrt a))))) is a
//      monotone function of a, so instead of sqrt + asin + log10 per pair the bin is found from a table of a-values
//      at the bin edges (built on the host in double, rounded up to float32): one MUFU.LG2 for a first guess, then
//      exact comparisons against the table.  The increment -- the step the reference's kernel could not do -- goes to
//      a histogram PRIVATE TO THE THREAD in shared memory (slot * 256 + thread: bank = lane, no conflicts, no
//      atomics); the CTA's columns are summed once at the end and added to the global histogram with one 64-bit atomic
//      per slot and CTA.
//
//  raster_kernel / compose_kernel   the frame of microbe_plotter.py:82-155 (plt.scatter of every microbe, coloured by
//      species) as a raster: per pixel the number of microbes of each species and the highest particle index (the
//      marker matplotlib draws last is the one on top), then an RGB image.
#include <cmath>

#include "lm_internal.cuh"

namespace lm {

constexpr int PDH_THREADS = 256;
constexpr int PDH_TJ = 1024;              // partners staged per tile
constexpr int PDH_MAX_SLOTS = 128;        // bins + 2 <= this

struct PdhArgs {
    const float *__restrict__ lat;
    const float *__restrict__ lon;
    int n;
    int slots;                            // bins + 2: bins 0..bins, then "beyond the last bin"
    float c0, c1;                         // first guess: slot ~ c0 + c1 * log2(a)
    long long n_iblocks, n_jchunks;
    unsigned long long *hist;             // [slots]
    float edge[PDH_MAX_SLOTS + 1];        // edge[k] = smallest float32 a that belongs to slot >= k; edge[0] = -inf, edge[slots] = +inf
};

__global__ void __launch_bounds__(PDH_THREADS, 2) pair_distance_hist_kernel(const __grid_constant__ PdhArgs A)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_raw);                                   // [slots][256]
    float4 *s_pts = reinterpret_cast<float4 *>(s_raw + (size_t)A.slots * PDH_THREADS * 4);   // [PDH_TJ] lat, lon, cos(lat), -
    float *s_edge = reinterpret_cast<float *>(s_pts + PDH_TJ);                               // [slots + 1]
    const int tid = threadIdx.x;
    const int slots = A.slots;
    for (int k = tid; k < slots * PDH_THREADS; k += PDH_THREADS) s_cnt[k] = 0u;
    for (int k = tid; k <= slots; k += PDH_THREADS) s_edge[k] = A.edge[k];
    uint32_t *my_cnt = s_cnt + tid;

    const long long n_tiles = A.n_iblocks * A.n_jchunks;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long bi = t / A.n_jchunks, cj = t - bi * A.n_jchunks;
        const long long i0 = bi * PDH_THREADS, j0 = cj * PDH_TJ;
        const int m = (int)min((long long)PDH_TJ, (long long)A.n - j0);       // partners in this chunk
        if (j0 + m <= i0 + 1) continue;                                       // chunk entirely at or below the first anchor (CTA-uniform)
        __syncthreads();                                                      // the previous tile has been read
        for (int jj = tid; jj < m; jj += PDH_THREADS) {
            const float la = __ldg(A.lat + j0 + jj), lo = __ldg(A.lon + j0 + jj);
            s_pts[jj] = make_float4(la, lo, cospif(la * (1.0f / 180.0f)), 0.f);
        }
        __syncthreads();
        const long long i = i0 + tid;
        if (i < A.n) {
            const float lat_i = __ldg(A.lat + i), lon_i = __ldg(A.lon + i);
            const float c_i = cospif(lat_i * (1.0f / 180.0f));
            const int jb = (int)max(0ll, i + 1 - j0);                         // first partner with j > i
#pragma unroll 4
            for (int jj = jb; jj < m; ++jj) {
                const float4 p = s_pts[jj];
                const float d1 = sinpif((p.x - lat_i) * (1.0f / 360.0f));
                const float d2 = sinpif((p.y - lon_i) * (1.0f / 360.0f));
                const float tt = __fmul_rn(__fmul_rn(__fmul_rn(d2, d2), c_i), p.z);     // d2 * d2 * c1 * c2, left to right
                const float a = __fadd_rn(__fmul_rn(d1, d1), tt);
                int g = __float2int_rn(fmaf(A.c1, __log2f(a), A.c0));                 // a = 0: -inf -> INT_MIN -> 0
                g = max(0, min(g, slots - 1));
                while (a >= s_edge[g + 1]) ++g;                                       // edge[slots] = +inf
                while (a < s_edge[g]) --g;                                            // edge[0] = -inf
                my_cnt[g * PDH_THREADS] += 1u;
            }
        }
    }
    __syncthreads();
    // column sums: warp w takes slots w, w + 8, ...; lane l adds the eight threads l, l + 32, ...
    const int lane = tid & 31, warp = tid >> 5;
    for (int s = warp; s < slots; s += PDH_THREADS / 32) {
        unsigned long long v = 0;
#pragma unroll
        for (int q = 0; q < PDH_THREADS / 32; ++q) v += s_cnt[s * PDH_THREADS + lane + 32 * q];
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(0xffffffffu, v, dd);
        if (lane == 0 && v) atomicAdd(A.hist + s, v);
    }
}

// smallest float32 >= x
static float f32_at_or_above(double x)
{
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

cudaError_t launch_pair_distance_hist(const float *lat, const float *lon, int64_t n, float radius_m, int bins,
                                      unsigned long long *hist, cudaStream_t s, int64_t *launches)
{
    const int slots = bins + 2;
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)slots * sizeof(unsigned long long), s);
    if (e != cudaSuccess || n < 2) return e;
    PdhArgs A;
    A.lat = lat; A.lon = lon; A.n = (int)n; A.slots = slots; A.hist = hist;
    A.n_iblocks = (n + PDH_THREADS - 1) / PDH_THREADS;
    A.n_jchunks = (n + PDH_TJ - 1) / PDH_TJ;
    // slot k >= 1 holds the pairs with 10 log10(d) in [k - 1/2, k + 1/2): d >= 10^((k - 1/2) / 10)
    //   <=> 2 R asin(sqrt a) >= d_k  <=>  a >= sin^2(d_k / 2R)   (d_k / 2R < pi / 2; beyond, no a reaches the slot)
    const double R = (double)radius_m;
    A.edge[0] = -INFINITY;
    for (int k = 1; k <= slots; ++k) {
        const double half_angle = pow(10.0, (k - 0.5) / 10.0) / (2.0 * R);
        if (k == slots || !(half_angle < M_PI / 2)) A.edge[k] = INFINITY;
        else { const double sn = sin(half_angle); A.edge[k] = f32_at_or_above(sn * sn); }
    }
    // first guess from the small-angle form d ~ 2 R sqrt(a):  10 log10(d) ~ 10 log10(2R) + 5 log10(2) log2(a)
    A.c0 = (float)(10.0 * log10(2.0 * R));
    A.c1 = (float)(5.0 * log10(2.0));
    const size_t smem = (size_t)slots * PDH_THREADS * 4 + (size_t)PDH_TJ * sizeof(float4) + (size_t)(slots + 1) * 4;
    e = cudaFuncSetAttribute(pair_distance_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long n_tiles = A.n_iblocks * A.n_jchunks;
    const int grid = (int)std::min<long long>(n_tiles, 2ll * kNumSMs);
    pair_distance_hist_kernel<<<grid, PDH_THREADS, smem, s>>>(A);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
struct RasterArgs {
    const float *__restrict__ lon;
    const float *__restrict__ lat;
    const int8_t *__restrict__ sp;
    int n;
    double lon_min, lat_min, sx, sy;      // pixel = floor((v - min) * s)
    int width, height;
    uint32_t *counts;                     // [3][height][width] microbes of species 1, 2, 3 per pixel
    int32_t *top;                         // [height][width] highest particle index in the pixel (-1: none)
};

__global__ void __launch_bounds__(256) raster_kernel(RasterArgs A)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= A.n) return;
    const double fx = floor(__dmul_rn(__dsub_rn((double)__ldg(A.lon + i), A.lon_min), A.sx));
    const double fy = floor(__dmul_rn(__dsub_rn((double)__ldg(A.lat + i), A.lat_min), A.sy));
    if (!(fx >= 0.0 && fx < (double)A.width && fy >= 0.0 && fy < (double)A.height)) return;   // outside the extent (or NaN)
    const size_t pix = (size_t)(A.height - 1 - (int)fy) * A.width + (int)fx;                  // row 0 = northern edge
    const int s = A.sp ? (int)__ldg(A.sp + i) : 1;
    if (s >= 1 && s <= 3) atomicAdd(A.counts + (size_t)(s - 1) * A.width * A.height + pix, 1u);
    atomicMax(A.top + pix, i);
}

struct ComposeArgs {
    const uint32_t *__restrict__ counts;
    const int32_t *__restrict__ top;
    const int8_t *__restrict__ sp;
    int n_pix;
    int mode;                             // 0: the microbe drawn last is on top; 1: the most numerous species
    uint32_t palette[4];                  // 0x00BBGGRR: background, rock, paper, scissors
    uint8_t *rgb;                         // [n_pix][3]
};

__global__ void __launch_bounds__(256) compose_kernel(const __grid_constant__ ComposeArgs A)
{
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= A.n_pix) return;
    int s = 0;
    if (A.mode == 0) {
        const int t = __ldg(A.top + p);
        if (t >= 0) { s = A.sp ? (int)__ldg(A.sp + t) : 1; if (s < 1 || s > 3) s = 0; }
    } else {
        const uint32_t c1 = __ldg(A.counts + p), c2 = __ldg(A.counts + A.n_pix + p), c3 = __ldg(A.counts + 2 * (size_t)A.n_pix + p);
        if (c1 | c2 | c3) s = (c1 >= c2 && c1 >= c3) ? 1 : (c2 >= c3 ? 2 : 3);            // ties: the lower species number
    }
    const uint32_t c = A.palette[s];
    A.rgb[3 * (size_t)p + 0] = (uint8_t)(c & 0xffu);
    A.rgb[3 * (size_t)p + 1] = (uint8_t)((c >> 8) & 0xffu);
    A.rgb[3 * (size_t)p + 2] = (uint8_t)((c >> 16) & 0xffu);
}

cudaError_t launch_raster(const float *lon, const float *lat, const int8_t *sp, int64_t n, double lon_min, double lon_max,
                          double lat_min, double lat_max, int width, int height, uint32_t *counts, int32_t *top,
                          cudaStream_t s, int64_t *launches)
{
    const size_t n_pix = (size_t)width * height;
    cudaError_t e = cudaMemsetAsync(counts, 0, 3 * n_pix * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(top, 0xff, n_pix * sizeof(int32_t), s);          // -1
    if (e != cudaSuccess || n <= 0) return e;
    RasterArgs A;
    A.lon = lon; A.lat = lat; A.sp = sp; A.n = (int)n;
    A.lon_min = lon_min; A.lat_min = lat_min;
    A.sx = (double)width / (lon_max - lon_min); A.sy = (double)height / (lat_max - lat_min);
    A.width = width; A.height = height; A.counts = counts; A.top = top;
    raster_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(A);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_compose(const uint32_t *counts, const int32_t *top, const int8_t *sp, int width, int height, int mode,
                           const uint8_t *palette_rgb, uint8_t *rgb, cudaStream_t s, int64_t *launches)
{
    ComposeArgs A;
    A.counts = counts; A.top = top; A.sp = sp; A.n_pix = width * height; A.mode = mode; A.rgb = rgb;
    for (int k = 0; k < 4; ++k)
        A.palette[k] = (uint32_t)palette_rgb[3 * k] | ((uint32_t)palette_rgb[3 * k + 1] << 8) | ((uint32_t)palette_rgb[3 * k + 2] << 16);
    if (A.n_pix <= 0) return cudaSuccess;
    compose_kernel<<<(unsigned)((A.n_pix + 255) / 256), 256, 0, s>>>(A);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace lm
