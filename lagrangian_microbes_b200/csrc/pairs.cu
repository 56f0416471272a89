// Radius pair search on the binned particle arrays + rock-paper-scissors resolution in the
// canonical cell-phase order.
//
// Replaces  kdt.query_pairs(r=interaction_radius, p=2)          (interaction_simulator.py:98)
// and       for pair in microbe_pairs: pair_interaction(...)    (interaction_simulator.py:104-105)
//           -> rock_paper_scissors_interaction                   (interactions.py:13-40)
//
// Predicate (identical to SciPy's for p=2, float32 positions widened to double):
//     s = fl(dx*dx); s = fl(s + fl(dy*dy));   pair <=> s <= fl(r*r)
// A float32 evaluation decides every pair whose squared distance is not within 4e-6 (relative) of
// r*r; the rest (a ~1e-5 fraction) take the exact double path, so the result is bit-exact while
// the inner loop stays in fp32.
//
// Two stages:
//
//  find_pairs_kernel   one thread per particle a, half stencil (rest of its own cell, E, NW, N, NE):
//      everything that does not depend on species is done here, once, densely -- distance tests,
//      the per-pair Philox draw reduced to three decision bits (u < pRS, u < pPR, u < pSP, as exact
//      integer compares of the 53-bit draw against ceil(p * 2^53)), the pair list (ids, i < j).
//      Per particle it leaves meta[a] = (offset, hits per direction) and hits[offset + k] =
//      b | decision_bits << 29 in (direction, b) order.
//
//  resolve_phase_kernel   the sequential part.  A *unit* is a cell (pairs inside it) or two
//      adjacent cells; units of one *phase* touch disjoint particles:
//          phase 0        same cell
//          phase 1,2      E neighbour, anchor cx even / odd
//          phase 3,4,5    NW, N, NE neighbour, anchor cy even
//          phase 6,7,8    NW, N, NE neighbour, anchor cy odd
//      so each phase is one conflict-free launch, one lane walks a unit's hit lists in (id_a, id_b)
//      order and applies interactions.py:13-40 with table look-ups: the reference's sequential
//      in-place semantics under the canonical total order (phase, unit, id_a, id_b)
//      (oracle/rps.py::cell_phase_order).  Units with more than HEAVY_TESTS candidate pairs (dense
//      clusters) are resolved by the whole warp: a row's hits are independent except through the
//      anchor particle's species, a 3-state value, so the row is a prefix scan over 3->3 maps.
#include "lm_internal.cuh"
#include "philox.cuh"

namespace lm {

constexpr int FIND_THREADS = 256;
constexpr int FIND_WARPS = FIND_THREADS / 32;
constexpr int STAGE = 512;               // staged hits per warp (dense Philox / coalesced stores)
constexpr uint32_t HIT_MASK = (1u << 29) - 1u;
constexpr int HEAVY_TESTS = 2048;        // candidate pairs above which a unit is resolved by the warp

struct FindArgs {
    const float *__restrict__ lon;
    const float *__restrict__ lat;
    const int32_t *__restrict__ id;
    const int32_t *__restrict__ cell_start;
    lm_grid g;
    int row0, rows_owned, rows_local;    // strip geometry (single GPU: 0, ncy, ncy)
    int n;                               // anchors = owned particles (ghost-row particles are partners only)
    float r2_lo, r2_hi;
    double r2;
    uint32_t seed_lo, seed_hi, step_lo, step_hi;
    unsigned long long thr[3];           // ceil(p * 2^53) for pRS, pPR, pSP
    uint32_t *__restrict__ hits;
    int4 *__restrict__ meta;
    int2 *__restrict__ pairs;
    unsigned long long cap_hits, cap_pairs;
    Counters *ctr;
};

__device__ __forceinline__ int cell_coord2(float v, double origin, double inv_h, int n)
{
    const double q = floor(__dmul_rn(__dsub_rn((double)v, origin), inv_h));   // == bin.cu::cell_coord
    if (!(q >= 0.0)) return 0;
    if (q >= (double)n) return n - 1;
    return (int)q;
}

__device__ __forceinline__ bool within_exact(float xa, float ya, float xb, float yb, double r2)
{
    const double dx = __dsub_rn((double)xa, (double)xb), dy = __dsub_rn((double)ya, (double)yb);
    const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    return s <= r2;
}

__device__ __forceinline__ bool within(const FindArgs &A, float xa, float ya, int b)
{
    const float xb = __ldg(A.lon + b), yb = __ldg(A.lat + b);
    const float dx = xa - xb, dy = ya - yb;
    const float d2 = fmaf(dx, dx, dy * dy);
    if (d2 > A.r2_hi) return false;
    if (d2 < A.r2_lo) return true;
    return within_exact(xa, ya, xb, yb, A.r2);
}

// three decision bits of the pair's draw: bit k set <=> u < p_k  (k = 0: pRS, 1: pPR, 2: pSP)
__device__ __forceinline__ uint32_t decision_bits(const FindArgs &A, int i, int j)
{
    uint32_t x[4];
    philox4x32_10((uint32_t)i, (uint32_t)j, A.step_lo, A.step_hi, A.seed_lo, A.seed_hi, x);
    const unsigned long long m = ((unsigned long long)(x[0] >> 5) << 26) | (unsigned long long)(x[1] >> 6);
    return (m < A.thr[0] ? 1u : 0u) | (m < A.thr[1] ? 2u : 0u) | (m < A.thr[2] ? 4u : 0u);
}

template <bool DO_RPS, bool EMIT>
__global__ void __launch_bounds__(FIND_THREADS) find_pairs_kernel(FindArgs A)
{
    __shared__ uint32_t s_b[FIND_WARPS][STAGE];
    __shared__ uint8_t s_owner[FIND_WARPS][STAGE];
    __shared__ unsigned int s_wtot[FIND_WARPS];
    __shared__ unsigned long long s_base;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * FIND_THREADS + threadIdx.x;
    const bool valid = a < A.n;

    float xa = 0.f, ya = 0.f;
    int my_id = 0;
    int r1_beg = 0, r1_end = 0, r2_beg = 0, r2_end = 0, sE = 0, sN = 0, sNE = 0;
    if (valid) {
        xa = __ldg(A.lon + a); ya = __ldg(A.lat + a);
        my_id = __ldg(A.id + a);
        const int ncx = A.g.ncx, ncy = A.rows_local;
        const int cx = cell_coord2(xa, A.g.x0, A.g.inv_h, ncx);
        const int cy = max(0, min(cell_coord2(ya, A.g.y0, A.g.inv_h, A.g.ncy) - A.row0, A.rows_owned - 1));   // == bin.cu
        const int c = cy * ncx + cx;
        const bool e_ok = cx + 1 < ncx;
        sE = __ldg(A.cell_start + c + 1);
        r1_beg = a + 1;
        r1_end = e_ok ? __ldg(A.cell_start + c + 2) : sE;
        if (cy + 1 < ncy) {
            const int up = c + ncx;
            sN = __ldg(A.cell_start + up);
            sNE = __ldg(A.cell_start + up + 1);
            r2_beg = (cx > 0) ? __ldg(A.cell_start + up - 1) : sN;
            r2_end = e_ok ? __ldg(A.cell_start + up + 2) : sNE;
        }
        if (r1_end - r1_beg > 65535 || r2_end - r2_beg > 65535) {     // 16-bit per-direction counts
            atomicAdd(&A.ctr->n_overflow, 1ull);
            r1_end = min(r1_end, r1_beg + 65535);
            r2_end = min(r2_end, r2_beg + 65535);
        }
    }

    // ---- pass 1: count hits per direction (S, E | NW, N, NE), 16 bits each
    unsigned long long cnt03 = 0;     // directions 0..3
    unsigned int cnt4 = 0;
    for (int b = r1_beg; b < r1_end; ++b)
        if (within(A, xa, ya, b)) cnt03 += (b >= sE) ? (1ull << 16) : 1ull;
    for (int b = r2_beg; b < r2_end; ++b)
        if (within(A, xa, ya, b)) {
            if (b >= sNE) ++cnt4;
            else cnt03 += (b >= sN) ? (1ull << 48) : (1ull << 32);
        }
    const unsigned int c0 = (unsigned int)(cnt03 & 0xffff), c1 = (unsigned int)((cnt03 >> 16) & 0xffff);
    const unsigned int c2 = (unsigned int)((cnt03 >> 32) & 0xffff), c3 = (unsigned int)(cnt03 >> 48);
    const unsigned int tot = c0 + c1 + c2 + c3 + cnt4;

    // ---- allocate: exclusive scan over the CTA, one atomic per CTA
    unsigned int inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const unsigned int wtot = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31) s_wtot[warp] = wtot;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int btot = 0;
#pragma unroll
        for (int w = 0; w < FIND_WARPS; ++w) btot += s_wtot[w];
        s_base = btot ? atomicAdd(&A.ctr->n_pairs, (unsigned long long)btot) : 0ull;
    }
    __syncthreads();
    unsigned long long wbase = s_base;
    for (int w = 0; w < warp; ++w) wbase += s_wtot[w];
    const unsigned int off = inc - tot;                    // within the warp
    if (DO_RPS && valid)
        A.meta[a] = make_int4((int)(unsigned int)(wbase + off), (int)(c0 | (c1 << 16)), (int)(c2 | (c3 << 16)), (int)cnt4);
    if (wtot == 0) return;                                 // warp-uniform
    if (!DO_RPS && !EMIT) return;

    if (wtot <= (unsigned int)STAGE) {
        // ---- pass 2 (staged): regenerate the hits into shared memory, then finish them densely
        unsigned int k = off;
        for (int b = r1_beg; b < r1_end; ++b)
            if (within(A, xa, ya, b)) { s_b[warp][k] = (uint32_t)b; s_owner[warp][k] = (uint8_t)lane; ++k; }
        for (int b = r2_beg; b < r2_end; ++b)
            if (within(A, xa, ya, b)) { s_b[warp][k] = (uint32_t)b; s_owner[warp][k] = (uint8_t)lane; ++k; }
        __syncwarp();
        for (unsigned int e0 = 0; e0 < wtot; e0 += 32) {
            const unsigned int e = e0 + lane;
            const bool act = e < wtot;
            const uint32_t b = act ? s_b[warp][e] : 0u;
            const int owner = act ? (int)s_owner[warp][e] : 0;
            const int ia = __shfl_sync(0xffffffffu, my_id, owner);
            if (act) {
                const int ib = __ldg(A.id + b);
                const int i = min(ia, ib), j = max(ia, ib);
                const unsigned long long g = wbase + e;
                if (DO_RPS && g < A.cap_hits) A.hits[g] = b | (decision_bits(A, i, j) << 29);
                if (EMIT && g < A.cap_pairs) A.pairs[g] = make_int2(i, j);
            }
        }
    } else {
        // ---- pass 2 (direct): a dense neighbourhood; every lane has many hits, finish them in place
        unsigned long long g = wbase + off;
        for (int pass = 0; pass < 2; ++pass) {
            const int beg = pass ? r2_beg : r1_beg, end = pass ? r2_end : r1_end;
            for (int b = beg; b < end; ++b)
                if (within(A, xa, ya, b)) {
                    const int ib = __ldg(A.id + b);
                    const int i = min(my_id, ib), j = max(my_id, ib);
                    if (DO_RPS && g < A.cap_hits) A.hits[g] = (uint32_t)b | (decision_bits(A, i, j) << 29);
                    if (EMIT && g < A.cap_pairs) A.pairs[g] = make_int2(i, j);
                    ++g;
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
enum UnitMode { MODE_SAME = 1, MODE_EAST = 2, MODE_CROSS = 3 };

struct ResolveArgs {
    int8_t *sp;
    const int32_t *__restrict__ cell_start;
    const int4 *__restrict__ meta;
    const uint32_t *__restrict__ hits;
    unsigned long long cap_hits;
    int ncx, ncy;
    long long n_units;
    int mode, parity, dir;   // dir in {-1,0,+1} for MODE_CROSS
    int d_idx;               // which of the five per-particle hit lists this phase consumes
};

__device__ __forceinline__ bool decode_unit(const ResolveArgs &A, long long u, int &anchor, int &other)
{
    const int ncx = A.ncx;
    if (A.mode == MODE_SAME) {
        anchor = other = (int)u;
    } else if (A.mode == MODE_EAST) {
        const int half = (ncx - A.parity) / 2;          // anchors per row with cx % 2 == parity, cx + 1 < ncx
        const int cy = (int)(u / half), i = (int)(u - (long long)cy * half);
        anchor = cy * ncx + 2 * i + A.parity;
        other = anchor + 1;
    } else {
        const int ry = (int)(u / ncx), cx = (int)(u - (long long)ry * ncx);
        const int cy = 2 * ry + A.parity;                // cy + 1 < ncy by construction of n_units
        const int ox = cx + A.dir;
        if (ox < 0 || ox >= ncx) return false;
        anchor = cy * ncx + cx;
        other = anchor + ncx + A.dir;
    }
    return true;
}

// offset and length of particle a's hit list for direction d
__device__ __forceinline__ void hit_list(const int4 m, int d, unsigned int &off, unsigned int &n)
{
    const unsigned int c0 = (unsigned int)m.y & 0xffffu, c1 = (unsigned int)m.y >> 16;
    const unsigned int c2 = (unsigned int)m.z & 0xffffu, c3 = (unsigned int)m.z >> 16;
    const unsigned int c4 = (unsigned int)m.w;
    off = (unsigned int)m.x;
    switch (d) {
        case 0: n = c0; break;
        case 1: off += c0; n = c1; break;
        case 2: off += c0 + c1; n = c2; break;
        case 3: off += c0 + c1 + c2; n = c3; break;
        default: off += c0 + c1 + c2 + c3; n = c4; break;
    }
}

// interactions.py:13-40 for species s1 != s2, both in {1,2,3}: the species both end up with.
// The forward winner (rock beats scissors, paper beats rock, scissors beats paper) wins iff its
// decision bit is set.
__device__ __forceinline__ int rps_apply(int s1, int s2, uint32_t dec)
{
    int d = s1 - s2;
    if (d < 0) d += 3;
    const int w = (d == 1) ? s1 : s2, l = (d == 1) ? s2 : s1;
    return ((dec >> (w - 1)) & 1u) ? w : l;
}

__device__ __forceinline__ bool is_rps(int s) { return s >= 1 && s <= 3; }

// 3->3 maps packed 2 bits per entry: M(s) = (M >> 2(s-1)) & 3
constexpr uint32_t MAP_ID = 1u | (2u << 2) | (3u << 4);
__device__ __forceinline__ uint32_t map_apply(uint32_t M, int s) { return (M >> (2 * (s - 1))) & 3u; }
__device__ __forceinline__ uint32_t map_compose(uint32_t second, uint32_t first)   // s -> second(first(s))
{
    return map_apply(second, (int)map_apply(first, 1)) | (map_apply(second, (int)map_apply(first, 2)) << 2) |
           (map_apply(second, (int)map_apply(first, 3)) << 4);
}

// Whole-warp resolution of one unit (all lanes call this with the same arguments).
__device__ void resolve_unit_warp(const ResolveArgs &A, int d_idx, int aBeg, int aEnd)
{
    const int lane = threadIdx.x & 31;
    for (int a = aBeg; a < aEnd; ++a) {
        unsigned int off, n;
        hit_list(A.meta[a], d_idx, off, n);
        if (n == 0) continue;
        __syncwarp();                                  // species written by earlier rows are visible
        int sa = ((volatile int8_t *)A.sp)[a];
        if (!is_rps(sa)) continue;                     // winner = None for every pair of this row
        const int sa0 = sa;
        for (unsigned int k0 = 0; k0 < n; k0 += 32) {
            const unsigned int k = k0 + lane;
            const bool act = k < n && (unsigned long long)off + k < A.cap_hits;
            uint32_t M = MAP_ID, dec = 0;
            int b = 0, sb = 0;
            if (act) {
                const uint32_t h = A.hits[off + k];
                b = (int)(h & HIT_MASK); dec = h >> 29;
                sb = ((volatile int8_t *)A.sp)[b];
                if (is_rps(sb)) {
                    M = 0;
#pragma unroll
                    for (int s = 1; s <= 3; ++s) M |= (uint32_t)((s == sb) ? s : rps_apply(s, sb, dec)) << (2 * (s - 1));
                }
            }
            uint32_t P = M;                            // inclusive scan of maps in lane (= id_b) order
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, P, d);
                if (lane >= d) P = map_compose(P, t);
            }
            uint32_t E = __shfl_up_sync(0xffffffffu, P, 1);
            if (lane == 0) E = MAP_ID;
            if (act && is_rps(sb)) {
                const int s_before = (int)map_apply(E, sa);
                if (s_before != sb) A.sp[b] = (int8_t)rps_apply(s_before, sb, dec);
            }
            sa = (int)map_apply(__shfl_sync(0xffffffffu, P, 31), sa);
        }
        if (lane == 0 && sa != sa0) A.sp[a] = (int8_t)sa;
    }
    __syncwarp();
}

// One unit = the pairs between the particles of `anchor` and those of `other` (or inside `anchor`), resolved in
// (id_a, id_b) order.  Called by all 32 lanes of a warp (valid = this lane has a unit): light units are walked
// by their lane, dense ones by the whole warp, one after the other.
__device__ __forceinline__ void resolve_unit(const ResolveArgs &A, bool valid, int anchor, int other, int d_idx)
{
    int aBeg = 0, aEnd = 0;
    bool heavy = false;
    if (valid) {
        aBeg = __ldg(A.cell_start + anchor);
        aEnd = __ldg(A.cell_start + anchor + 1);
        if (aEnd > aBeg) {
            const int nb = (anchor == other) ? (aEnd - aBeg) : (__ldg(A.cell_start + other + 1) - __ldg(A.cell_start + other));
            if (nb == 0 || (anchor == other && nb < 2)) aEnd = aBeg;
            else heavy = (long long)(aEnd - aBeg) * nb > HEAVY_TESTS;
        }
    }
    if (!heavy) {
        // one lane, one unit: rows in id order, each row's hits in id order
        for (int a = aBeg; a < aEnd; ++a) {
            unsigned int off, n;
            hit_list(__ldg(A.meta + a), d_idx, off, n);
            if (n == 0) continue;
            int sa = A.sp[a];
            const int sa0 = sa;
            for (unsigned int k = 0; k < n && (unsigned long long)off + k < A.cap_hits; ++k) {
                const uint32_t h = __ldg(A.hits + off + k);
                const int b = (int)(h & HIT_MASK);
                const int sb = A.sp[b];
                if (sa != sb && is_rps(sa) && is_rps(sb)) {
                    sa = rps_apply(sa, sb, h >> 29);
                    A.sp[b] = (int8_t)sa;
                }
            }
            if (sa != sa0) A.sp[a] = (int8_t)sa;
        }
    }
    // dense clusters: the warp resolves them together, one after the other
    unsigned hm = __ballot_sync(0xffffffffu, heavy);
    while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        resolve_unit_warp(A, d_idx, __shfl_sync(0xffffffffu, aBeg, src), __shfl_sync(0xffffffffu, aEnd, src));
    }
}

// One phase per launch, one lane per unit: the general path (any grid shape).
__global__ void __launch_bounds__(256) resolve_phase_kernel(ResolveArgs A)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int anchor = 0, other = 0;
    const bool valid = u < A.n_units && decode_unit(A, u, anchor, other);
    resolve_unit(A, valid, anchor, other, A.d_idx);
}

// Three phases per launch for grids with many long rows.  The phases of one group couple cells of ONE row
// (group 0: phases 0,1,2 = same cell, east even, east odd) or of ONE pair of rows (group 1: phases 3,4,5 =
// NW, N, NE from an even anchor row; group 2: phases 6,7,8 from an odd one), so a CTA that owns the row
// (pair) can run its three phases back to back with a block barrier in between: 3 launches instead of 9,
// and every particle's meta / hit list is read once per group instead of once per phase.
constexpr int ROW_THREADS = 1024;
template <int GROUP>
__global__ void __launch_bounds__(ROW_THREADS) resolve_rows_kernel(ResolveArgs A)
{
    const int ncx = A.ncx;
    const int cy = (GROUP == 0) ? (int)blockIdx.x : 2 * (int)blockIdx.x + (GROUP - 1);
    const int row = cy * ncx;
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
        const int n_units = (GROUP == 0 && ph > 0) ? (ncx - (ph - 1)) / 2 : ncx;
        const int d_idx = (GROUP == 0) ? (ph > 0 ? 1 : 0) : 2 + ph;
        for (int u0 = 0; u0 < n_units; u0 += blockDim.x) {     // block-uniform trip count
            const int u = u0 + (int)threadIdx.x;
            bool valid = u < n_units;
            int anchor = 0, other = 0;
            if (valid) {
                if (GROUP == 0) {
                    anchor = (ph == 0) ? row + u : row + 2 * u + (ph - 1);
                    other = (ph == 0) ? anchor : anchor + 1;
                } else {
                    const int ox = u + ph - 1;
                    valid = ox >= 0 && ox < ncx;
                    anchor = row + u;
                    other = row + ncx + ox;
                }
            }
            resolve_unit(A, valid, anchor, other, d_idx);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
cudaError_t launch_find(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int n, double r,
                        const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s)
{
    h->rps_cap = rps ? h->max_pairs : -1;
    if (n <= 0) return cudaSuccess;
    FindArgs F;
    F.lon = lon; F.lat = lat; F.id = id; F.cell_start = h->cell_start;
    F.g = h->grid; F.n = n;
    F.row0 = h->strip.row0; F.rows_owned = h->strip.rows_owned; F.rows_local = h->strip.rows_local;
    F.r2 = r * r;
    F.r2_lo = (float)(F.r2 * (1.0 - 4e-6));
    F.r2_hi = (float)(F.r2 * (1.0 + 4e-6));
    F.seed_lo = F.seed_hi = F.step_lo = F.step_hi = 0;
    F.thr[0] = F.thr[1] = F.thr[2] = 0;
    if (rps) {
        F.seed_lo = rps->seed_lo; F.seed_hi = rps->seed_hi; F.step_lo = rps->step_lo; F.step_hi = rps->step_hi;
        const double p[3] = {rps->pRS, rps->pPR, rps->pSP};
        for (int k = 0; k < 3; ++k) {
            // u = m * 2^-53 with integer m < 2^53:  u < p  <=>  m < ceil(p * 2^53)   (the scaling is exact)
            if (!(p[k] > 0.0)) F.thr[k] = 0;
            else if (p[k] >= 1.0) F.thr[k] = 1ull << 53;
            else F.thr[k] = (unsigned long long)ceil(p[k] * 9007199254740992.0);
        }
    }
    F.hits = h->hits; F.meta = h->meta;
    F.pairs = pairs_out;
    F.cap_hits = rps ? (unsigned long long)h->max_pairs : 0ull;
    F.cap_pairs = (pairs_out && cap > 0) ? (unsigned long long)cap : 0ull;
    F.ctr = h->ctr;
    const bool emit = F.cap_pairs > 0;
    const int grid = (n + FIND_THREADS - 1) / FIND_THREADS;
    if (rps) {
        if (emit) find_pairs_kernel<true, true><<<grid, FIND_THREADS, 0, s>>>(F);
        else find_pairs_kernel<true, false><<<grid, FIND_THREADS, 0, s>>>(F);
    } else {
        if (emit) find_pairs_kernel<false, true><<<grid, FIND_THREADS, 0, s>>>(F);
        else find_pairs_kernel<false, false><<<grid, FIND_THREADS, 0, s>>>(F);
    }
    ++h->launches;
    return cudaGetLastError();
}

// phases [first, last] of the canonical cell-phase order on the local rows.  Strip boundaries sit on even
// global rows, so local and global row parities agree; same-cell and east units cover the owned rows,
// cross-row units may reach into the ghost row (local row rows_owned), only in phases 6-8.
cudaError_t launch_resolve_phases(lm_handle_s *h, int8_t *sp, int first, int last, cudaStream_t s)
{
    if (h->rps_cap < 0) return cudaSuccess;
    ResolveArgs R;
    R.sp = sp; R.cell_start = h->cell_start; R.meta = h->meta; R.hits = h->hits;
    R.cap_hits = (unsigned long long)h->max_pairs;
    R.ncx = h->grid.ncx; R.ncy = h->strip.rows_local;
    R.mode = R.parity = R.dir = R.d_idx = 0; R.n_units = 0;
    const long long ncx = R.ncx, rows_owned = h->strip.rows_owned, rows_local = h->strip.rows_local;
    // row-fused path: whole groups of three phases, enough rows to fill the GPU, rows long enough for a CTA
    const bool fused = h->resolve_mode == 1 ||
                       (h->resolve_mode == 0 && rows_owned >= 2 * kNumSMs && ncx >= ROW_THREADS / 2);
    int ph = first;
    while (ph <= last) {
        cudaError_t e;
        if (fused && ph % 3 == 0 && ph + 2 <= last) {
            const int group = ph / 3;
            const int threads = (int)(ncx >= ROW_THREADS ? ROW_THREADS : ((ncx + 31) / 32) * 32);
            const long long rows = group == 0 ? rows_owned : (rows_local - (group - 1)) / 2;
            if (rows > 0) {
                if (group == 0) resolve_rows_kernel<0><<<(unsigned)rows, threads, 0, s>>>(R);
                else if (group == 1) resolve_rows_kernel<1><<<(unsigned)rows, threads, 0, s>>>(R);
                else resolve_rows_kernel<2><<<(unsigned)rows, threads, 0, s>>>(R);
                ++h->launches;
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
            }
            ph += 3;
            continue;
        }
        long long n_units;
        if (ph == 0) { R.mode = MODE_SAME; R.parity = 0; R.dir = 0; R.d_idx = 0; n_units = ncx * rows_owned; }
        else if (ph <= 2) {
            R.mode = MODE_EAST; R.parity = ph - 1; R.dir = 0; R.d_idx = 1;
            n_units = ((ncx - R.parity) / 2) * rows_owned;
        } else {
            R.mode = MODE_CROSS; R.parity = (ph - 3) / 3; R.dir = (ph - 3) % 3 - 1; R.d_idx = 3 + R.dir;
            n_units = ((rows_local - R.parity) / 2) * ncx;
        }
        ++ph;
        if (n_units <= 0) continue;
        R.n_units = n_units;
        resolve_phase_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, s>>>(R);
        ++h->launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_pairs(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                         double r, const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s)
{
    cudaError_t e = launch_find(h, lon, lat, id, n, r, rps, pairs_out, cap, s);
    if (e != cudaSuccess || !rps || n <= 0) return e;
    return launch_resolve_phases(h, sp, 0, 8, s);
}

}  // namespace lm
