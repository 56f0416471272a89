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
// The other Minkowski norms SciPy evaluates without pow() (interaction_norm, interaction_simulator.py:27,98)
// are template variants of the same kernel (LM_OPT_NORM):  p=1  fl(|dx| + |dy|) <= r,  p=inf  max(|dx|, |dy|) <= r.
// Every norm with p >= 1 bounds |dx| and |dy| by r, so the same half stencil of cells with edge >= r finds them.
//
// Two stages:
//
//  find_pairs_kernel   one thread per particle a, ONE pass over its half stencil (rest of its own cell,
//      E, NW, N, NE): hits are staged per lane in shared memory, then finished densely by the whole warp --
//      everything that does not depend on species is done here, once: the per-pair Philox draw reduced to
//      three decision bits (u < pRS, u < pPR, u < pSP, as exact integer compares of the 53-bit draw against
//      ceil(p * 2^53)), the pair list (ids, i < j), and the hand-off to the resolver.
//
//      Hand-off layout, built for the resolver's access pattern.  A SEGMENT is the run of particles of one
//      cell inside one 32-particle chunk (= one warp of this kernel); a cell has one segment, two if it
//      straddles a chunk boundary (9 %), more only in dense clusters.  hits[] is allocated in blocks of one
//      CTA (256 particles); inside a block the entries are DIRECTION-MAJOR (0 same cell, 1 E, 2 NW, 3 N, 4 NE),
//      then by warp, segment, anchor, partner -- so the pairs of one resolver unit (cell x direction) are one
//      contiguous stream in canonical (a, b) order, and a resolver phase, which consumes one direction,
//      streams 1/5 of the array instead of dragging every sector through the L2.
//          entry = b_rel (24 bits, index relative to the partner cell's start) | a_rel (5 bits, anchor index
//                  inside the segment) << 24 | decision bits << 29
//          rec[d][cell]   = (first entry, count) of the cell's FIRST segment in direction d
//          rec2[d][chunk] = the same for the segment that continues a cell at the start of a chunk
//      Both tables are read coalesced (consecutive lanes <-> consecutive cells).
//
//  resolve kernels   the sequential part.  A *unit* is a cell (pairs inside it) or two adjacent cells;
//      units of one *phase* touch disjoint particles:
//          phase 0        same cell
//          phase 1,2      E neighbour, anchor cx even / odd
//          phase 3,4,5    NW, N, NE neighbour, anchor cy even
//          phase 6,7,8    NW, N, NE neighbour, anchor cy odd
//      so each phase is conflict-free; a lane walks the entry streams of its units (eight per lane, back to
//      back, as a small state machine so that lanes with short units do not wait for long ones) and applies
//      interactions.py:13-40 with table look-ups: the reference's sequential in-place semantics under the
//      canonical total order (phase, unit, id_a, id_b) (oracle/rps.py::cell_phase_order).  Units with more
//      than HEAVY_TESTS candidate pairs (dense clusters) are resolved by the whole warp: the hits of one
//      anchor are independent except through the anchor's species, a 3-state value, so a run of them is a
//      prefix scan over 3->3 maps.
#include <algorithm>

#include "lm_internal.cuh"
#include "philox.cuh"

namespace lm {

constexpr int FIND_THREADS = 256;
constexpr int FIND_WARPS = FIND_THREADS / 32;
constexpr uint32_t B_REL_MASK = (1u << 24) - 1u;     // entry: b_rel | a_rel << 24 | decision bits << 29
constexpr unsigned int HEAVY_MIN = 160;              // a segment below this many pairs is never worth a whole warp
constexpr int CELL_BITS = 21;                        // particles per cell < 2^21 (packed per-direction hit counters)

struct FindArgs {
    const float *__restrict__ lon;
    const float *__restrict__ lat;
    const int32_t *__restrict__ id;
    const int32_t *__restrict__ cell_start;
    lm_grid g;
    int row0, rows_owned, rows_local;    // strip geometry (single GPU: 0, ncy, ncy)
    int n;                               // anchors = owned particles (ghost-row particles are partners only)
    int force_two_pass;                  // LM_OPT_FIND_PATH = 1 (tests)
    float r2_lo, r2_hi;                  // float32 pre-filter window around the threshold (r*r for p=2, r for p=1 and p=inf)
    double r2;                           // the exact threshold
    uint32_t pair_key;                   // key of this step's per-pair Philox2x32 stream (philox.cuh)
    unsigned long long thr[3];           // ceil(p * 2^53) for pRS, pPR, pSP
    uint32_t *__restrict__ hits;
    uint2 *__restrict__ rec;             // [5][rec_stride]
    uint2 *__restrict__ rec2;            // [5][rec2_stride]
    long long rec_stride, rec2_stride;
    int2 *__restrict__ pairs;
    unsigned long long cap_words, cap_pairs;
    Counters *ctr;
    // LM_OPT_INTERACT_MODE = 2 (hybrid): HEAVY units (csrc/interact.cu::unit_is_light, part of the cell-round order) are left
    // out here and queued per phase for interact_heavy_kernel
    int hybrid;
    int2 *heavy_list;                    // [9][2][heavy_cap]
    unsigned int *heavy_cnt;             // [9][4]
    unsigned int heavy_cap;
    unsigned long long heavy_min;        // a unit is heavy when its candidate pairs (m_a * m_b; one cell: m (m - 1) / 2) exceed this
};

__device__ __forceinline__ int cell_coord2(float v, double origin, double inv_h, int n)
{
    const double q = floor(__dmul_rn(__dsub_rn((double)v, origin), inv_h));   // == bin.cu::cell_coord
    if (!(q >= 0.0)) return 0;
    if (q >= (double)n) return n - 1;
    return (int)q;
}

template <int NORM>
__device__ __forceinline__ bool within_exact(float xa, float ya, float xb, float yb, double thr)
{
    const double dx = __dsub_rn((double)xa, (double)xb), dy = __dsub_rn((double)ya, (double)yb);
    if (NORM == LM_NORM_2) return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= thr;
    if (NORM == LM_NORM_1) return __dadd_rn(fabs(dx), fabs(dy)) <= thr;
    return fmax(fabs(dx), fabs(dy)) <= thr;
}

template <int NORM>
__device__ __forceinline__ bool within(const FindArgs &A, float xa, float ya, int b)
{
    const float xb = __ldg(A.lon + b), yb = __ldg(A.lat + b);
    const float dx = xa - xb, dy = ya - yb;
    const float d2 = NORM == LM_NORM_2 ? fmaf(dx, dx, dy * dy) : (NORM == LM_NORM_1 ? fabsf(dx) + fabsf(dy) : fmaxf(fabsf(dx), fabsf(dy)));
    if (d2 > A.r2_hi) return false;
    if (d2 < A.r2_lo) return true;
    return within_exact<NORM>(xa, ya, xb, yb, A.r2);
}

// three decision bits of the pair's draw: bit k set <=> u < p_k  (k = 0: pRS, 1: pPR, 2: pSP)
__device__ __forceinline__ uint32_t decision_bits(const FindArgs &A, int i, int j)
{
    const unsigned long long m = pair_draw_m((uint32_t)i, (uint32_t)j, A.pair_key);
    return (m < A.thr[0] ? 1u : 0u) | (m < A.thr[1] ? 2u : 0u) | (m < A.thr[2] ? 4u : 0u);
}

__device__ __forceinline__ unsigned int warp_incl_scan(unsigned int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// bits [0, t) of a 64-bit mask
__device__ __forceinline__ unsigned long long bits_below(int t)
{
    return t <= 0 ? 0ull : (t >= 64 ? ~0ull : ((1ull << t) - 1ull));
}

constexpr int FILL_CAP = 1024;       // hits of one warp that the dense finishing stage can take
constexpr int OWN_WORDS = 12;        // per lane, for the finishing stage: rel[5], base0, sE, begNW, sN, sNE, beg0, n1

template <bool DO_RPS, bool EMIT, int NORM>
__global__ void __launch_bounds__(FIND_THREADS, 5) find_pairs_kernel(FindArgs A)
{
    __shared__ uint16_t s_fill_all[FIND_WARPS][FILL_CAP];      // compact hit e of the warp -> owner lane | candidate index << 5
    __shared__ uint32_t s_own_all[FIND_WARPS][OWN_WORDS][32];
    __shared__ unsigned int s_wtot[FIND_WARPS][5];             // hits of warp w in direction d
    __shared__ unsigned int s_dbase[FIND_WARPS][5];            // first slot of (warp, direction) inside the CTA's block
    __shared__ unsigned long long s_base;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * FIND_THREADS + threadIdx.x;
    const bool valid = a < A.n;
    uint16_t *s_fill = s_fill_all[warp];
    uint32_t (*s_own)[32] = s_own_all[warp];

    float xa = 0.f, ya = 0.f;
    int my_id = 0, c = -1 - lane;        // invalid lanes get distinct negative "cells"
    // candidate ranges per direction: same cell [beg0, sE), E [sE, endE), NW [begNW, sN), N [sN, sNE), NE [sNE, endNE)
    int beg0 = 0, sE = 0, endE = 0, begNW = 0, sN = 0, sNE = 0, endNE = 0, base0 = 0;
    if (valid) {
        xa = __ldg(A.lon + a); ya = __ldg(A.lat + a);
        my_id = __ldg(A.id + a);
        const int ncx = A.g.ncx, ncy = A.rows_local;
        const int cx = cell_coord2(xa, A.g.x0, A.g.inv_h, ncx);
        const int cy = max(0, min(cell_coord2(ya, A.g.y0, A.g.inv_h, A.g.ncy) - A.row0, A.rows_owned - 1));   // == bin.cu
        c = cy * ncx + cx;
        const bool e_ok = cx + 1 < ncx;
        base0 = __ldg(A.cell_start + c);
        sE = __ldg(A.cell_start + c + 1);
        beg0 = a + 1;
        endE = e_ok ? __ldg(A.cell_start + c + 2) : sE;
        begNW = sN = sNE = endNE = endE;                 // no row to the north: empty ranges that keep the starts monotone
        if (cy + 1 < ncy) {
            const int up = c + ncx;
            sN = __ldg(A.cell_start + up);
            sNE = __ldg(A.cell_start + up + 1);
            begNW = (cx > 0) ? __ldg(A.cell_start + up - 1) : sN;
            endNE = e_ok ? __ldg(A.cell_start + up + 2) : sNE;
        }
        constexpr int CELL_MAX = (1 << CELL_BITS) - 1;
        if (sE - base0 > CELL_MAX || endE - sE > CELL_MAX || sN - begNW > CELL_MAX || sNE - sN > CELL_MAX ||
            endNE - sNE > CELL_MAX) {                                          // packed counters / relative partner index
            atomicAdd(&A.ctr->n_overflow, 1ull);
            beg0 = sE = endE = begNW = sN = sNE = endNE = 0;
        }
    }

    // ---- hybrid mode: which of my five units are heavy?  (two cells: m_a * m_b > 1,024; one cell: m (m - 1) / 2 > 1,024.)  Their
    // hits are dropped below; the first microbe of the cell queues them for the rounds-of-matchings kernel.
    unsigned int heavy_dirs = 0;
    if (A.hybrid && valid) {
        const unsigned int ma = (unsigned int)(sE - base0);
        const unsigned int mb[5] = {ma, (unsigned int)(endE - sE), (unsigned int)(sN - begNW), (unsigned int)(sNE - sN),
                                    (unsigned int)(endNE - sNE)};
        if ((unsigned long long)ma * (ma - 1u) / 2ull > A.heavy_min) heavy_dirs |= 1u;
#pragma unroll
        for (int d = 1; d < 5; ++d)
            if ((unsigned long long)ma * mb[d] > A.heavy_min) heavy_dirs |= 1u << d;
        if (heavy_dirs && a == base0) {
            const int ncx = A.g.ncx, cx = c % ncx, cy = c / ncx;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                if (!((heavy_dirs >> d) & 1u)) continue;
                const int other = d == 0 ? c : (d == 1 ? c + 1 : c + ncx + (d - 3));
                const int ph = d == 0 ? 0 : (d == 1 ? 1 + (cx & 1) : 3 * (cy & 1) + 1 + d);
                const unsigned long long M = d == 0 ? (unsigned long long)(ma + (ma & 1u)) : (unsigned long long)max(ma, mb[d]);
                const unsigned long long slots = d == 0 ? (M - 1ull) * (M >> 1) : (unsigned long long)ma * M;
                // the whole CTA takes it: many slots (interact.cu::IT_MEGA_MIN) or more microbes than a warp stages (HV_WARP_CAP)
                const int big = (slots >= 8192ull || (d == 0 ? ma : ma + mb[d]) > 256u) ? 1 : 0;
                const unsigned int k = atomicAdd(&A.heavy_cnt[4 * ph + big], 1u);
                if (k < A.heavy_cap) A.heavy_list[(size_t)(2 * ph + big) * A.heavy_cap + k] = make_int2(c, other);
                else atomicAdd(&A.ctr->n_heavy_overflow, 1u);
            }
        }
    }
    // ---- the candidates of a lane: the concatenation of two index ranges (same row: rest of its cell + E;
    // next row: NW, N, NE), i = 0 .. ntot-1.  ONE loop, so a warp iterates max-of-sums, not sum-of-maxes, of
    // its lanes' candidate counts; hits are recorded as bits of a 64-bit mask -- nothing else happens in the
    // loop.  Warps with a lane of more than 128 candidates (dense clusters) count per direction instead and
    // regenerate the hits in a second pass.
    const int n1 = endE - beg0, ntot = n1 + (endNE - begNW);
    const int t1 = sE - beg0, t3 = n1 + (sN - begNW), t4 = n1 + (sNE - begNW);     // first candidate of E | N | NE (NW: n1)
    const bool warp_big = __any_sync(0xffffffffu, ntot > 128) || A.force_two_pass;
    unsigned long long mask = 0, mask_hi = 0;                  // hit bits of candidates 0..63 | 64..127
    unsigned int cnt[5], tot;
    if (!warp_big && !__any_sync(0xffffffffu, ntot > 64)) {
        // the common case: every lane of the warp has at most 64 candidates
        int b = beg0;
        unsigned long long bit = 1ull;
        for (int i = 0; i < ntot; ++i) {
            if (i == n1) b = begNW;
            if (within<NORM>(A, xa, ya, b)) mask |= bit;
            ++b;
            bit <<= 1;
        }
        if (heavy_dirs) {                                      // hybrid mode: the hits of heavy units are not mine
            const int lo[5] = {0, t1, n1, t3, t4}, hi[5] = {t1, n1, t3, t4, ntot};
#pragma unroll
            for (int d = 0; d < 5; ++d)
                if ((heavy_dirs >> d) & 1u) mask &= ~(bits_below(hi[d]) & ~bits_below(lo[d]));
        }
        const unsigned int p1 = __popcll(mask & bits_below(t1)), p2 = __popcll(mask & bits_below(n1));
        const unsigned int p3 = __popcll(mask & bits_below(t3)), p4 = __popcll(mask & bits_below(t4));
        tot = __popcll(mask);
        cnt[0] = p1; cnt[1] = p2 - p1; cnt[2] = p3 - p2; cnt[3] = p4 - p3; cnt[4] = tot - p4;
    } else if (!warp_big) {
        int b = beg0;
        const int n_lo = min(ntot, 64);
        unsigned long long bit = 1ull;
        for (int i = 0; i < n_lo; ++i) {
            if (i == n1) b = begNW;
            if (within<NORM>(A, xa, ya, b)) mask |= bit;
            ++b;
            bit <<= 1;
        }
        bit = 1ull;
        for (int i = 64; i < ntot; ++i) {
            if (i == n1) b = begNW;
            if (within<NORM>(A, xa, ya, b)) mask_hi |= bit;
            ++b;
            bit <<= 1;
        }
        if (heavy_dirs) {
            const int lo[5] = {0, t1, n1, t3, t4}, hi[5] = {t1, n1, t3, t4, ntot};
#pragma unroll
            for (int d = 0; d < 5; ++d)
                if ((heavy_dirs >> d) & 1u) {
                    mask &= ~(bits_below(hi[d]) & ~bits_below(lo[d]));
                    mask_hi &= ~(bits_below(hi[d] - 64) & ~bits_below(lo[d] - 64));
                }
        }
        const unsigned int p1 = __popcll(mask & bits_below(t1)) + __popcll(mask_hi & bits_below(t1 - 64));
        const unsigned int p2 = __popcll(mask & bits_below(n1)) + __popcll(mask_hi & bits_below(n1 - 64));
        const unsigned int p3 = __popcll(mask & bits_below(t3)) + __popcll(mask_hi & bits_below(t3 - 64));
        const unsigned int p4 = __popcll(mask & bits_below(t4)) + __popcll(mask_hi & bits_below(t4 - 64));
        tot = __popcll(mask) + __popcll(mask_hi);
        cnt[0] = p1; cnt[1] = p2 - p1; cnt[2] = p3 - p2; cnt[3] = p4 - p3; cnt[4] = tot - p4;
    } else {
        unsigned long long acc0 = 0, acc1 = 0;                 // hit counters, CELL_BITS each: d0 d1 d2 | d3 d4
        for (int i = 0; i < ntot; ++i) {
            const int b = (i < n1) ? beg0 + i : begNW + (i - n1);
            const int d = (i >= t1) + (i >= n1) + (i >= t3) + (i >= t4);
            if (!((heavy_dirs >> d) & 1u) && within<NORM>(A, xa, ya, b)) {
                if (d < 3) acc0 += 1ull << (CELL_BITS * d); else acc1 += 1ull << (CELL_BITS * (d - 3));
            }
        }
        constexpr unsigned int CM = (1u << CELL_BITS) - 1u;
        cnt[0] = (unsigned int)acc0 & CM; cnt[1] = (unsigned int)(acc0 >> CELL_BITS) & CM;
        cnt[2] = (unsigned int)(acc0 >> (2 * CELL_BITS)) & CM;
        cnt[3] = (unsigned int)acc1 & CM; cnt[4] = (unsigned int)(acc1 >> CELL_BITS) & CM;
        tot = cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4];
    }

    // ---- layout: segments = runs of equal cell among the lanes; one block of hits[] per CTA, direction-major
    const int c_prev = __shfl_up_sync(0xffffffffu, c, 1);
    const bool head = valid && (lane == 0 || c != c_prev);
    const unsigned int heads = __ballot_sync(0xffffffffu, head);
    const unsigned int valid_mask = __ballot_sync(0xffffffffu, valid);
    unsigned int incl[5], total[5];
    unsigned int wpairs = 0;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        incl[d] = warp_incl_scan(cnt[d], lane);
        total[d] = __shfl_sync(0xffffffffu, incl[d], 31);
        wpairs += total[d];
    }
    if (lane < 5) s_wtot[warp][lane] = lane == 0 ? total[0] : (lane == 1 ? total[1] : (lane == 2 ? total[2] : (lane == 3 ? total[3] : total[4])));
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int w = 0; w < FIND_WARPS; ++w) { s_dbase[w][d] = run; run += s_wtot[w][d]; }
        s_base = run ? atomicAdd(&A.ctr->n_pairs, (unsigned long long)run) : 0ull;     // one atomic per CTA
    }
    __syncthreads();
    if (!DO_RPS && !EMIT) return;
    if (valid_mask == 0u || wpairs == 0) {                                  // warp-uniform
        // segments without pairs still need their (empty) records
        if (DO_RPS && head) {
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                if (a == base0) A.rec[d * A.rec_stride + c] = make_uint2(0u, 0u);
                else A.rec2[d * A.rec2_stride + (a >> 5)] = make_uint2(0u, 0u);
            }
        }
        return;
    }
    const unsigned long long cta_base = s_base;

    const unsigned int incl_tot = incl[0] + incl[1] + incl[2] + incl[3] + incl[4];
    const unsigned int excl_tot = incl_tot - tot;
    int H = 0;                                                               // head lane of my segment
    unsigned long long rel[5];                                               // hit k of this lane, of direction d, lands at rel[d] + k
    {
        const unsigned int below = heads & (0xffffffffu >> (31 - lane));     // heads at or below my lane
        H = below ? 31 - __clz(below) : 0;
        const unsigned int above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
        const int Hn = above ? __ffs(above) - 1 : 32;                        // head lane of the next segment
        unsigned int start_d = 0;
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const unsigned int excl = incl[d] - cnt[d];
            const unsigned long long first = cta_base + s_dbase[warp][d] + excl;   // my first hit of direction d
            rel[d] = first - start_d;
            start_d += cnt[d];
            if (DO_RPS) {
                const unsigned int pN_s = __shfl_sync(0xffffffffu, excl, Hn & 31);
                const unsigned int T = ((Hn < 32) ? pN_s : total[d]) - excl;    // for a head lane: the segment's count
                if (head) {
                    const uint2 r = make_uint2((uint32_t)first, T);
                    if (a == base0) A.rec[d * A.rec_stride + c] = r;              // the cell's first segment
                    else A.rec2[d * A.rec2_stride + (a >> 5)] = r;              // a cell continued from the previous chunk
                }
            }
        }
    }

    if (!warp_big && wpairs <= (unsigned int)FILL_CAP) {
        // ---- finish densely: lanes <-> hits of the whole warp (id look-up, Philox, hand-off entry, pair)
#pragma unroll
        for (int d = 0; d < 5; ++d) s_own[d][lane] = (uint32_t)(rel[d] - cta_base);  // relative to the CTA's block (may wrap: mod 2^32)
        s_own[5][lane] = (uint32_t)base0; s_own[6][lane] = (uint32_t)sE; s_own[7][lane] = (uint32_t)begNW;
        s_own[8][lane] = (uint32_t)sN; s_own[9][lane] = (uint32_t)sNE; s_own[10][lane] = (uint32_t)beg0;
        s_own[11][lane] = (uint32_t)n1;
        {
            unsigned int kk = excl_tot;
            for (unsigned int m = (unsigned int)mask; m; m &= m - 1) s_fill[kk++] = (uint16_t)(lane | ((__ffs((int)m) - 1) << 5));
            for (unsigned int m = (unsigned int)(mask >> 32); m; m &= m - 1) s_fill[kk++] = (uint16_t)(lane | ((__ffs((int)m) + 31) << 5));
            for (unsigned int m = (unsigned int)mask_hi; m; m &= m - 1) s_fill[kk++] = (uint16_t)(lane | ((__ffs((int)m) + 63) << 5));
            for (unsigned int m = (unsigned int)(mask_hi >> 32); m; m &= m - 1) s_fill[kk++] = (uint16_t)(lane | ((__ffs((int)m) + 95) << 5));
        }
        __syncwarp();
        for (unsigned int e0 = 0; e0 < wpairs; e0 += 32) {
            const unsigned int e = e0 + lane;
            const bool act = e < wpairs;
            const unsigned int f = act ? s_fill[e] : 0u;
            const int owner = (int)(f & 31u), i = (int)(f >> 5);
            const int ia = __shfl_sync(0xffffffffu, my_id, owner);
            const int oH = __shfl_sync(0xffffffffu, H, owner);
            const unsigned int oex = __shfl_sync(0xffffffffu, excl_tot, owner);
            if (act) {
                const int o_n1 = (int)s_own[11][owner];
                const int b = (i < o_n1) ? (int)s_own[10][owner] + i : (int)s_own[7][owner] + (i - o_n1);
                const int ib = __ldg(A.id + b);
                const int lo = min(ia, ib), hi = max(ia, ib);
                // direction from the partner's index: the cell starts are non-decreasing (empty ranges collapse)
                const int oE = (int)s_own[6][owner], oNW = (int)s_own[7][owner], oN = (int)s_own[8][owner], oNE = (int)s_own[9][owner];
                const int d = (b >= oE) + (b >= oNW) + (b >= oN) + (b >= oNE);
                const unsigned long long dst = cta_base + (uint32_t)(s_own[d][owner] + (e - oex));
                if (DO_RPS && dst < A.cap_words) {
                    const int bbase = d == 0 ? (int)s_own[5][owner] : (d == 1 ? oE : (d == 2 ? oNW : (d == 3 ? oN : oNE)));
                    A.hits[dst] = (uint32_t)(b - bbase) | ((uint32_t)(owner - oH) << 24) | (decision_bits(A, lo, hi) << 29);
                }
                if (EMIT && dst < A.cap_pairs) A.pairs[dst] = make_int2(lo, hi);
            }
        }
    } else {
        // ---- a dense neighbourhood: every lane has many hits, regenerate them and finish them in place
        const int base[5] = {base0, sE, begNW, sN, sNE};
        unsigned int kk = 0;
        for (int i = 0; i < ntot; ++i) {
            const int b = (i < n1) ? beg0 + i : begNW + (i - n1);
            const int d = (i >= t1) + (i >= n1) + (i >= t3) + (i >= t4);
            if (!((heavy_dirs >> d) & 1u) && within<NORM>(A, xa, ya, b)) {
                const int ib = __ldg(A.id + b);
                const int lo = min(my_id, ib), hi = max(my_id, ib);
                const unsigned long long r_d = d == 0 ? rel[0] : (d == 1 ? rel[1] : (d == 2 ? rel[2] : (d == 3 ? rel[3] : rel[4])));
                const unsigned long long dst = r_d + kk;
                if (DO_RPS && dst < A.cap_words) {
                    const int b_d = d == 0 ? base[0] : (d == 1 ? base[1] : (d == 2 ? base[2] : (d == 3 ? base[3] : base[4])));
                    A.hits[dst] = (uint32_t)(b - b_d) | ((uint32_t)(lane - H) << 24) | (decision_bits(A, lo, hi) << 29);
                }
                if (EMIT && dst < A.cap_pairs) A.pairs[dst] = make_int2(lo, hi);
                ++kk;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
enum UnitMode { MODE_SAME = 1, MODE_EAST = 2, MODE_CROSS = 3 };

constexpr int RES_THREADS = 128;
constexpr int MAX_UNITS_PER_LANE = 8;                      // a warp streams up to 256 consecutive units of one row

struct ResolveArgs {
    int8_t *sp;
    const int32_t *__restrict__ cell_start;
    const uint32_t *__restrict__ hits;
    const uint2 *__restrict__ rec;       // + d * rec_stride already applied
    const uint2 *__restrict__ rec2;      // + d * rec2_stride already applied
    const unsigned long long *n_pairs;   // pairs found by the search (snapshot: the counters are reset by the next step)
    unsigned long long cap_words;
    int ncx;
    int units_per_row, warps_per_row;
    int upl;                 // units per lane: 1, 2, 4 or 8 (fewer when the grid has few cells, to keep enough warps in flight)
    long long n_warps;
    int mode, parity, dir;   // dir in {-1,0,+1} for MODE_CROSS
    unsigned int heavy_min;  // a segment with fewer pairs than this is never handed to the whole warp (LM_OPT_RESOLVE_HEAVY_MIN)
};

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

// 3->3 maps of species, one byte per entry: byte s-1 holds M(s)-1 (byte 3 = 3, unused).  Composition is one PRMT:
// byte i of compose(second, first) = second[first[i]].
constexpr uint32_t MAP_ID = 0x03020100u;
__device__ __forceinline__ int map_apply(uint32_t M, int s) { return (int)((M >> (8 * (s - 1))) & 0xffu) + 1; }
__device__ __forceinline__ uint32_t map_compose(uint32_t second, uint32_t first)   // s -> second(first(s))
{
    const uint32_t sel = (first & 0xfu) | ((first >> 4) & 0xf0u) | ((first >> 8) & 0xf00u) | 0x3000u;
    return __byte_perm(second, 0u, sel);
}
// the map "interact with a partner of species sb (in 1..3), draw dec"
__device__ __forceinline__ uint32_t map_of_partner(int sb, uint32_t dec)
{
    uint32_t M = 0x03000000u;
#pragma unroll
    for (int q = 1; q <= 3; ++q) M |= (uint32_t)(((q == sb) ? q : rps_apply(q, sb, dec)) - 1) << (8 * (q - 1));
    return M;
}

// Whole-warp resolution of one unit (all lanes call this with the same arguments): its entry streams, segment
// after segment, in chunks of 32 entries; inside a chunk, one run of equal anchor after the other.
__device__ void resolve_unit_warp(const ResolveArgs &A, int cell, int cs0, int cs1, int oBeg)
{
    const int lane = threadIdx.x & 31;
    int cur_a = -1, sa = 0, sa0 = 0;
    __syncwarp();
    int a0 = cs0;
    uint2 R = __ldg(A.rec + cell);
    while (true) {
        const uint32_t *ent = A.hits + R.x;
        for (unsigned int k0 = 0; k0 < R.y; k0 += 32) {
            const unsigned int k = k0 + lane;
            const bool act = k < R.y;
            const uint32_t en = act ? __ldg(ent + k) : 0u;
            const int ar = (int)((en >> 24) & 31u);
            unsigned int todo = __ballot_sync(0xffffffffu, act);
            while (todo) {
                const int lead = __ffs(todo) - 1;
                const int ar0 = __shfl_sync(0xffffffffu, ar, lead);
                const unsigned int m = __ballot_sync(0xffffffffu, act && ar == ar0) & todo;   // entries are sorted by anchor
                todo &= ~m;
                const int a = a0 + ar0;
                if (a != cur_a) {
                    if (cur_a >= 0 && sa != sa0 && lane == 0) A.sp[cur_a] = (int8_t)sa;
                    __syncwarp();
                    cur_a = a;
                    sa = sa0 = ((volatile int8_t *)A.sp)[a];
                }
                if (!is_rps(sa)) continue;                     // winner = None for every pair of this anchor
                const bool mine = (m >> lane) & 1u;
                uint32_t M = MAP_ID, dec = 0;
                int b = 0, sb = 0;
                if (mine) {
                    b = oBeg + (int)(en & B_REL_MASK); dec = en >> 29;
                    sb = ((volatile int8_t *)A.sp)[b];
                    if (is_rps(sb)) M = map_of_partner(sb, dec);
                }
                uint32_t P = M;                                // inclusive scan of maps in lane (= id_b) order
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, P, dd);
                    if (lane >= dd) P = map_compose(P, t);
                }
                uint32_t E = __shfl_up_sync(0xffffffffu, P, 1);
                if (lane == 0) E = MAP_ID;
                if (mine && is_rps(sb)) {
                    const int s_before = map_apply(E, sa);
                    if (s_before != sb) A.sp[b] = (int8_t)rps_apply(s_before, sb, dec);
                }
                sa = map_apply(__shfl_sync(0xffffffffu, P, 31), sa);
                __syncwarp();                                  // partner species written above are visible to later runs
            }
        }
        a0 = (a0 | 31) + 1;                                    // the cell continues in the next 32-particle chunk?
        if (a0 >= cs1) break;
        R = __ldg(A.rec2 + (a0 >> 5));
    }
    if (cur_a >= 0 && sa != sa0 && lane == 0) A.sp[cur_a] = (int8_t)sa;
    __syncwarp();
}

// The same with the species of the unit's particles staged in shared memory: a dense unit is m_a x m_b / 2
// SEQUENTIAL pairs for one warp, and against global memory every run of 32 of them pays an L2 round trip (config 2
// in its stirred state: one same-cell unit of a few hundred microbes held up the whole phase for 0.45 ms).  `buf`
// (cap bytes) is the warp's descriptor buffer, free once stage B is done.  Falls back to the global-memory version
// when the two cells do not fit.
__device__ void resolve_unit_warp_staged(const ResolveArgs &A, int cell, int other, int8_t *buf, int cap)
{
    const int lane = threadIdx.x & 31;
    const int cs0 = __ldg(A.cell_start + cell), cs1 = __ldg(A.cell_start + cell + 1);
    const bool same = cell == other;
    const int oBeg = same ? cs0 : __ldg(A.cell_start + other), oEnd = same ? cs1 : __ldg(A.cell_start + other + 1);
    const int m_a = cs1 - cs0, m_b = same ? 0 : oEnd - oBeg;
    if (m_a + m_b > cap) { resolve_unit_warp(A, cell, cs0, cs1, oBeg); return; }
    int8_t *s_a = buf, *s_b = same ? buf : buf + m_a;
    __syncwarp();
    for (int i = lane; i < m_a; i += 32) s_a[i] = A.sp[cs0 + i];
    for (int i = lane; i < m_b; i += 32) s_b[i] = A.sp[oBeg + i];
    __syncwarp();
    int cur_a = -1, sa = 0, sa0 = 0;
    int a0 = cs0;
    uint2 R = __ldg(A.rec + cell);
    while (true) {
        const uint32_t *ent = A.hits + R.x;
        uint32_t en_next = lane < (int)R.y ? __ldg(ent + lane) : 0u;
        for (unsigned int k0 = 0; k0 < R.y; k0 += 32) {
            const unsigned int k = k0 + lane;
            const bool act = k < R.y;
            const uint32_t en = en_next;
            en_next = (k + 32 < R.y) ? __ldg(ent + k + 32) : 0u;      // the next chunk is on its way while this one is resolved
            const int ar = (int)((en >> 24) & 31u);
            unsigned int todo = __ballot_sync(0xffffffffu, act);
            while (todo) {
                const int lead = __ffs(todo) - 1;
                const int ar0 = __shfl_sync(0xffffffffu, ar, lead);
                const unsigned int m = __ballot_sync(0xffffffffu, act && ar == ar0) & todo;   // entries are sorted by anchor
                todo &= ~m;
                const int a = a0 - cs0 + ar0;                  // index into s_a
                if (a != cur_a) {
                    if (cur_a >= 0 && sa != sa0 && lane == 0) s_a[cur_a] = (int8_t)sa;
                    __syncwarp();
                    cur_a = a;
                    sa = sa0 = ((volatile int8_t *)s_a)[a];
                }
                if (!is_rps(sa)) continue;                     // winner = None for every pair of this anchor
                const bool mine = (m >> lane) & 1u;
                uint32_t M = MAP_ID, dec = 0;
                int b = 0, sb = 0;
                if (mine) {
                    b = (int)(en & B_REL_MASK); dec = en >> 29;   // index into s_b
                    sb = ((volatile int8_t *)s_b)[b];
                    if (is_rps(sb)) M = map_of_partner(sb, dec);
                }
                uint32_t P = M;                                // inclusive scan of maps in lane (= id_b) order
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, P, dd);
                    if (lane >= dd) P = map_compose(P, t);
                }
                uint32_t E = __shfl_up_sync(0xffffffffu, P, 1);
                if (lane == 0) E = MAP_ID;
                if (mine && is_rps(sb)) {
                    const int s_before = map_apply(E, sa);
                    if (s_before != sb) s_b[b] = (int8_t)rps_apply(s_before, sb, dec);
                }
                sa = map_apply(__shfl_sync(0xffffffffu, P, 31), sa);
                __syncwarp();                                  // partner species written above are visible to later runs
            }
        }
        a0 = (a0 | 31) + 1;                                    // the cell continues in the next 32-particle chunk?
        if (a0 >= cs1) break;
        R = __ldg(A.rec2 + (a0 >> 5));
    }
    if (cur_a >= 0 && sa != sa0 && lane == 0) s_a[cur_a] = (int8_t)sa;
    __syncwarp();
    for (int i = lane; i < m_a; i += 32) A.sp[cs0 + i] = s_a[i];
    for (int i = lane; i < m_b; i += 32) A.sp[oBeg + i] = s_b[i];
    __syncwarp();
}

// One phase.  A warp takes 32 * upl (upl = 4 or 8) consecutive units of one cell row, lane l the units l, l + 32, ...
// Pair counts per unit are heavy-tailed (same-cell units: ~m^2/2 for m microbes in the cell) and most units
// are short, so "one lane walks one unit" leaves the warp waiting for its longest unit with a handful of lanes
// active.  Instead, in two warp-uniform stages:
//   A  every lane looks at its units (loads batched four units at a time, all coalesced) and pushes one
//      descriptor per non-empty (segment, direction) entry stream -- (first entry, count, first anchor,
//      partner cell start) -- onto the WARP's list in shared memory; dense clusters are resolved right away
//      by the whole warp;
//   B  the lanes pull units from that list (a shared-memory ticket) and walk their entry streams, one pair
//      per iteration: all lanes execute the same instruction stream, and a lane that finishes a short unit
//      takes the next one instead of idling.
constexpr int HEAD_CAP = 32 * MAX_UNITS_PER_LANE;  // at most one head descriptor per unit
constexpr int CONT_CAP = 64;                       // descriptors of continuation segments (more: warp path)
constexpr unsigned int NO_LINK = 0xffffu;            // descriptor .y = count (16 bits) | link << 16
constexpr int HEAVY_CAP = 16;                      // dense units queued per warp (more: resolved on the spot, against global memory;
                                                   // with the relative limit at most ~10 segments of a warp can exceed it)
template <int BATCH, int PB>
__global__ void __launch_bounds__(RES_THREADS) resolve_phase_kernel(ResolveArgs A)
{
    __shared__ uint4 s_desc_all[RES_THREADS / 32][HEAD_CAP + CONT_CAP];   // x first entry | y count + link << 16 | z first anchor | w partner cell start
    __shared__ unsigned int s_ctr_all[RES_THREADS / 32][4];               // heads, continuations, next ticket, dense units
    __shared__ int s_heavy_all[RES_THREADS / 32][2 * HEAVY_CAP];           // (cell, other) of the dense units of this warp
    if (*A.n_pairs > A.cap_words) return;              // the hand-off overflowed: reported by lm_sync_stats
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * RES_THREADS + threadIdx.x) >> 5;
    if (wid >= A.n_warps) return;                      // warp-uniform
    uint4 *s_desc = s_desc_all[threadIdx.x >> 5];
    unsigned int *s_ctr = s_ctr_all[threadIdx.x >> 5];
    int *s_heavy = s_heavy_all[threadIdx.x >> 5];
    const int row = (int)(wid / A.warps_per_row);
    const int u_base = (int)(wid - (long long)row * A.warps_per_row) * (32 * A.upl) + lane;
    const int ncx = A.ncx;
    const int cy = (A.mode == MODE_CROSS) ? 2 * row + A.parity : row;
    const int row_cell = cy * ncx;
    if (lane < 4) s_ctr[lane] = 0;
    __syncwarp();

    // ---- how loaded is this warp?  (first segments only: one coalesced 8-byte load per unit, the same lines stage A
    // reads again.)  A segment is "dense" -- handed to the whole warp -- not when it has many pairs: in a crowded
    // region every unit has many, the lanes are evenly loaded and walking the streams lane by lane is the efficient
    // way; but when it has many more than a lane's fair share of this warp's pairs.
    unsigned int limit;
    {
        unsigned int mine = 0;
        for (int j = 0; j < A.upl; ++j) {
            const int u = u_base + 32 * j;
            if (u < A.units_per_row) {
                const int c = (A.mode == MODE_EAST) ? row_cell + 2 * u + A.parity : row_cell + u;
                // three independent loads (the record of an empty cell is stale and is not counted): one round trip
                const unsigned int cnt = __ldg(A.rec + c).y;
                if (__ldg(A.cell_start + c + 1) > __ldg(A.cell_start + c)) mine += cnt;
            }
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, dd);
        limit = min(max(A.heavy_min, 3u * (mine / 32u)), 0xffffu);
    }

    // ---- stage A: BATCH units at a time, all loads independent and coalesced across the lanes
#pragma unroll 1
    for (int j0 = 0; j0 < A.upl; j0 += BATCH) {
        int cell[BATCH], other[BATCH];
        bool on[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int u = u_base + 32 * (j0 + q);
            on[q] = u < A.units_per_row;
            if (A.mode == MODE_SAME) { cell[q] = row_cell + u; other[q] = cell[q]; }
            else if (A.mode == MODE_EAST) { cell[q] = row_cell + 2 * u + A.parity; other[q] = cell[q] + 1; }
            else {
                const int ox = u + A.dir;
                on[q] = on[q] && ox >= 0 && ox < ncx;
                cell[q] = row_cell + u; other[q] = cell[q] + ncx + A.dir;
            }
        }
        int cs0[BATCH], cs1[BATCH], ob[BATCH];
        uint2 R[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            cs0[q] = on[q] ? __ldg(A.cell_start + cell[q]) : 0;
            cs1[q] = on[q] ? __ldg(A.cell_start + cell[q] + 1) : 0;
            ob[q] = on[q] ? __ldg(A.cell_start + other[q]) : 0;
            R[q] = on[q] ? __ldg(A.rec + cell[q]) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            bool heavy = false;
            if (on[q] && cs1[q] > cs0[q]) {
                uint2 r = R[q];
                int a0 = cs0[q];
                int head_slot = -1, prev_slot = -1;
                while (true) {
                    if (r.y > limit) { heavy = true; break; }
                    if (r.y) {
                        int slot;
                        if (head_slot < 0) slot = head_slot = (int)atomicAdd(&s_ctr[0], 1u);          // < HEAD_CAP by construction
                        else {
                            const unsigned int cs = atomicAdd(&s_ctr[1], 1u);
                            if (cs >= (unsigned int)CONT_CAP) { heavy = true; break; }
                            slot = HEAD_CAP + (int)cs;
                            s_desc[prev_slot].y = (s_desc[prev_slot].y & 0xffffu) | ((unsigned int)slot << 16);
                        }
                        s_desc[slot] = make_uint4(r.x, r.y | (NO_LINK << 16), (unsigned int)a0, (unsigned int)ob[q]);
                        prev_slot = slot;
                    }
                    a0 = (a0 | 31) + 1;                        // the cell goes on in the next 32-particle chunk?
                    if (a0 >= cs1[q]) break;
                    r = __ldg(A.rec2 + (a0 >> 5));
                }
                if (heavy && head_slot >= 0) s_desc[head_slot].y = NO_LINK << 16;      // void what was pushed: the whole unit goes to the warp
            }
            // dense units wait until stage B is done: then the descriptor buffer is free to stage their species in
            unsigned int hm = __ballot_sync(0xffffffffu, heavy);
            while (hm) {
                const int src = __ffs(hm) - 1;
                hm &= hm - 1;
                const int hc = __shfl_sync(0xffffffffu, cell[q], src), ho = __shfl_sync(0xffffffffu, other[q], src);
                const unsigned int slot = s_ctr[3];            // warp-uniform
                __syncwarp();                                  // every lane has read the counter before lane 0 advances it
                if (slot < (unsigned int)HEAVY_CAP) {
                    if (lane == 0) { s_heavy[2 * slot] = hc; s_heavy[2 * slot + 1] = ho; s_ctr[3] = slot + 1; }
                    __syncwarp();
                } else {
                    resolve_unit_warp(A, hc, __shfl_sync(0xffffffffu, cs0[q], src), __shfl_sync(0xffffffffu, cs1[q], src),
                                      __shfl_sync(0xffffffffu, ob[q], src));
                }
            }
        }
    }
    __syncwarp();
    const unsigned int n_heads = s_ctr[0];

    // ---- stage B: a lane holds the unit it is walking and the ticket of its next one (whose first sector is
    // already on its way).  PB == 1: one pair per iteration, the two species loads of a pair issued together.
    // PB > 1 (LM_OPT_RESOLVE_BATCH): up to PB consecutive pairs of the unit per iteration -- their entries, then all
    // their species, are loaded before the first of them is resolved, so a unit of m pairs costs ~2 m / PB dependent
    // memory round trips instead of 2 m (a phase lasts as long as its longest lane-walked unit).  The sequential
    // semantics are kept by forwarding inside the batch: a species loaded early is replaced when an earlier pair of
    // the batch rewrote that microbe (same partner b again, or -- same-cell units -- the partner becoming the anchor).
    // Earlier batches' stores precede these loads in program order, and no pair writes anything but its own b
    // (the anchor's species lives in a register and is written back when the anchor changes; a retired anchor is
    // never a partner or an anchor again: entries are sorted by anchor, partners of a same-cell unit lie above it).
    int cur_a = -1, sa = 0, sa0 = 0, oBeg = 0;
    unsigned int k = 0, k_end = 0, a0 = 0, link = NO_LINK;
    unsigned int nxt = atomicAdd(&s_ctr[2], 1u);
    bool active = nxt < n_heads;
    while (__any_sync(0xffffffffu, active)) {
        if (active && k == k_end) {
            unsigned int take = link;
            if (take == NO_LINK) {
                take = nxt;
                if (take < n_heads) {
                    nxt = atomicAdd(&s_ctr[2], 1u);
                    if (nxt < n_heads) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.hits + s_desc[nxt].x));
                } else active = false;
            }
            if (active) {
                const uint4 D = s_desc[take];
                k = D.x; k_end = D.x + (D.y & 0xffffu); link = D.y >> 16; a0 = D.z; oBeg = (int)D.w;
            }
        }
        if (PB == 1) {
            if (active && k < k_end) {
                const uint32_t en = __ldg(A.hits + k);
                const int a = (int)a0 + (int)((en >> 24) & 31u), b = oBeg + (int)(en & B_REL_MASK);
                const bool new_a = a != cur_a;
                if (new_a && cur_a >= 0 && sa != sa0) A.sp[cur_a] = (int8_t)sa;     // never aliases the two loads below
                int sa_l = 0;
                if (new_a) sa_l = A.sp[a];
                const int sb = A.sp[b];
                if (new_a) { cur_a = a; sa = sa0 = sa_l; }
                if (sa != sb && is_rps(sa) && is_rps(sb)) {
                    sa = rps_apply(sa, sb, en >> 29);
                    A.sp[b] = (int8_t)sa;
                }
                ++k;
            }
        } else if (active && k < k_end) {
            const int nb = min((int)(k_end - k), PB);
            uint32_t en[PB];
            int av[PB], bv[PB], sbv[PB], sav[PB];
#pragma unroll
            for (int j = 0; j < PB; ++j) en[j] = j < nb ? __ldg(A.hits + k + j) : 0u;
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                av[j] = (int)a0 + (int)((en[j] >> 24) & 31u);
                bv[j] = oBeg + (int)(en[j] & B_REL_MASK);
            }
#pragma unroll
            for (int j = 0; j < PB; ++j) {                       // every load of the batch before its first store
                sbv[j] = j < nb ? (int)A.sp[bv[j]] : 0;
                sav[j] = (j < nb && av[j] != (j == 0 ? cur_a : av[j - 1])) ? (int)A.sp[av[j]] : 0;
            }
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                if (j < nb) {
                    if (av[j] != cur_a) {
                        if (cur_a >= 0 && sa != sa0) A.sp[cur_a] = (int8_t)sa;
                        cur_a = av[j]; sa = sa0 = sav[j];
                    }
                    int sb = sbv[j];
                    if (sa != sb && is_rps(sa) && is_rps(sb)) {
                        sa = rps_apply(sa, sb, en[j] >> 29);
                        sb = sa;
                        A.sp[bv[j]] = (int8_t)sa;
#pragma unroll
                        for (int i = j + 1; i < PB; ++i) {       // later pairs of the batch that loaded this microbe too early
                            if (bv[i] == bv[j]) sbv[i] = sb;
                            if (av[i] == bv[j]) sav[i] = sb;
                        }
                    }
                }
            }
            k += (unsigned int)nb;
        }
    }
    if (cur_a >= 0 && sa != sa0) A.sp[cur_a] = (int8_t)sa;

    // ---- dense units, one after the other, with their species staged in the (now free) descriptor buffer
    __syncwarp();
    const unsigned int n_heavy = s_ctr[3];
    for (unsigned int q = 0; q < n_heavy; ++q)
        resolve_unit_warp_staged(A, s_heavy[2 * q], s_heavy[2 * q + 1], reinterpret_cast<int8_t *>(s_desc),
                                 (int)((HEAD_CAP + CONT_CAP) * sizeof(uint4)));
}

// ---------------------------------------------------------------------------------------------------
template <int NORM>
static void launch_find_norm(const FindArgs &F, bool rps, bool emit, int grid, cudaStream_t s)
{
    if (rps) {
        if (emit) find_pairs_kernel<true, true, NORM><<<grid, FIND_THREADS, 0, s>>>(F);
        else find_pairs_kernel<true, false, NORM><<<grid, FIND_THREADS, 0, s>>>(F);
    } else {
        if (emit) find_pairs_kernel<false, true, NORM><<<grid, FIND_THREADS, 0, s>>>(F);
        else find_pairs_kernel<false, false, NORM><<<grid, FIND_THREADS, 0, s>>>(F);
    }
}

cudaError_t launch_find(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int n, double r,
                        const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s)
{
    h->rps_cap = rps ? h->max_pairs : -1;
    if (h->interact_mode == 2) {                           // hybrid: the heavy units' kernel needs these (launch_resolve_phases)
        h->ia_lon = lon; h->ia_lat = lat; h->ia_id = id; h->ia_n = n; h->ia_r = r;
        h->ia_have_rps = rps != nullptr;
        if (rps) h->ia_rps = *rps;
        h->ia_pairs = pairs_out; h->ia_cap = cap;
        cudaError_t e0 = cudaMemsetAsync(h->heavy_cnt, 0, 36 * sizeof(unsigned int), s);
        if (e0 != cudaSuccess) return e0;
    }
    if (h->interact_mode == 1) {
        // fused tile kernel: the tile phases and the vertical-boundary phases now, the horizontal-boundary phases
        // (the only ones that cross a strip boundary) with launch_resolve_phases(.., 6, 8)
        h->rps_cap = -1;                                   // no hand-off to overflow
        h->ia_lon = lon; h->ia_lat = lat; h->ia_id = id; h->ia_n = n; h->ia_r = r;
        h->ia_have_rps = rps != nullptr;
        if (rps) h->ia_rps = *rps;
        h->ia_pairs = pairs_out; h->ia_cap = cap;
        return launch_interact(h, lon, lat, id, h->sp[h->cur], n, r, rps, pairs_out, cap, 0, 11, s);
    }
    if (n <= 0) return cudaSuccess;
    FindArgs F;
    F.lon = lon; F.lat = lat; F.id = id; F.cell_start = h->cell_start;
    F.g = h->grid; F.n = n;
    F.force_two_pass = h->find_path == 1 ? 1 : 0;
    F.row0 = h->strip.row0; F.rows_owned = h->strip.rows_owned; F.rows_local = h->strip.rows_local;
    F.r2 = h->norm == LM_NORM_2 ? r * r : r;           // SciPy: tub = r*r for p=2, pow(r, 1) for p=1, r for p=inf
    F.r2_lo = (float)(F.r2 * (1.0 - 4e-6));
    F.r2_hi = (float)(F.r2 * (1.0 + 4e-6));
    F.pair_key = 0;
    F.thr[0] = F.thr[1] = F.thr[2] = 0;
    if (rps) {
        F.pair_key = rps->pair_key;
        const double p[3] = {rps->pRS, rps->pPR, rps->pSP};
        for (int k = 0; k < 3; ++k) {
            // u = m * 2^-53 with integer m < 2^53:  u < p  <=>  m < ceil(p * 2^53)   (the scaling is exact)
            if (!(p[k] > 0.0)) F.thr[k] = 0;
            else if (p[k] >= 1.0) F.thr[k] = 1ull << 53;
            else F.thr[k] = (unsigned long long)ceil(p[k] * 9007199254740992.0);
        }
    }
    F.hits = h->hits; F.rec = h->rec; F.rec2 = h->rec2; F.rec_stride = h->max_cells; F.rec2_stride = h->max_particles / 32 + 2;
    F.pairs = pairs_out;
    F.cap_words = rps ? (unsigned long long)h->max_pairs : 0ull;
    F.cap_pairs = (pairs_out && cap > 0) ? (unsigned long long)cap : 0ull;
    F.ctr = h->ctr;
    F.hybrid = h->interact_mode == 2 ? 1 : 0;
    F.heavy_list = h->heavy_list; F.heavy_cnt = h->heavy_cnt; F.heavy_cap = (unsigned int)h->heavy_cap;
    // measured on B200 (profiles/r2g_heavy_sweep.jsonl): 1,024 -- up to 32 x 32 microbes -- is the best compromise: BASELINE config 3
    // stirred (12 microbes per cell) 10.9 -> 7.4 ms per step against the round-1 pipeline alone; 256 floods the queue with
    // medium units that leave half a warp idle (9.9 ms), 16,384 and up leave the long lane walks in (8.1 - 10.5 ms)
    F.heavy_min = h->heavy_min > 0 ? (unsigned long long)h->heavy_min : 1024ull;
    const bool emit = F.cap_pairs > 0;
    const int grid = (n + FIND_THREADS - 1) / FIND_THREADS;
    ++h->launches;
    if (h->norm == LM_NORM_1) launch_find_norm<LM_NORM_1>(F, rps != nullptr, emit, grid, s);
    else if (h->norm == LM_NORM_INF) launch_find_norm<LM_NORM_INF>(F, rps != nullptr, emit, grid, s);
    else launch_find_norm<LM_NORM_2>(F, rps != nullptr, emit, grid, s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the resolver's overflow guard reads a snapshot: the counters themselves are reset by the next step while the
    // phases of this one may still be running on the side stream
    return cudaMemcpyAsync(h->n_pairs_snap, &h->ctr->n_pairs, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s);
}

// ---------------------------------------------------------------------------------------------------
// Tiled resolver (LM_OPT_RESOLVE_MODE = 1; experimental, default off).  ONE launch for a whole range of phases.
//
// The nine phase launches above are latency-bound: every pair costs a dependent chain entry -> species (L2 round
// trips, 1 useful byte per 32-byte sector), and every phase waits for its slowest warp.  But the canonical order only
// couples NEIGHBOURING cells, phase by phase: what an earlier phase did wrong at a cell can reach at most one cell
// further per coupling phase.  East units couple columns (2k, 2k+1) in phase 1 and (2k+1, 2k+2) in phase 2; the cross
// phases couple rows (2k, 2k+1) in 3-5 and (2k+1, 2k+2) in 6-8, and columns by one in 3, 5, 6, 8.  So a CTA that loads a
// tile of cells plus a halo of TILE_HX = 6 columns and TILE_HY = 2 rows on every side (tile origin on even rows and
// columns), copies the species of those cells from a SNAPSHOT taken before the first phase into shared memory, and runs
// every unit that lies completely inside the loaded region, phase after phase with __syncthreads() in between, ends up
// with exactly the sequential result on the INTERIOR of its tile -- the halo is computed redundantly (and possibly
// wrongly at its rim), and only the interior is written back.  (Halo widths: tests/test_resolver_tiling_cpu.py runs
// this scheme in NumPy against the global sequential order.)  Cost: the entries of the halo are read by more than one
// CTA (loaded / interior = 1.5x at 64 x 16), which is bandwidth the path has to spare; gain: species accesses are
// shared-memory accesses, there is one launch instead of nine, and a tile's slow unit delays only its own CTA.
// A tile whose loaded region holds more microbes than fit the CTA's shared memory works on a private slice of a global
// scratch array instead (same code through a generic pointer); a cell is loaded by at most four tiles, which bounds it.
constexpr int TILE_HX = 6, TILE_HY = 2;                    // halo; the tile itself (TILE_X x TILE_Y cells) is a template parameter:
                                                           // 64 x 16 by default, LM_OPT_RESOLVE_TILE_SHAPE picks another for A/B runs
constexpr int TILE_THREADS = 256;
constexpr int TILE_HEAVY_Q = 128;                          // dense units queued per phase for whole warps
constexpr int TILE_MEGA_Q = 16;                            // ... and for the whole CTA (knots of hundreds of microbes per cell)
constexpr unsigned int TILE_MEGA_MIN = 8192;               // ~128 microbes in one cell: 64 partners per anchor, where 256-wide scans win

struct TileArgs {
    const int8_t *__restrict__ sp_in;    // snapshot of the species before the first phase of this launch
    int8_t *sp_out;                      // the live species (only interiors are written)
    int8_t *scratch;                     // tiles that do not fit in shared memory
    unsigned long long *scratch_used;
    const int32_t *__restrict__ cell_start;
    const uint32_t *__restrict__ hits;
    const uint2 *__restrict__ rec;       // [5][rec_stride]
    const uint2 *__restrict__ rec2;      // [5][rec2_stride]
    long long rec_stride, rec2_stride;
    const unsigned long long *n_pairs;
    unsigned long long cap_words;
    int ncx, rows_owned, rows_local;
    int tiles_x;
    int first, last;                     // phases
    int smem_cap;                        // bytes of species a CTA can hold in shared memory
    unsigned int heavy_min;
    unsigned int mega_min;               // a unit with more pairs than this is resolved by the whole CTA
};

// Whole-warp resolution of one unit against the tile's private species (resolve_unit_warp above, with the two cells'
// species reached through spA / spB: index = particle index in storage order).
__device__ void resolve_unit_warp_tile(const uint32_t *__restrict__ hits, const uint2 *__restrict__ rec_d,
                                       const uint2 *__restrict__ rec2_d, int cell, int cs0, int cs1, int oBeg,
                                       int8_t *spA, int8_t *spB)
{
    const int lane = threadIdx.x & 31;
    int cur_a = -1, sa = 0, sa0 = 0;
    __syncwarp();
    int a0 = cs0;
    uint2 R = __ldg(rec_d + cell);
    while (true) {
        const uint32_t *ent = hits + R.x;
        for (unsigned int k0 = 0; k0 < R.y; k0 += 32) {
            const unsigned int k = k0 + lane;
            const bool act = k < R.y;
            const uint32_t en = act ? __ldg(ent + k) : 0u;
            const int ar = (int)((en >> 24) & 31u);
            unsigned int todo = __ballot_sync(0xffffffffu, act);
            while (todo) {
                const int lead = __ffs(todo) - 1;
                const int ar0 = __shfl_sync(0xffffffffu, ar, lead);
                const unsigned int m = __ballot_sync(0xffffffffu, act && ar == ar0) & todo;   // entries are sorted by anchor
                todo &= ~m;
                const int a = a0 + ar0;
                if (a != cur_a) {
                    if (cur_a >= 0 && sa != sa0 && lane == 0) spA[cur_a] = (int8_t)sa;
                    __syncwarp();
                    cur_a = a;
                    sa = sa0 = ((volatile int8_t *)spA)[a];
                }
                if (!is_rps(sa)) continue;                     // winner = None for every pair of this anchor
                const bool mine = (m >> lane) & 1u;
                uint32_t M = MAP_ID, dec = 0;
                int b = 0, sb = 0;
                if (mine) {
                    b = oBeg + (int)(en & B_REL_MASK); dec = en >> 29;
                    sb = ((volatile int8_t *)spB)[b];
                    if (is_rps(sb)) M = map_of_partner(sb, dec);
                }
                uint32_t P = M;                                // inclusive scan of maps in lane (= id_b) order
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, P, dd);
                    if (lane >= dd) P = map_compose(P, t);
                }
                uint32_t E = __shfl_up_sync(0xffffffffu, P, 1);
                if (lane == 0) E = MAP_ID;
                if (mine && is_rps(sb)) {
                    const int s_before = map_apply(E, sa);
                    if (s_before != sb) spB[b] = (int8_t)rps_apply(s_before, sb, dec);
                }
                sa = map_apply(__shfl_sync(0xffffffffu, P, 31), sa);
                __syncwarp();                                  // partner species written above are visible to later runs
            }
        }
        a0 = (a0 | 31) + 1;                                    // the cell continues in the next 32-particle chunk?
        if (a0 >= cs1) break;
        R = __ldg(rec2_d + (a0 >> 5));
    }
    if (cur_a >= 0 && sa != sa0 && lane == 0) spA[cur_a] = (int8_t)sa;
    __syncwarp();
}

// Whole-CTA resolution of one unit (every thread of the CTA calls this with the same arguments).  For the knots a
// long run produces -- several hundred microbes in one cell, 10^4..10^5 pairs in ONE unit -- the anchors are inherently
// sequential, but each anchor's run of partners is a prefix scan over 3->3 maps: a warp takes it 32 partners at a
// time (resolve_unit_warp_tile), the CTA takes 256 at a time, warp scans joined through shared memory.  Entries are
// read 256 per chunk; the distinct anchors of a chunk (at most 32: a_rel is a 5-bit index) are visited in ascending
// order through a presence mask, the anchor's species is carried in registers by every thread.
__device__ void resolve_unit_cta_tile(const uint32_t *__restrict__ hits, const uint2 *__restrict__ rec_d,
                                      const uint2 *__restrict__ rec2_d, int cell, int cs0, int cs1, int oBeg,
                                      int8_t *spA, int8_t *spB)
{
    __shared__ uint32_t s_wmap[TILE_THREADS / 32];
    __shared__ unsigned int s_present;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cur_a = -1, sa = 0, sa0 = 0;
    int a0 = cs0;
    uint2 R = __ldg(rec_d + cell);
    __syncthreads();
    while (true) {
        const uint32_t *ent = hits + R.x;
        for (unsigned int k0 = 0; k0 < R.y; k0 += TILE_THREADS) {
            const unsigned int k = k0 + tid;
            const bool act = k < R.y;
            const uint32_t en = act ? __ldg(ent + k) : 0u;
            const int ar = (int)((en >> 24) & 31u);
            if (tid == 0) s_present = 0u;
            __syncthreads();
            if (act) atomicOr(&s_present, 1u << ar);
            __syncthreads();
            unsigned int todo = s_present;
            __syncthreads();                                   // everyone has the mask before it is reset for the next chunk
            while (todo) {                                     // CTA-uniform
                const int ar0 = __ffs(todo) - 1;
                todo &= todo - 1;
                const int a = a0 + ar0;
                if (a != cur_a) {
                    if (cur_a >= 0 && sa != sa0 && tid == 0) spA[cur_a] = (int8_t)sa;
                    __syncthreads();
                    cur_a = a;
                    sa = sa0 = ((volatile int8_t *)spA)[a];
                }
                if (!is_rps(sa)) continue;                     // winner = None for every pair of this anchor (uniform)
                const bool mine = act && ar == ar0;
                uint32_t M = MAP_ID, dec = 0;
                int b = 0, sb = 0;
                if (mine) {
                    b = oBeg + (int)(en & B_REL_MASK); dec = en >> 29;
                    sb = ((volatile int8_t *)spB)[b];
                    if (is_rps(sb)) M = map_of_partner(sb, dec);
                }
                uint32_t P = M;                                // inclusive scan of maps inside the warp, in partner order
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, P, dd);
                    if (lane >= dd) P = map_compose(P, t);
                }
                uint32_t E = __shfl_up_sync(0xffffffffu, P, 1);
                if (lane == 0) E = MAP_ID;
                if (lane == 31) s_wmap[warp] = P;
                __syncthreads();
                uint32_t before = MAP_ID, total = MAP_ID;      // maps of the warps before mine | of all warps
#pragma unroll
                for (int w = 0; w < TILE_THREADS / 32; ++w) {
                    const uint32_t mw = s_wmap[w];
                    if (w < warp) before = map_compose(mw, before);
                    total = map_compose(mw, total);
                }
                if (mine && is_rps(sb)) {
                    const int s_before = map_apply(map_compose(E, before), sa);
                    if (s_before != sb) spB[b] = (int8_t)rps_apply(s_before, sb, dec);
                }
                sa = map_apply(total, sa);
                __syncthreads();                               // partner species written above are visible to later anchors
            }
        }
        a0 = (a0 | 31) + 1;                                    // the cell continues in the next 32-particle chunk?
        if (a0 >= cs1) break;
        R = __ldg(rec2_d + (a0 >> 5));
    }
    if (cur_a >= 0 && sa != sa0 && tid == 0) spA[cur_a] = (int8_t)sa;
    __syncthreads();
}

// geometry of phase ph: which direction table it reads and which cells anchor a unit
struct PhaseGeom { int mode, parity, dir, d_idx; };
__device__ __forceinline__ PhaseGeom phase_geom(int ph)
{
    PhaseGeom g;
    if (ph == 0) { g.mode = MODE_SAME; g.parity = 0; g.dir = 0; g.d_idx = 0; }
    else if (ph <= 2) { g.mode = MODE_EAST; g.parity = ph - 1; g.dir = 0; g.d_idx = 1; }
    else { g.mode = MODE_CROSS; g.parity = (ph - 3) / 3; g.dir = (ph - 3) % 3 - 1; g.d_idx = 3 + g.dir; }
    return g;
}

template <int TILE_X, int TILE_Y>
__global__ void __launch_bounds__(TILE_THREADS, 4) resolve_tiled_kernel(TileArgs A)
{
    constexpr int TILE_LX = TILE_X + 2 * TILE_HX, TILE_LY = TILE_Y + 2 * TILE_HY;         // loaded columns, rows
    constexpr int TILE_UPT = (TILE_LX * TILE_LY + TILE_THREADS - 1) / TILE_THREADS;       // units per thread and phase (64 x 16: 6)
    extern __shared__ __align__(16) int8_t s_species[];    // [smem_cap]
    __shared__ int s_p0[TILE_LY], s_cnt[TILE_LY], s_delta[TILE_LY];   // loaded row: first particle, particles, index delta
    __shared__ int s_cs[TILE_LY][TILE_LX + 1];             // cell_start of the loaded cells (+ one past the end of each row)
    __shared__ long long s_off;                            // >= 0: this tile works in the global scratch
    __shared__ int s_heavy[TILE_HEAVY_Q];
    __shared__ int s_mega[TILE_MEGA_Q];
    __shared__ unsigned int s_nheavy, s_phase_pairs, s_nmega;
    if (*A.n_pairs > A.cap_words) return;                  // the hand-off overflowed: reported by lm_sync_stats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncx = A.ncx;
    const int tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x;
    const int ix0 = tx * TILE_X, ix1 = min(ncx, ix0 + TILE_X);
    const int iy0 = ty * TILE_Y, iy1 = min(A.rows_local, iy0 + TILE_Y);
    const int lx0 = max(0, ix0 - TILE_HX), lx1 = min(ncx, ix1 + TILE_HX);
    const int ly0 = max(0, iy0 - TILE_HY), ly1 = min(A.rows_local, iy1 + TILE_HY);
    const int lw = lx1 - lx0, lh = ly1 - ly0;
    const unsigned int inv_lw = ((1u << 20) + (unsigned int)lw - 1u) / (unsigned int)lw;   // u / lw == (u * inv_lw) >> 20 while u * lw < 2^20
    static_assert((long long)TILE_LX * TILE_LY * TILE_LX < (1ll << 20) && TILE_LX * TILE_LY < 4096, "reciprocal division range");
    static_assert(TILE_X % 2 == 0 && TILE_Y % 2 == 0, "tile origins on even rows and columns");
    for (int k = tid; k < lh * (lw + 1); k += TILE_THREADS) {
        const int t = k / (lw + 1), x = k - t * (lw + 1);
        s_cs[t][x] = __ldg(A.cell_start + (long long)(ly0 + t) * ncx + lx0 + x);
    }
    __syncthreads();
    if (tid < lh) { s_p0[tid] = s_cs[tid][0]; s_cnt[tid] = s_cs[tid][lw] - s_cs[tid][0]; }
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int t = 0; t < lh; ++t) { s_delta[t] = off - s_p0[t]; off += s_cnt[t]; }
        s_off = off > A.smem_cap ? (long long)atomicAdd(A.scratch_used, (unsigned long long)((off + 15) & ~15)) : -1ll;
    }
    __syncthreads();
    int8_t *tsp = s_off >= 0 ? A.scratch + s_off : s_species;
    for (int t = 0; t < lh; ++t) {
        const int p0 = s_p0[t], cnt = s_cnt[t];
        int8_t *dst = tsp + s_delta[t] + p0;
        for (int i = tid; i < cnt; i += TILE_THREADS) dst[i] = __ldg(A.sp_in + p0 + i);
    }
    __syncthreads();

    for (int ph = A.first; ph <= A.last; ++ph) {
        const PhaseGeom G = phase_geom(ph);
        const uint2 *rec_d = A.rec + (size_t)G.d_idx * A.rec_stride;
        const uint2 *rec2_d = A.rec2 + (size_t)G.d_idx * A.rec2_stride;
        if (tid == 0) { s_nheavy = 0u; s_phase_pairs = 0u; s_nmega = 0u; }
        __syncthreads();
        // ---- pass 1: the units of this thread (u = tid, tid + 256, ...: at most TILE_UPT) and their pair counts.
        // "Dense" is relative, as in resolve_phase_kernel: in a crowded tile every unit is long, the lanes are evenly
        // loaded and the lane walk is the efficient way; only a unit far above a lane's fair share goes to a warp.
        uint2 rr[TILE_UPT];
        unsigned int tot[TILE_UPT];
        unsigned int mine = 0;
#pragma unroll
        for (int k = 0; k < TILE_UPT; ++k) {
            rr[k] = make_uint2(0u, 0u);
            tot[k] = 0u;
            const int u = tid + k * TILE_THREADS;
            if (u < lw * lh) {
                const int ry = (int)(((unsigned int)u * inv_lw) >> 20), cy = ly0 + ry, cx = lx0 + (u - ry * lw);   // ry = u / lw
                bool on;
                if (G.mode == MODE_SAME) on = cy < A.rows_owned;
                else if (G.mode == MODE_EAST) on = cy < A.rows_owned && (cx & 1) == G.parity && cx + 1 < lx1;
                else on = (cy & 1) == G.parity && cy + 1 < ly1 && cx + G.dir >= lx0 && cx + G.dir < lx1;
                if (on) {
                    const int cell = cy * ncx + cx;
                    const int cs0 = s_cs[ry][cx - lx0], cs1 = s_cs[ry][cx - lx0 + 1];
                    if (cs1 > cs0) {                           // an empty cell's record is stale
                        rr[k] = __ldg(rec_d + cell);
                        unsigned int total = rr[k].y;
                        for (int a0 = (cs0 | 31) + 1; a0 < cs1; a0 = (a0 | 31) + 1) total += __ldg(rec2_d + (a0 >> 5)).y;
                        tot[k] = total;
                        mine += total;
                    }
                }
            }
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, dd);
        if (lane == 0 && mine) atomicAdd(&s_phase_pairs, mine);
        __syncthreads();
        const unsigned int limit = max(A.heavy_min, 4u * (s_phase_pairs / (unsigned int)TILE_THREADS));
        // ---- pass 2: walk them (or queue the dense ones for whole warps)
#pragma unroll
        for (int k = 0; k < TILE_UPT; ++k) {
            if (tot[k] == 0u) continue;
            const int u = tid + k * TILE_THREADS;
            if (tot[k] > A.mega_min) {
                const unsigned int q = atomicAdd(&s_nmega, 1u);
                if (q < (unsigned int)TILE_MEGA_Q) { s_mega[q] = u; continue; }
            }
            if (tot[k] > limit) {
                const unsigned int q = atomicAdd(&s_nheavy, 1u);
                if (q < (unsigned int)TILE_HEAVY_Q) { s_heavy[q] = u; continue; }
            }
            const int ry = (int)(((unsigned int)u * inv_lw) >> 20), cy = ly0 + ry, cx = lx0 + (u - ry * lw);   // ry = u / lw
            int oy = cy, ox = cx;
            if (G.mode == MODE_EAST) ox = cx + 1;
            else if (G.mode == MODE_CROSS) { oy = cy + 1; ox = cx + G.dir; }
            const int cs0 = s_cs[ry][cx - lx0], cs1 = s_cs[ry][cx - lx0 + 1];
            const int ob = s_cs[oy - ly0][ox - lx0];
            uint2 r = rr[k];
            int8_t *spA = tsp + s_delta[ry], *spB = tsp + s_delta[oy - ly0];
            int cur_a = -1, sa = 0, sa0 = 0, a0 = cs0;
            while (true) {
                const uint32_t *ent = A.hits + r.x;
                for (unsigned int kk = 0; kk < r.y; kk += 4) {
                    const int nb = (int)min(4u, r.y - kk);
                    uint32_t en[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) en[j] = j < nb ? __ldg(ent + kk + j) : 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < nb) {
                            const int a = a0 + (int)((en[j] >> 24) & 31u), b = ob + (int)(en[j] & B_REL_MASK);
                            if (a != cur_a) {
                                if (cur_a >= 0 && sa != sa0) spA[cur_a] = (int8_t)sa;
                                cur_a = a; sa = sa0 = spA[a];
                            }
                            const int sb = spB[b];
                            if (sa != sb && is_rps(sa) && is_rps(sb)) {
                                sa = rps_apply(sa, sb, en[j] >> 29);
                                spB[b] = (int8_t)sa;
                            }
                        }
                    }
                }
                a0 = (a0 | 31) + 1;                            // the cell goes on in the next 32-particle chunk?
                if (a0 >= cs1) break;
                r = __ldg(rec2_d + (a0 >> 5));
            }
            if (cur_a >= 0 && sa != sa0) spA[cur_a] = (int8_t)sa;
        }
        __syncthreads();
        const unsigned int n_heavy = min(s_nheavy, (unsigned int)TILE_HEAVY_Q);
        for (unsigned int q = warp; q < n_heavy; q += TILE_THREADS / 32) {      // dense units: one warp each
            const int u = s_heavy[q];
            const int ry = (int)(((unsigned int)u * inv_lw) >> 20), cy = ly0 + ry, cx = lx0 + (u - ry * lw);   // ry = u / lw
            int oy = cy, ox = cx;
            if (G.mode == MODE_EAST) ox = cx + 1;
            else if (G.mode == MODE_CROSS) { oy = cy + 1; ox = cx + G.dir; }
            resolve_unit_warp_tile(A.hits, rec_d, rec2_d, cy * ncx + cx, s_cs[ry][cx - lx0], s_cs[ry][cx - lx0 + 1],
                                   s_cs[oy - ly0][ox - lx0], tsp + s_delta[ry], tsp + s_delta[oy - ly0]);
        }
        __syncthreads();
        const unsigned int n_mega = min(s_nmega, (unsigned int)TILE_MEGA_Q);
        for (unsigned int q = 0; q < n_mega; ++q) {                             // knots: the whole CTA, one after the other
            const int u = s_mega[q];
            const int ry = (int)(((unsigned int)u * inv_lw) >> 20), cy = ly0 + ry, cx = lx0 + (u - ry * lw);   // ry = u / lw
            int oy = cy, ox = cx;
            if (G.mode == MODE_EAST) ox = cx + 1;
            else if (G.mode == MODE_CROSS) { oy = cy + 1; ox = cx + G.dir; }
            resolve_unit_cta_tile(A.hits, rec_d, rec2_d, cy * ncx + cx, s_cs[ry][cx - lx0], s_cs[ry][cx - lx0 + 1],
                                  s_cs[oy - ly0][ox - lx0], tsp + s_delta[ry], tsp + s_delta[oy - ly0]);
        }
        __syncthreads();                                       // the phase is complete on the whole loaded region
    }

    // the interior of the tile is exact: write it back
    for (int y = iy0; y < iy1; ++y) {
        const int p0 = s_cs[y - ly0][ix0 - lx0], p1 = s_cs[y - ly0][ix1 - lx0];
        const int8_t *src = tsp + s_delta[y - ly0];
        for (int p = p0 + tid; p < p1; p += TILE_THREADS) A.sp_out[p] = src[p];
    }
}

template <int PB>
static void launch_resolve_pb(const ResolveArgs &R, int upl, unsigned int blocks, cudaStream_t s)
{
    if (upl >= 4) resolve_phase_kernel<4, PB><<<blocks, RES_THREADS, 0, s>>>(R);
    else if (upl == 2) resolve_phase_kernel<2, PB><<<blocks, RES_THREADS, 0, s>>>(R);
    else resolve_phase_kernel<1, PB><<<blocks, RES_THREADS, 0, s>>>(R);
}

// phases [first, last] of the canonical cell-phase order on the local rows.  Strip boundaries sit on even
// global rows, so local and global row parities agree; same-cell and east units cover the owned rows,
// cross-row units may reach into the ghost row (local row rows_owned), only in phases 6-8.
template <int TX, int TY>
static cudaError_t launch_tiled_shape(TileArgs &T, cudaStream_t s)
{
    T.tiles_x = (T.ncx + TX - 1) / TX;
    const long long tiles = (long long)T.tiles_x * ((T.rows_local + TY - 1) / TY);
    cudaError_t e = cudaFuncSetAttribute(resolve_tiled_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem_cap);
    if (e != cudaSuccess) return e;
    resolve_tiled_kernel<TX, TY><<<(unsigned)tiles, TILE_THREADS, (size_t)T.smem_cap, s>>>(T);
    return cudaGetLastError();
}

// LM_OPT_RESOLVE_MODE = 1: phases [first, last] as ONE launch of resolve_tiled_kernel on a snapshot of the species
static cudaError_t launch_resolve_tiled(lm_handle_s *h, int8_t *sp, int first, int last, cudaStream_t s)
{
    const int64_t n_all = std::min<int64_t>(h->max_particles, h->n + ((h->has_north || h->has_south) ? h->ghost_cap : 0));
    cudaError_t e = cudaMemcpyAsync(h->sp_snap, sp, (size_t)n_all, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(h->tile_scratch_used, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    TileArgs T;
    T.sp_in = h->sp_snap; T.sp_out = sp; T.scratch = h->tile_scratch; T.scratch_used = h->tile_scratch_used;
    T.cell_start = h->cell_start; T.hits = h->hits; T.rec = h->rec; T.rec2 = h->rec2;
    T.rec_stride = h->max_cells; T.rec2_stride = h->max_particles / 32 + 2;
    T.n_pairs = h->n_pairs_snap; T.cap_words = (unsigned long long)h->max_pairs;
    T.ncx = h->grid.ncx; T.rows_owned = h->strip.rows_owned; T.rows_local = h->strip.rows_local;
    T.first = first; T.last = last;
    T.smem_cap = h->resolve_tile_smem;
    T.heavy_min = h->resolve_heavy_min > 0 ? (unsigned int)h->resolve_heavy_min : 4u * HEAVY_MIN;
    T.mega_min = h->resolve_mega_min > 0 ? (unsigned int)h->resolve_mega_min : TILE_MEGA_MIN;
    if (T.ncx <= 0 || T.rows_local <= 0) return cudaSuccess;
    switch (h->resolve_tile_shape) {                       // LM_OPT_RESOLVE_TILE_SHAPE
        case 1: e = launch_tiled_shape<32, 16>(T, s); break;
        case 2: e = launch_tiled_shape<128, 16>(T, s); break;
        case 3: e = launch_tiled_shape<64, 32>(T, s); break;
        default: e = launch_tiled_shape<64, 16>(T, s); break;
    }
    ++h->launches;
    return e;
}

cudaError_t launch_resolve_phases(lm_handle_s *h, int8_t *sp, int first, int last, cudaStream_t s)
{
    if (h->interact_mode == 1) {
        if (last < 6) return cudaSuccess;
        return launch_interact(h, h->ia_lon, h->ia_lat, h->ia_id, sp, h->ia_n, h->ia_r, h->ia_have_rps ? &h->ia_rps : nullptr,
                               h->ia_pairs, h->ia_cap, 12, 14, s);
    }
    const bool hybrid = h->interact_mode == 2;             // the heavy units of every phase follow its light units
    const bool light = h->rps_cap >= 0;                    // false: pair search only (the heavy units' pairs still have to be found)
    if (!light && !hybrid) return cudaSuccess;
    if (h->resolve_mode == 1 && !hybrid) return launch_resolve_tiled(h, sp, first, last, s);
    ResolveArgs R;
    R.sp = sp; R.cell_start = h->cell_start; R.hits = h->hits;
    R.n_pairs = h->n_pairs_snap;
    R.cap_words = (unsigned long long)h->max_pairs;
    R.ncx = h->grid.ncx;
    R.heavy_min = h->resolve_heavy_min > 0 ? (unsigned int)h->resolve_heavy_min : HEAVY_MIN;
    const long long ncx = R.ncx, rows_owned = h->strip.rows_owned, rows_local = h->strip.rows_local;
    for (int ph = first; ph <= last; ++ph) {
        long long rows;
        int d_idx;
        if (ph == 0) { R.mode = MODE_SAME; R.parity = 0; R.dir = 0; d_idx = 0; R.units_per_row = (int)ncx; rows = rows_owned; }
        else if (ph <= 2) {
            R.mode = MODE_EAST; R.parity = ph - 1; R.dir = 0; d_idx = 1;
            R.units_per_row = (int)((ncx - R.parity) / 2); rows = rows_owned;
        } else {
            R.mode = MODE_CROSS; R.parity = (ph - 3) / 3; R.dir = (ph - 3) % 3 - 1; d_idx = 3 + R.dir;
            R.units_per_row = (int)ncx; rows = (rows_local - R.parity) / 2;       // anchor rows cy = 2 i + parity, cy + 1 < rows_local
        }
        R.rec = h->rec + (size_t)d_idx * h->max_cells;
        R.rec2 = h->rec2 + (size_t)d_idx * (h->max_particles / 32 + 2);
        if (rows <= 0 || R.units_per_row <= 0) continue;
        // the heavy units of the phase run beside its light units (disjoint microbes) on the handle's side stream: fork here,
        // join after the light launch -- the phase costs the longer of the two, and a lone knot does not hold up the SMs
        const bool forked = hybrid && light && h->side_stream && h->side_stream != s;
        if (hybrid) {
            cudaError_t eh = cudaSuccess;
            if (forked) {
                eh = cudaEventRecord(h->ev_fork, s);
                if (eh == cudaSuccess) eh = cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0);
                if (eh == cudaSuccess) eh = launch_interact_heavy(h, sp, ph, h->side_stream);
                if (eh == cudaSuccess) eh = cudaEventRecord(h->ev_join, h->side_stream);
            } else eh = launch_interact_heavy(h, sp, ph, s);
            if (eh != cudaSuccess) return eh;
        }
        if (!light) continue;
        // units per lane, measured on B200 (profiles/): 8 wins on large sparse grids (12.5 M microbes, 5.2 M
        // cells, 2 pairs per unit: 0.81 vs 0.90 ms), 4 on smaller / denser ones (10 M microbes, 1.2 M cells, 27
        // pairs per unit: 2.2 vs 3.6 ms) where eight long streams per lane leave too few warps in flight
        const long long units = rows * R.units_per_row;
        int upl = units >= 3000000 ? MAX_UNITS_PER_LANE : (units >= 600000 ? 4 : (units >= 300000 ? 2 : 1));   // small grids: keep >= ~5k warps
        if (h->resolve_upl == 1 || h->resolve_upl == 2 || h->resolve_upl == 4 || h->resolve_upl == 8) upl = h->resolve_upl;
        R.upl = upl;
        R.warps_per_row = (R.units_per_row + 32 * upl - 1) / (32 * upl);
        R.n_warps = rows * R.warps_per_row;
        const long long blocks = (R.n_warps * 32 + RES_THREADS - 1) / RES_THREADS;
        const int pb = h->resolve_batch;                       // LM_OPT_RESOLVE_BATCH: pairs per stage-B iteration
        if (pb == 8) launch_resolve_pb<8>(R, upl, (unsigned)blocks, s);
        else if (pb == 4) launch_resolve_pb<4>(R, upl, (unsigned)blocks, s);
        else launch_resolve_pb<1>(R, upl, (unsigned)blocks, s);
        ++h->launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (forked) {
            e = cudaStreamWaitEvent(s, h->ev_join, 0);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_pairs(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                         double r, const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s)
{
    if (h->interact_mode == 1) return launch_interact(h, lon, lat, id, sp, n, r, rps, pairs_out, cap, 0, 14, s);
    cudaError_t e = launch_find(h, lon, lat, id, n, r, rps, pairs_out, cap, s);
    if (e != cudaSuccess || n <= 0 || (!rps && h->interact_mode != 2)) return e;
    return launch_resolve_phases(h, sp, 0, 8, s);
}

}  // namespace lm
