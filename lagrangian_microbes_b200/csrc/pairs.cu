// Radius pair search on the binned particle arrays, optionally fused with rock-paper-scissors
// resolution in the canonical cell-phase order.
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
// Work decomposition.  A *unit* is a cell (pairs inside it) or two adjacent cells (pairs across
// them): same cell, E, NW, N, NE -- the half stencil.  Units of one *phase* touch disjoint
// particles:
//     phase 0        same cell
//     phase 1,2      E neighbour, anchor cx even / odd
//     phase 3,4,5    NW, N, NE neighbour, anchor cy even
//     phase 6,7,8    NW, N, NE neighbour, anchor cy odd
// so with RPS fused in, each phase is one conflict-free launch and a unit is resolved sequentially
// by one lane in (anchor id, other id) order -- the reference's sequential in-place semantics under
// the canonical total order (phase, unit, id_a, id_b); see DESIGN.md §4.3 and
// oracle/rps.py::cell_phase_order.  Without RPS all five directions run in a single launch.
//
// Lanes of a warp pull units from the warp's contiguous chunk as they go idle (unit sizes are
// Poisson-distributed; a static unit-per-thread mapping would idle ~3/4 of the lanes), test one
// candidate pair per iteration, and append hits to a per-warp shared-memory buffer that is flushed
// to the global pair list with one atomic per ~200 pairs.
#include "lm_internal.cuh"
#include "philox.cuh"

namespace lm {

constexpr int PAIR_WARPS = 8;            // warps per CTA
constexpr int PAIR_BUF = 256;            // pairs buffered per warp (2 KB)
constexpr int UNITS_PER_WARP = 256;      // contiguous units owned by one warp
constexpr int REFILL_IDLE = 12;          // refill when at least this many lanes are idle

enum UnitMode { MODE_ALL = 0, MODE_SAME = 1, MODE_EAST = 2, MODE_CROSS = 3 };

struct PairArgs {
    const float *__restrict__ lon;
    const float *__restrict__ lat;
    const int32_t *__restrict__ id;
    int8_t *sp;
    const int32_t *__restrict__ cell_start;
    int ncx, ncy;
    float r2_lo, r2_hi;
    double r2;
    RpsDev rps;
    int2 *pairs;
    unsigned long long cap;
    Counters *ctr;
    long long n_units;
    int mode, parity, dir;   // dir in {-1,0,+1} for MODE_CROSS
};

// Decode unit u of the launch into anchor / other particle ranges.  Returns false for units that
// cannot contain a pair.
__device__ __forceinline__ bool decode_unit(const PairArgs &A, long long u, int &aBeg, int &aEnd, int &bBeg, int &bEnd,
                                            bool &same)
{
    int anchor, other;
    const int ncx = A.ncx, ncy = A.ncy;
    if (A.mode == MODE_ALL) {
        const int c = (int)(u / 5), d = (int)(u - 5ll * c);
        const int cy = c / ncx, cx = c - cy * ncx;
        anchor = c;
        if (d == 0) other = c;
        else if (d == 1) { if (cx + 1 >= ncx) return false; other = c + 1; }
        else {
            const int ox = cx + d - 3;   // d = 2,3,4 -> NW, N, NE
            if (cy + 1 >= ncy || ox < 0 || ox >= ncx) return false;
            other = c + ncx + d - 3;
        }
    } else if (A.mode == MODE_SAME) {
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
    aBeg = __ldg(A.cell_start + anchor);
    aEnd = __ldg(A.cell_start + anchor + 1);
    if (aBeg == aEnd) return false;
    same = (anchor == other);
    if (same) {
        if (aEnd - aBeg < 2) return false;
        bBeg = aBeg; bEnd = aEnd;
    } else {
        bBeg = __ldg(A.cell_start + other);
        bEnd = __ldg(A.cell_start + other + 1);
        if (bBeg == bEnd) return false;
    }
    return true;
}

__device__ __forceinline__ bool within_exact(float xa, float ya, float xb, float yb, double r2)
{
    const double dx = __dsub_rn((double)xa, (double)xb), dy = __dsub_rn((double)ya, (double)yb);
    const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    return s <= r2;
}

// interactions.py:13-40 for a pair whose species differ.  Returns the species both particles end
// up with (the rule always leaves them equal), or -1 when either is not rock/paper/scissors
// (winner = None: the draw is consumed, nothing changes).
__device__ __forceinline__ int rps_outcome(int s1, int s2, double u, const RpsDev &R)
{
    if (s1 < 1 || s1 > 3 || s2 < 1 || s2 > 3) return -1;
    // forward winner: rock beats scissors, paper beats rock, scissors beats paper
    int d = s1 - s2;
    if (d < 0) d += 3;
    const int w = (d == 1) ? s1 : s2, l = (d == 1) ? s2 : s1;
    const double p = (w == 1) ? R.pRS : ((w == 2) ? R.pPR : R.pSP);
    return (u < p) ? w : l;
}

template <bool DO_RPS, bool EMIT>
__global__ void __launch_bounds__(PAIR_WARPS * 32) pair_units_kernel(PairArgs A)
{
    __shared__ int2 s_buf[EMIT ? PAIR_WARPS : 1][EMIT ? PAIR_BUF : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long wid = (long long)blockIdx.x * PAIR_WARPS + warp;
    long long u_next = wid * UNITS_PER_WARP;
    const long long u_end = min(u_next + UNITS_PER_WARP, A.n_units);
    if (u_next >= u_end) return;   // warp-uniform

    bool busy = false, same = false;
    int a = 0, aEnd = 0, b = 0, bBeg = 0, bEnd = 0;
    float xa = 0.f, ya = 0.f;
    int nbuf = 0;                       // warp-uniform
    unsigned long long my_hits = 0;     // per lane (used when !EMIT)

    for (;;) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, busy);
        const int n_idle = 32 - __popc(busy_mask);
        if (u_next < u_end && (n_idle >= REFILL_IDLE || busy_mask == 0u)) {
            const int rank = __popc(~busy_mask & lt_mask);
            const long long myu = u_next + rank;
            if (!busy && myu < u_end) {
                int aBeg;
                if (decode_unit(A, myu, aBeg, aEnd, bBeg, bEnd, same)) {
                    busy = true;
                    a = aBeg;
                    xa = __ldg(A.lon + a); ya = __ldg(A.lat + a);
                    b = same ? a + 1 : bBeg;
                }
            }
            u_next += n_idle;
            continue;
        }
        if (busy_mask == 0u) break;

        bool hit = false;
        int2 pr = make_int2(0, 0);
        if (busy) {
            const float xb = __ldg(A.lon + b), yb = __ldg(A.lat + b);
            const float dx = xa - xb, dy = ya - yb;
            const float d2 = fmaf(dx, dx, dy * dy);
            if (d2 <= A.r2_hi) hit = (d2 < A.r2_lo) ? true : within_exact(xa, ya, xb, yb, A.r2);
            if (hit) {
                const int ia = __ldg(A.id + a), ib = __ldg(A.id + b);
                pr = (ia < ib) ? make_int2(ia, ib) : make_int2(ib, ia);
                if (DO_RPS) {
                    // the unit's particles are owned by this lane for the whole launch
                    const int s1 = A.sp[a], s2 = A.sp[b];
                    if (s1 != s2) {
                        const double u = pair_uniform((uint32_t)pr.x, (uint32_t)pr.y, A.rps.step_lo, A.rps.step_hi,
                                                      A.rps.seed_lo, A.rps.seed_hi);
                        const int ns = rps_outcome(s1, s2, u, A.rps);
                        if (ns >= 0) {
                            if (ns != s1) A.sp[a] = (int8_t)ns;
                            if (ns != s2) A.sp[b] = (int8_t)ns;
                        }
                    }
                }
                if (!EMIT) ++my_hits;
            }
            // advance to the next candidate of the unit
            if (++b == bEnd) {
                ++a;
                if (a == aEnd || (same && a + 1 == aEnd)) busy = false;
                else {
                    xa = __ldg(A.lon + a); ya = __ldg(A.lat + a);
                    b = same ? a + 1 : bBeg;
                }
            }
        }
        if (EMIT) {
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hm) {
                if (hit) s_buf[warp][nbuf + __popc(hm & lt_mask)] = pr;
                nbuf += __popc(hm);
                if (nbuf > PAIR_BUF - 32) {
                    __syncwarp();
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&A.ctr->n_pairs, (unsigned long long)nbuf);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int i = lane; i < nbuf; i += 32)
                        if (base + i < A.cap) A.pairs[base + i] = s_buf[warp][i];
                    __syncwarp();
                    nbuf = 0;
                }
            }
        }
    }
    if (EMIT) {
        if (nbuf > 0) {
            __syncwarp();
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&A.ctr->n_pairs, (unsigned long long)nbuf);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int i = lane; i < nbuf; i += 32)
                if (base + i < A.cap) A.pairs[base + i] = s_buf[warp][i];
        }
    } else {
        for (int d = 16; d > 0; d >>= 1) my_hits += __shfl_xor_sync(0xffffffffu, my_hits, d);
        if (lane == 0 && my_hits) atomicAdd(&A.ctr->n_pairs, my_hits);
    }
}

static cudaError_t launch_units(const PairArgs &A, bool do_rps, bool emit, cudaStream_t s, int64_t *launches)
{
    if (A.n_units <= 0) return cudaSuccess;
    const long long per_block = (long long)PAIR_WARPS * UNITS_PER_WARP;
    const long long grid = (A.n_units + per_block - 1) / per_block;
    const dim3 g((unsigned)grid), b(PAIR_WARPS * 32);
    if (do_rps) {
        if (emit) pair_units_kernel<true, true><<<g, b, 0, s>>>(A);
        else pair_units_kernel<true, false><<<g, b, 0, s>>>(A);
    } else {
        if (emit) pair_units_kernel<false, true><<<g, b, 0, s>>>(A);
        else pair_units_kernel<false, false><<<g, b, 0, s>>>(A);
    }
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_pairs(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                         double r, const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    PairArgs A;
    A.lon = lon; A.lat = lat; A.id = id; A.sp = sp;
    A.cell_start = h->cell_start;
    A.ncx = h->grid.ncx; A.ncy = h->grid.ncy;
    A.r2 = r * r;
    A.r2_lo = (float)(A.r2 * (1.0 - 4e-6));
    A.r2_hi = (float)(A.r2 * (1.0 + 4e-6));
    if (rps) A.rps = *rps;
    else A.rps = RpsDev{0, 0, 0, 0, 0, 0, 0};
    A.pairs = pairs_out;
    A.cap = (pairs_out && cap > 0) ? (unsigned long long)cap : 0ull;
    A.ctr = h->ctr;
    const bool emit = A.cap > 0;
    const long long ncx = A.ncx, ncy = A.ncy;
    cudaError_t e = cudaSuccess;
    if (!rps) {
        A.mode = MODE_ALL; A.parity = 0; A.dir = 0;
        A.n_units = 5ll * ncx * ncy;
        return launch_units(A, false, emit, s, &h->launches);
    }
    // phase 0
    A.mode = MODE_SAME; A.parity = 0; A.dir = 0; A.n_units = ncx * ncy;
    e = launch_units(A, true, emit, s, &h->launches);
    if (e != cudaSuccess) return e;
    // phases 1, 2
    for (int q = 0; q < 2; ++q) {
        A.mode = MODE_EAST; A.parity = q; A.dir = 0;
        const long long half = (ncx - q) / 2;
        A.n_units = half * ncy;
        e = launch_units(A, true, emit, s, &h->launches);
        if (e != cudaSuccess) return e;
    }
    // phases 3..8
    for (int q = 0; q < 2; ++q) {
        const long long rows = (ncy - q) / 2;     // anchors rows cy = q, q+2, ... with cy + 1 < ncy
        for (int d = -1; d <= 1; ++d) {
            A.mode = MODE_CROSS; A.parity = q; A.dir = d;
            A.n_units = rows * ncx;
            e = launch_units(A, true, emit, s, &h->launches);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

}  // namespace lm
