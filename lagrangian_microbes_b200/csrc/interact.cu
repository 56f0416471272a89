// Fused radius pair search + rock-paper-scissors resolution on the binned particle arrays: ONE pass, no hand-off.
//
// Replaces  kdt.query_pairs(r=interaction_radius, p=interaction_norm)   (interaction_simulator.py:93-98)
// and       for pair in microbe_pairs: pair_interaction(...)            (interaction_simulator.py:104-105)
//           -> rock_paper_scissors_interaction                          (interactions.py:13-40)
//
// Why this replaces the round-1 pipeline (find_pairs_kernel -> hits[] / rec[] hand-off -> nine resolve_phase launches,
// csrc/pairs.cu, still selectable as LM_OPT_INTERACT_MODE = 0): that pipeline wrote 2.5x its algorithmic bytes to hand
// every pair from the search to the resolver, paid a full Philox draw for every pair whether or not the rule consumes
// it, and resolved a unit's pairs strictly one after the other in (id_a, id_b) order -- so the several-hundred-microbe
// cells a long run collects serialised 10^4..10^5 pairs on one warp (BASELINE config 2 past step 2,000: 2.4 ms of a
// 5 ms step; profiles/r2a_config2_full_mode0.jsonl).  The tiled variant of that resolver lost 2-2.7x on hardware
// (profiles/r2a_tiled_sweep.jsonl).  Here the sequential order itself is chosen so that it parallelises:
//
// CANONICAL ORDER ("tile-round order", oracle/rps.py::tile_round_order).  Cells are grouped into tiles of 32 x 16
// cells.  A unit is one cell or two adjacent cells (half stencil E, NW, N, NE).  Phases 0-8: the units inside one tile
// (0 same cell | 1, 2 east, cx even / odd | 3, 4 north-west, cy even / odd | 5, 6 north | 7, 8 north-east); phases 9-14:
// the units across a tile boundary (east | NW, NE across a vertical boundary | NW, N, NE across a horizontal one).
// Units of one phase touch disjoint microbes.  Inside a unit:
//   LIGHT units (two cells: m_a * m_b <= 256 and m_b <= 64; one cell: m <= 23) -- (rank in the anchor cell, rank in the
//       other cell) lexicographic, ranks by particle id; a handful of pairs, walked by one lane;
//   HEAVY units -- ROUNDS OF MATCHINGS: two cells, M = max(m_a, m_b): round k pairs rank i with rank (i + k) mod M; one
//       cell: the circle method of round-robin tournaments.  No microbe occurs twice in a round, so a round is
//       order-free: a cell of 600 microbes is 600 rounds of 300 independent pairs instead of 180,000 sequential ones.
// The result is by construction the reference's sequential in-place loop run in that total order (checked against the
// unmodified reference function).
//
// KERNELS.  interact_tile_kernel: one CTA per tile, the tile's microbes (positions, ids, species, cell) staged in
// shared memory.  Per direction (same cell, E, NW, N, NE):
//   F  every microbe is a lane: distance tests against the partner cell, hits recorded as (anchor, partner) records in
//      shared memory, contiguous per anchor; then the records are finished DENSELY, lanes <-> records of the whole tile:
//      ids, pair emission (coalesced, one atomic per CTA and direction), and the Philox draw for the records whose two
//      species differ at that moment -- so the draw, the most expensive operation of the step, runs at full lane
//      occupancy and only where the rule can consume it (interactions.py:17-20);
//   R  per parity (one phase each): a lane walks the records of its unit in canonical order against the species in
//      shared memory (a record without a draw whose species have meanwhile come to differ draws on the spot).
// Heavy units go to a queue and are resolved round by round by a whole warp (above 8,192 slots by the whole CTA).  A
// tile that does not fit in shared memory, or whose records overflow, runs the same order through a plain lane walk
// (walk_light) -- as do the units of the boundary phases in interact_cross_kernel (a few per cent of the pairs).
//
// Predicate: exactly SciPy's (float32 positions widened to double, s = fl(dx*dx); s = fl(s + fl(dy*dy)); s <= fl(r*r));
// a float32 evaluation decides everything farther than 4e-6 (relative) from the threshold.  p = 1 / inf likewise.
#include <algorithm>

#include "lm_internal.cuh"
#include "philox.cuh"

namespace lm {
namespace {

constexpr int IT_TW = LM_TILE_W, IT_TH = LM_TILE_H;
constexpr int IT_THREADS = 256;
constexpr int IT_CELLS = IT_TW * IT_TH;
constexpr int IT_UPT = IT_CELLS / IT_THREADS;            // cells a thread looks at when the units of a phase are listed
constexpr int IT_STAGE = 512;                             // pairs staged per CTA between flushes (walk / heavy paths)
constexpr unsigned int IT_MEGA_MIN = 8192;                // slots from which the whole CTA takes a heavy unit
constexpr int IT_MAX_TILE_CAP = 6144;                     // microbes a tile can stage (13-bit local indices in a record)
constexpr unsigned int REC_IDX = 0x1fffu;                 // record: anchor (13 bits) | partner << 13 | draw bits << 26 | has draw << 29
constexpr unsigned int REC_HAS_DRAW = 1u << 29;
constexpr unsigned int FULL = 0xffffffffu;
static_assert(IT_TW == 32 && IT_TH % 2 == 0 && IT_CELLS % IT_THREADS == 0, "tile geometry");

struct IArgs {
    const float *__restrict__ lon;
    const float *__restrict__ lat;
    const int32_t *__restrict__ id;
    int8_t *sp;                          // null: pair search only
    const int32_t *__restrict__ cell_start;
    int ncx, rows_owned, rows_local;     // strip-local rows (single GPU: ncy, ncy)
    int tiles_x, tiles_y;
    int norm;
    float r2_lo, r2_hi;                  // float32 pre-filter window around the threshold
    double r2;                           // the exact threshold (r*r for p = 2, r for p = 1 and p = inf)
    uint32_t pair_key;                   // key of this step's per-pair Philox2x32 stream (philox.cuh)
    unsigned long long thr[3];           // ceil(p * 2^53) for pRS, pPR, pSP
    int2 *__restrict__ pairs;            // null: count only
    unsigned long long cap_pairs;
    Counters *ctr;
    int tile_cap;                        // microbes a tile can stage in shared memory
    int rec_cap;                         // records (hits of one direction) a tile can hold in shared memory
    int draw_batch;                      // lane walk: parked lanes that trigger the warp's Philox rounds
    int force_walk;                      // LM_OPT_TILE_PATH = 1: staged tiles take the lane walk too (tests)
    int phase;                           // interact_cross_kernel: 9..14
};

// ---- the rule (interactions.py:13-40) for species s1 != s2, both in {1,2,3}: the species both end up with --------
__device__ __forceinline__ int rps_apply(int s1, int s2, uint32_t dec)
{
    int d = s1 - s2;
    if (d < 0) d += 3;
    const int w = (d == 1) ? s1 : s2, l = (d == 1) ? s2 : s1;    // w: the forward winner (R > S, P > R, S > P)
    return ((dec >> (w - 1)) & 1u) ? w : l;
}
__device__ __forceinline__ bool is_rps(int s) { return s >= 1 && s <= 3; }

// three decision bits of the pair's draw: bit k set <=> u < p_k  (k = 0: pRS, 1: pPR, 2: pSP); u = m * 2^-53
__device__ __forceinline__ uint32_t decision_bits(const IArgs &A, int i, int j)
{
    const unsigned long long m = pair_draw_m((uint32_t)i, (uint32_t)j, A.pair_key);
    return (m < A.thr[0] ? 1u : 0u) | (m < A.thr[1] ? 2u : 0u) | (m < A.thr[2] ? 4u : 0u);
}

__device__ __forceinline__ bool within_exact(int norm, float2 a, float2 b, double thr)
{
    const double dx = __dsub_rn((double)a.x, (double)b.x), dy = __dsub_rn((double)a.y, (double)b.y);
    if (norm == LM_NORM_2) return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= thr;
    if (norm == LM_NORM_1) return __dadd_rn(fabs(dx), fabs(dy)) <= thr;
    return fmax(fabs(dx), fabs(dy)) <= thr;
}

__device__ __forceinline__ bool within(const IArgs &A, float2 a, float2 b)
{
    const float dx = a.x - b.x, dy = a.y - b.y;
    const float d2 = A.norm == LM_NORM_2 ? fmaf(dx, dx, dy * dy)
                                         : (A.norm == LM_NORM_1 ? fabsf(dx) + fabsf(dy) : fmaxf(fabsf(dx), fabsf(dy)));
    if (d2 > A.r2_hi) return false;
    if (d2 < A.r2_lo) return true;
    return within_exact(A.norm, a, b, A.r2);
}

// ---- light / heavy: part of the definition of the canonical order (oracle/rps.py::tile_round_order) -------------
__device__ __forceinline__ bool unit_is_light(bool same, unsigned int ma, unsigned int mb)
{
    return same ? ma <= 23u : (ma * mb <= 256u && mb <= 64u);
}
__device__ __forceinline__ bool unit_has_pairs(bool same, unsigned int ma, unsigned int mb)
{
    return same ? ma >= 2u : (ma >= 1u && mb >= 1u);
}
// slots of a heavy unit walked in rounds (decides warp vs whole CTA)
__device__ __forceinline__ unsigned int unit_slots(bool same, unsigned int ma, unsigned int mb)
{
    if (same) { const unsigned int M = ma + (ma & 1u); return (M - 1u) * (M >> 1); }
    return ma * max(ma, mb);
}

// ---- where a tile's microbes live: shared memory (staged) or the global arrays ----------------------------------
struct SmemView {
    static constexpr bool kStaged = true;
    const float2 *pos;
    const int32_t *id;
    int8_t *sp;
    const uint16_t *pcell;               // (row in tile) << 5 | column in tile
    uint16_t *hstart;                    // first record of the microbe in the current direction
    uint8_t *hcnt;                       // its records
    uint32_t *rec;                       // records of the current direction
    __device__ __forceinline__ float2 P(int i) const { return pos[i]; }
    __device__ __forceinline__ int I(int i) const { return id[i]; }
    __device__ __forceinline__ int S(int i) const { return ((volatile int8_t *)sp)[i]; }
    __device__ __forceinline__ void W(int i, int s) const { ((volatile int8_t *)sp)[i] = (int8_t)s; }
};
struct GlobalView {
    static constexpr bool kStaged = false;
    const float *lon, *lat;
    const int32_t *id;
    int8_t *sp;
    __device__ __forceinline__ float2 P(int i) const { return make_float2(__ldg(lon + i), __ldg(lat + i)); }
    __device__ __forceinline__ int I(int i) const { return __ldg(id + i); }
    __device__ __forceinline__ int S(int i) const { return ((volatile int8_t *)sp)[i]; }
    __device__ __forceinline__ void W(int i, int s) const { ((volatile int8_t *)sp)[i] = (int8_t)s; }
};

// ---- CTA-wide working set in shared memory ----------------------------------------------------------------------
struct Shared {
    int2 stage[IT_STAGE];                 // found pairs waiting for the next flush (walk / heavy paths)
    uint4 unit[IT_CELLS];                 // units of the current phase: x a_base | y b_base | z m_a | w m_b (b_base == a_base: one cell)
    unsigned int n_light, n_heavy;        // light units grow from the front of unit[], heavy ones from the back
    unsigned int ticket, hticket;
    unsigned int stage_cnt;
    unsigned int rec_used;                // records of the current direction
    unsigned long long flush_base;
};

// One found pair per calling lane (`hit`); all 32 lanes of the warp call this converged.
__device__ __forceinline__ void stage_pairs(const IArgs &A, Shared &sh, bool hit, int lo, int hi)
{
    const unsigned int mh = __ballot_sync(FULL, hit);
    if (!mh) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mh) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(&sh.stage_cnt, (unsigned int)__popc(mh));
    base = __shfl_sync(FULL, base, leader);
    if (hit && A.pairs) {
        const unsigned int k = base + __popc(mh & ((1u << lane) - 1u));
        if (k < (unsigned int)IT_STAGE) sh.stage[k] = make_int2(lo, hi);
        else {                                                        // the stage is full: straight to the list
            const unsigned long long g = atomicAdd(&A.ctr->n_pairs, 1ull);
            if (g < A.cap_pairs) A.pairs[g] = make_int2(lo, hi);
        }
    }
}

// Every thread of the CTA calls this at the same point; the caller has a __syncthreads() before it.
__device__ void flush_pairs(const IArgs &A, Shared &sh, bool force)
{
    const unsigned int cnt = sh.stage_cnt;                            // CTA-uniform ...
    __syncthreads();                                                  // ... because nobody stages before everybody has read it
    if (cnt == 0 || (!force && cnt < (unsigned int)(IT_STAGE / 2))) return;
    const unsigned int n_buf = A.pairs ? min(cnt, (unsigned int)IT_STAGE) : cnt;   // the overflow went out directly
    if (threadIdx.x == 0) sh.flush_base = atomicAdd(&A.ctr->n_pairs, (unsigned long long)n_buf);
    __syncthreads();
    if (A.pairs) {
        const unsigned long long base = sh.flush_base;
        for (unsigned int k = threadIdx.x; k < n_buf; k += IT_THREADS)
            if (base + k < A.cap_pairs) A.pairs[base + k] = sh.stage[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.stage_cnt = 0;
    __syncthreads();
}

// ---- light units walked by one lane: (anchor rank, partner rank) lexicographic ----------------------------------
struct CrossWalk {
    int a_base, b_base, ma, mb, i, j;
    __device__ __forceinline__ void init(const uint4 &u) { a_base = (int)u.x; b_base = (int)u.y; ma = (int)u.z; mb = (int)u.w; i = 0; j = 0; }
    __device__ __forceinline__ int a() const { return a_base + i; }
    __device__ __forceinline__ int b() const { return b_base + j; }
    __device__ __forceinline__ bool next() { if (++j == mb) { j = 0; if (++i == ma) return false; } return true; }   // false: done
};
struct SameWalk {
    int base, m, i, j;
    __device__ __forceinline__ void init(const uint4 &u) { base = (int)u.x; m = (int)u.z; i = 0; j = 1; }
    __device__ __forceinline__ int a() const { return base + i; }
    __device__ __forceinline__ int b() const { return base + j; }
    __device__ __forceinline__ bool next() { if (++j == m) { ++i; j = i + 1; if (j >= m) return false; } return true; }
};

// Every lane walks its own units; a lane whose pair needs a draw parks, and the warp runs the Philox rounds when enough
// lanes are parked (or nobody can advance), so that the draw is paid at a useful lane occupancy.
template <class V, class Walk, bool DO_RPS>
__device__ void walk_light(const IArgs &A, Shared &sh, const V &v)
{
    const unsigned int n_units = sh.n_light;
    Walk w;
    bool active = false, parked = false;
    int pa = 0, pb = 0, plo = 0, phi = 0;                            // the parked pair: local indices, ids (lo < hi)
    unsigned int t = threadIdx.x;                                     // the first unit of a lane is fixed: the ticket starts at IT_THREADS
    if (t < n_units) { w.init(sh.unit[t]); active = true; }
    while (true) {
        const bool can = active && !parked;
        const unsigned int m_can = __ballot_sync(FULL, can);
        const unsigned int m_park = DO_RPS ? __ballot_sync(FULL, parked) : 0u;
        if (!m_can && !m_park) break;
        if (DO_RPS && m_park && (!m_can || __popc(m_park) >= A.draw_batch)) {
            if (parked) {
                const uint32_t dec = decision_bits(A, plo, phi);
                const int sa = v.S(pa), sb = v.S(pb);                 // unchanged since the lane parked: units are disjoint
                const int s = rps_apply(sa, sb, dec);
                if (s != sa) v.W(pa, s); else v.W(pb, s);
                parked = false;
            }
            continue;
        }
        bool hit = false;
        int ia = 0, ib = 0;
        if (can) {
            ia = w.a(); ib = w.b();
            hit = within(A, v.P(ia), v.P(ib));
            if (!w.next()) {                                          // next unit (its first pair is taken next iteration)
                t = atomicAdd(&sh.ticket, 1u);
                if (t < n_units) w.init(sh.unit[t]); else active = false;
            }
        }
        int lo = 0, hi = 0;
        if (hit) { const int x = v.I(ia), y = v.I(ib); lo = min(x, y); hi = max(x, y); }
        stage_pairs(A, sh, hit, lo, hi);
        if (DO_RPS && hit) {
            const int sa = v.S(ia), sb = v.S(ib);
            if (sa != sb && is_rps(sa) && is_rps(sb)) { parked = true; pa = ia; pb = ib; plo = lo; phi = hi; }
        }
    }
}

// ---- heavy units: rounds of matchings ---------------------------------------------------------------------------
// One slot of a cooperative (warp / CTA) round: everything inline, the lanes of a round hold disjoint pairs.
template <class V, bool DO_RPS>
__device__ __forceinline__ void slot_inline(const IArgs &A, Shared &sh, const V &v, bool valid, int ia, int ib)
{
    bool hit = false;
    if (valid) hit = within(A, v.P(ia), v.P(ib));
    int lo = 0, hi = 0;
    if (hit) { const int x = v.I(ia), y = v.I(ib); lo = min(x, y); hi = max(x, y); }
    stage_pairs(A, sh, hit, lo, hi);
    if (DO_RPS && hit) {
        const int sa = v.S(ia), sb = v.S(ib);
        if (sa != sb && is_rps(sa) && is_rps(sb)) {
            const int s = rps_apply(sa, sb, decision_bits(A, lo, hi));
            if (s != sa) v.W(ia, s); else v.W(ib, s);
        }
    }
}

// A unit round by round with `nthr` threads (32: one warp, __syncwarp between rounds; IT_THREADS: the CTA,
// __syncthreads between rounds -- then every thread of the CTA calls this with the same unit).
//   two cells, M = max(m_a, m_b): round k in [0, M) pairs rank i of the anchor cell with rank (i + k) mod M of the other
//   one cell, M = m rounded up to even (rank M - 1 is a phantom when m is odd): round k in [0, M - 1) pairs rank M - 1
//       with rank k and rank (k + j) mod (M - 1) with rank (k - j) mod (M - 1) for j in [1, M / 2)
template <class V, bool DO_RPS, bool CTA>
__device__ void unit_rounds(const IArgs &A, Shared &sh, const V &v, const uint4 u)
{
    const int nthr = CTA ? IT_THREADS : 32;
    const int me = CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const int a_base = (int)u.x, b_base = (int)u.y, ma = (int)u.z, mb = (int)u.w;
    if (a_base == b_base) {
        const int m = ma, M = m + (m & 1), n1 = M - 1, half = M >> 1;
        for (int k = 0; k < n1; ++k) {
            for (int j0 = 0; j0 < half; j0 += nthr) {                 // uniform trip count: stage_pairs needs whole warps
                const int jj = j0 + me;
                int ra = 0, rb = 0;
                bool valid = jj < half;
                if (valid) {
                    if (jj == 0) { ra = M - 1; rb = k; }
                    else { ra = k + jj; if (ra >= n1) ra -= n1; rb = k - jj; if (rb < 0) rb += n1; }
                    valid = ra < m && rb < m;
                }
                slot_inline<V, DO_RPS>(A, sh, v, valid, a_base + min(ra, rb), a_base + max(ra, rb));
            }
            if (CTA) { __syncthreads(); flush_pairs(A, sh, false); } else __syncwarp();
        }
    } else {
        const int M = max(ma, mb);
        for (int k = 0; k < M; ++k) {
            for (int i0 = 0; i0 < ma; i0 += nthr) {
                const int i = i0 + me;
                int j = i + k;
                if (j >= M) j -= M;
                slot_inline<V, DO_RPS>(A, sh, v, i < ma && j < mb, a_base + i, b_base + j);
            }
            if (CTA) { __syncthreads(); flush_pairs(A, sh, false); } else __syncwarp();
        }
    }
}

// The units listed in sh.unit[] (light from the front, heavy from the back); every thread of the CTA calls this.
template <class V, bool DO_RPS>
__device__ void run_units(const IArgs &A, Shared &sh, const V &v, bool same)
{
    __syncthreads();                                                  // the list is complete
    if (sh.n_light) {
        if (same) walk_light<V, SameWalk, DO_RPS>(A, sh, v);
        else walk_light<V, CrossWalk, DO_RPS>(A, sh, v);
    }
    const unsigned int n_heavy = sh.n_heavy;                          // CTA-uniform
    if (n_heavy) {
        // whole warps pull the heavy units; the very large ones are left to the whole CTA afterwards
        const int lane = threadIdx.x & 31;
        while (true) {
            unsigned int hq = 0;
            if (lane == 0) hq = atomicAdd(&sh.hticket, 1u);
            hq = __shfl_sync(FULL, hq, 0);
            if (hq >= n_heavy) break;
            const uint4 u = sh.unit[IT_CELLS - 1 - hq];
            if (unit_slots(u.x == u.y, u.z, u.w) >= IT_MEGA_MIN) continue;
            unit_rounds<V, DO_RPS, false>(A, sh, v, u);
        }
        __syncthreads();
        for (unsigned int hq = 0; hq < n_heavy; ++hq) {
            const uint4 u = sh.unit[IT_CELLS - 1 - hq];
            if (unit_slots(u.x == u.y, u.z, u.w) < IT_MEGA_MIN) continue;
            unit_rounds<V, DO_RPS, true>(A, sh, v, u);
        }
    }
    __syncthreads();
    flush_pairs(A, sh, false);
}

// Append a unit to the list; all 32 lanes of the warp call this converged (`light` / `heavy`: this lane has one).
__device__ __forceinline__ void push_unit(Shared &sh, bool light, bool heavy, uint4 u)
{
    const int lane = threadIdx.x & 31;
    const unsigned int ml = __ballot_sync(FULL, light), mh = __ballot_sync(FULL, heavy);
    if (ml) {
        const int leader = __ffs(ml) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&sh.n_light, (unsigned int)__popc(ml));
        base = __shfl_sync(FULL, base, leader);
        if (light) sh.unit[base + __popc(ml & ((1u << lane) - 1u))] = u;
    }
    if (mh) {
        const int leader = __ffs(mh) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&sh.n_heavy, (unsigned int)__popc(mh));
        base = __shfl_sync(FULL, base, leader);
        if (heavy) sh.unit[IT_CELLS - 1 - (base + __popc(mh & ((1u << lane) - 1u)))] = u;
    }
}

// ---- stage I: the units inside one tile ---------------------------------------------------------------------------
// direction of a group: 0 same cell, 1 E, 2 NW, 3 N, 4 NE; phases: 0 | 1, 2 | 3, 4 | 5, 6 | 7, 8 (parity of cx for E, of cy else)
struct TileGeom {
    const int (*cs)[IT_TW + 1];          // cell_start of the tile's cells (global particle indices), one more per row
    const int *delta;                    // local index = global index + delta[row]
    int w, hgt;
};

// the unit anchored at cell (t, x) in direction g: false if its partner cell is not in this tile
__device__ __forceinline__ bool unit_of_cell(const TileGeom &G, int g, int t, int x, uint4 &u)
{
    int to = t, xo = x;
    if (g == 1) xo = x + 1;
    else if (g >= 2) { to = t + 1; xo = x + (g - 3); }
    if (t >= G.hgt || x >= G.w || to >= G.hgt || xo < 0 || xo >= G.w) return false;
    const int a0 = G.cs[t][x], a1 = G.cs[t][x + 1], b0 = G.cs[to][xo], b1 = G.cs[to][xo + 1];
    u = make_uint4((unsigned int)(a0 + G.delta[t]), (unsigned int)(b0 + G.delta[to]), (unsigned int)(a1 - a0), (unsigned int)(b1 - b0));
    return true;
}

// F of one direction on a staged tile: records of every light unit of direction g, contiguous per anchor.  Returns
// false (CTA-uniform) if they do not fit -- nothing has been emitted then, and the direction takes the lane walk.
template <bool DO_RPS>
__device__ bool find_direction(const IArgs &A, Shared &sh, const SmemView &v, const TileGeom &G, int g, int total)
{
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) sh.rec_used = 0;
    __syncthreads();
    for (int p0 = 0; p0 < total; p0 += IT_THREADS) {
        const int p = p0 + tid;
        int nc = 0, beg = 0;
        float2 pa = make_float2(0.f, 0.f);
        if (p < total) {
            const int pc = v.pcell[p];
            uint4 u;
            if (unit_of_cell(G, g, pc >> 5, pc & 31, u) && unit_is_light(g == 0, u.z, u.w)) {
                if (g == 0) { beg = p + 1; nc = (int)(u.x + u.z) - beg; }
                else { beg = (int)u.y; nc = (int)u.w; }
                pa = v.pos[p];
            }
        }
        // hit bits of candidates 0..31 | 32..63 (two 32-bit words: cheaper to build and to walk than one 64-bit mask)
        unsigned int m_lo = 0, m_hi = 0;
        {
            const int n_lo = min(nc, 32);
            if (A.norm == LM_NORM_2) {                                // the reference's norm: no selects in the loop
                for (int i = 0; i < n_lo; ++i) {
                    const float2 pb = v.pos[beg + i];
                    const float dx = pa.x - pb.x, dy = pa.y - pb.y, d2 = fmaf(dx, dx, dy * dy);
                    if (d2 <= A.r2_hi && (d2 < A.r2_lo || within_exact(LM_NORM_2, pa, pb, A.r2))) m_lo |= 1u << i;
                }
            } else {
                for (int i = 0; i < n_lo; ++i)
                    if (within(A, pa, v.pos[beg + i])) m_lo |= 1u << i;
            }
            for (int i = 32; i < nc; ++i)
                if (within(A, pa, v.pos[beg + i])) m_hi |= 1u << (i - 32);
        }
        const unsigned int cnt = (unsigned int)(__popc(m_lo) + __popc(m_hi));
        unsigned int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int up = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += up;
        }
        const unsigned int wtotal = __shfl_sync(FULL, incl, 31);
        unsigned int wbase = 0;
        if (lane == 0 && wtotal) wbase = atomicAdd(&sh.rec_used, wtotal);
        wbase = __shfl_sync(FULL, wbase, 0);
        unsigned int pos = wbase + incl - cnt;
        if (p < total) { v.hstart[p] = (uint16_t)min(pos, 0xffffu); v.hcnt[p] = (uint8_t)cnt; }
        if (wbase + wtotal <= (unsigned int)A.rec_cap) {
            const unsigned int lo_part = (unsigned int)p | ((unsigned int)beg << 13);
            for (; m_lo; m_lo &= m_lo - 1) v.rec[pos++] = lo_part + ((unsigned int)(__ffs((int)m_lo) - 1) << 13);
            for (; m_hi; m_hi &= m_hi - 1) v.rec[pos++] = lo_part + ((unsigned int)(__ffs((int)m_hi) + 31) << 13);
        }
    }
    __syncthreads();
    const unsigned int n_rec = sh.rec_used;                           // CTA-uniform
    if (n_rec > (unsigned int)A.rec_cap) return false;
    // ---- finish densely: lanes <-> records of the whole tile
    if (n_rec) {
        if (tid == 0) sh.flush_base = atomicAdd(&A.ctr->n_pairs, (unsigned long long)n_rec);
        __syncthreads();
        const unsigned long long base = sh.flush_base;
        for (unsigned int e = tid; e < n_rec; e += IT_THREADS) {
            unsigned int r = v.rec[e];
            const int a = (int)(r & REC_IDX), b = (int)((r >> 13) & REC_IDX);
            const int x = v.id[a], y = v.id[b];
            const int lo = min(x, y), hi = max(x, y);
            if (A.pairs && base + e < A.cap_pairs) A.pairs[base + e] = make_int2(lo, hi);
            if (DO_RPS) {
                const int sa = v.sp[a], sb = v.sp[b];
                if (sa != sb && is_rps(sa) && is_rps(sb)) v.rec[e] = r | (decision_bits(A, lo, hi) << 26) | REC_HAS_DRAW;
            }
        }
    }
    __syncthreads();
    return true;
}

// R: one lane resolves one light unit from its records, in canonical order (anchor rank, partner rank).  The records
// of one anchor are contiguous; those of a cell's anchors are contiguous too when the anchors sat in one warp of F
// (same 32-aligned block of local indices: 9 in 10 cells) -- then the lane walks ONE run of records and never looks at
// an anchor without records.
__device__ __forceinline__ void resolve_one(const IArgs &A, const SmemView &v, uint32_t e, int a, int &sa)
{
    const int b = (int)((e >> 13) & REC_IDX);
    const int sb = v.sp[b];
    if (sa != sb && is_rps(sa) && is_rps(sb)) {
        uint32_t dec;
        if (e & REC_HAS_DRAW) dec = (e >> 26) & 7u;
        else { const int x = v.id[a], y = v.id[b]; dec = decision_bits(A, min(x, y), max(x, y)); }   // came to differ since F
        const int s = rps_apply(sa, sb, dec);
        if (s != sa) sa = s; else v.sp[b] = (int8_t)s;
    }
}

// u.x first record | u.y records (0xffffffff: the cell's anchors straddle two warps of F, walk anchor by anchor) | u.z first anchor | u.w anchors
__device__ __forceinline__ void resolve_records(const IArgs &A, const SmemView &v, const uint4 &u)
{
    if (u.y != 0xffffffffu) {
        int cur = -1, sa = 0, sa0 = 0;
        const uint32_t *r = v.rec + u.x;
        for (unsigned int h = 0; h < u.y; ++h) {
            const uint32_t e = r[h];
            const int a = (int)(e & REC_IDX);
            if (a != cur) {
                if (cur >= 0 && sa != sa0) v.sp[cur] = (int8_t)sa;   // a same-cell partner may be this very microbe later: write first
                cur = a; sa = sa0 = v.sp[a];
            }
            resolve_one(A, v, e, a, sa);
        }
        if (cur >= 0 && sa != sa0) v.sp[cur] = (int8_t)sa;
        return;
    }
    const int a_end = (int)(u.z + u.w);
    for (int a = (int)u.z; a < a_end; ++a) {
        const int hc = v.hcnt[a];
        if (!hc) continue;
        const uint32_t *r = v.rec + v.hstart[a];
        const int sa0 = v.sp[a];
        int sa = sa0;
        for (int h = 0; h < hc; ++h) resolve_one(A, v, r[h], a, sa);
        if (sa != sa0) v.sp[a] = (int8_t)sa;
    }
}

// every lane takes units with records from the list (dynamic ticket) until it is empty
__device__ void walk_records(const IArgs &A, Shared &sh, const SmemView &v)
{
    const unsigned int n_units = sh.n_light;
    unsigned int t = threadIdx.x;
    while (t < n_units) {
        resolve_records(A, v, sh.unit[t]);
        t = atomicAdd(&sh.ticket, 1u);
    }
}

template <class V, bool DO_RPS>
__device__ void tile_phases(const IArgs &A, Shared &sh, const V &v, const TileGeom &G, int total)
{
    const int tid = threadIdx.x;
    for (int g = 0; g < 5; ++g) {
        bool records = false;
        if constexpr (V::kStaged) {
            if (!A.force_walk) records = find_direction<DO_RPS>(A, sh, v, G, g, total);
        }
        for (int par = 0; par < (g == 0 ? 1 : 2); ++par) {            // one phase
            if (tid == 0) { sh.n_light = 0; sh.n_heavy = 0; sh.ticket = IT_THREADS; sh.hticket = 0; }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < IT_UPT; ++q) {
                const int e = tid + q * IT_THREADS;
                const int t = e >> 5, x = e & 31;                     // IT_TW == 32
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                bool have = (g == 0 || ((g == 1 ? x : t) & 1) == par) && unit_of_cell(G, g, t, x, u);
                have = have && unit_has_pairs(g == 0, u.z, u.w);
                const bool light = have && unit_is_light(g == 0, u.z, u.w);
                bool listed = light;
                uint4 w4 = u;
                if constexpr (V::kStaged) {
                    if (records) {
                        listed = false;
                        if (DO_RPS && light) {
                            // the unit's records: one run if its anchors sat in one warp of F, else anchor by anchor
                            const int a0 = (int)u.x, al = (int)(u.x + u.z) - 1;
                            if ((a0 >> 5) == (al >> 5)) {
                                const unsigned int rs = v.hstart[a0], re = (unsigned int)v.hstart[al] + v.hcnt[al];
                                w4 = make_uint4(rs, re - rs, u.x, u.z);
                                listed = re > rs;
                            } else {
                                w4 = make_uint4(0u, 0xffffffffu, u.x, u.z);
                                listed = true;
                            }
                        }
                    }
                }
                push_unit(sh, listed, have && !light, listed ? w4 : u);
            }
            if constexpr (V::kStaged) {
                if (records) {
                    __syncthreads();                                   // the list is complete
                    if (DO_RPS && sh.n_light) walk_records(A, sh, v);
                    __syncthreads();
                    if (tid == 0) sh.n_light = 0;                      // run_units: only the heavy units are left
                }
            }
            run_units<V, DO_RPS>(A, sh, v, g == 0);
        }
    }
}

template <bool DO_RPS>
__global__ void __launch_bounds__(IT_THREADS) interact_tile_kernel(IArgs A)
{
    // float2 pos[cap] | uint32 rec[rec_cap] | int32 id[cap] | uint16 pcell[cap] | uint16 hstart[cap] | int8 sp[cap] | uint8 hcnt[cap]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ Shared sh;
    __shared__ int s_cs[IT_TH][IT_TW + 1];
    __shared__ int s_delta[IT_TH];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x;
    const int x0 = tx * IT_TW, y0 = ty * IT_TH;
    TileGeom G;
    G.cs = s_cs; G.delta = s_delta;
    G.w = min(IT_TW, A.ncx - x0); G.hgt = min(IT_TH, A.rows_owned - y0);
    const int w = G.w, hgt = G.hgt;
    for (int k = tid; k < hgt * (w + 1); k += IT_THREADS) {
        const int t = k / (w + 1), x = k - t * (w + 1);
        s_cs[t][x] = __ldg(A.cell_start + (long long)(y0 + t) * A.ncx + x0 + x);
    }
    if (tid == 0) sh.stage_cnt = 0;
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int t = 0; t < hgt; ++t) { s_delta[t] = off - s_cs[t][0]; off += s_cs[t][w] - s_cs[t][0]; }
        s_total = off;
    }
    __syncthreads();
    const int total = s_total;
    if (total < 2) return;                                            // CTA-uniform: no pair inside this tile
    if (total <= A.tile_cap) {
        const int cap = A.tile_cap;
        float2 *s_pos = reinterpret_cast<float2 *>(s_dyn);
        uint32_t *s_rec = reinterpret_cast<uint32_t *>(s_pos + cap);
        int32_t *s_id = reinterpret_cast<int32_t *>(s_rec + A.rec_cap);
        uint16_t *s_pcell = reinterpret_cast<uint16_t *>(s_id + cap);
        uint16_t *s_hstart = s_pcell + cap;
        int8_t *s_sp = reinterpret_cast<int8_t *>(s_hstart + cap);
        uint8_t *s_hcnt = reinterpret_cast<uint8_t *>(s_sp + cap);
        for (int t = 0; t < hgt; ++t) {
            const int p0 = s_cs[t][0], cnt = s_cs[t][w] - p0, d = s_delta[t] + p0;
            for (int i = tid; i < cnt; i += IT_THREADS) {
                const int gi = p0 + i;
                s_pos[d + i] = make_float2(__ldg(A.lon + gi), __ldg(A.lat + gi));
                s_id[d + i] = __ldg(A.id + gi);
                if (DO_RPS) s_sp[d + i] = A.sp[gi];
                int lo = 0, hi = w;                                   // the cell: cs[t][lo] <= gi < cs[t][lo + 1]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_cs[t][mid] <= gi) lo = mid; else hi = mid; }
                s_pcell[d + i] = (uint16_t)((t << 5) | lo);
            }
        }
        __syncthreads();
        SmemView v{s_pos, s_id, s_sp, s_pcell, s_hstart, s_hcnt, s_rec};
        tile_phases<SmemView, DO_RPS>(A, sh, v, G, total);
        if (DO_RPS) {
            for (int t = 0; t < hgt; ++t) {
                const int p0 = s_cs[t][0], cnt = s_cs[t][w] - p0, d = s_delta[t] + p0;
                for (int i = tid; i < cnt; i += IT_THREADS) A.sp[p0 + i] = s_sp[d + i];
            }
        }
    } else {
        __syncthreads();
        if (tid < hgt) s_delta[tid] = 0;                              // local index == global index
        __syncthreads();
        GlobalView v{A.lon, A.lat, A.id, A.sp};
        tile_phases<GlobalView, DO_RPS>(A, sh, v, G, total);
    }
    __syncthreads();
    flush_pairs(A, sh, true);
}

// ---- stage II: the units of ONE boundary phase (9..14), from global memory --------------------------------------
//   9  east       anchor (bx*32 - 1, cy)             bx = 1 .. tiles_x - 1, every owned row
//  10  NW, 11 NE  across a vertical boundary only:   anchor column bx*32 (NW) / bx*32 - 1 (NE), rows with cy % 16 != 15
//  12  NW, 13 N, 14 NE across a horizontal boundary: anchor row by*16 - 1, by = 1 .. (rows_local - 1) / 16, every column
// (the partner row of the last boundary may be the ghost row of a strip: rows_local = rows_owned + 1)
template <bool DO_RPS>
__global__ void __launch_bounds__(IT_THREADS) interact_cross_kernel(IArgs A, long long n_units)
{
    __shared__ Shared sh;
    const int tid = threadIdx.x;
    if (tid == 0) { sh.stage_cnt = 0; sh.n_light = 0; sh.n_heavy = 0; sh.ticket = IT_THREADS; sh.hticket = 0; }
    __syncthreads();
    const int nbx = A.tiles_x - 1;
    const int ph = A.phase;
    const long long base = (long long)blockIdx.x * IT_CELLS;
#pragma unroll
    for (int q = 0; q < IT_UPT; ++q) {
        const long long e = base + tid + q * IT_THREADS;
        bool have = e < n_units;
        int cx = 0, cy = 0, ox = 0, oy = 0;
        if (have) {
            if (ph <= 11) {
                cy = (int)(e / nbx);
                const int bx = 1 + (int)(e - (long long)cy * nbx);
                if (ph == 9) { cx = bx * IT_TW - 1; ox = cx + 1; oy = cy; }
                else {
                    cx = ph == 10 ? bx * IT_TW : bx * IT_TW - 1;
                    ox = ph == 10 ? cx - 1 : cx + 1; oy = cy + 1;
                    have = (cy % IT_TH) != IT_TH - 1 && oy < A.rows_local;
                }
            } else {
                const int by = 1 + (int)(e / A.ncx);
                cx = (int)(e - (long long)(by - 1) * A.ncx);
                cy = by * IT_TH - 1; oy = cy + 1; ox = cx + (ph - 13);
                have = ox >= 0 && ox < A.ncx;
            }
        }
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (have) {
            const long long ca = (long long)cy * A.ncx + cx, cb = (long long)oy * A.ncx + ox;
            const int a0 = __ldg(A.cell_start + ca), a1 = __ldg(A.cell_start + ca + 1);
            const int b0 = __ldg(A.cell_start + cb), b1 = __ldg(A.cell_start + cb + 1);
            u = make_uint4((unsigned int)a0, (unsigned int)b0, (unsigned int)(a1 - a0), (unsigned int)(b1 - b0));
        }
        have = have && unit_has_pairs(false, u.z, u.w);
        const bool light = have && unit_is_light(false, u.z, u.w);
        push_unit(sh, light, have && !light, u);
    }
    GlobalView v{A.lon, A.lat, A.id, A.sp};
    run_units<GlobalView, DO_RPS>(A, sh, v, false);
    __syncthreads();
    flush_pairs(A, sh, true);
}

// ---- LM_OPT_INTERACT_MODE = 2 (hybrid): the HEAVY units of one phase of the cell-round order, from a device-wide queue.
// The pair search of the round-1 pipeline (csrc/pairs.cu) leaves the heavy units out and queues them per phase; here
// whole warps take them round by round (the large ones the whole CTA).  The microbes of the unit's one or two cells are
// staged in shared memory first -- a round is then a shared-memory round trip, not a global one: the rounds are what is
// sequential, a cell of 340 microbes has 339 of them -- distance tests, pair emission, draws and species updates inline;
// every lane of a round holds a pair of its own.
constexpr int HV_WARP_CAP = 256;                          // microbes of a unit one warp stages (13 B each)
constexpr int HV_CTA_CAP = 4096;                          // ... the whole CTA (larger units work on the global arrays)
constexpr int HV_WARP_BYTES = HV_WARP_CAP * 13;
constexpr int HV_CTA_BYTES = HV_CTA_CAP * 13;
constexpr int HV_DYN_BYTES = HV_CTA_BYTES > 8 * HV_WARP_BYTES ? HV_CTA_BYTES : 8 * HV_WARP_BYTES;
static_assert(HV_WARP_BYTES % 16 == 0 && (HV_WARP_CAP * 13) % 16 == 0 && (HV_CTA_CAP * 13) % 16 == 0, "alignment of the staging buffers");
struct UnitView {
    static constexpr bool kStaged = true;
    const float2 *pos;
    const int32_t *id;
    int8_t *sp;
    __device__ __forceinline__ float2 P(int i) const { return pos[i]; }
    __device__ __forceinline__ int I(int i) const { return id[i]; }
    __device__ __forceinline__ int S(int i) const { return ((volatile int8_t *)sp)[i]; }
    __device__ __forceinline__ void W(int i, int s) const { ((volatile int8_t *)sp)[i] = (int8_t)s; }
};

// (Tried and measured, profiles/r2k_config2_probe_two_sweep_rounds.jsonl: evaluating the slots of a chunk of rounds first --
// distance test, emission and draw do not depend on species -- into one code byte per slot, and keeping only the rule in
// the sequential rounds.  Slower: BASELINE config 2's worst step 1.39 -> 2.07 ms.  One warp issues the same instructions
// either way, and the second sweep plus the code bytes come on top; the round is bound by issue, not by its dependences.)
// ug: x, y first microbe of the anchor / other cell (global indices; equal: one cell) | z, w their sizes.  buf: room for
// `cap` microbes.  Called by all lanes of a warp (CTA = false) or all threads of the CTA (CTA = true) with the same unit.
template <bool DO_RPS, bool CTA>
__device__ void heavy_unit(const IArgs &A, Shared &sh, const uint4 ug, unsigned char *buf, int cap)
{
    const int nthr = CTA ? IT_THREADS : 32;
    const int me = CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const bool same = ug.x == ug.y;
    const int ma = (int)ug.z, mb = same ? 0 : (int)ug.w, n = ma + mb;
    if (n > cap) {
        GlobalView g{A.lon, A.lat, A.id, A.sp};
        unit_rounds<GlobalView, DO_RPS, CTA>(A, sh, g, ug);
        return;
    }
    float2 *s_pos = reinterpret_cast<float2 *>(buf);
    int32_t *s_id = reinterpret_cast<int32_t *>(s_pos + cap);
    int8_t *s_sp = reinterpret_cast<int8_t *>(s_id + cap);
    if (CTA) __syncthreads(); else __syncwarp();                      // the buffer's previous unit is done with
    for (int i = me; i < n; i += nthr) {
        const int g = i < ma ? (int)ug.x + i : (int)ug.y + (i - ma);
        s_pos[i] = make_float2(__ldg(A.lon + g), __ldg(A.lat + g));
        s_id[i] = __ldg(A.id + g);
        if (DO_RPS) s_sp[i] = A.sp[g];
    }
    if (CTA) __syncthreads(); else __syncwarp();
    UnitView v{s_pos, s_id, s_sp};
    unit_rounds<UnitView, DO_RPS, CTA>(A, sh, v, make_uint4(0u, same ? 0u : (unsigned int)ma, ug.z, ug.w));
    if (CTA) __syncthreads(); else __syncwarp();
    if (DO_RPS)
        for (int i = me; i < n; i += nthr) A.sp[i < ma ? (int)ug.x + i : (int)ug.y + (i - ma)] = s_sp[i];
}

template <bool DO_RPS>
__global__ void __launch_bounds__(IT_THREADS) interact_heavy_kernel(IArgs A, const int2 *__restrict__ list_w, const int2 *__restrict__ list_m,
                                                                    unsigned int *cnt, unsigned int cap)
{
    // cnt: [0] queued warp units | [1] queued CTA units | [2] warp ticket | [3] CTA ticket   (of this phase)
    extern __shared__ __align__(16) unsigned char s_buf[];            // HV_DYN_BYTES: 8 warp buffers, or one for the CTA
    __shared__ Shared sh;
    __shared__ unsigned int s_t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_w = min(cnt[0], cap), n_m = min(cnt[1], cap);
    if (n_w == 0 && n_m == 0) return;
    if (tid == 0) sh.stage_cnt = 0;
    __syncthreads();
    auto unit_of = [&](int2 cc) {
        const int a0 = __ldg(A.cell_start + cc.x), a1 = __ldg(A.cell_start + cc.x + 1);
        const int b0 = __ldg(A.cell_start + cc.y), b1 = __ldg(A.cell_start + cc.y + 1);
        return make_uint4((unsigned int)a0, (unsigned int)b0, (unsigned int)(a1 - a0), (unsigned int)(b1 - b0));
    };
    while (true) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(&cnt[2], 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_w) break;
        heavy_unit<DO_RPS, false>(A, sh, unit_of(list_w[t]), s_buf + (size_t)warp * HV_WARP_BYTES, HV_WARP_CAP);
    }
    __syncthreads();
    flush_pairs(A, sh, false);
    while (true) {
        if (tid == 0) s_t = atomicAdd(&cnt[3], 1u);
        __syncthreads();
        const unsigned int t = s_t;
        __syncthreads();
        if (t >= n_m) break;
        heavy_unit<DO_RPS, true>(A, sh, unit_of(list_m[t]), s_buf, HV_CTA_CAP);
    }
    __syncthreads();
    flush_pairs(A, sh, true);
}

static void fill_args(IArgs &A, lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, double r,
                      const RpsDev *rps, int2 *pairs_out, int64_t cap);

}  // namespace

// hybrid mode: the queued heavy units of phase `ph` (0..8, cell-phase numbering); arguments as saved by launch_find
cudaError_t launch_interact_heavy(lm_handle_s *h, int8_t *sp, int ph, cudaStream_t s)
{
    if (h->ia_n <= 0 || !h->heavy_list) return cudaSuccess;
    IArgs A;
    fill_args(A, h, h->ia_lon, h->ia_lat, h->ia_id, sp, h->ia_r, h->ia_have_rps ? &h->ia_rps : nullptr, h->ia_pairs, h->ia_cap);
    const int2 *lw = h->heavy_list + (size_t)(2 * ph) * h->heavy_cap, *lm_ = h->heavy_list + (size_t)(2 * ph + 1) * h->heavy_cap;
    unsigned int *cnt = h->heavy_cnt + 4 * ph;
    const unsigned int grid = 2 * kNumSMs;
    const size_t dyn = (size_t)HV_DYN_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(interact_heavy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(interact_heavy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (h->ia_have_rps) interact_heavy_kernel<true><<<grid, IT_THREADS, dyn, s>>>(A, lw, lm_, cnt, (unsigned int)h->heavy_cap);
    else interact_heavy_kernel<false><<<grid, IT_THREADS, dyn, s>>>(A, lw, lm_, cnt, (unsigned int)h->heavy_cap);
    ++h->launches;
    return cudaGetLastError();
}

namespace {
static void fill_args(IArgs &A, lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, double r,
                      const RpsDev *rps, int2 *pairs_out, int64_t cap)
{
    A.lon = lon; A.lat = lat; A.id = id; A.sp = rps ? sp : nullptr; A.cell_start = h->cell_start;
    A.ncx = h->grid.ncx; A.rows_owned = h->strip.rows_owned; A.rows_local = h->strip.rows_local;
    A.tiles_x = (A.ncx + IT_TW - 1) / IT_TW; A.tiles_y = (A.rows_owned + IT_TH - 1) / IT_TH;
    A.norm = h->norm;
    A.r2 = h->norm == LM_NORM_2 ? r * r : r;           // SciPy: tub = r*r for p=2, pow(r, 1) for p=1, r for p=inf
    A.r2_lo = (float)(A.r2 * (1.0 - 4e-6));
    A.r2_hi = (float)(A.r2 * (1.0 + 4e-6));
    A.pair_key = 0;
    A.thr[0] = A.thr[1] = A.thr[2] = 0;
    if (rps) {
        A.pair_key = rps->pair_key;
        const double p[3] = {rps->pRS, rps->pPR, rps->pSP};
        for (int k = 0; k < 3; ++k) {
            // u = m * 2^-53 with integer m < 2^53:  u < p  <=>  m < ceil(p * 2^53)   (the scaling is exact)
            if (!(p[k] > 0.0)) A.thr[k] = 0;
            else if (p[k] >= 1.0) A.thr[k] = 1ull << 53;
            else A.thr[k] = (unsigned long long)ceil(p[k] * 9007199254740992.0);
        }
    }
    A.pairs = (pairs_out && cap > 0) ? pairs_out : nullptr;
    A.cap_pairs = A.pairs ? (unsigned long long)cap : 0ull;
    A.ctr = h->ctr;
    A.draw_batch = h->draw_batch > 0 ? h->draw_batch : 8;
    A.force_walk = h->tile_path == 1 ? 1 : 0;
    A.phase = 0;
    A.tile_cap = A.rec_cap = 0;
}
}  // namespace

// Phases [first, last] of 0..14 (0..8 are ONE launch: any of them asks for all nine).
cudaError_t launch_interact(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                            double r, const RpsDev *rps, int2 *pairs_out, int64_t cap, int first, int last,
                            cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    IArgs A;
    fill_args(A, h, lon, lat, id, sp, r, rps, pairs_out, cap);
    // shared memory per tile: room for 1.5 x the mean occupancy of a tile (at least 2,048 microbes, at most 6,144) and for
    // two records per staged microbe and direction (the mean is below one); fuller tiles / directions take the lane walk
    const long long tiles = (long long)A.tiles_x * A.tiles_y;
    long long want = h->tile_cap > 0 ? h->tile_cap : std::max<long long>(2048, 3 * ((long long)n / std::max<long long>(1, tiles)) / 2 + 256);
    want = std::min<long long>(want, IT_MAX_TILE_CAP);
    A.tile_cap = (int)((want + 255) / 256 * 256);
    A.rec_cap = h->tile_rec_cap > 0 ? h->tile_rec_cap : std::min(2 * A.tile_cap, 16384);
    const size_t dyn = (size_t)A.tile_cap * 18 + (size_t)A.rec_cap * 4;
    cudaError_t e = cudaSuccess;
    if (first <= 8) {
        auto k = rps ? interact_tile_kernel<true> : interact_tile_kernel<false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        k<<<(unsigned int)tiles, IT_THREADS, dyn, s>>>(A);
        ++h->launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    for (int ph = std::max(first, 9); ph <= last; ++ph) {
        long long n_units;
        if (ph <= 11) n_units = (long long)(A.tiles_x - 1) * A.rows_owned;
        else n_units = (long long)((A.rows_local - 1) / IT_TH) * A.ncx;
        if (n_units <= 0) continue;
        A.phase = ph;
        const unsigned int grid = (unsigned int)((n_units + IT_CELLS - 1) / IT_CELLS);
        if (rps) interact_cross_kernel<true><<<grid, IT_THREADS, 0, s>>>(A, n_units);
        else interact_cross_kernel<false><<<grid, IT_THREADS, 0, s>>>(A, n_units);
        ++h->launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace lm
