// Binning: put the particle records into (cell key, particle id) order and build cell_start[].
//
// This is the B200 stand-in for  cKDTree(np.array(microbe_locations))  (interaction_simulator.py:93):
// the spatial index the radius query runs on.  A uniform grid with cell edge h > r replaces the
// kd-tree.  Counting sort instead of a radix sort: the key range (cells) is of the
// order of the particle count, so one histogram + one scan + one scatter move ~50 B per particle
// where an LSD radix sort of (key, payload) moves > 100 B.
//
//   1. bin_count    key[p] = cy*ncx + cx, cell_count[key]++           (REDG atomics, L2)
//   2. scan         cell_start = exclusive_scan(cell_count)            (3 small kernels over cells)
//   3. bin_scatter  slot = cursor[key]++ ; slots[slot] = (p, id[p])    (arbitrary order inside a cell)
//   4. bin_reorder  rank each slot among its cell mates by id, move the record to cell_start + rank
//
// Step 4 makes the storage order canonical -- (cell, id) -- independent of atomic arrival order,
// of the previous storage order and (multi-GPU) of how particles were distributed over ranks.
// The fused RPS resolver's pair order is defined on it (oracle/rps.py::cell_phase_order).
#include <climits>
#include "lm_internal.cuh"

namespace lm {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// cell = clamp(floor((double(v) - origin) * inv_h), 0, n-1); NaN -> 0.  Mirrored bit for bit by
// oracle/pairs.py::cell_index.
__device__ __forceinline__ int cell_coord(float v, double origin, double inv_h, int n, bool &clamped)
{
    const double q = floor(__dmul_rn(__dsub_rn((double)v, origin), inv_h));
    if (!(q >= 0.0)) { clamped = true; return 0; }
    if (q >= (double)n) { clamped = true; return n - 1; }
    return (int)q;
}

// A particle whose global row lies outside [row0, row0 + rows_owned) has left the strip (multi-GPU latitude
// strips): with st.migrate it is packed into the send buffer of the neighbour it moved towards
// (record k at [1 + k], running count in [0].x) and dropped from the local state (key = -1).
__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ lon, const float *__restrict__ lat,
                                                        const int8_t *__restrict__ sp, const int32_t *__restrict__ id,
                                                        int first, int n, lm_grid g, Strip st,
                                                        int32_t *__restrict__ keys, int32_t *__restrict__ cell_count,
                                                        int4 *__restrict__ send_south, int4 *__restrict__ send_north,
                                                        int send_cap, Counters *ctr)
{
    const int p = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= first + n) return;
    bool clamped = false;
    const float x = lon[p], y = lat[p];
    const int cx = cell_coord(x, g.x0, g.inv_h, g.ncx, clamped);
    int cy = cell_coord(y, g.y0, g.inv_h, g.ncy, clamped) - st.row0;
    if (cy < 0 || cy >= st.rows_owned) {
        int4 *buf = (cy < 0) ? send_south : send_north;
        if (st.migrate && buf) {
            const unsigned int slot = atomicAdd(reinterpret_cast<unsigned int *>(&buf[0].x), 1u);
            if (slot < (unsigned int)send_cap) {
                buf[1 + slot] = make_int4(__float_as_int(x), __float_as_int(y), id ? id[p] : p, sp ? (int)sp[p] : 0);
                keys[p] = -1;
                return;
            }
            // this step's message is full: the particle stays one more step (the host clamps the count)
        }
        // Held by a strip that does not own its row: an arrival that crossed more than one strip in a step, a
        // leaver that did not fit the message, or a particle loaded on the wrong rank.  Kept in an edge row
        // and reported; the next binning pass sends it on (StripSet.settle loops until there are none).
        if (st.rows_owned < g.ncy) atomicAdd(&ctr->n_misrouted, 1u);
        cy = (cy < 0) ? 0 : st.rows_owned - 1;
    }
    const int key = cy * g.ncx + cx;
    keys[p] = key;
    atomicAdd(cell_count + key, 1);
    if (clamped) atomicAdd(&ctr->n_clamped, 1ull);
}

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a block of SCAN_THREADS; returns the block total in `total`
__device__ __forceinline__ int block_exclusive_scan(int v, int &total)
{
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0;
        s = warp_inclusive_scan(s);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int warp_off = (w == 0) ? 0 : warp_sums[w - 1];
    total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int32_t *__restrict__ cnt, int ncells,
                                                                      int32_t *__restrict__ tile_sums)
{
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
    if (base + SCAN_ITEMS <= ncells) {
        const int4 *p4 = reinterpret_cast<const int4 *>(cnt + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            const int4 v = p4[k];
            s += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < ncells) s += cnt[base + k];
    }
    int total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums in place
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(int32_t *__restrict__ tile_sums, int ntiles)
{
    int carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const int v = (i < ntiles) ? tile_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

// cell_start[i] = cell_cursor[i] = exclusive prefix; zeroes cell_count for the next step
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int32_t *__restrict__ cnt, int ncells,
                                                                  const int32_t *__restrict__ tile_offsets,
                                                                  int32_t *__restrict__ cell_start,
                                                                  int32_t *__restrict__ cell_cursor, int n_total)
{
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const bool full = base + SCAN_ITEMS <= ncells;          // the arrays are 16-byte aligned and base is a multiple of 16
    int v[SCAN_ITEMS];
    int s = 0;
    if (full) {
        const int4 *p4 = reinterpret_cast<const int4 *>(cnt + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            const int4 q = p4[k];
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < ncells) ? cnt[base + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
    int total;
    int run = block_exclusive_scan(s, total) + tile_offsets[blockIdx.x];
    if (full) {
        int4 *cs4 = reinterpret_cast<int4 *>(cell_start + base), *cc4 = reinterpret_cast<int4 *>(cell_cursor + base);
        int4 *z4 = reinterpret_cast<int4 *>(cnt + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            int4 o;
            o.x = run; run += v[4 * k];
            o.y = run; run += v[4 * k + 1];
            o.z = run; run += v[4 * k + 2];
            o.w = run; run += v[4 * k + 3];
            cs4[k] = o;
            cc4[k] = o;
            z4[k] = make_int4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < ncells) {
                cell_start[base + k] = run;
                cell_cursor[base + k] = run;
                cnt[base + k] = 0;
            }
            run += v[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cell_start[ncells] = n_total;
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const int32_t *__restrict__ keys,
                                                          const int32_t *__restrict__ id, int n,
                                                          int32_t *__restrict__ cell_cursor, int2 *__restrict__ slots)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int key = keys[p];
    if (key < 0) return;                     // left the strip
    const int slot = atomicAdd(cell_cursor + key, 1);
    slots[slot] = make_int2(p, id ? id[p] : p);
}

__global__ void __launch_bounds__(256) bin_reorder_kernel(const int2 *__restrict__ slots,
                                                          const int32_t *__restrict__ keys,
                                                          const int32_t *__restrict__ cell_start,
                                                          const float *__restrict__ lon, const float *__restrict__ lat,
                                                          const int8_t *__restrict__ sp, int n,
                                                          float *__restrict__ lon_o, float *__restrict__ lat_o,
                                                          int8_t *__restrict__ sp_o, int32_t *__restrict__ id_o)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int2 me = slots[s];
    const int key = keys[me.x];
    const int cs = cell_start[key], ce = cell_start[key + 1];
    int rank = 0;
    for (int t = cs; t < ce; ++t) rank += (slots[t].y < me.y);
    const int dst = cs + rank;
    lon_o[dst] = lon[me.x];
    lat_o[dst] = lat[me.x];
    if (sp_o) sp_o[dst] = sp ? sp[me.x] : (int8_t)0;
    id_o[dst] = me.y;
}

cudaError_t launch_bin_count(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id,
                             int first, int n, bool migrate, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    Strip st = h->strip;
    st.migrate = migrate ? 1 : 0;
    bin_count_kernel<<<(n + 255) / 256, 256, 0, s>>>(lon, lat, sp, id, first, n, h->grid, st, h->keys, h->cell_count,
                                                      h->has_south ? h->mig_send[0] : nullptr,
                                                      h->has_north ? h->mig_send[1] : nullptr, (int)h->send_cap, h->ctr);
    ++h->launches;
    return cudaGetLastError();
}

// scan + scatter + reorder: n_in source records (leavers have key -1), n_out of them stay
cudaError_t launch_bin_finish(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id,
                              int n_in, int n_out, float *lon_o, float *lat_o, int8_t *sp_o, int32_t *id_o, cudaStream_t s)
{
    const int ncells = h->grid.ncx * h->strip.rows_owned;
    const int ntiles = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    h->cs_idx ^= 1;                                  // the previous table may still be read by pending RPS phases
    h->cell_start = h->cell_start_buf[h->cs_idx];
    // cell_count is left zeroed by scan_apply (and by lm_create / lm_set_grid)
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(h->cell_count, ncells, h->block_sums);
    scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, s>>>(h->block_sums, ntiles);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(h->cell_count, ncells, h->block_sums, h->cell_start,
                                                       h->cell_cursor, n_out);
    h->launches += 3;
    if (n_in > 0) {
        bin_scatter_kernel<<<(n_in + 255) / 256, 256, 0, s>>>(h->keys, id, n_in, h->cell_cursor, h->slots);
        ++h->launches;
    }
    if (n_out > 0) {
        bin_reorder_kernel<<<(n_out + 255) / 256, 256, 0, s>>>(h->slots, h->keys, h->cell_start, lon, lat, sp, n_out,
                                                                lon_o, lat_o, sp_o, id_o);
        ++h->launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_bin(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                       float *lon_o, float *lat_o, int8_t *sp_o, int32_t *id_o, cudaStream_t s)
{
    cudaError_t e = launch_bin_count(h, lon, lat, sp, id, 0, n, false, s);
    if (e != cudaSuccess) return e;
    return launch_bin_finish(h, lon, lat, sp, id, n, n, lon_o, lat_o, sp_o, id_o, s);
}

// out[id[p]] = value[p]   (the reference's per-step record is in particle-id order:
// interaction_simulator.py:108-110, particle_advecter.py:233-235)
// The record leaves in particle-id order (the reference's [particle][time] columns), the state lives in cell order:
// one random 4-byte store per value.  A store that misses L2 costs a whole 32-byte sector twice (fill + write-back),
// so the ids are taken in ``passes`` windows: window w only stores ids in [w n / passes, (w + 1) n / passes), whose
// targets (n / passes values per array) stay resident in L2 until every byte of a sector has been written.  The
// sources are streamed (evict-first) so that they do not push the window out.
__global__ void __launch_bounds__(256) scatter_by_id_kernel(const float *__restrict__ lon, const float *__restrict__ lat,
                                                            const int8_t *__restrict__ sp,
                                                            const int32_t *__restrict__ id, int n, int id_lo, int id_hi,
                                                            float *__restrict__ lon_o, float *__restrict__ lat_o,
                                                            int8_t *__restrict__ sp_o)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int d = __ldcs(id + p);
    if (d < id_lo || d >= id_hi) return;
    if (lon_o) lon_o[d] = __ldcs(lon + p);
    if (lat_o) lat_o[d] = __ldcs(lat + p);
    if (sp_o) sp_o[d] = sp[p];
}

cudaError_t launch_scatter_by_id(const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                                 float *lon_o, float *lat_o, int8_t *sp_o, cudaStream_t s, int64_t *launches, int passes,
                                 int id_span)
{
    if (n <= 0) return cudaSuccess;
    const int block = 256;
    if (passes <= 1 || id_span <= 0) {
        scatter_by_id_kernel<<<(n + block - 1) / block, block, 0, s>>>(lon, lat, sp, id, n, INT_MIN, INT_MAX, lon_o, lat_o, sp_o);
        ++*launches;
        return cudaGetLastError();
    }
    for (int w = 0; w < passes; ++w) {
        const int lo = w == 0 ? INT_MIN : (int)((long long)id_span * w / passes);
        const int hi = w == passes - 1 ? INT_MAX : (int)((long long)id_span * (w + 1) / passes);
        scatter_by_id_kernel<<<(n + block - 1) / block, block, 0, s>>>(lon, lat, sp, id, n, lo, hi, lon_o, lat_o, sp_o);
        ++*launches;
    }
    return cudaGetLastError();
}

// order-preserving float <-> uint encoding for atomicMin/Max
__device__ __forceinline__ unsigned int enc_f(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// species histogram + bounding box (analysis.py:31-35 counts species per step on the host).
// Minima are stored as the maximum of the complemented encoding so that zeroed counters are the
// identity for all four bbox slots.
__global__ void __launch_bounds__(256) stats_kernel(const float *__restrict__ lon, const float *__restrict__ lat,
                                                    const int8_t *__restrict__ sp, int n, Counters *ctr)
{
    __shared__ unsigned int s_cnt[4];
    __shared__ unsigned int s_box[4];
    if (threadIdx.x < 4) {
        s_cnt[threadIdx.x] = 0;
        s_box[threadIdx.x] = 0;
    }
    __syncthreads();
    unsigned int c[4] = {0, 0, 0, 0};
    unsigned int b[4] = {0, 0, 0, 0};   // ~enc(min lon), enc(max lon), ~enc(min lat), enc(max lat)
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int s = sp ? sp[p] : 0;
        c[(s >= 1 && s <= 3) ? s : 0]++;
        const unsigned int ex = enc_f(lon[p]), ey = enc_f(lat[p]);
        b[0] = max(b[0], ~ex); b[1] = max(b[1], ex);
        b[2] = max(b[2], ~ey); b[3] = max(b[3], ey);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned int v = c[k], m = b[k];
        for (int d = 16; d > 0; d >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, d);
            m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        }
        if ((threadIdx.x & 31) == 0) {
            if (v) atomicAdd(&s_cnt[k], v);
            atomicMax(&s_box[k], m);
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        if (s_cnt[threadIdx.x]) atomicAdd(&ctr->species[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        atomicMax(&ctr->bbox_enc[threadIdx.x], s_box[threadIdx.x]);
    }
}

cudaError_t launch_stats(const float *lon, const float *lat, const int8_t *sp, int n, Counters *ctr, cudaStream_t s,
                         int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    const int block = 256;
    int grid = (n + block - 1) / block;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    stats_kernel<<<grid, block, 0, s>>>(lon, lat, sp, n, ctr);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace lm
