// TEST INFRASTRUCTURE ONLY -- C entry points that drive the emulated kernels of csrc/pairs.cu and csrc/analysis.cu
// on host arrays (tests/test_kernels_emulated.py).  The handle is assembled by hand: the launchers under test
// (launch_find, launch_resolve_phases -> the nine phase kernels or the tiled kernel, the analysis launchers) are the
// product's own code, compiled from the untouched sources.
#include <cstdio>

#include "lm_internal.cuh"
#include "philox.cuh"

using namespace lm;

template <class T> static T *zalloc(size_t count) { return static_cast<T *>(calloc(count ? count : 1, sizeof(T))); }

extern "C" {

// Pair search + RPS phases [first, last] on a binned state (arrays in storage order; cell_start has cells + 1
// entries for rows_local rows).  mode 0: nine phase launches; 1: tiled.  Returns the number of pairs found.
// species is updated in place; pairs_out (cap rows) may be null; hand-off copies are returned when the pointers
// are not null (hits: max_pairs words, rec: 5 * cells uint2, rec2: 5 * (n / 32 + 2) uint2).
long long emu_interact(const float *lon, const float *lat, const int32_t *id, const int32_t *cell_start, int8_t *species,
                       int n_owned, int n_all, double x0, double y0, double inv_h, int ncx, int ncy, int row0, int rows_owned,
                       int rows_local, double r, double pRS, double pPR, double pSP, unsigned long long seed,
                       unsigned long long step, int mode, int first, int last, int tile_smem, int heavy_min, int batch, int upl,
                       int find_path, int mega_min, int32_t *pairs_out, long long cap, long long max_pairs, uint32_t *hits_out,
                       uint32_t *rec_out, uint32_t *rec2_out)
{
    lm_handle_s *h = zalloc<lm_handle_s>(1);
    h->max_particles = n_all > 0 ? n_all : 1;
    h->max_cells = (int64_t)ncx * rows_local;
    h->max_pairs = max_pairs;
    h->grid.x0 = x0; h->grid.y0 = y0; h->grid.inv_h = inv_h; h->grid.ncx = ncx; h->grid.ncy = ncy;
    h->have_grid = true;
    h->strip.row0 = row0; h->strip.rows_owned = rows_owned; h->strip.rows_local = rows_local;
    h->has_north = rows_local > rows_owned; h->has_south = row0 > 0;
    h->ghost_cap = n_all - n_owned;
    h->n = n_owned;
    h->norm = LM_NORM_2;
    h->cell_start = const_cast<int32_t *>(cell_start);
    h->ctr = zalloc<Counters>(1);
    h->n_pairs_snap = zalloc<unsigned long long>(1);
    h->hits = zalloc<uint32_t>((size_t)max_pairs + 4);
    h->rec = zalloc<uint2>(5 * (size_t)h->max_cells);
    h->rec2 = zalloc<uint2>(5 * ((size_t)h->max_particles / 32 + 2));
    // stale records must not matter: fill the tables with garbage
    for (size_t k = 0; k < 5 * (size_t)h->max_cells; ++k) h->rec[k] = make_uint2(0x7fff0000u + (unsigned)k, 12345u);
    for (size_t k = 0; k < 5 * ((size_t)h->max_particles / 32 + 2); ++k) h->rec2[k] = make_uint2(0x7ffe0000u + (unsigned)k, 54321u);
    h->find_path = find_path;
    h->resolve_batch = batch; h->resolve_upl = upl; h->resolve_heavy_min = heavy_min;
    h->resolve_mode = mode; h->resolve_tile_smem = tile_smem; h->resolve_mega_min = mega_min;
    h->sp_snap = zalloc<int8_t>((size_t)h->max_particles);
    h->tile_scratch = zalloc<int8_t>(4 * (size_t)h->max_particles + (size_t)h->max_cells + 64);
    h->tile_scratch_used = zalloc<unsigned long long>(1);
    RpsDev rd;
    rd.pRS = pRS; rd.pPR = pPR; rd.pSP = pSP;
    rd.seed_lo = (uint32_t)seed; rd.seed_hi = (uint32_t)(seed >> 32);
    rd.step_lo = (uint32_t)step; rd.step_hi = (uint32_t)(step >> 32);
    rd.pair_key = pair_stream_key(seed, step);
    cudaError_t e = launch_find(h, lon, lat, id, n_owned, r, &rd, reinterpret_cast<int2 *>(pairs_out), pairs_out ? cap : 0, nullptr);
    if (e == cudaSuccess && n_owned > 0) e = launch_resolve_phases(h, species, first, last, nullptr);
    const long long found = (long long)h->ctr->n_pairs;
    if (hits_out) memcpy(hits_out, h->hits, (size_t)max_pairs * 4);
    if (rec_out) memcpy(rec_out, h->rec, 5 * (size_t)h->max_cells * 8);
    if (rec2_out) memcpy(rec2_out, h->rec2, 5 * ((size_t)h->max_particles / 32 + 2) * 8);
    const long long launches = h->launches;
    free(h->ctr); free(h->n_pairs_snap); free(h->hits); free(h->rec); free(h->rec2); free(h->sp_snap);
    free(h->tile_scratch); free(h->tile_scratch_used); free(h);
    return e == cudaSuccess ? found + (launches << 48) : -1;
}

// The fused tile kernel (csrc/interact.cu): phases [first, last] of the tile-round order (0..14) on a binned state.
// species may be null (pair search only); pairs_out may be null (count only).  Returns pairs found + launches << 48.
long long emu_interact_tile(const float *lon, const float *lat, const int32_t *id, const int32_t *cell_start, int8_t *species,
                            int n_owned, int ncx, int ncy, int row0, int rows_owned, int rows_local, double r, int norm,
                            double pRS, double pPR, double pSP, unsigned long long seed, unsigned long long step, int first,
                            int last, int tile_cap, int draw_batch, int tile_rec_cap, int tile_path, int32_t *pairs_out, long long cap)
{
    lm_handle_s *h = zalloc<lm_handle_s>(1);
    h->grid.ncx = ncx; h->grid.ncy = ncy;
    h->have_grid = true;
    h->strip.row0 = row0; h->strip.rows_owned = rows_owned; h->strip.rows_local = rows_local;
    h->n = n_owned;
    h->norm = norm;
    h->cell_start = const_cast<int32_t *>(cell_start);
    h->ctr = zalloc<Counters>(1);
    h->tile_cap = tile_cap; h->draw_batch = draw_batch; h->tile_rec_cap = tile_rec_cap; h->tile_path = tile_path;
    RpsDev rd;
    rd.pRS = pRS; rd.pPR = pPR; rd.pSP = pSP;
    rd.seed_lo = (uint32_t)seed; rd.seed_hi = (uint32_t)(seed >> 32);
    rd.step_lo = (uint32_t)step; rd.step_hi = (uint32_t)(step >> 32);
    rd.pair_key = pair_stream_key(seed, step);
    cudaError_t e = launch_interact(h, lon, lat, id, species, n_owned, r, species ? &rd : nullptr,
                                    reinterpret_cast<int2 *>(pairs_out), pairs_out ? cap : 0, first, last, nullptr);
    const long long found = (long long)h->ctr->n_pairs, launches = h->launches;
    free(h->ctr); free(h);
    return e == cudaSuccess ? found + (launches << 48) : -1;
}

int emu_pair_distance_hist(const float *lat, const float *lon, long long n, float radius_m, int bins, unsigned long long *hist)
{
    return launch_pair_distance_hist(lat, lon, n, radius_m, bins, hist, nullptr, nullptr);
}

int emu_raster(const float *lon, const float *lat, const int8_t *sp, long long n, double lon_min, double lon_max, double lat_min,
               double lat_max, int width, int height, uint32_t *counts, int32_t *top, int mode, const uint8_t *palette,
               uint8_t *rgb)
{
    cudaError_t e = launch_raster(lon, lat, sp, n, lon_min, lon_max, lat_min, lat_max, width, height, counts, top, nullptr, nullptr);
    if (e != cudaSuccess) return e;
    return launch_compose(counts, top, sp, width, height, mode, palette, rgb, nullptr, nullptr);
}

}  // extern "C"
