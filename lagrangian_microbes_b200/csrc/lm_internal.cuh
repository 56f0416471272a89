// Internal declarations shared by the translation units of liblm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lm_b200.h"

namespace lm {

constexpr int kNumSMs = 148;   // B200

struct FieldDev {
    const float *U, *V, *lon, *lat;
    int T, Y, X;
    float lon0, lat0, inv_dx, inv_dy;   // index guess: i = (x - lon0) * inv_dx
};

// device-side counters, reset at the start of every step / interact call
struct Counters {
    unsigned long long n_pairs;
    unsigned long long n_oob;
    unsigned long long n_clamped;
    unsigned long long species[4];
    unsigned int bbox_enc[4];   // order-preserving uint encodings: min lon, max lon, min lat, max lat
    unsigned long long n_overflow;   // neighbourhoods too dense for the 16-bit per-direction hit counts
};

struct RpsDev {
    double pRS, pPR, pSP;
    uint32_t seed_lo, seed_hi, step_lo, step_hi;
};

}  // namespace lm

struct lm_handle_s {
    int device;
    int64_t max_particles, max_cells, max_pairs;
    // velocity field (borrowed)
    lm::FieldDev field;
    bool have_field;
    // cell grid
    lm_grid grid;
    bool have_grid;
    // resident particle state, ping-pong, kept in (cell, id) order
    float *lon[2], *lat[2];
    int8_t *sp[2];
    int32_t *id[2];
    int cur;
    int64_t n;
    bool binned;           // state[cur] is binned with the current grid
    // binning workspace
    int32_t *keys;         // [max_particles] cell key of each particle (input order)
    int2 *slots;           // [max_particles] (source index, id) in arbitrary within-cell order
    int32_t *cell_count;   // [max_cells]
    int32_t *cell_start;   // [max_cells + 1]
    int32_t *cell_cursor;  // [max_cells]
    int32_t *block_sums;   // [ceil(max_cells / SCAN_TILE) + 1]
    // counters
    lm::Counters *ctr;     // device
    int64_t emit_cap;      // capacity of the pair buffer passed to the last call (-1: none)
    int64_t rps_cap;       // capacity of hits[] if the last call resolved RPS (-1: it did not)
    // pair search -> resolver hand-off
    uint32_t *hits;        // [max_pairs] partner index | decision bits << 29, grouped per particle
    int4 *meta;            // [max_particles] (offset into hits, 5 x 16-bit hit counts per direction)
    // explicit-order resolver workspace
    unsigned long long *head;   // [max_particles]
    int32_t *pending[2];        // [max_pairs] each
    unsigned int *pending_cnt;  // [2]
    // host-copy pipeline
    float *stage_lon[2], *stage_lat[2];
    int8_t *stage_sp[2];
    cudaStream_t copy_stream;
    cudaEvent_t ev_scatter[2], ev_copied[2];
    cudaEvent_t ev_phase[5];   // LM_STEP_TIMING: step start | advect done | bin done | pairs done | stats done
    bool timed;
    int stage_idx;
    int64_t launches;
};

namespace lm {

// ---- launchers (each returns cudaGetLastError() of its launches) --------------------------------
cudaError_t launch_advect(const FieldDev &f, float *lon, float *lat, int n, const lm_stage_times &st, float dt,
                          Counters *ctr, cudaStream_t s, int64_t *launches);
cudaError_t launch_diffuse(float *lon, float *lat, const int32_t *ids, int n, double amp, uint64_t seed, uint64_t step,
                           cudaStream_t s, int64_t *launches);
// bins (lon,lat,sp,id)[src] into (cell,id) order in dst; sp / id may be null (id -> source index)
cudaError_t launch_bin(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                       float *lon_o, float *lat_o, int8_t *sp_o, int32_t *id_o, cudaStream_t s);
cudaError_t launch_scatter_by_id(const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                                 float *lon_o, float *lat_o, int8_t *sp_o, cudaStream_t s, int64_t *launches);
cudaError_t launch_stats(const float *lon, const float *lat, const int8_t *sp, int n, Counters *ctr, cudaStream_t s,
                         int64_t *launches);
// pair search (+ optional fused RPS in canonical cell-phase order) on binned arrays
cudaError_t launch_pairs(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                         double r, const RpsDev *rps /* null = find only */, int2 *pairs_out, int64_t cap,
                         cudaStream_t s);
cudaError_t launch_pair_uniforms(const int2 *pairs, int64_t np, uint64_t seed, uint64_t step, double *u,
                                 cudaStream_t s);
int resolve_explicit(lm_handle_s *h, const int2 *pairs, const double *u, int64_t np, int8_t *species, int64_t n,
                     double pRS, double pPR, double pSP, int32_t *rounds_out, cudaStream_t s);

void set_last_cuda_error(cudaError_t e, const char *where);

}  // namespace lm
