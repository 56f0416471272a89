// Internal declarations shared by the translation units of liblm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lm_b200.h"

namespace lm {

constexpr int kNumSMs = 148;   // B200

struct FieldDev {
    const float *U, *V, *lon, *lat;
    const float4 *UV4;                   // LM_OPT_ADVECT_MODE = 1: [T - 1][Y][X] (u_t, v_t, u_t+1, v_t+1) per grid point and time interval
    int T, Y, X;
    float lon0, lat0, lon1, lat1, inv_dx, inv_dy;   // axis ends; index guess: i = (x - lon0) * inv_dx
};

// device-side counters, reset at the start of every step / interact call
struct Counters {
    unsigned long long n_pairs;
    unsigned long long n_oob;
    unsigned long long n_clamped;
    unsigned long long species[4];
    unsigned int bbox_enc[4];   // order-preserving uint encodings: min lon, max lon, min lat, max lat
    unsigned long long n_overflow;   // neighbourhoods too dense for the 16-bit per-direction hit counts
    unsigned int n_leave[2];         // particles that left the strip this step: [0] southwards, [1] northwards
    unsigned int n_misrouted;        // arrivals that belong to neither this strip nor ... (moved > 1 strip in a step)
    unsigned int n_xfer_overflow;    // migration / ghost records that did not fit the exchange buffers
    unsigned int n_heavy_overflow;   // heavy units that did not fit the queue (hybrid mode): species invalid
};

// Strip geometry of one handle inside the GLOBAL cell grid (multi-GPU latitude strips).
// Single GPU: row0 = 0, rows_owned = rows_local = ncy, no neighbours.
struct Strip {
    int row0;         // first global cell row owned by this handle (even)
    int rows_owned;   // rows owned: particles of these rows live here
    int rows_local;   // rows_owned + 1 when a ghost row (first row of the strip to the north) is appended
    int migrate;      // bin_count: pack particles outside the strip into the send buffers (else clamp)
};

// layout of the ghost-row message (int32 words): header | cell_start row [row_cap + 1] | lon | lat | id [ghost_cap each]
constexpr int GHOST_HDR = 4;

struct RpsDev {
    double pRS, pPR, pSP;
    uint32_t seed_lo, seed_hi, step_lo, step_hi;
    uint32_t pair_key;     // philox.cuh::pair_stream_key(seed, step): key of this step's per-pair Philox2x32 stream
};

}  // namespace lm

struct lm_handle_s {
    int device;
    int64_t max_particles, max_cells, max_pairs;
    // velocity field (borrowed)
    lm::FieldDev field;
    bool have_field;
    float4 *uv4;           // the handle's interleaved copy of the field for the float32 RK4 (built lazily on the advecting stream)
    size_t uv4_elems;
    bool uv4_stale;
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
    int32_t *cell_start;   // [max_cells + 1] the current one of cell_start_buf[2] (double-buffered: the RPS phases of step k
                           //                 may still read theirs while step k+1 builds its own)
    int32_t *cell_start_buf[2];
    int cs_idx;
    int32_t *cell_cursor;  // [max_cells]
    int32_t *block_sums;   // [ceil(max_cells / SCAN_TILE) + 1]
    // counters
    lm::Counters *ctr;     // device
    bool ctr_reported;     // the counters' current contents have been returned by lm_sync_stats
    unsigned int *sticky;  // device: capacity faults latched before every counter reset (csrc/api.cu::latch_faults_kernel)
    int64_t emit_cap;      // capacity of the pair buffer passed to the last call (-1: none)
    int64_t rps_cap;       // capacity of hits[] if the last call resolved RPS (-1: it did not)
    // pair search -> resolver hand-off
    uint32_t *hits;        // [max_pairs + 4] hand-off entries (layout: csrc/pairs.cu)
    uint2 *rec;            // [5][max_cells] (first entry, count) of each cell's first segment, per direction
    uint2 *rec2;           // [5][max_particles / 32 + 2] the same for a cell continued at the start of a chunk
    // RPS phases of step k on a side stream, under the advection of step k+1 (single handle, no strips)
    cudaStream_t side_stream;
    cudaEvent_t ev_find_done, ev_resolve_done;
    cudaEvent_t ev_fork, ev_join;   // hybrid mode: the heavy units of a phase run on side_stream beside the phase's light units
    bool resolve_pending;  // ev_resolve_done has not been waited for yet
    bool resolve_on_side;  // this step's phases go to the side stream
    int overlap;           // LM_OPT_OVERLAP (default 1)
    unsigned long long *n_pairs_snap;   // device copy of ctr->n_pairs taken after the search (the resolver's overflow guard)
    int resolve_upl;       // LM_OPT_RESOLVE_UPL: units per lane in the resolver (0 = auto)
    int find_path;         // LM_OPT_FIND_PATH: 0 auto | 1 every warp takes the two-pass (dense cluster) path
    int resolve_heavy_min; // LM_OPT_RESOLVE_HEAVY_MIN: 0 = default (160)
    int resolve_batch;     // LM_OPT_RESOLVE_BATCH: pairs per lane and iteration in the resolver's stream walk (1, 4, 8)
    int interact_mode;     // LM_OPT_INTERACT_MODE: 2 (default) hybrid: round-1 pipeline for the light units + a device-wide queue of heavy
                           //                         units resolved in rounds of matchings (cell-round order) |
                           //                       1 fused tile kernel, tile-round order (csrc/interact.cu) |
                           //                       0 round-1 pipeline: pair search -> hand-off -> nine phase launches (csrc/pairs.cu)
    int2 *heavy_list;      // [9 phases][2: warp units | CTA units][heavy_cap] (anchor cell, other cell) queued by the pair search
    unsigned int *heavy_cnt;   // [9][4] queued warp units | queued CTA units | warp ticket | CTA ticket
    int64_t heavy_cap;
    int record_debug;      // LM_OPT_RECORD_DEBUG (measurement only)
    int scatter_passes;    // LM_OPT_SCATTER_PASSES: id windows of the record scatter (0 = auto, see scatter_windows)
    int64_t heavy_min;     // LM_OPT_HEAVY_MIN: candidate pairs above which a unit is heavy in the hybrid mode (0 = default, 1024: part of the
                           // definition of the cell-round order; other values are for A/B measurements)
    int draw_batch;        // LM_OPT_DRAW_BATCH: parked lanes that trigger a warp's Philox rounds (0 = default, 20)
    int tile_cap;          // LM_OPT_TILE_CAP: microbes a tile stages in shared memory (0 = from the mean occupancy)
    int tile_rec_cap;      // LM_OPT_TILE_REC_CAP: records (hits of one direction) a tile holds in shared memory (0 = 2 x tile_cap)
    int tile_path;         // LM_OPT_TILE_PATH: 0 records in shared memory where they fit (default) | 1 lane walk everywhere (tests)
    // arguments of the interaction in flight (fused tile kernel: the boundary phases 12-14 run in lm_step_interact_end)
    const float *ia_lon, *ia_lat;
    const int32_t *ia_id;
    int ia_n;
    double ia_r;
    lm::RpsDev ia_rps;
    bool ia_have_rps;
    int2 *ia_pairs;
    int64_t ia_cap;
    int advect_mode;       // LM_OPT_ADVECT_MODE: 0 bit-faithful to the float32 restatement of Parcels' kernel | 1 float32 arithmetic
    int norm;              // LM_OPT_NORM: LM_NORM_2 (default) | LM_NORM_1 | LM_NORM_INF
    // tiled resolver (LM_OPT_RESOLVE_MODE = 1; csrc/pairs.cu): allocated when the mode is first switched on
    int resolve_mode;      // 0: nine phase launches (default) | 1: one tiled launch per phase range
    int resolve_tile_smem; // LM_OPT_RESOLVE_TILE_SMEM: bytes of species a tile keeps in shared memory
    int resolve_tile_shape; // LM_OPT_RESOLVE_TILE_SHAPE: 0 = 64 x 16 cells (default), 1 = 32 x 16, 2 = 128 x 16, 3 = 64 x 32
    int resolve_mega_min;  // LM_OPT_RESOLVE_MEGA_MIN: pairs above which a unit goes to the whole CTA (0 = default, 8192)
    bool resolve_all_in_begin;   // this step's phases 6-8 already ran with 0-5 (tiled, single handle)
    int8_t *sp_snap;       // [max_particles] species before the first phase of a tiled launch
    int8_t *tile_scratch;  // [4 * max_particles + 16 * tiles] tiles that do not fit in shared memory
    unsigned long long *tile_scratch_used;
    // explicit-order resolver workspace
    unsigned long long *head;   // [max_particles]
    int32_t *pending[2];        // [max_pairs] each
    unsigned int *pending_cnt;  // [2]
    // host-copy pipeline
    float *stage_lon[2], *stage_lat[2];
    int8_t *stage_sp[2];
    cudaStream_t copy_stream;
    cudaEvent_t ev_scatter[2], ev_copied[2];
    cudaEvent_t ev_phase[6];   // LM_STEP_TIMING: step start | advect done | bin done | pair search done | RPS done | stats done
    bool timed;
    int stage_idx;
    // per-step record requested for the next step (lm_record_next_step): scattered + copied on copy_stream
    float *rec_lon_host, *rec_lat_host;
    int8_t *rec_sp_host;
    int32_t *rec_ids_host;           // lm_record_next_step_ids: the record in storage order, ids beside it
    bool rec_by_ids;
    int64_t rec_count;               // particles in the last record by ids
    bool rec_armed, rec_active;      // armed: next step records; active: this step is recording
    int rec_slot;
    cudaEvent_t ev_pos_ready, ev_pos_scattered, ev_sp_ready;
    cudaEvent_t ev_rec_reads[2];     // per record slot: the copy stream has read ids / species of the state buffers of that step
    int rec_reads_age[2];            // 2 when a record has been issued: the re-binning TWO steps on overwrites those buffers and waits
    bool pos_scatter_pending;        // the in-place advection of the next step must wait for ev_pos_scattered
    int64_t launches;
    // ---- latitude-strip decomposition (lm_strip_alloc / lm_set_strip); all zero for a single GPU
    lm::Strip strip;
    bool has_south, has_north;
    int64_t send_cap, ghost_cap, row_cap;
    int4 *mig_send[2], *mig_recv[2];     // [send_cap + 1] records (lon bits, lat bits, id, species); [0].x = count
    int32_t *ghost_send, *ghost_recv;    // GHOST_HDR + row_cap + 1 + 3 * ghost_cap words
    int8_t *gsp_send, *gsp_recv;         // [ghost_cap] species of the ghost row, north -> south after phase 5
    int8_t *gret_send, *gret_recv;       // [ghost_cap] species of the ghost row, south -> north after phase 8
    int32_t *xfer_counts_host;           // pinned + mapped: n_leave[2], n_arrive[2], stored by the device (xfer_counts_kernel)
    int32_t *xfer_counts_dev;            // the same memory as the device sees it
    long long peer_wait_cycles;          // LM_OPT_PEER_WAIT_CYCLES: SM clocks after which a wait for a neighbour's message gives up
    uint32_t *stats_host, *stats_dev;    // pinned + mapped: Counters + the sticky fault word, stored by the device (lm_sync_stats)
    // peer-memory exchange (lm_strip_peer_connect): the neighbours' receive buffers and flag words, mapped into this process
    struct Peer {
        bool connected, ipc;
        int4 *mig_recv;                  // the neighbour's mig_recv on the side that faces this strip
        int32_t *ghost_recv;
        int8_t *gsp_recv, *gret_recv;
        unsigned int *flags;
        void *base[LM_PEER_BUFFERS];     // what cudaIpcOpenMemHandle returned (to close)
    } peer[2];
    unsigned int *xflags;                // [8] mine: mig from S | mig from N | ghost | gsp | gret | mig ack from S | mig ack from N
    unsigned int xseq;                   // sequence number of the staged step in flight
    int n_moved_in, n_moved_out;         // last step (host copies)
    // staged step (lm_step_move .. lm_step_finish)
    int32_t step_flags;
    int stage;                           // 0 idle | 1 moved | 2 binned | 3 interact begun | 4 interact ended
    bool step_moved;
    double step_r;
    lm::RpsDev step_rps;
    int2 *step_pairs;
    int64_t step_cap;
};

namespace lm {

// ---- launchers (each returns cudaGetLastError() of its launches) --------------------------------
cudaError_t launch_advect(const FieldDev &f, float *lon, float *lat, int n, const lm_stage_times &st, float dt,
                          Counters *ctr, cudaStream_t s, int64_t *launches, int mode = 0);
cudaError_t launch_interleave_field(const FieldDev &f, float4 *uv4, cudaStream_t s, int64_t *launches);
cudaError_t launch_diffuse(float *lon, float *lat, const int32_t *ids, int n, double amp, uint64_t seed, uint64_t step,
                           cudaStream_t s, int64_t *launches);
// bins (lon,lat,sp,id)[src] into (cell,id) order in dst; sp / id may be null (id -> source index)
cudaError_t launch_bin(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                       float *lon_o, float *lat_o, int8_t *sp_o, int32_t *id_o, cudaStream_t s);
// the same in two halves, with migration between strips in the middle (csrc/bin.cu, csrc/strip.cu)
cudaError_t launch_bin_count(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id,
                             int first, int n, bool migrate, cudaStream_t s);
cudaError_t launch_bin_finish(lm_handle_s *h, const float *lon, const float *lat, const int8_t *sp, const int32_t *id,
                              int n_in, int n_out, float *lon_o, float *lat_o, int8_t *sp_o, int32_t *id_o,
                              cudaStream_t s);
cudaError_t launch_unpack_arrivals(lm_handle_s *h, int dir, int n_arrive, int first, float *lon, float *lat, int8_t *sp,
                                   int32_t *id, cudaStream_t s);
cudaError_t launch_ghost_pack(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, cudaStream_t s);
// peer-memory exchange (csrc/strip.cu)
cudaError_t launch_words_to_host(const uint32_t *src, const uint32_t *last, uint32_t *host_dev, int words, cudaStream_t s, int64_t *launches);
cudaError_t launch_xfer_counts(const void *send0, const void *send1, const void *recv0, const void *recv1, int32_t *host_counts_dev,
                               cudaStream_t s, int64_t *launches);
cudaError_t launch_peer_signal(unsigned int *flag, unsigned int seq, cudaStream_t s, int64_t *launches);
cudaError_t launch_peer_wait(const unsigned int *flag, unsigned int seq, unsigned int *sticky, long long limit_cycles, cudaStream_t s,
                             int64_t *launches);
cudaError_t launch_peer_push_mig(const int4 *src, int4 *dst, int64_t send_cap, cudaStream_t s, int64_t *launches);
cudaError_t launch_ghost_unpack(lm_handle_s *h, float *lon, float *lat, int32_t *id, int n_owned, cudaStream_t s);
// species of the first owned row -> gsp_send (pack) / gsp_recv -> ghost particles (unpack); and the way back
cudaError_t launch_row0_species_pack(lm_handle_s *h, const int8_t *sp, cudaStream_t s);
cudaError_t launch_ghost_species_unpack(lm_handle_s *h, int8_t *sp, int n_owned, cudaStream_t s);
cudaError_t launch_ghost_species_pack(lm_handle_s *h, const int8_t *sp, int n_owned, cudaStream_t s);
cudaError_t launch_row0_species_unpack(lm_handle_s *h, int8_t *sp, cudaStream_t s);
cudaError_t launch_scatter_by_id(const float *lon, const float *lat, const int8_t *sp, const int32_t *id, int n,
                                 float *lon_o, float *lat_o, int8_t *sp_o, cudaStream_t s, int64_t *launches, int passes = 1,
                                 int id_span = 0);
cudaError_t launch_stats(const float *lon, const float *lat, const int8_t *sp, int n, Counters *ctr, cudaStream_t s,
                         int64_t *launches);
// pair search (+ optional fused RPS in canonical cell-phase order) on binned arrays
cudaError_t launch_pairs(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                         double r, const RpsDev *rps /* null = find only */, int2 *pairs_out, int64_t cap,
                         cudaStream_t s);
// the two halves: find (+ hand-off lists when rps != null), then resolve phases [first, last] of 0..8
cudaError_t launch_find(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int n, double r,
                        const RpsDev *rps, int2 *pairs_out, int64_t cap, cudaStream_t s);
cudaError_t launch_resolve_phases(lm_handle_s *h, int8_t *sp, int first, int last, cudaStream_t s);
// fused tile kernel (csrc/interact.cu): phases [first, last] of the tile-round order, 0..14 (0..8 are one launch)
cudaError_t launch_interact(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, int8_t *sp, int n,
                            double r, const RpsDev *rps, int2 *pairs_out, int64_t cap, int first, int last,
                            cudaStream_t s);
cudaError_t launch_interact_heavy(lm_handle_s *h, int8_t *sp, int phase, cudaStream_t s);
cudaError_t launch_pair_uniforms(const int2 *pairs, int64_t np, uint64_t seed, uint64_t step, double *u,
                                 cudaStream_t s);
int resolve_explicit(lm_handle_s *h, const int2 *pairs, const double *u, int64_t np, int8_t *species, int64_t n,
                     double pRS, double pPR, double pSP, int32_t *rounds_out, cudaStream_t s);

// analysis reductions on a snapshot (csrc/analysis.cu)
cudaError_t launch_pair_distance_hist(const float *lat, const float *lon, int64_t n, float radius_m, int bins,
                                      unsigned long long *hist, cudaStream_t s, int64_t *launches);
cudaError_t launch_raster(const float *lon, const float *lat, const int8_t *sp, int64_t n, double lon_min, double lon_max,
                          double lat_min, double lat_max, int width, int height, uint32_t *counts, int32_t *top,
                          cudaStream_t s, int64_t *launches);
cudaError_t launch_compose(const uint32_t *counts, const int32_t *top, const int8_t *sp, int width, int height, int mode,
                           const uint8_t *palette_rgb, uint8_t *rgb, cudaStream_t s, int64_t *launches);

// lossless delta packing of the position record (csrc/record.cu)
cudaError_t launch_record_delta_pack(const float *prev_lon, const float *prev_lat, const float *lon, const float *lat, int64_t n,
                                     int16_t *dlon, int16_t *dlat, uint32_t *esc, int64_t esc_cap, uint32_t *esc_count,
                                     cudaStream_t s);

int64_t record_delta_unpack_host(const float *prev_lon, const float *prev_lat, const int16_t *dlon, const int16_t *dlat,
                                 const uint32_t *esc, int64_t n_esc, int64_t n, float *lon_out, float *lat_out, int n_threads);

void set_last_cuda_error(cudaError_t e, const char *where);

}  // namespace lm
