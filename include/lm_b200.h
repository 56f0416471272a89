/* lm_b200.h -- C ABI of the B200-native lagrangian-microbes hot path.
 *
 * The reference (ali-ramadhan/lagrangian-microbes) is pure Python with no FFI of its own; its
 * hot path is three call sites into third-party native code plus one Python loop:
 *
 *   (A3) pset.execute(parcels.AdvectionRK4, runtime=dt, dt=dt, ...)     particle_advecter.py:222-223
 *   (A5) the random-walk diffusion kick                                  particle_advecter.py:240-242
 *   (P1/P2) cKDTree(locations).query_pairs(r, p)                         interaction_simulator.py:93,98
 *   (R1/R2) for pair in pairs: rock_paper_scissors_interaction(...)      interaction_simulator.py:104-105,
 *                                                                        interactions.py:13-40
 *
 * Each entry point below names the call site it replaces.  A maintainer of the reference would
 * bind these with ctypes (see INTEGRATION.md for the stubs).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every function returns an int status: LM_OK (0) or a negative LM_E* code; nothing throws.
 *   - Unless a parameter says "host", pointers are DEVICE pointers owned by the caller
 *     (e.g. torch tensors' data_ptr()); all work is enqueued on the given stream (a cudaStream_t
 *     passed as void*; NULL = default stream) and is asynchronous unless stated otherwise.
 *   - No function allocates device memory per call: workspaces are sized at lm_create.
 *   - A handle is bound to one device and may be used from one host thread at a time.
 *   - Particle ids are int32 (N < 2^31 per handle); species are int8 with ROCK=1, PAPER=2,
 *     SCISSORS=3 (interactions.py:5).
 */
#ifndef LM_B200_H
#define LM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_OK 0
#define LM_EINVAL (-1)    /* bad argument */
#define LM_ENOMEM (-2)    /* allocation failed at lm_create */
#define LM_ECUDA (-3)     /* CUDA runtime error (lm_last_cuda_error has the text) */
#define LM_ENOSPC (-4)    /* capacity exceeded (particles / cells / pairs) */
#define LM_ESTATE (-5)    /* call order violated (no field / grid / state yet) */
#define LM_ENOCONV (-6)   /* explicit-order resolver exceeded its round limit */

typedef struct lm_handle_s *lm_handle;

/* Per-stage time decisions of one RK4 step (stages sample at t, t+dt/2, t+dt/2, t+dt).  All
 * particles share one clock, so Parcels' cached time index / interpolate-or-hold decision /
 * float time fraction are the same for every particle; the host computes them once per step
 * (host mirror: lagrangian_microbes_b200/particle_advecter.py::StageClock). */
typedef struct {
    int32_t ti[4];       /* lower time level of each stage */
    int32_t interp[4];   /* 1: f0 + (f1 - f0) * frac;  0: hold f0 */
    float frac[4];       /* (float)((t - t0) / (t1 - t0)) */
} lm_stage_times;

/* Uniform cell grid used for binning: cell = clamp(floor((double(v) - origin) * inv_h), 0, n-1).
 * 1/inv_h must exceed the interaction radius. */
typedef struct {
    double x0, y0, inv_h;
    int32_t ncx, ncy;
} lm_grid;

/* Rock-paper-scissors parameters (interactions.py:47-51) + the per-pair random stream key. */
typedef struct {
    double pRS, pPR, pSP;
    uint64_t seed, step;   /* u(i,j) = Philox4x32-10(counter=(i,j,step), key=seed) -> 53-bit double */
} lm_rps_params;

/* Counters of the last lm_step / lm_interact (host-readable after lm_sync_stats). */
typedef struct {
    int64_t n_pairs;          /* pairs found (may exceed the emit capacity) */
    int64_t n_out_of_bounds;  /* particles that left the velocity grid (left unchanged) */
    int64_t n_clamped;        /* particles outside the cell grid (binned into edge cells) */
    int64_t species_count[4]; /* [0]=other, [1]=rock, [2]=paper, [3]=scissors, after the step */
    float bbox[4];            /* lon_min, lon_max, lat_min, lat_max of the particles */
    int64_t n_particles;      /* particles owned by this handle after the step */
    int64_t n_moved_in;       /* strips: particles that arrived from / left to the neighbour strips in the */
    int64_t n_moved_out;      /*         last step that re-binned */
    int64_t n_misrouted;      /* strips: arrivals belonging to neither strip (moved more than one strip per step) */
} lm_stats;

/* Latitude strip of one handle inside the GLOBAL cell grid (multi-GPU): the handle owns
 * the particles of cell rows [row0, row0 + rows_owned).  row0 must be even; has_south / has_north say whether
 * a neighbouring strip exists (they must be consistent with row0 == 0 / row0 + rows_owned == ncy). */
typedef struct {
    int32_t row0, rows_owned;
    int32_t has_south, has_north;
} lm_strip;

/* Exchange buffers of a strip (device memory owned by the handle; sizes are fixed so the transfers can be
 * posted without negotiating counts -- the live counts travel in the messages' headers).  The caller moves,
 * between the stages of a step (see lm_step_move):
 *     mig_send[0]  -> the southern neighbour's mig_recv[1]        mig_send[1] -> the northern one's mig_recv[0]
 *     ghost_send   -> the southern neighbour's ghost_recv
 *     gsp_send     -> the southern neighbour's gsp_recv
 *     gret_send    -> the northern neighbour's gret_recv
 * i.e. index 0 = south side, 1 = north side of THIS strip, for send and recv alike. */
typedef struct {
    void *mig_send[2], *mig_recv[2];
    int64_t mig_bytes;
    void *ghost_send, *ghost_recv;
    int64_t ghost_bytes;
    void *gsp_send, *gsp_recv, *gret_send, *gret_recv;
    int64_t species_bytes;
} lm_strip_buffers;

/* Peer-memory exchange (NVLink / NVSwitch peer stores instead of NCCL send/recv).  A strip exports the addresses of its
 * five RECEIVE buffers and of its flag words -- as CUDA IPC handles for a neighbour in another process, or as plain device
 * pointers for a neighbour in the same process -- and connects to its neighbours' exports.  Once connected, the kernels that
 * pack the ghost row and the boundary species write STRAIGHT into the neighbour's receive buffer, the migrants are copied
 * there by one kernel that moves the live records only, each message is followed by a flag store (sequence number) into the
 * neighbour's flag words, and the stage that consumes a message first spins on its flag: stream-ordered on both sides, no
 * collective library, no host involvement.  lm_step_push(kind) replaces the caller's transfer of `kind` (LM_XCHG_*). */
#define LM_IPC_HANDLE_BYTES 64
#define LM_PEER_BUFFERS 6 /* mig_recv[0], mig_recv[1], ghost_recv, gsp_recv, gret_recv, flags */
typedef struct {
    unsigned char ipc[LM_PEER_BUFFERS][LM_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of each buffer */
    void *ptr[LM_PEER_BUFFERS];                                /* the same buffers as device pointers (same process only) */
} lm_peer_export;
#define LM_XCHG_MIG 0
#define LM_XCHG_GHOST 1
#define LM_XCHG_GSP 2
#define LM_XCHG_GRET 3

int lm_version(void);
const char *lm_error_string(int code);
const char *lm_last_cuda_error(void);

/* ---- lifetime ------------------------------------------------------------------------------ */
/* max_particles < 2^31 (< 2^29 when max_pairs > 0), max_cells < 2^31; max_pairs (< 2^32) sizes the pair-search ->
 * RPS hand-off (4 B per pair + 40 B per cell of record tables) and may be 0 for a handle that only advects or
 * only searches.  All device memory of the handle is allocated here. */
int lm_create(lm_handle *out, int device, int64_t max_particles, int64_t max_cells, int64_t max_pairs);
int lm_destroy(lm_handle h);

/* ---- inputs -------------------------------------------------------------------------------- */
/* Replaces Grid/Field/FieldSet construction (particle_advecter.py:177-184).  Borrowed device
 * pointers: U, V float32 [T][Y][X] (NaN already zeroed), lon float32[X] ascending, lat float32[Y]
 * ascending.  They must stay valid while the handle uses them. */
int lm_set_field(lm_handle h, const float *U, const float *V, const float *lon, const float *lat,
                 int32_t T, int32_t Y, int32_t X);
/* Swap the U/V data (same lon/lat axes) for another set of T time levels -- for callers that stream
 * a window of snapshots through a small device buffer.  No synchronisation. */
int lm_update_field_data(lm_handle h, const float *U, const float *V, int32_t T);
int lm_set_grid(lm_handle h, const lm_grid *grid);
int lm_get_grid(lm_handle h, lm_grid *grid_out /* host */);

/* ---- stateless operators on caller-owned arrays in particle-id order ------------------------ */
/* (A3) One AdvectionRK4 step in place on lon/lat float32[n].  Particles that leave the velocity
 * grid are left unchanged and counted in lm_stats.n_out_of_bounds, which accumulates over calls
 * until lm_reset_stats (Parcels raises OutOfBoundsError; the host mirror raises after the sync). */
int lm_advect_rk4(lm_handle h, float *lon, float *lat, int64_t n, const lm_stage_times *st /* host */,
                  float dt, void *stream);
/* (A5) lat += U(-1,1)*amp, then lon += U(-1,1)*amp; Philox keyed (seed, step, id = array index). */
int lm_diffuse(lm_handle h, float *lon, float *lat, int64_t n, double amp_deg, uint64_t seed, uint64_t step,
               void *stream);
/* The same with explicit particle ids (device int32[n]; NULL = array index): a process that holds only some of the
 * reference's N_procs tiles (particle_advecter.py:143-148) kicks its particles exactly as one process holding all would. */
int lm_diffuse_ids(lm_handle h, float *lon, float *lat, const int32_t *ids, int64_t n, double amp_deg, uint64_t seed,
                   uint64_t step, void *stream);
/* (P1/P2) All pairs (i<j, array indices) with dx*dx + dy*dy <= r*r evaluated exactly as SciPy does
 * (float32 positions widened to double; other norms: LM_OPT_NORM).  pairs_out int32[cap][2] in unspecified order;
 * *n_pairs_out (device int64) receives the number found; LM_ENOSPC is reported by
 * lm_sync_stats when it exceeded cap (the list is then incomplete: size it from the reported count and repeat). */
int lm_find_pairs(lm_handle h, const float *lon, const float *lat, int64_t n, double r,
                  int32_t *pairs_out, int64_t cap, int64_t *n_pairs_out, void *stream);
/* (P1/P2 + R1/R2) Pair search fused with RPS resolution in the canonical cell-phase order
 * (DESIGN.md §4.3).  species int8[n] updated in place; pairs_out may be NULL (cap = 0). */
int lm_interact_rps(lm_handle h, const float *lon, const float *lat, int8_t *species, int64_t n, double r,
                    const lm_rps_params *prm /* host */, int32_t *pairs_out, int64_t cap,
                    int64_t *n_pairs_out, void *stream);
/* Per-pair uniforms of the stream above for an explicit pair list (rows i<j). */
int lm_pair_uniforms(const int32_t *pairs, int64_t n_pairs, uint64_t seed, uint64_t step, double *u_out,
                     void *stream);
/* (R1/R2) The reference's sequential in-place loop for an EXPLICIT pair order: pair k = pairs[k],
 * random draw u[k] consumed iff the species differ when pair k is reached.  Resolved in
 * conflict-free rounds (a pair fires when it is the lowest-rank pending pair at both of its
 * particles).  Synchronous; *rounds_out (host) receives the number of rounds. */
int lm_resolve_rps(lm_handle h, const int32_t *pairs, const double *u, int64_t n_pairs, int8_t *species,
                   int64_t n, double pRS, double pPR, double pSP, int32_t *rounds_out /* host */, void *stream);

/* ---- resident pipeline (state owned by the handle, kept in (cell, id) order) ------------------ */
/* Upload n particles (arrays in id order; ids NULL = 0..n-1; device pointers) and bin them.  On a strip
 * (lm_set_strip) the particles are only loaded -- ids are then required and global -- and the first step bins
 * them and sends those that belong to a neighbouring strip on their way. */
int lm_state_set(lm_handle h, const float *lon, const float *lat, const int8_t *species, const int32_t *ids,
                 int64_t n, void *stream);
int64_t lm_state_size(lm_handle h);
/* One fused step k = prm->step:  [diffuse: the kick the reference applies at the end of iteration
 * k-1, keyed (seed, k-1)] -> [advect] -> bin -> pair search + RPS keyed (seed, k) (-> emit pairs).
 * flags: LM_STEP_* below.  st may be NULL when LM_STEP_ADVECT is not set.  Asynchronous; with LM_OPT_OVERLAP
 * (default) the RPS phases run on an internal stream and later calls order themselves after them. */
#define LM_STEP_ADVECT 1
#define LM_STEP_DIFFUSE 2
#define LM_STEP_INTERACT 4
#define LM_STEP_EMIT_PAIRS 8
#define LM_STEP_STATS 16
#define LM_STEP_TIMING 32 /* record CUDA events between the phases (read with lm_phase_times) */
int lm_step(lm_handle h, int32_t flags, const lm_stage_times *st /* host */, float dt, double diffuse_amp_deg,
            double r, const lm_rps_params *prm /* host */, int32_t *pairs_out, int64_t cap, void *stream);
/* ---- the same step in five stages, for latitude strips (one handle per GPU) ----------------------------
 * The reference has no counterpart: its interaction phase is one serial process
 * (interaction_simulator.py:82-117) and only advection is tiled (particle_advecter.py:143-148).
 * lm_step == move, bin, interact_begin, interact_end, finish back to back.  With strips the caller moves the
 * exchange buffers between the stages (NCCL send/recv between neighbours, or device copies):
 *     lm_step_move            [diffuse] [advect], cell keys, leavers packed      -> exchange mig_send
 *     lm_step_bin             arrivals unpacked, binning, first row packed         -> exchange ghost_send
 *                             (synchronises the stream once to read the migration counts)
 *     lm_step_interact_begin  ghost row appended, pair search, RPS phases 0-5      -> exchange gsp_send
 *     lm_step_interact_end    ghost species refreshed, RPS phases 6-8             -> exchange gret_send
 *     lm_step_finish          first row's species taken back, [stats]
 * The result (pair set, species, positions) is bit-identical to a single handle running lm_step on all
 * particles with the same grid. */
int lm_strip_alloc(lm_handle h, int64_t send_cap, int64_t ghost_cap, int32_t row_cap);
int lm_set_strip(lm_handle h, const lm_strip *strip /* host */);   /* after lm_set_grid (which resets it) */
int lm_strip_buffers_get(lm_handle h, lm_strip_buffers *out /* host */);
/* peer-memory exchange: see lm_peer_export.  side: 0 = the southern neighbour, 1 = the northern one; use_ipc: open the IPC
 * handles (neighbour in another process) or take the pointers (same process).  Connect resets the sequence numbers: every
 * strip of the set connects before the first step.  lm_step_push: after the stage that produced `kind` (MIG: lm_step_move;
 * GHOST: lm_step_bin; GSP: lm_step_interact_begin; GRET: lm_step_interact_end). */
int lm_strip_peer_export(lm_handle h, lm_peer_export *out /* host */);
int lm_strip_peer_connect(lm_handle h, int32_t side, const lm_peer_export *peer /* host */, int32_t use_ipc);
int lm_step_push(lm_handle h, int32_t kind, void *stream);
int lm_step_move(lm_handle h, int32_t flags, const lm_stage_times *st /* host */, float dt, double diffuse_amp_deg,
                 const lm_rps_params *prm /* host */, void *stream);
int lm_step_bin(lm_handle h, void *stream);
int lm_step_interact_begin(lm_handle h, double r, int32_t *pairs_out, int64_t cap, void *stream);
int lm_step_interact_end(lm_handle h, void *stream);
int lm_step_finish(lm_handle h, void *stream);
/* Scatter the state back to id order: out[id] = value (device outputs, any may be NULL). */
int lm_state_get(lm_handle h, float *lon_out, float *lat_out, int8_t *species_out, void *stream);
/* Same, into pinned HOST buffers (device scatter + async D2H on the stream). */
int lm_state_get_host(lm_handle h, float *lon_host, float *lat_host, int8_t *species_host, void *stream);
/* The per-step record without holding up the step: the NEXT lm_step also writes, into pinned HOST buffers (any may
 * be NULL) in particle-id order, the positions after its advection (sent while the pair search runs) and the species
 * after its interactions (sent while the next step advects) -- what the reference stores per iteration
 * (particle_advecter.py:233-235, interaction_simulator.py:108-110).  Complete after lm_host_copies_sync; alternate
 * two sets of buffers.  Single handle only (strips: lm_state_view + the ids). */
int lm_record_next_step(lm_handle h, float *lon_host, float *lat_host, int8_t *species_host);
/* The same for a handle that holds ANY subset of the particles (a latitude strip): the record of the next step in STORAGE
 * order with the particle ids beside it -- ids_host int32, lon_host / lat_host float32 (optional), species_host int8
 * (optional), pinned, each with room for the handle's max_particles.  Positions and ids are copied under the pair search,
 * species after the RPS phases, all on the library's copy stream (lm_host_copies_sync); lm_record_count = the number of
 * particles of the last such record (known when lm_step_bin has returned).  What a per-strip output file holds, like the
 * reference's per-tile chunk pickles (particle_advecter.py:201-214). */
int lm_record_next_step_ids(lm_handle h, int32_t *ids_host, float *lon_host, float *lat_host, int8_t *species_host);
int64_t lm_record_count(lm_handle h);
/* Wait for every D2H copy issued by lm_state_get_host (they run on an internal copy stream so
 * that the next step's kernels overlap them; up to two may be in flight). */
int lm_host_copies_sync(lm_handle h);
/* Raw view of the resident arrays in storage order (for halo exchange / tests). */
int lm_state_view(lm_handle h, float **lon, float **lat, int8_t **species, int32_t **ids, int32_t **cell_start);

/* ---- status -------------------------------------------------------------------------------- */
/* Synchronise the stream and copy the device counters; returns LM_ENOSPC if pairs overflowed the
 * emit capacity (or another capacity: hand-off, exchange buffers), LM_ESTATE for misrouted particles, LM_OK otherwise.
 * Faults are STICKY: a step latches the faults of the previous step before it zeroes the counters, so a fault in any
 * step since the last lm_sync_stats / lm_reset_stats is reported here (once), whether or not that step asked for
 * LM_STEP_STATS. */
int lm_sync_stats(lm_handle h, lm_stats *out /* host */, void *stream);
/* Zero the device counters (lm_step and the pair-search operators do this themselves). */
int lm_reset_stats(lm_handle h, void *stream);
/* Tuning knobs (defaults are right for production; tests use them to force a code path).
 *   LM_OPT_FIND_PATH  0 = auto; 1 = every warp of the pair search takes the two-pass path that is normally
 *                   reserved for dense clusters (more than 64 candidate partners for one microbe). */
#define LM_OPT_FIND_PATH 2
/*   LM_OPT_RESOLVE_UPL  units (cell x direction) per lane in the RPS resolver: 0 = auto (as many as keeps the GPU
 *                   full of warps), or 1, 2, 4, 8. */
#define LM_OPT_RESOLVE_UPL 3
/*   LM_OPT_OVERLAP    1 (default): on a single handle the RPS phases of a step run on an internal stream and overlap
 *                   the advection of the next lm_step; every entry point that reads species or the state orders
 *                   itself after them, lm_join does so explicitly.  0: everything on the caller's stream. */
#define LM_OPT_OVERLAP 4
/*   LM_OPT_NORM       the Minkowski norm of the radius query, query_pairs(r, p=interaction_norm)
 *                   (interaction_simulator.py:27,98): LM_NORM_2 (default; the only one the reference's scripts use),
 *                   LM_NORM_1 or LM_NORM_INF, each evaluated exactly as SciPy does for that p:
 *                       p=2    fl(fl(dx*dx) + fl(dy*dy)) <= fl(r*r)
 *                       p=1    fl(|dx| + |dy|) <= r
 *                       p=inf  max(|dx|, |dy|) <= r
 *                   (dx, dy: differences of the float32 coordinates widened to double).  Other p need pow() and are
 *                   not offered.  Applies to lm_find_pairs, lm_interact_rps, lm_step and the staged step. */
#define LM_OPT_NORM 5
#define LM_NORM_INF 0
#define LM_NORM_1 1
#define LM_NORM_2 2
/*   LM_OPT_RESOLVE_HEAVY_MIN  RPS resolver: a (cell x direction) unit with fewer pairs than this is always walked by one
 *                   lane, never handed to a whole warp (0 = default, 160). */
#define LM_OPT_RESOLVE_HEAVY_MIN 6
/*   LM_OPT_RESOLVE_BATCH  RPS resolver: pairs a lane loads ahead per iteration of its stream walk: 1, 4 (default)
 *                   or 8.  Results are identical for every value; only the number of dependent memory round trips per
 *                   unit changes. */
#define LM_OPT_RESOLVE_BATCH 7
/*   LM_OPT_RESOLVE_MODE  RPS resolver: 0 (default) nine phase launches on the live species; 1 (EXPERIMENTAL, not yet
 *                   measured) one launch per phase range over tiles of 64 x 16 cells: each CTA copies the species of
 *                   its tile plus a halo of 6 columns / 2 rows from a snapshot into shared memory, runs every phase
 *                   there and writes back the tile's interior (DESIGN.md 4.3 "Tiled resolver").  Same results.
 *                   Switching it on allocates 5 bytes per particle + 1 per cell (the only allocation outside
 *                   lm_create); not allowed between the stages of a step (LM_ESTATE).
 *   LM_OPT_RESOLVE_TILE_SMEM  bytes of species a tile may keep in shared memory (default 32768); fuller tiles work on a
 *                   private slice of global memory instead.
 *   LM_OPT_RESOLVE_MEGA_MIN  tiled resolver: a unit with more pairs than this (0 = default, 8192: about 128 microbes in
 *                   one cell) is resolved by the whole CTA, 256 partners of an anchor per scan, instead of by one warp. */
#define LM_OPT_RESOLVE_MODE 8
#define LM_OPT_RESOLVE_TILE_SMEM 9
#define LM_OPT_RESOLVE_MEGA_MIN 10
/*   LM_OPT_RESOLVE_TILE_SHAPE  tiled resolver: cells per tile, 0 = 64 x 16 (default), 1 = 32 x 16, 2 = 128 x 16,
 *                   3 = 64 x 32 (the halo is 6 columns / 2 rows in every case).  Same results; for A/B measurements. */
#define LM_OPT_RESOLVE_TILE_SHAPE 11
/*   LM_OPT_ADVECT_MODE  RK4 step (replaces pset.execute(AdvectionRK4), particle_advecter.py:222-223):
 *                   0 (default) bit-faithful to the float32 restatement of Parcels' JIT kernel (oracle/rk4.py);
 *                   1 the same step in float32 FMA arithmetic with MUFU reciprocal / cosine: positions within 1e-6
 *                   relative of the float64 RK4 (north_star's tolerance; tests/test_gpu_advect_fast.py), ~4x fewer
 *                   instructions.  The stored state is float32 in both modes. */
#define LM_OPT_ADVECT_MODE 12
/*   LM_OPT_INTERACT_MODE  pair search + RPS resolution (replaces query_pairs + the pair loop,
 *                   interaction_simulator.py:93-105):
 *                   1 (default) ONE fused pass per tile of LM_TILE_W x LM_TILE_H cells in shared memory, pairs resolved
 *                     in the tile-round order (rounds of matchings; csrc/interact.cu, oracle/rps.py::tile_round_order);
 *                   0 the round-1 pipeline: pair search -> hand-off -> nine phase launches in the cell-phase order
 *                     (csrc/pairs.cu, oracle/rps.py::cell_phase_order).  Same pair set; the species differ because the
 *                     (arbitrary but fixed) sequential order differs -- each is exact against the reference rule run in
 *                     its own order.
 *   LM_OPT_DRAW_BATCH  fused pass, lane walk: lanes waiting for a Philox draw that make their warp run the draws (0 = default 8)
 *   LM_OPT_TILE_CAP    fused pass: microbes a tile stages in shared memory (0 = 1.5 x the mean tile occupancy); fuller
 *                     tiles work on the global arrays through the same code */
#define LM_OPT_INTERACT_MODE 13
#define LM_OPT_DRAW_BATCH 14
#define LM_OPT_TILE_CAP 15
/*   LM_OPT_TILE_REC_CAP  fused pass: records (found pairs of one direction) a tile holds in shared memory (0 = twice the
 *                     staged microbes); a direction with more takes the lane walk (same results)
 *   LM_OPT_TILE_PATH   fused pass: 0 (default) records in shared memory wherever they fit; 1 the lane walk everywhere --
 *                     same results, for tests and A/B measurements */
#define LM_OPT_TILE_REC_CAP 16
#define LM_OPT_TILE_PATH 17
/*   LM_OPT_HEAVY_MIN   hybrid mode: candidate pairs (m_a * m_b; one cell: m (m - 1) / 2) above which a unit is HEAVY, i.e.
 *                     left out of the pair search's hand-off and resolved in rounds of matchings.  0 = default, 1,024 -- the
 *                     value is part of the definition of the cell-round order (oracle/rps.py::cell_round_order takes
 *                     it as a parameter); other values are for A/B measurements. */
#define LM_OPT_HEAVY_MIN 18
/*   LM_OPT_SCATTER_PASSES  the record leaves in particle-id order: the scatter from cell order takes the ids in this many
 *                     windows so that a window's targets stay in L2 until their sectors are complete (0 = auto: windows of
 *                     about 64 MB; 1 = one pass) */
#define LM_OPT_SCATTER_PASSES 19
/*   LM_OPT_RECORD_DEBUG  MEASUREMENT ONLY (the record is then incomplete): bit 0 leaves out the D2H copies of the in-step
 *                     record, bit 1 the scatter to id order -- what each costs the end-to-end loop (tools/scatter_probe.py) */
#define LM_OPT_RECORD_DEBUG 20
/*   LM_OPT_PEER_WAIT_CYCLES  peer-memory exchange: SM clocks a stage spins for a neighbour's message before it gives up
 *                     (default 1.2e11, about a minute): the neighbour is gone -- the wait latches a fault that the next
 *                     lm_sync_stats reports as LM_ESTATE, and later waits of the handle return at once, so a lost rank ends
 *                     the run with an error instead of leaving the device spinning */
#define LM_OPT_PEER_WAIT_CYCLES 21
/* tile of the fused interaction pass, in cells: part of the definition of its canonical pair order.  Strip boundaries
 * (lm_set_strip) must sit on multiples of LM_TILE_H rows in this mode. */
#define LM_TILE_W 32
#define LM_TILE_H 16
int lm_set_option(lm_handle h, int32_t option, int64_t value);
/* Make `stream` wait for work of the last lm_step that is still running on the handle's internal stream
 * (LM_OPT_OVERLAP).  Only needed before the caller reads the resident arrays through pointers obtained earlier,
 * or before it records an end-of-run timing event. */
int lm_join(lm_handle h, void *stream);
/* Number of kernels this library launched since the handle was created. */
int64_t lm_launch_count(lm_handle h);
/* Device time of the phases of the last lm_step run with LM_STEP_TIMING (synchronises on it):
 * ms_out[0] diffuse+advect, [1] binning, [2] pair search, [3] RPS resolution (with strips: including the
 * waits for the halo exchanges), [4] stats.  Host float[5]. */
int lm_phase_times(lm_handle h, float *ms_out);

/* ---- analysis reductions on a snapshot (handle-free) ---------------- */
/* The reference's pair-distance histogram (sandbox/pairwise_distance_histogram_distributed.jl:33-44, :54-64; its
 * unfinished CUDA kernel: sandbox/pairwise_distance_histogram_gpu.jl:14-43): for every unordered pair i < j of the
 * n points (float32 degrees; the reference calls it once per species on that species' microbes)
 *     d   = haversine_distance32(lat_i, lon_i, lat_j, lon_j, radius_m)        float32, the reference's operation order
 *     bin = round(10 * log10(max(1, d)))
 * hist_out: device uint64[bins + 2], zeroed by the call: [b] = pairs of bin b for b = 0..bins ([0]: d < 10^0.05 m,
 * which the reference's 1-based hist[bin] cannot hold), [bins + 1] = pairs beyond the last bin.  The entries always
 * sum to n (n - 1) / 2.  The bin is decided on the float32 haversine argument `a` against bin edges computed in
 * double, so a pair whose distance is within float32 rounding of an edge may land in the neighbouring bin compared
 * with a float32 evaluation of asin / log10 (tests/test_gpu_zz_analysis.py states the band).  n < 2^24,
 * 1 <= bins <= LM_PDH_MAX_BINS (the reference uses 70).  Compute-bound: n^2 / 2 pair evaluations. */
#define LM_PDH_MAX_BINS 126
int lm_pair_distance_hist(const float *lat, const float *lon, int64_t n, float radius_m, int32_t bins,
                          uint64_t *hist_out, void *stream);
/* The frame of microbe_plotter.py:82-155 (plt.scatter of every microbe coloured by species, :132-146) as a raster of
 * width x height pixels over [lon_min, lon_max) x [lat_min, lat_max), row 0 = northern edge:
 *     column = floor((double(lon) - lon_min) * (width / (lon_max - lon_min))),  row likewise from the north;
 * microbes outside the extent are skipped.  counts_out: device uint32[3][height][width], microbes of species 1, 2, 3
 * per pixel; top_out: device int32[height][width], the highest array index in the pixel (-1: none) -- matplotlib draws
 * the markers of one scatter call in array order, so that microbe's marker is the visible one.  species may be NULL
 * (everything counts as species 1).  Both outputs are initialised by the call. */
int lm_rasterize(const float *lon, const float *lat, const int8_t *species, int64_t n, double lon_min, double lon_max,
                 double lat_min, double lat_max, int32_t width, int32_t height, uint32_t *counts_out, int32_t *top_out,
                 void *stream);
/* RGB image (device uint8[height][width][3]) from lm_rasterize's outputs.  palette_rgb: HOST uint8[4][3] --
 * background, rock, paper, scissors (interactions.py:8-10: red, limegreen, blue).  mode LM_FRAME_LAST_DRAWN: the colour
 * of the microbe drawn last (needs top + the species array given to lm_rasterize); LM_FRAME_PLURALITY: the colour of
 * the most numerous species in the pixel (ties: the lower species number; needs counts). */
#define LM_FRAME_LAST_DRAWN 0
#define LM_FRAME_PLURALITY 1
int lm_compose_frame(const uint32_t *counts, const int32_t *top, const int8_t *species, int32_t width, int32_t height,
                     int32_t mode, const uint8_t *palette_rgb /* host */, uint8_t *rgb_out, void *stream);

/* ---- packed per-step record (on-GPU delta-pack; handle-free) ---------- */
/* The position record of one step (what the reference stores per iteration as float32: particle_advecter.py:233-235,
 * interaction_simulator.py:108-110) as the difference to the previous step's record, in float32 ulps, LOSSLESS:
 *     key(x) = bit pattern of x mapped monotonically to uint32 (negative: ~bits, else bits | 0x80000000)
 *     d      = key(cur[i]) - key(prev[i]);  |d| <= 32767: dlon_out[i] / dlat_out[i] = d
 *              else the int16 is LM_DELTA_ESCAPE and {2 i + (0 lon | 1 lat), raw bits of cur[i]} is appended to esc_out
 * All pointers are device pointers in particle-id order.  esc_out: uint32[esc_cap][2], entries in arbitrary order;
 * esc_count: device uint32, zeroed by the call, counts EVERY escape -- a value above esc_cap means the list is
 * incomplete and the caller falls back to the plain record for that step.  4 B instead of 8 B per microbe-step over
 * PCIe; decoded on the host by io.py::unpack_delta_record.  n < 2^31. */
#define LM_DELTA_ESCAPE (-32768)
int lm_record_delta_pack(const float *prev_lon, const float *prev_lat, const float *lon, const float *lat, int64_t n,
                         int16_t *dlon_out, int16_t *dlat_out, uint32_t *esc_out, int64_t esc_cap, uint32_t *esc_count,
                         void *stream);
/* Host decoder of lm_record_delta_pack: all pointers are HOST arrays, no device work.  lon_out / lat_out (float32[n]) may
 * alias prev_lon / prev_lat (decoding in place).  esc: uint32[n_esc][2].  n_threads host threads (>= 1) share the arrays.
 * Bit-exact inverse of the packing.  LM_EINVAL if the number of LM_DELTA_ESCAPE markers differs from n_esc (the escape
 * list overflowed: resend that step plain) or an escape entry points outside the arrays. */
int lm_record_delta_unpack_host(const float *prev_lon, const float *prev_lat, const int16_t *dlon, const int16_t *dlat,
                                const uint32_t *esc, int64_t n_esc, int64_t n, float *lon_out, float *lat_out,
                                int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* LM_B200_H */
