// C ABI of liblm_b200.so -- see include/lm_b200.h for the contract of every entry point and the
// reference call site each one replaces.
#include <cstdio>
#include <cstring>
#include <new>

#include "lm_internal.cuh"

namespace lm {

static thread_local char g_cuda_err[512] = "";

void set_last_cuda_error(cudaError_t e, const char *where)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

}  // namespace lm

using namespace lm;

#define LM_CUDA(call)                                  \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) {                      \
            set_last_cuda_error(e__, #call);           \
            return LM_ECUDA;                           \
        }                                              \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static bool dev_alloc(T **p, int64_t count)
{
    *p = nullptr;
    if (count <= 0) count = 1;
    return cudaMalloc(reinterpret_cast<void **>(p), (size_t)count * sizeof(T)) == cudaSuccess;
}

extern "C" {

int lm_version(void) { return 100; }

const char *lm_error_string(int code)
{
    switch (code) {
        case LM_OK: return "ok";
        case LM_EINVAL: return "invalid argument";
        case LM_ENOMEM: return "out of device memory at lm_create";
        case LM_ECUDA: return "CUDA runtime error";
        case LM_ENOSPC: return "capacity exceeded";
        case LM_ESTATE: return "call order violated (field / grid / state not set)";
        case LM_ENOCONV: return "explicit-order resolver exceeded its round limit";
        default: return "unknown error";
    }
}

const char *lm_last_cuda_error(void) { return g_cuda_err; }

int lm_destroy(lm_handle h)
{
    if (!h) return LM_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->lon[k]); cudaFree(h->lat[k]); cudaFree(h->sp[k]); cudaFree(h->id[k]);
        cudaFree(h->pending[k]);
        cudaFree(h->stage_lon[k]); cudaFree(h->stage_lat[k]); cudaFree(h->stage_sp[k]);
        if (h->ev_scatter[k]) cudaEventDestroy(h->ev_scatter[k]);
        if (h->ev_copied[k]) cudaEventDestroy(h->ev_copied[k]);
    }
    cudaFree(h->keys); cudaFree(h->slots); cudaFree(h->cell_count); cudaFree(h->cell_start);
    cudaFree(h->cell_cursor); cudaFree(h->block_sums); cudaFree(h->ctr); cudaFree(h->head);
    cudaFree(h->pending_cnt);
    cudaFree(h->hits); cudaFree(h->meta);
    for (int k = 0; k < 5; ++k)
        if (h->ev_phase[k]) cudaEventDestroy(h->ev_phase[k]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return LM_OK;
}

int lm_create(lm_handle *out, int device, int64_t max_particles, int64_t max_cells, int64_t max_pairs)
{
    if (!out || max_particles <= 0 || max_cells <= 0 || max_pairs < 0) return LM_EINVAL;
    if (max_particles >= (1ll << 31) - 64 || max_cells >= (1ll << 31) - 64) return LM_EINVAL;
    *out = nullptr;
    LM_CUDA(cudaSetDevice(device));
    lm_handle h = new (std::nothrow) lm_handle_s();
    if (!h) return LM_ENOMEM;
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_particles = max_particles;
    h->max_cells = max_cells;
    h->max_pairs = max_pairs;
    h->emit_cap = h->rps_cap = -1;
    bool ok = true;
    for (int k = 0; k < 2; ++k) {
        ok = ok && dev_alloc(&h->lon[k], max_particles) && dev_alloc(&h->lat[k], max_particles);
        ok = ok && dev_alloc(&h->sp[k], max_particles) && dev_alloc(&h->id[k], max_particles);
        ok = ok && dev_alloc(&h->pending[k], max_pairs);
        ok = ok && dev_alloc(&h->stage_lon[k], max_particles) && dev_alloc(&h->stage_lat[k], max_particles);
        ok = ok && dev_alloc(&h->stage_sp[k], max_particles);
    }
    ok = ok && dev_alloc(&h->keys, max_particles) && dev_alloc(&h->slots, max_particles);
    ok = ok && dev_alloc(&h->cell_count, max_cells) && dev_alloc(&h->cell_start, max_cells + 1);
    ok = ok && dev_alloc(&h->cell_cursor, max_cells) && dev_alloc(&h->block_sums, max_cells / 4096 + 2);
    ok = ok && dev_alloc(&h->hits, max_pairs) && dev_alloc(&h->meta, max_particles);
    ok = ok && dev_alloc(&h->ctr, 1) && dev_alloc(&h->head, max_particles) && dev_alloc(&h->pending_cnt, 4);
    if (ok) ok = cudaMemset(h->cell_count, 0, (size_t)max_cells * sizeof(int32_t)) == cudaSuccess;
    if (ok) ok = cudaMemset(h->ctr, 0, sizeof(Counters)) == cudaSuccess;
    if (ok) ok = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < 2; ++k) {
        ok = ok && cudaEventCreateWithFlags(&h->ev_scatter[k], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_copied[k], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int k = 0; ok && k < 5; ++k) ok = ok && cudaEventCreate(&h->ev_phase[k]) == cudaSuccess;
    if (!ok) {
        cudaError_t e = cudaGetLastError();
        set_last_cuda_error(e, "lm_create");
        lm_destroy(h);
        return LM_ENOMEM;
    }
    *out = h;
    return LM_OK;
}

int lm_set_field(lm_handle h, const float *U, const float *V, const float *lon, const float *lat, int32_t T, int32_t Y,
                 int32_t X)
{
    if (!h || !U || !V || !lon || !lat || T < 1 || Y < 2 || X < 2) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    float ends[4];
    LM_CUDA(cudaMemcpy(&ends[0], lon, sizeof(float), cudaMemcpyDeviceToHost));
    LM_CUDA(cudaMemcpy(&ends[1], lon + X - 1, sizeof(float), cudaMemcpyDeviceToHost));
    LM_CUDA(cudaMemcpy(&ends[2], lat, sizeof(float), cudaMemcpyDeviceToHost));
    LM_CUDA(cudaMemcpy(&ends[3], lat + Y - 1, sizeof(float), cudaMemcpyDeviceToHost));
    if (!(ends[1] > ends[0]) || !(ends[3] > ends[2])) return LM_EINVAL;   // ascending axes required
    h->field.U = U; h->field.V = V; h->field.lon = lon; h->field.lat = lat;
    h->field.T = T; h->field.Y = Y; h->field.X = X;
    h->field.lon0 = ends[0]; h->field.lat0 = ends[2];
    h->field.inv_dx = (float)(X - 1) / (ends[1] - ends[0]);
    h->field.inv_dy = (float)(Y - 1) / (ends[3] - ends[2]);
    h->have_field = true;
    return LM_OK;
}

int lm_update_field_data(lm_handle h, const float *U, const float *V, int32_t T)
{
    if (!h || !U || !V || T < 1) return LM_EINVAL;
    if (!h->have_field) return LM_ESTATE;
    h->field.U = U; h->field.V = V; h->field.T = T;
    return LM_OK;
}

int lm_set_grid(lm_handle h, const lm_grid *g)
{
    if (!h || !g || g->ncx < 1 || g->ncy < 1 || !(g->inv_h > 0.0)) return LM_EINVAL;
    if ((int64_t)g->ncx * g->ncy > h->max_cells) return LM_ENOSPC;
    h->grid = *g;
    h->have_grid = true;
    h->binned = false;
    return LM_OK;
}

int lm_get_grid(lm_handle h, lm_grid *g)
{
    if (!h || !g) return LM_EINVAL;
    if (!h->have_grid) return LM_ESTATE;
    *g = h->grid;
    return LM_OK;
}

static int reset_counters(lm_handle h, cudaStream_t s)
{
    LM_CUDA(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), s));
    return LM_OK;
}

static int check_radius(lm_handle h, double r)
{
    if (!(r >= 0.0)) return LM_EINVAL;
    if (!h->have_grid) return LM_ESTATE;
    if (!(1.0 / h->grid.inv_h > r)) return LM_EINVAL;   // cell edge must exceed the radius
    return LM_OK;
}

int lm_advect_rk4(lm_handle h, float *lon, float *lat, int64_t n, const lm_stage_times *st, float dt, void *stream)
{
    if (!h || !lon || !lat || !st || n < 0 || n > h->max_particles) return LM_EINVAL;
    if (!h->have_field) return LM_ESTATE;
    for (int k = 0; k < 4; ++k)
        if (st->ti[k] < 0 || st->ti[k] + (st->interp[k] ? 1 : 0) >= h->field.T) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    // n_out_of_bounds accumulates over calls until lm_reset_stats
    LM_CUDA(launch_advect(h->field, lon, lat, (int)n, *st, dt, h->ctr, as_stream(stream), &h->launches));
    return LM_OK;
}

int lm_reset_stats(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    h->emit_cap = h->rps_cap = -1;
    return reset_counters(h, as_stream(stream));
}

int lm_diffuse(lm_handle h, float *lon, float *lat, int64_t n, double amp_deg, uint64_t seed, uint64_t step,
               void *stream)
{
    if (!h || !lon || !lat || n < 0 || n > h->max_particles) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    LM_CUDA(launch_diffuse(lon, lat, nullptr, (int)n, amp_deg, seed, step, as_stream(stream), &h->launches));
    return LM_OK;
}

int lm_state_set(lm_handle h, const float *lon, const float *lat, const int8_t *species, const int32_t *ids, int64_t n,
                 void *stream)
{
    if (!h || !lon || !lat || n < 0) return LM_EINVAL;
    if (n > h->max_particles) return LM_ENOSPC;
    if (!h->have_grid) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    int rc = reset_counters(h, s);
    if (rc) return rc;
    const int d = h->cur ^ 1;
    LM_CUDA(launch_bin(h, lon, lat, species, ids, (int)n, h->lon[d], h->lat[d], h->sp[d], h->id[d], s));
    h->cur = d;
    h->n = n;
    h->binned = true;
    return LM_OK;
}

int64_t lm_state_size(lm_handle h) { return h ? h->n : 0; }

static int finish_pairs(lm_handle h, int64_t cap, int64_t *n_pairs_out, cudaStream_t s)
{
    h->emit_cap = cap;
    if (n_pairs_out)
        LM_CUDA(cudaMemcpyAsync(n_pairs_out, &h->ctr->n_pairs, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    return LM_OK;
}

int lm_find_pairs(lm_handle h, const float *lon, const float *lat, int64_t n, double r, int32_t *pairs_out, int64_t cap,
                  int64_t *n_pairs_out, void *stream)
{
    if (!h) return LM_EINVAL;
    int rc = check_radius(h, r);
    if (rc) return rc;
    rc = lm_state_set(h, lon, lat, nullptr, nullptr, n, stream);
    if (rc) return rc;
    cudaStream_t s = as_stream(stream);
    const int c = h->cur;
    LM_CUDA(launch_pairs(h, h->lon[c], h->lat[c], h->id[c], nullptr, (int)n, r, nullptr,
                         reinterpret_cast<int2 *>(pairs_out), pairs_out ? cap : 0, s));
    return finish_pairs(h, pairs_out ? cap : -1, n_pairs_out, s);
}

static RpsDev to_dev(const lm_rps_params *p)
{
    RpsDev d;
    d.pRS = p->pRS; d.pPR = p->pPR; d.pSP = p->pSP;
    d.seed_lo = (uint32_t)p->seed; d.seed_hi = (uint32_t)(p->seed >> 32);
    d.step_lo = (uint32_t)p->step; d.step_hi = (uint32_t)(p->step >> 32);
    return d;
}

int lm_interact_rps(lm_handle h, const float *lon, const float *lat, int8_t *species, int64_t n, double r,
                    const lm_rps_params *prm, int32_t *pairs_out, int64_t cap, int64_t *n_pairs_out, void *stream)
{
    if (!h || !species || !prm) return LM_EINVAL;
    int rc = check_radius(h, r);
    if (rc) return rc;
    rc = lm_state_set(h, lon, lat, species, nullptr, n, stream);
    if (rc) return rc;
    cudaStream_t s = as_stream(stream);
    const int c = h->cur;
    const RpsDev rd = to_dev(prm);
    LM_CUDA(launch_pairs(h, h->lon[c], h->lat[c], h->id[c], h->sp[c], (int)n, r, &rd,
                         reinterpret_cast<int2 *>(pairs_out), pairs_out ? cap : 0, s));
    LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], h->sp[c], h->id[c], (int)n, nullptr, nullptr, species, s,
                                 &h->launches));
    return finish_pairs(h, pairs_out ? cap : -1, n_pairs_out, s);
}

int lm_pair_uniforms(const int32_t *pairs, int64_t n_pairs, uint64_t seed, uint64_t step, double *u_out, void *stream)
{
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !u_out))) return LM_EINVAL;
    LM_CUDA(launch_pair_uniforms(reinterpret_cast<const int2 *>(pairs), n_pairs, seed, step, u_out, as_stream(stream)));
    return LM_OK;
}

int lm_resolve_rps(lm_handle h, const int32_t *pairs, const double *u, int64_t n_pairs, int8_t *species, int64_t n,
                   double pRS, double pPR, double pSP, int32_t *rounds_out, void *stream)
{
    if (!h || n_pairs < 0 || n < 0 || (n_pairs > 0 && (!pairs || !u || !species))) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    return resolve_explicit(h, reinterpret_cast<const int2 *>(pairs), u, n_pairs, species, n, pRS, pPR, pSP,
                            rounds_out, as_stream(stream));
}

int lm_step(lm_handle h, int32_t flags, const lm_stage_times *st, float dt, double diffuse_amp_deg, double r,
            const lm_rps_params *prm, int32_t *pairs_out, int64_t cap, void *stream)
{
    if (!h) return LM_EINVAL;
    if (!h->have_grid) return LM_ESTATE;
    if (h->n <= 0) return LM_OK;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int n = (int)h->n;
    int rc = reset_counters(h, s);
    if (rc) return rc;
    int c = h->cur;
    bool moved = false;
    const bool timing = (flags & LM_STEP_TIMING) != 0;
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[0], s));
    if (flags & LM_STEP_DIFFUSE) {
        // The reference kicks the particles at the END of an iteration, after the positions were
        // stored (particle_advecter.py:233-242): the stored positions -- the ones interactions see --
        // are pre-kick.  So the kick of iteration step-1 is applied here, before this step's advection.
        if (!prm) return LM_EINVAL;
        LM_CUDA(launch_diffuse(h->lon[c], h->lat[c], h->id[c], n, diffuse_amp_deg, prm->seed, prm->step - 1, s,
                               &h->launches));
        moved = true;
    }
    if (flags & LM_STEP_ADVECT) {
        if (!st) return LM_EINVAL;
        if (!h->have_field) return LM_ESTATE;
        for (int k = 0; k < 4; ++k)
            if (st->ti[k] < 0 || st->ti[k] + (st->interp[k] ? 1 : 0) >= h->field.T) return LM_EINVAL;
        LM_CUDA(launch_advect(h->field, h->lon[c], h->lat[c], n, *st, dt, h->ctr, s, &h->launches));
        moved = true;
    }
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[1], s));
    if (moved || !h->binned) {
        const int d = c ^ 1;
        LM_CUDA(launch_bin(h, h->lon[c], h->lat[c], h->sp[c], h->id[c], n, h->lon[d], h->lat[d], h->sp[d], h->id[d], s));
        h->cur = c = d;
        h->binned = true;
    }
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[2], s));
    if (flags & LM_STEP_INTERACT) {
        if (!prm) return LM_EINVAL;
        rc = check_radius(h, r);
        if (rc) return rc;
        const RpsDev rd = to_dev(prm);
        const bool emit = (flags & LM_STEP_EMIT_PAIRS) && pairs_out && cap > 0;
        LM_CUDA(launch_pairs(h, h->lon[c], h->lat[c], h->id[c], h->sp[c], n, r, &rd,
                             emit ? reinterpret_cast<int2 *>(pairs_out) : nullptr, emit ? cap : 0, s));
        h->emit_cap = emit ? cap : -1;
    }
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[3], s));
    if (flags & LM_STEP_STATS) LM_CUDA(launch_stats(h->lon[c], h->lat[c], h->sp[c], n, h->ctr, s, &h->launches));
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[4], s));
    h->timed = timing;
    return LM_OK;
}

int lm_phase_times(lm_handle h, float *ms_out)
{
    if (!h || !ms_out) return LM_EINVAL;
    if (!h->timed) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    LM_CUDA(cudaEventSynchronize(h->ev_phase[4]));
    for (int k = 0; k < 4; ++k) LM_CUDA(cudaEventElapsedTime(ms_out + k, h->ev_phase[k], h->ev_phase[k + 1]));
    return LM_OK;
}

int lm_state_get(lm_handle h, float *lon_out, float *lat_out, int8_t *species_out, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    const int c = h->cur;
    LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], h->sp[c], h->id[c], (int)h->n, lon_out, lat_out, species_out,
                                 as_stream(stream), &h->launches));
    return LM_OK;
}

int lm_state_get_host(lm_handle h, float *lon_host, float *lat_host, int8_t *species_host, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int c = h->cur, k = h->stage_idx;
    const size_t n = (size_t)h->n;
    // the staging buffers of slot k were last read by the copy issued two calls ago
    LM_CUDA(cudaStreamWaitEvent(s, h->ev_copied[k], 0));
    LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], h->sp[c], h->id[c], (int)n, lon_host ? h->stage_lon[k] : nullptr,
                                 lat_host ? h->stage_lat[k] : nullptr, species_host ? h->stage_sp[k] : nullptr, s,
                                 &h->launches));
    LM_CUDA(cudaEventRecord(h->ev_scatter[k], s));
    LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_scatter[k], 0));
    if (lon_host) LM_CUDA(cudaMemcpyAsync(lon_host, h->stage_lon[k], n * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
    if (lat_host) LM_CUDA(cudaMemcpyAsync(lat_host, h->stage_lat[k], n * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
    if (species_host) LM_CUDA(cudaMemcpyAsync(species_host, h->stage_sp[k], n, cudaMemcpyDeviceToHost, h->copy_stream));
    LM_CUDA(cudaEventRecord(h->ev_copied[k], h->copy_stream));
    h->stage_idx = k ^ 1;
    return LM_OK;
}

int lm_host_copies_sync(lm_handle h)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaStreamSynchronize(h->copy_stream));
    return LM_OK;
}

int lm_state_view(lm_handle h, float **lon, float **lat, int8_t **species, int32_t **ids, int32_t **cell_start)
{
    if (!h) return LM_EINVAL;
    const int c = h->cur;
    if (lon) *lon = h->lon[c];
    if (lat) *lat = h->lat[c];
    if (species) *species = h->sp[c];
    if (ids) *ids = h->id[c];
    if (cell_start) *cell_start = h->cell_start;
    return LM_OK;
}

static float dec_f(unsigned int e)
{
    const unsigned int u = (e & 0x80000000u) ? (e ^ 0x80000000u) : ~e;
    float f;
    memcpy(&f, &u, sizeof(f));
    return f;
}

int lm_sync_stats(lm_handle h, lm_stats *out, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    Counters c;
    LM_CUDA(cudaMemcpyAsync(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost, as_stream(stream)));
    LM_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (out) {
        out->n_pairs = (int64_t)c.n_pairs;
        out->n_out_of_bounds = (int64_t)c.n_oob;
        out->n_clamped = (int64_t)c.n_clamped;
        for (int k = 0; k < 4; ++k) out->species_count[k] = (int64_t)c.species[k];
        out->bbox[0] = dec_f(~c.bbox_enc[0]);
        out->bbox[1] = dec_f(c.bbox_enc[1]);
        out->bbox[2] = dec_f(~c.bbox_enc[2]);
        out->bbox[3] = dec_f(c.bbox_enc[3]);
    }
    if (h->emit_cap >= 0 && (int64_t)c.n_pairs > h->emit_cap) return LM_ENOSPC;   // pair list truncated
    if (h->rps_cap >= 0 && (int64_t)c.n_pairs > h->rps_cap) return LM_ENOSPC;     // RPS hand-off buffer too small: species invalid
    if (c.n_overflow) return LM_ENOSPC;
    return LM_OK;
}

int64_t lm_launch_count(lm_handle h) { return h ? h->launches : 0; }

}  // extern "C"
