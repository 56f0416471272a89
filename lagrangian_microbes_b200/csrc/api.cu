// C ABI of liblm_b200.so -- see include/lm_b200.h for the contract of every entry point and the
// reference call site each one replaces.
#include <cstdio>
#include <cstring>
#include <new>
#include <algorithm>

#include "lm_internal.cuh"
#include "philox.cuh"

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
    cudaFree(h->keys); cudaFree(h->slots); cudaFree(h->cell_count); cudaFree(h->cell_start_buf[0]); cudaFree(h->cell_start_buf[1]);
    cudaFree(h->n_pairs_snap);
    cudaFree(h->sticky);
    cudaFree(h->heavy_list); cudaFree(h->heavy_cnt);
    cudaFree(h->uv4);
    cudaFree(h->sp_snap); cudaFree(h->tile_scratch); cudaFree(h->tile_scratch_used);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_find_done) cudaEventDestroy(h->ev_find_done);
    if (h->ev_resolve_done) cudaEventDestroy(h->ev_resolve_done);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_pos_ready) cudaEventDestroy(h->ev_pos_ready);
    if (h->ev_pos_scattered) cudaEventDestroy(h->ev_pos_scattered);
    if (h->ev_sp_ready) cudaEventDestroy(h->ev_sp_ready);
    for (int k = 0; k < 2; ++k)
        if (h->ev_rec_reads[k]) cudaEventDestroy(h->ev_rec_reads[k]);
    cudaFree(h->cell_cursor); cudaFree(h->block_sums); cudaFree(h->ctr); cudaFree(h->head);
    cudaFree(h->pending_cnt);
    cudaFree(h->hits); cudaFree(h->rec); cudaFree(h->rec2);
    for (int k = 0; k < 2; ++k) { cudaFree(h->mig_send[k]); cudaFree(h->mig_recv[k]); }
    cudaFree(h->ghost_send); cudaFree(h->ghost_recv);
    cudaFree(h->gsp_send); cudaFree(h->gsp_recv); cudaFree(h->gret_send); cudaFree(h->gret_recv);
    if (h->xfer_counts_host) cudaFreeHost(h->xfer_counts_host);
    if (h->stats_host) cudaFreeHost(h->stats_host);
    for (int sd = 0; sd < 2; ++sd)
        if (h->peer[sd].connected && h->peer[sd].ipc)
            for (int k = 0; k < LM_PEER_BUFFERS; ++k) if (h->peer[sd].base[k]) cudaIpcCloseMemHandle(h->peer[sd].base[k]);
    cudaFree(h->xflags);
    for (int k = 0; k < 6; ++k)
        if (h->ev_phase[k]) cudaEventDestroy(h->ev_phase[k]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return LM_OK;
}

int lm_create(lm_handle *out, int device, int64_t max_particles, int64_t max_cells, int64_t max_pairs)
{
    if (!out || max_particles <= 0 || max_cells <= 0 || max_pairs < 0) return LM_EINVAL;
    if (max_particles >= (1ll << 31) - 64 || max_cells >= (1ll << 31) - 64) return LM_EINVAL;
    if (max_pairs > 0 && max_particles >= (1ll << 29)) return LM_EINVAL;     // staged hits carry a 29-bit partner index
    if (max_pairs >= (1ll << 32) - 8) return LM_EINVAL;                      // 32-bit entry offsets
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
    ok = ok && dev_alloc(&h->cell_count, max_cells) && dev_alloc(&h->cell_start_buf[0], max_cells + 1);
    ok = ok && dev_alloc(&h->cell_start_buf[1], max_cells + 1) && dev_alloc(&h->n_pairs_snap, 1);
    ok = ok && cudaHostAlloc(reinterpret_cast<void **>(&h->stats_host), sizeof(Counters) + 16, cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->stats_dev), h->stats_host, 0) == cudaSuccess;
    h->cell_start = h->cell_start_buf[0];
    h->overlap = 1;
    h->norm = LM_NORM_2;
    h->advect_mode = 0;
    h->interact_mode = 2;
    h->draw_batch = 0;
    h->tile_cap = 0;
    h->tile_rec_cap = 0;
    h->tile_path = 0;
    h->scatter_passes = 0;
    h->peer_wait_cycles = 120000000000ll;          // about a minute of SM clocks
    h->record_debug = 0;
    h->resolve_tile_smem = 32768;
    h->resolve_batch = 4;      // measured on B200 (profiles/r1y_sweep_resolve.jsonl): 4 beats 1 and 8 on every workload
    ok = ok && dev_alloc(&h->cell_cursor, max_cells) && dev_alloc(&h->block_sums, max_cells / 4096 + 2);
    ok = ok && dev_alloc(&h->hits, max_pairs + 4);
    if (max_pairs > 0) ok = ok && dev_alloc(&h->rec, 5 * max_cells) && dev_alloc(&h->rec2, 5 * (max_particles / 32 + 2));
    ok = ok && dev_alloc(&h->ctr, 1) && dev_alloc(&h->head, max_particles) && dev_alloc(&h->pending_cnt, 4);
    if (ok) ok = cudaMemset(h->cell_count, 0, (size_t)max_cells * sizeof(int32_t)) == cudaSuccess;
    if (ok) ok = cudaMemset(h->ctr, 0, sizeof(Counters)) == cudaSuccess;
    if (ok) ok = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) ok = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_find_done, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_resolve_done, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_pos_ready, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_pos_scattered, cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaEventCreateWithFlags(&h->ev_sp_ready, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < 2 && ok; ++k) ok = cudaEventCreateWithFlags(&h->ev_rec_reads[k], cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaMemset(h->n_pairs_snap, 0, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && dev_alloc(&h->sticky, 1);
    // hybrid mode: queues of heavy units, per phase; a heavy unit needs >= 17 microbes in two cells, so N / 8 per phase is ample
    h->heavy_cap = std::max<int64_t>(1024, max_particles / 8);
    ok = ok && dev_alloc(&h->heavy_list, (size_t)18 * h->heavy_cap) && dev_alloc(&h->heavy_cnt, 36);
    if (ok) ok = cudaMemset(h->heavy_cnt, 0, 36 * sizeof(unsigned int)) == cudaSuccess;
    if (ok) ok = cudaMemset(h->sticky, 0, sizeof(unsigned int)) == cudaSuccess;
    for (int k = 0; ok && k < 2; ++k) {
        ok = ok && cudaEventCreateWithFlags(&h->ev_scatter[k], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_copied[k], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int k = 0; ok && k < 6; ++k) ok = ok && cudaEventCreate(&h->ev_phase[k]) == cudaSuccess;
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
    h->field.lon1 = ends[1]; h->field.lat1 = ends[3];
    h->field.inv_dx = (float)(X - 1) / (ends[1] - ends[0]);
    h->field.inv_dy = (float)(Y - 1) / (ends[3] - ends[2]);
    h->field.UV4 = nullptr;
    h->uv4_stale = true;
    h->have_field = true;
    return LM_OK;
}

int lm_update_field_data(lm_handle h, const float *U, const float *V, int32_t T)
{
    if (!h || !U || !V || T < 1) return LM_EINVAL;
    if (!h->have_field) return LM_ESTATE;
    h->field.U = U; h->field.V = V; h->field.T = T;
    h->field.UV4 = nullptr;
    h->uv4_stale = true;
    return LM_OK;
}

int lm_set_grid(lm_handle h, const lm_grid *g)
{
    if (!h || !g || g->ncx < 1 || g->ncy < 1 || !(g->inv_h > 0.0)) return LM_EINVAL;
    // a handle with strip buffers may be given a GLOBAL grid larger than its own cell table: it then needs
    // lm_set_strip before anything else (rows_owned = 0 until then)
    const bool fits = (int64_t)g->ncx * g->ncy <= h->max_cells;
    if (!fits && !h->send_cap) return LM_ENOSPC;
    h->grid = *g;
    h->have_grid = true;
    h->binned = false;
    // a new grid drops any strip: the handle owns all rows until lm_set_strip says otherwise
    h->strip.row0 = 0;
    h->strip.rows_owned = h->strip.rows_local = fits ? g->ncy : 0;
    h->strip.migrate = 0;
    h->has_south = h->has_north = false;
    h->stage = 0;
    return LM_OK;
}

int lm_strip_alloc(lm_handle h, int64_t send_cap, int64_t ghost_cap, int32_t row_cap)
{
    if (!h || send_cap < 1 || ghost_cap < 1 || row_cap < 1) return LM_EINVAL;
    if (h->send_cap) return LM_ESTATE;   // once per handle
    LM_CUDA(cudaSetDevice(h->device));
    const int64_t ghost_words = GHOST_HDR + (int64_t)row_cap + 1 + 3 * ghost_cap;
    bool ok = true;
    for (int k = 0; k < 2; ++k) ok = ok && dev_alloc(&h->mig_send[k], send_cap + 1) && dev_alloc(&h->mig_recv[k], send_cap + 1);
    ok = ok && dev_alloc(&h->ghost_send, ghost_words) && dev_alloc(&h->ghost_recv, ghost_words);
    ok = ok && dev_alloc(&h->gsp_send, ghost_cap) && dev_alloc(&h->gsp_recv, ghost_cap);
    ok = ok && dev_alloc(&h->gret_send, ghost_cap) && dev_alloc(&h->gret_recv, ghost_cap);
    ok = ok && cudaHostAlloc(reinterpret_cast<void **>(&h->xfer_counts_host), 4 * sizeof(int32_t), cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->xfer_counts_dev), h->xfer_counts_host, 0) == cudaSuccess;
    ok = ok && dev_alloc(&h->xflags, 8);
    if (ok) ok = cudaMemset(h->xflags, 0, 8 * sizeof(unsigned int)) == cudaSuccess;
    for (int k = 0; ok && k < 2; ++k) {
        ok = ok && cudaMemset(h->mig_send[k], 0, (size_t)(send_cap + 1) * sizeof(int4)) == cudaSuccess;
        ok = ok && cudaMemset(h->mig_recv[k], 0, (size_t)(send_cap + 1) * sizeof(int4)) == cudaSuccess;
    }
    if (ok) ok = cudaMemset(h->ghost_send, 0, (size_t)ghost_words * 4) == cudaSuccess;
    if (ok) ok = cudaMemset(h->ghost_recv, 0, (size_t)ghost_words * 4) == cudaSuccess;
    if (ok) ok = cudaMemset(h->gsp_send, 0, (size_t)ghost_cap) == cudaSuccess && cudaMemset(h->gsp_recv, 0, (size_t)ghost_cap) == cudaSuccess;
    if (ok) ok = cudaMemset(h->gret_send, 0, (size_t)ghost_cap) == cudaSuccess && cudaMemset(h->gret_recv, 0, (size_t)ghost_cap) == cudaSuccess;
    if (!ok) {
        set_last_cuda_error(cudaGetLastError(), "lm_strip_alloc");
        return LM_ENOMEM;
    }
    h->send_cap = send_cap; h->ghost_cap = ghost_cap; h->row_cap = row_cap;
    return LM_OK;
}

int lm_set_strip(lm_handle h, const lm_strip *st)
{
    if (!h || !st) return LM_EINVAL;
    if (!h->have_grid || !h->send_cap) return LM_ESTATE;
    const lm_grid &g = h->grid;
    if (st->row0 < 0 || st->rows_owned < 1 || st->row0 + st->rows_owned > g.ncy) return LM_EINVAL;
    if (st->row0 & 1) return LM_EINVAL;                                  // boundaries on even rows
    if (st->has_south && st->row0 == 0) return LM_EINVAL;
    if (st->has_north && st->row0 + st->rows_owned >= g.ncy) return LM_EINVAL;
    if (!st->has_south && st->row0 != 0) return LM_EINVAL;
    if (!st->has_north && st->row0 + st->rows_owned != g.ncy) return LM_EINVAL;
    if (g.ncx > h->row_cap) return LM_ENOSPC;
    const int rows_local = st->rows_owned + (st->has_north ? 1 : 0);
    if ((int64_t)rows_local * g.ncx > h->max_cells) return LM_ENOSPC;
    LM_CUDA(cudaSetDevice(h->device));
    h->strip.row0 = st->row0;
    h->strip.rows_owned = st->rows_owned;
    h->strip.rows_local = rows_local;
    h->strip.migrate = 0;
    h->has_south = st->has_south != 0;
    h->has_north = st->has_north != 0;
    h->binned = false;
    h->stage = 0;
    return LM_OK;
}

// ---- peer-memory exchange (include/lm_b200.h: lm_peer_export) ----------------------------------------------------------
int lm_strip_peer_export(lm_handle h, lm_peer_export *out)
{
    if (!h || !out) return LM_EINVAL;
    if (!h->send_cap) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    void *buf[LM_PEER_BUFFERS] = {h->mig_recv[0], h->mig_recv[1], h->ghost_recv, h->gsp_recv, h->gret_recv, h->xflags};
    memset(out, 0, sizeof(*out));
    for (int k = 0; k < LM_PEER_BUFFERS; ++k) {
        out->ptr[k] = buf[k];
        cudaIpcMemHandle_t hd;
        static_assert(sizeof(cudaIpcMemHandle_t) == LM_IPC_HANDLE_BYTES, "IPC handle size");
        if (cudaIpcGetMemHandle(&hd, buf[k]) == cudaSuccess) memcpy(out->ipc[k], &hd, sizeof(hd));
        else cudaGetLastError();               // no IPC on this platform: the pointers still serve a same-process neighbour
    }
    return LM_OK;
}

int lm_strip_peer_connect(lm_handle h, int32_t side, const lm_peer_export *peer, int32_t use_ipc)
{
    if (!h || !peer || side < 0 || side > 1) return LM_EINVAL;
    if (!h->send_cap) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    lm_handle_s::Peer &P = h->peer[side];
    if (P.connected) return LM_ESTATE;
    void *m[LM_PEER_BUFFERS];
    for (int k = 0; k < LM_PEER_BUFFERS; ++k) {
        P.base[k] = nullptr;
        if (use_ipc) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, peer->ipc[k], sizeof(hd));
            LM_CUDA(cudaIpcOpenMemHandle(&m[k], hd, cudaIpcMemLazyEnablePeerAccess));
            P.base[k] = m[k];
        } else m[k] = peer->ptr[k];
        if (!m[k]) return LM_EINVAL;
    }
    // my southern neighbour receives what I send south in ITS northern slot (mig_recv[1]) and vice versa
    P.mig_recv = static_cast<int4 *>(m[side == 0 ? 1 : 0]);
    P.ghost_recv = static_cast<int32_t *>(m[2]);
    P.gsp_recv = static_cast<int8_t *>(m[3]);
    P.gret_recv = static_cast<int8_t *>(m[4]);
    P.flags = static_cast<unsigned int *>(m[5]);
    P.ipc = use_ipc != 0;
    P.connected = true;
    h->xseq = 0;
    LM_CUDA(cudaMemset(h->xflags, 0, 8 * sizeof(unsigned int)));
    return LM_OK;
}

// flag words of a strip: [0] migrants from the south arrived | [1] from the north | [2] ghost row | [3] its species (gsp) |
// [4] the species coming back (gret) | [5] the southern neighbour has consumed my migrants | [6] the northern one has
int lm_step_push(lm_handle h, int32_t kind, void *stream)
{
    if (!h || kind < LM_XCHG_MIG || kind > LM_XCHG_GRET) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const unsigned int seq = h->xseq;
    if (kind == LM_XCHG_MIG) {
        for (int d = 0; d < 2; ++d) {
            if (!(d ? h->has_north : h->has_south) || !h->peer[d].connected) continue;
            // the neighbour's buffer is free once it has consumed the previous message (its acknowledgement lands in my flags)
            if (seq > 1) LM_CUDA(launch_peer_wait(h->xflags + 5 + d, seq - 1, h->sticky, h->peer_wait_cycles, s, &h->launches));
            LM_CUDA(launch_peer_push_mig(h->mig_send[d], h->peer[d].mig_recv, h->send_cap, s, &h->launches));
            LM_CUDA(launch_peer_signal(h->peer[d].flags + (d == 0 ? 1 : 0), seq, s, &h->launches));   // I am its northern / southern side
        }
        return LM_OK;
    }
    // the other messages were written into the neighbour's buffer by their pack kernels: only the flag is left
    const bool interact = (h->step_flags & LM_STEP_INTERACT) != 0;
    if (!interact) return LM_OK;
    if (kind == LM_XCHG_GHOST && h->has_south && h->peer[0].connected) LM_CUDA(launch_peer_signal(h->peer[0].flags + 2, seq, s, &h->launches));
    if (kind == LM_XCHG_GSP && h->has_south && h->peer[0].connected) LM_CUDA(launch_peer_signal(h->peer[0].flags + 3, seq, s, &h->launches));
    if (kind == LM_XCHG_GRET && h->has_north && h->peer[1].connected) LM_CUDA(launch_peer_signal(h->peer[1].flags + 4, seq, s, &h->launches));
    return LM_OK;
}

int lm_strip_buffers_get(lm_handle h, lm_strip_buffers *out)
{
    if (!h || !out) return LM_EINVAL;
    if (!h->send_cap) return LM_ESTATE;
    for (int k = 0; k < 2; ++k) { out->mig_send[k] = h->mig_send[k]; out->mig_recv[k] = h->mig_recv[k]; }
    out->mig_bytes = (h->send_cap + 1) * (int64_t)sizeof(int4);
    out->ghost_send = h->ghost_send; out->ghost_recv = h->ghost_recv;
    out->ghost_bytes = (GHOST_HDR + h->row_cap + 1 + 3 * h->ghost_cap) * 4;
    out->gsp_send = h->gsp_send; out->gsp_recv = h->gsp_recv;
    out->gret_send = h->gret_send; out->gret_recv = h->gret_recv;
    out->species_bytes = h->ghost_cap;
    return LM_OK;
}

int lm_get_grid(lm_handle h, lm_grid *g)
{
    if (!h || !g) return LM_EINVAL;
    if (!h->have_grid) return LM_ESTATE;
    *g = h->grid;
    return LM_OK;
}

// Capacity faults of a step must not vanish with the counters the next step zeroes: before every reset they are
// latched into h->sticky (bit 0 pair list truncated | 1 RPS hand-off overflowed: species invalid | 2 cell too full for
// the packed counters | 3 migration / ghost buffer overflow | 4 misrouted particles), which only lm_sync_stats (which
// reports them) and lm_reset_stats clear.
__global__ void latch_faults_kernel(const Counters *c, unsigned int *sticky, long long emit_cap, long long rps_cap)
{
    unsigned int f = 0;
    if (emit_cap >= 0 && (long long)c->n_pairs > emit_cap) f |= 1u;
    if (rps_cap >= 0 && (long long)c->n_pairs > rps_cap) f |= 2u;
    if (c->n_overflow) f |= 4u;
    if (c->n_xfer_overflow || c->n_heavy_overflow) f |= 8u;
    if (c->n_misrouted) f |= 16u;
    if (f) atomicOr(sticky, f);
}

// LM_OPT_ADVECT_MODE = 1 samples an interleaved copy of the field (one 16-byte load per corner instead of four 4-byte
// loads): (re)built on the advecting stream after lm_set_field / lm_update_field_data.  A single time level has no
// interval to interleave: the kernel then reads U and V directly.
static int ensure_uv4(lm_handle h, cudaStream_t s)
{
    if (h->advect_mode != 1 || !h->have_field) return LM_OK;
    if (h->field.T < 2) { h->field.UV4 = nullptr; return LM_OK; }
    if (!h->uv4_stale && h->field.UV4) return LM_OK;
    const size_t need = (size_t)(h->field.T - 1) * h->field.Y * h->field.X;
    if (need > h->uv4_elems) {
        LM_CUDA(cudaStreamSynchronize(s));               // nothing may still read the old copy
        cudaFree(h->uv4);
        h->uv4 = nullptr; h->uv4_elems = 0;
        if (cudaMalloc(reinterpret_cast<void **>(&h->uv4), need * sizeof(float4)) != cudaSuccess) { cudaGetLastError(); return LM_ENOMEM; }
        h->uv4_elems = need;
    }
    LM_CUDA(launch_interleave_field(h->field, h->uv4, s, &h->launches));
    h->field.UV4 = h->uv4;
    h->uv4_stale = false;
    return LM_OK;
}

static int reset_counters(lm_handle h, cudaStream_t s)
{
    if (!h->ctr_reported) {          // counters nobody has read through lm_sync_stats: keep their faults
        latch_faults_kernel<<<1, 1, 0, s>>>(h->ctr, h->sticky, (long long)h->emit_cap, (long long)h->rps_cap);
        LM_CUDA(cudaGetLastError());
    }
    h->ctr_reported = false;
    LM_CUDA(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), s));
    return LM_OK;
}

// RPS phases of the last step may still be running on the side stream: make `s` wait for them before it
// touches species / the hand-off / the state.
// id windows of a scatter into ``arrays`` float32 targets: each window's targets are to stay L2-resident
static int scatter_windows(const lm_handle_s *h, int arrays)
{
    if (h->scatter_passes > 0) return h->scatter_passes;
    if (arrays == 0) return 1;
    const int64_t bytes = (int64_t)h->n * 4 * arrays;
    const int64_t window = 64ll << 20;   // measured: 12.5 M microbes, lon + lat (100 MB): 0.34 ms in one pass, 0.26 in two, 0.36 in four
    return (int)std::max<int64_t>(1, std::min<int64_t>(16, (bytes + window - 1) / window));
}

static int join_side(lm_handle h, cudaStream_t s)
{
    if (h->resolve_pending) {
        LM_CUDA(cudaStreamWaitEvent(s, h->ev_resolve_done, 0));
        h->resolve_pending = false;
    }
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
    { const int rcu = ensure_uv4(h, as_stream(stream)); if (rcu) return rcu; }
    LM_CUDA(launch_advect(h->field, lon, lat, (int)n, *st, dt, h->ctr, as_stream(stream), &h->launches, h->advect_mode));
    return LM_OK;
}

int lm_reset_stats(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    h->emit_cap = h->rps_cap = -1;
    LM_CUDA(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), as_stream(stream)));
    LM_CUDA(cudaMemsetAsync(h->sticky, 0, sizeof(unsigned int), as_stream(stream)));
    return LM_OK;
}

int lm_diffuse(lm_handle h, float *lon, float *lat, int64_t n, double amp_deg, uint64_t seed, uint64_t step,
               void *stream)
{
    if (!h || !lon || !lat || n < 0 || n > h->max_particles) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    LM_CUDA(launch_diffuse(lon, lat, nullptr, (int)n, amp_deg, seed, step, as_stream(stream), &h->launches));
    return LM_OK;
}

int lm_diffuse_ids(lm_handle h, float *lon, float *lat, const int32_t *ids, int64_t n, double amp_deg, uint64_t seed,
                   uint64_t step, void *stream)
{
    if (!h || !lon || !lat || n < 0 || n > h->max_particles) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    LM_CUDA(launch_diffuse(lon, lat, ids, (int)n, amp_deg, seed, step, as_stream(stream), &h->launches));
    return LM_OK;
}

int lm_state_set(lm_handle h, const float *lon, const float *lat, const int8_t *species, const int32_t *ids, int64_t n,
                 void *stream)
{
    if (!h || !lon || !lat || n < 0) return LM_EINVAL;
    if (n > h->max_particles) return LM_ENOSPC;
    if (!h->have_grid || h->strip.rows_owned < 1) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    { const int rcj = join_side(h, s); if (rcj) return rcj; }
    for (int k = 0; k < 2; ++k) LM_CUDA(cudaStreamWaitEvent(s, h->ev_copied[k], 0));   // records still reading the old state
    h->pos_scatter_pending = false;
    int rc = reset_counters(h, s);
    if (rc) return rc;
    h->stage = 0;
    if (h->has_south || h->has_north) {
        // strips: the particles are only loaded; the first step bins them and sends those that belong to a
        // neighbouring strip on their way (repeat a flags = 0 step until lm_stats.n_misrouted == 0 everywhere)
        if (!ids) return LM_EINVAL;                        // global ids are required
        const int c = h->cur;
        const size_t m = (size_t)n;
        LM_CUDA(cudaMemcpyAsync(h->lon[c], lon, m * sizeof(float), cudaMemcpyDeviceToDevice, s));
        LM_CUDA(cudaMemcpyAsync(h->lat[c], lat, m * sizeof(float), cudaMemcpyDeviceToDevice, s));
        LM_CUDA(cudaMemcpyAsync(h->id[c], ids, m * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        if (species) LM_CUDA(cudaMemcpyAsync(h->sp[c], species, m, cudaMemcpyDeviceToDevice, s));
        else LM_CUDA(cudaMemsetAsync(h->sp[c], 0, m, s));
        h->n = n;
        h->binned = false;
        return LM_OK;
    }
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
    if (h->has_south || h->has_north) return LM_ESTATE;   // stateless operators work on a whole domain
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
    d.pair_key = pair_stream_key(p->seed, p->step);
    return d;
}

int lm_interact_rps(lm_handle h, const float *lon, const float *lat, int8_t *species, int64_t n, double r,
                    const lm_rps_params *prm, int32_t *pairs_out, int64_t cap, int64_t *n_pairs_out, void *stream)
{
    if (!h || !species || !prm) return LM_EINVAL;
    if (h->has_south || h->has_north) return LM_ESTATE;
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

static inline bool in_strip_mode(lm_handle h) { return h->has_south || h->has_north; }

int lm_step_move(lm_handle h, int32_t flags, const lm_stage_times *st, float dt, double diffuse_amp_deg,
                 const lm_rps_params *prm, void *stream)
{
    if (!h) return LM_EINVAL;
    if (!h->have_grid || h->strip.rows_owned < 1) return LM_ESTATE;
    if ((flags & (LM_STEP_DIFFUSE | LM_STEP_INTERACT)) && !prm) return LM_EINVAL;
    if (flags & LM_STEP_ADVECT) {
        if (!st) return LM_EINVAL;
        if (!h->have_field) return LM_ESTATE;
        for (int k = 0; k < 4; ++k)
            if (st->ti[k] < 0 || st->ti[k] + (st->interp[k] ? 1 : 0) >= h->field.T) return LM_EINVAL;
    }
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int n = (int)h->n;
    int rc = reset_counters(h, s);
    if (rc) return rc;
    if (in_strip_mode(h)) { rc = join_side(h, s); if (rc) return rc; }     // leavers are packed with their species
    if (h->pos_scatter_pending) {       // the previous step's record still reads the positions this step updates in place
        LM_CUDA(cudaStreamWaitEvent(s, h->ev_pos_scattered, 0));
        h->pos_scatter_pending = false;
    }
    h->rec_active = h->rec_armed && (h->rec_ids_host || !in_strip_mode(h));
    h->rec_by_ids = h->rec_active && h->rec_ids_host != nullptr;
    h->rec_armed = false;
    const int c = h->cur;
    bool moved = false;
    const bool timing = (flags & LM_STEP_TIMING) != 0;
    h->timed = false;
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[0], s));
    if (flags & LM_STEP_DIFFUSE) {
        // The reference kicks the particles at the END of an iteration, after the positions were
        // stored (particle_advecter.py:233-242): the stored positions -- the ones interactions see --
        // are pre-kick.  So the kick of iteration step-1 is applied here, before this step's advection.
        LM_CUDA(launch_diffuse(h->lon[c], h->lat[c], h->id[c], n, diffuse_amp_deg, prm->seed, prm->step - 1, s,
                               &h->launches));
        moved = true;
    }
    if (flags & LM_STEP_ADVECT) {
        { const int rcu = ensure_uv4(h, s); if (rcu) return rcu; }
        LM_CUDA(launch_advect(h->field, h->lon[c], h->lat[c], n, *st, dt, h->ctr, s, &h->launches, h->advect_mode));
        moved = true;
    }
    if (timing) LM_CUDA(cudaEventRecord(h->ev_phase[1], s));
    h->step_moved = moved || !h->binned;
    if (h->step_moved) {
        for (int d = 0; d < 2; ++d)
            if (d ? h->has_north : h->has_south) LM_CUDA(cudaMemsetAsync(h->mig_send[d], 0, sizeof(int4), s));
        LM_CUDA(launch_bin_count(h, h->lon[c], h->lat[c], h->sp[c], h->id[c], 0, n, in_strip_mode(h), s));
    }
    h->step_flags = flags;
    if (prm) h->step_rps = to_dev(prm);
    ++h->xseq;                            // peer-memory exchange: the messages of this step carry this number
    h->stage = 1;
    return LM_OK;
}

int lm_step_bin(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    if (h->stage != 1) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    int c = h->cur;
    for (int d = 0; d < 2; ++d)          // peer-memory exchange: the neighbours' migrants of this step have landed
        if ((d ? h->has_north : h->has_south) && h->peer[d].connected) LM_CUDA(launch_peer_wait(h->xflags + d, h->xseq, h->sticky, h->peer_wait_cycles, s, &h->launches));
    if (h->step_moved) {
        // the re-binning gathers species: the previous step's RPS phases must be done
        { const int rcj = join_side(h, s); if (rcj) return rcj; }
        int n_in = (int)h->n, n_out = (int)h->n;
        h->n_moved_in = h->n_moved_out = 0;
        if (in_strip_mode(h)) {
            // the one host synchronisation of a multi-GPU step: how many left, how many arrived
            int32_t *cnt = h->xfer_counts_host;
            cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
            LM_CUDA(launch_xfer_counts(h->has_south ? h->mig_send[0] : nullptr, h->has_north ? h->mig_send[1] : nullptr,
                                       h->has_south ? h->mig_recv[0] : nullptr, h->has_north ? h->mig_recv[1] : nullptr,
                                       h->xfer_counts_dev, s, &h->launches));
            LM_CUDA(cudaStreamSynchronize(s));
            for (int k = 0; k < 4; ++k) {       // a full message holds send_cap records; the rest stayed behind
                if (cnt[k] < 0) return LM_EINVAL;
                if (cnt[k] > h->send_cap) cnt[k] = (int32_t)h->send_cap;
            }
            const int n_arr = cnt[2] + cnt[3];
            if ((int64_t)n_in + n_arr > h->max_particles) return LM_ENOSPC;
            int first = n_in;
            for (int d = 0; d < 2; ++d) {
                LM_CUDA(launch_unpack_arrivals(h, d, cnt[2 + d], first, h->lon[c], h->lat[c], h->sp[c], h->id[c], s));
                first += cnt[2 + d];
            }
            LM_CUDA(launch_bin_count(h, h->lon[c], h->lat[c], h->sp[c], h->id[c], n_in, n_arr, false, s));
            n_out = n_in - cnt[0] - cnt[1] + n_arr;
            n_in += n_arr;
            h->n_moved_out = cnt[0] + cnt[1];
            h->n_moved_in = n_arr;
        }
        const int d = c ^ 1;
        // buffers [d] held the state two steps ago: a species record of that step (scattered on the copy stream, behind its
        // position copies) may still be reading them when no record of the step in between orders the streams (stride >= 2)
        for (int k = 0; k < 2; ++k)
            if (h->rec_reads_age[k] > 0 && --h->rec_reads_age[k] == 0) LM_CUDA(cudaStreamWaitEvent(s, h->ev_rec_reads[k], 0));
        LM_CUDA(launch_bin_finish(h, h->lon[c], h->lat[c], h->sp[c], h->id[c], n_in, n_out, h->lon[d], h->lat[d],
                                  h->sp[d], h->id[d], s));
        h->cur = c = d;
        h->n = n_out;
        h->binned = true;
    }
    for (int d = 0; d < 2; ++d)          // ... and are consumed (or were not needed): the neighbour may overwrite them
        if ((d ? h->has_north : h->has_south) && h->peer[d].connected)
            LM_CUDA(launch_peer_signal(h->peer[d].flags + 5 + (d == 0 ? 1 : 0), h->xseq, s, &h->launches));
    if (h->step_flags & LM_STEP_TIMING) LM_CUDA(cudaEventRecord(h->ev_phase[2], s));
    if (h->rec_active && h->rec_by_ids) {
        // the record in STORAGE order with the ids beside it (strips: the owned particles of this strip).  lon / lat are
        // advected in place by the next step, so they are staged (device to device, on the copy stream); ids and species
        // stay where they are until the re-binning two steps on (sp_scatter_age) and are copied from there.
        const int k = h->rec_slot = h->stage_idx;
        h->stage_idx = k ^ 1;
        const size_t nn = (size_t)h->n;
        LM_CUDA(cudaEventRecord(h->ev_pos_ready, s));
        LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_pos_ready, 0));
        if (nn && h->rec_lon_host) LM_CUDA(cudaMemcpyAsync(h->stage_lon[k], h->lon[c], nn * sizeof(float), cudaMemcpyDeviceToDevice, h->copy_stream));
        if (nn && h->rec_lat_host) LM_CUDA(cudaMemcpyAsync(h->stage_lat[k], h->lat[c], nn * sizeof(float), cudaMemcpyDeviceToDevice, h->copy_stream));
        LM_CUDA(cudaEventRecord(h->ev_pos_scattered, h->copy_stream));
        h->pos_scatter_pending = true;
        if (nn && !(h->record_debug & 1)) {
            LM_CUDA(cudaMemcpyAsync(h->rec_ids_host, h->id[c], nn * sizeof(int32_t), cudaMemcpyDeviceToHost, h->copy_stream));
            if (h->rec_lon_host) LM_CUDA(cudaMemcpyAsync(h->rec_lon_host, h->stage_lon[k], nn * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
            if (h->rec_lat_host) LM_CUDA(cudaMemcpyAsync(h->rec_lat_host, h->stage_lat[k], nn * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
        }
        h->rec_count = (int64_t)nn;
    } else if (h->rec_active) {
        // positions and ids of this step are final: scatter them to id order and send them to the host on the copy
        // stream, under the pair search (which only reads them)
        const int k = h->rec_slot = h->stage_idx;
        h->stage_idx = k ^ 1;
        const size_t nn = (size_t)h->n;
        LM_CUDA(cudaEventRecord(h->ev_pos_ready, s));
        LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_pos_ready, 0));
        if (!(h->record_debug & 2))
            LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], nullptr, h->id[c], (int)nn, h->rec_lon_host ? h->stage_lon[k] : nullptr,
                                         h->rec_lat_host ? h->stage_lat[k] : nullptr, nullptr, h->copy_stream, &h->launches,
                                         scatter_windows(h, (h->rec_lon_host ? 1 : 0) + (h->rec_lat_host ? 1 : 0)), (int)nn));
        LM_CUDA(cudaEventRecord(h->ev_pos_scattered, h->copy_stream));
        h->pos_scatter_pending = true;
        if (!(h->record_debug & 1)) {
            if (h->rec_lon_host) LM_CUDA(cudaMemcpyAsync(h->rec_lon_host, h->stage_lon[k], nn * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
            if (h->rec_lat_host) LM_CUDA(cudaMemcpyAsync(h->rec_lat_host, h->stage_lat[k], nn * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
        }
    }
    // the halo is only needed by (and only sized for) interacting steps; routing passes skip it
    if (h->has_south && (h->step_flags & LM_STEP_INTERACT)) LM_CUDA(launch_ghost_pack(h, h->lon[c], h->lat[c], h->id[c], s));
    h->stage = 2;
    return LM_OK;
}

int lm_step_interact_begin(lm_handle h, double r, int32_t *pairs_out, int64_t cap, void *stream)
{
    if (!h) return LM_EINVAL;
    if (h->stage != 2) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    { const int rcj = join_side(h, s); if (rcj) return rcj; }
    const int c = h->cur, n = (int)h->n;
    const bool interact = (h->step_flags & LM_STEP_INTERACT) != 0;
    if (h->has_north && interact && h->peer[1].connected) LM_CUDA(launch_peer_wait(h->xflags + 2, h->xseq, h->sticky, h->peer_wait_cycles, s, &h->launches));
    if (h->has_north && interact) LM_CUDA(launch_ghost_unpack(h, h->lon[c], h->lat[c], h->id[c], n, s));
    h->rps_cap = -1;
    h->emit_cap = -1;
    if (h->step_flags & LM_STEP_INTERACT) {
        const int rc = check_radius(h, r);
        if (rc) return rc;
        const bool emit = (h->step_flags & LM_STEP_EMIT_PAIRS) && pairs_out && cap > 0;
        LM_CUDA(launch_find(h, h->lon[c], h->lat[c], h->id[c], n, r, &h->step_rps,
                            emit ? reinterpret_cast<int2 *>(pairs_out) : nullptr, emit ? cap : 0, s));
        h->emit_cap = emit ? cap : -1;
        if (h->step_flags & LM_STEP_TIMING) LM_CUDA(cudaEventRecord(h->ev_phase[3], s));
        // Species never feed back into advection, so on a single handle the nine RPS phases -- latency
        // bound, half-empty warps -- go to a side stream and run under the issue-bound advection of the NEXT step;
        // whatever touches species, the hand-off or the state next waits for them (join_side).
        h->resolve_on_side = h->overlap && h->interact_mode == 0 && !in_strip_mode(h) && !(h->step_flags & (LM_STEP_TIMING | LM_STEP_STATS)) && n > 0;
        cudaStream_t rs = s;
        if (h->resolve_on_side) {
            LM_CUDA(cudaEventRecord(h->ev_find_done, s));
            LM_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_find_done, 0));
            rs = h->side_stream;
        }
        // the tiled resolver takes all nine phases in one launch when no halo exchange has to happen after phase 5
        h->resolve_all_in_begin = h->interact_mode == 0 && h->resolve_mode == 1 && !in_strip_mode(h);   // (the hybrid mode interleaves phases: nine launches)
        if (n > 0) LM_CUDA(launch_resolve_phases(h, h->sp[c], 0, h->resolve_all_in_begin ? 8 : 5, rs));
    }
    if (h->has_south && interact) LM_CUDA(launch_row0_species_pack(h, h->sp[c], s));
    h->stage = 3;
    return LM_OK;
}

int lm_step_interact_end(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    if (h->stage != 3) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int c = h->cur, n = (int)h->n;
    const bool interact = (h->step_flags & LM_STEP_INTERACT) != 0;
    if (h->has_north && interact && h->peer[1].connected) LM_CUDA(launch_peer_wait(h->xflags + 3, h->xseq, h->sticky, h->peer_wait_cycles, s, &h->launches));
    if (h->has_north && interact) LM_CUDA(launch_ghost_species_unpack(h, h->sp[c], n, s));
    if (interact && n > 0) {
        if (!h->resolve_all_in_begin) LM_CUDA(launch_resolve_phases(h, h->sp[c], 6, 8, h->resolve_on_side ? h->side_stream : s));
        h->resolve_all_in_begin = false;
        if (h->resolve_on_side) {
            LM_CUDA(cudaEventRecord(h->ev_resolve_done, h->side_stream));
            h->resolve_pending = true;
            h->resolve_on_side = false;
        }
    }
    if (h->has_north && interact) LM_CUDA(launch_ghost_species_pack(h, h->sp[c], n, s));
    if (h->step_flags & LM_STEP_TIMING) {
        if (!interact) LM_CUDA(cudaEventRecord(h->ev_phase[3], s));
        LM_CUDA(cudaEventRecord(h->ev_phase[4], s));
    }
    h->stage = 4;
    return LM_OK;
}

int lm_step_finish(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    if (h->stage != 4) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int c = h->cur, n = (int)h->n;
    if (h->has_south && (h->step_flags & LM_STEP_INTERACT) && h->peer[0].connected)
        LM_CUDA(launch_peer_wait(h->xflags + 4, h->xseq, h->sticky, h->peer_wait_cycles, s, &h->launches));
    if (h->has_south && (h->step_flags & LM_STEP_INTERACT)) LM_CUDA(launch_row0_species_unpack(h, h->sp[c], s));
    if (h->step_flags & LM_STEP_STATS) LM_CUDA(launch_stats(h->lon[c], h->lat[c], h->sp[c], n, h->ctr, s, &h->launches));
    if (h->step_flags & LM_STEP_TIMING) {
        LM_CUDA(cudaEventRecord(h->ev_phase[5], s));
        h->timed = true;
    }
    if (h->rec_active) {
        // species after this step's interactions (interaction_simulator.py:108-110): after the RPS phases, wherever
        // they run; the copy stream does not hold up the caller's stream
        const int k = h->rec_slot;
        if (h->rec_by_ids) {
            if (h->rec_sp_host && n > 0) {
                if (h->resolve_pending) LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_resolve_done, 0));
                else {
                    LM_CUDA(cudaEventRecord(h->ev_sp_ready, s));
                    LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_sp_ready, 0));
                }
                if (!(h->record_debug & 1))
                    LM_CUDA(cudaMemcpyAsync(h->rec_sp_host, h->sp[c], (size_t)n, cudaMemcpyDeviceToHost, h->copy_stream));
            }
            LM_CUDA(cudaEventRecord(h->ev_rec_reads[k], h->copy_stream));      // ids and species of buffers [c] have been read
            h->rec_reads_age[k] = 2;
        } else if (h->rec_sp_host) {
            if (h->resolve_pending) LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_resolve_done, 0));
            else {
                LM_CUDA(cudaEventRecord(h->ev_sp_ready, s));
                LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_sp_ready, 0));
            }
            if (!(h->record_debug & 2))
                LM_CUDA(launch_scatter_by_id(nullptr, nullptr, h->sp[c], h->id[c], n, nullptr, nullptr, h->stage_sp[k], h->copy_stream,
                                             &h->launches));
            LM_CUDA(cudaEventRecord(h->ev_rec_reads[k], h->copy_stream));
            h->rec_reads_age[k] = 2;
            if (!(h->record_debug & 1))
                LM_CUDA(cudaMemcpyAsync(h->rec_sp_host, h->stage_sp[k], (size_t)n, cudaMemcpyDeviceToHost, h->copy_stream));
        }
        LM_CUDA(cudaEventRecord(h->ev_copied[k], h->copy_stream));
        h->rec_active = false;
    }
    h->stage = 0;
    return LM_OK;
}

int lm_step(lm_handle h, int32_t flags, const lm_stage_times *st, float dt, double diffuse_amp_deg, double r,
            const lm_rps_params *prm, int32_t *pairs_out, int64_t cap, void *stream)
{
    if (!h) return LM_EINVAL;
    if (in_strip_mode(h)) return LM_ESTATE;   // strips need the exchanges between the stages: use lm_step_move ...
    int rc = lm_step_move(h, flags, st, dt, diffuse_amp_deg, prm, stream);
    if (rc == LM_OK) rc = lm_step_bin(h, stream);
    if (rc == LM_OK) rc = lm_step_interact_begin(h, r, pairs_out, cap, stream);
    if (rc == LM_OK) rc = lm_step_interact_end(h, stream);
    if (rc == LM_OK) rc = lm_step_finish(h, stream);
    if (rc != LM_OK) h->stage = 0;
    return rc;
}

int lm_phase_times(lm_handle h, float *ms_out)
{
    if (!h || !ms_out) return LM_EINVAL;
    if (!h->timed) return LM_ESTATE;
    LM_CUDA(cudaSetDevice(h->device));
    LM_CUDA(cudaEventSynchronize(h->ev_phase[5]));
    for (int k = 0; k < 5; ++k) LM_CUDA(cudaEventElapsedTime(ms_out + k, h->ev_phase[k], h->ev_phase[k + 1]));
    return LM_OK;
}

int lm_state_get(lm_handle h, float *lon_out, float *lat_out, int8_t *species_out, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    const int c = h->cur;
    { const int rcj = join_side(h, as_stream(stream)); if (rcj) return rcj; }
    LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], h->sp[c], h->id[c], (int)h->n, lon_out, lat_out, species_out,
                                 as_stream(stream), &h->launches, scatter_windows(h, (lon_out ? 1 : 0) + (lat_out ? 1 : 0)), (int)h->n));
    return LM_OK;
}

int lm_state_get_host(lm_handle h, float *lon_host, float *lat_host, int8_t *species_host, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = as_stream(stream);
    const int c = h->cur, k = h->stage_idx;
    const size_t n = (size_t)h->n;
    { const int rcj = join_side(h, s); if (rcj) return rcj; }
    // the staging buffers of slot k were last read by the copy issued two calls ago
    LM_CUDA(cudaStreamWaitEvent(s, h->ev_copied[k], 0));
    LM_CUDA(launch_scatter_by_id(h->lon[c], h->lat[c], h->sp[c], h->id[c], (int)n, lon_host ? h->stage_lon[k] : nullptr,
                                 lat_host ? h->stage_lat[k] : nullptr, species_host ? h->stage_sp[k] : nullptr, s,
                                 &h->launches, scatter_windows(h, (lon_host ? 1 : 0) + (lat_host ? 1 : 0)), (int)n));
    LM_CUDA(cudaEventRecord(h->ev_scatter[k], s));
    LM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_scatter[k], 0));
    if (lon_host) LM_CUDA(cudaMemcpyAsync(lon_host, h->stage_lon[k], n * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
    if (lat_host) LM_CUDA(cudaMemcpyAsync(lat_host, h->stage_lat[k], n * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
    if (species_host) LM_CUDA(cudaMemcpyAsync(species_host, h->stage_sp[k], n, cudaMemcpyDeviceToHost, h->copy_stream));
    LM_CUDA(cudaEventRecord(h->ev_copied[k], h->copy_stream));
    h->stage_idx = k ^ 1;
    return LM_OK;
}

int lm_record_next_step(lm_handle h, float *lon_host, float *lat_host, int8_t *species_host)
{
    if (!h) return LM_EINVAL;
    if (h->has_south || h->has_north) return LM_ESTATE;      // strips: ids travel with the record, see lm_state_view
    h->rec_lon_host = lon_host; h->rec_lat_host = lat_host; h->rec_sp_host = species_host;
    h->rec_ids_host = nullptr;
    h->rec_armed = lon_host || lat_host || species_host;
    return LM_OK;
}

int lm_record_next_step_ids(lm_handle h, int32_t *ids_host, float *lon_host, float *lat_host, int8_t *species_host)
{
    if (!h || !ids_host) return LM_EINVAL;
    h->rec_lon_host = lon_host; h->rec_lat_host = lat_host; h->rec_sp_host = species_host;
    h->rec_ids_host = ids_host;
    h->rec_armed = true;
    return LM_OK;
}

int64_t lm_record_count(lm_handle h) { return h ? h->rec_count : -1; }

int lm_host_copies_sync(lm_handle h)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaStreamSynchronize(h->copy_stream));
    return LM_OK;
}

int lm_state_view(lm_handle h, float **lon, float **lat, int8_t **species, int32_t **ids, int32_t **cell_start)
{
    if (!h) return LM_EINVAL;
    if (h->resolve_pending) {            // raw pointers escape: finish the side-stream work on the host side
        LM_CUDA(cudaEventSynchronize(h->ev_resolve_done));
        h->resolve_pending = false;
    }
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
    { const int rcj = join_side(h, as_stream(stream)); if (rcj) return rcj; }
    // the counters reach the host as stores of a small kernel into mapped pinned memory: a D2H copy would wait on the copy
    // engine behind the bulk copies of the per-step record (see xfer_counts_kernel)
    static_assert(sizeof(Counters) % 4 == 0, "Counters is copied word by word");
    LM_CUDA(launch_words_to_host(reinterpret_cast<const uint32_t *>(h->ctr), h->sticky, h->stats_dev, (int)(sizeof(Counters) / 4),
                                 as_stream(stream), &h->launches));
    LM_CUDA(cudaMemsetAsync(h->sticky, 0, sizeof(unsigned int), as_stream(stream)));      // reported once
    LM_CUDA(cudaStreamSynchronize(as_stream(stream)));
    Counters c;
    unsigned int sticky = 0;
    memcpy(&c, h->stats_host, sizeof(c));
    memcpy(&sticky, h->stats_host + sizeof(Counters) / 4, sizeof(sticky));
    h->ctr_reported = true;
    if (out) {
        out->n_pairs = (int64_t)c.n_pairs;
        out->n_out_of_bounds = (int64_t)c.n_oob;
        out->n_clamped = (int64_t)c.n_clamped;
        for (int k = 0; k < 4; ++k) out->species_count[k] = (int64_t)c.species[k];
        out->bbox[0] = dec_f(~c.bbox_enc[0]);
        out->bbox[1] = dec_f(c.bbox_enc[1]);
        out->bbox[2] = dec_f(~c.bbox_enc[2]);
        out->bbox[3] = dec_f(c.bbox_enc[3]);
        out->n_particles = h->n;
        out->n_moved_in = h->n_moved_in;
        out->n_moved_out = h->n_moved_out;
        out->n_misrouted = (int64_t)c.n_misrouted;
    }
    if (c.n_xfer_overflow) return LM_ENOSPC;                                       // migration / ghost buffers too small
    if (c.n_heavy_overflow) return LM_ENOSPC;                                      // heavy-unit queue too small: species invalid
    if (c.n_misrouted) return LM_ESTATE;   // particles held by a strip that does not own them (see bin.cu)
    if (h->emit_cap >= 0 && (int64_t)c.n_pairs > h->emit_cap) return LM_ENOSPC;   // pair list truncated
    if (h->rps_cap >= 0 && (int64_t)c.n_pairs > h->rps_cap) return LM_ENOSPC;     // RPS hand-off buffer too small: species invalid
    if (c.n_overflow) return LM_ENOSPC;
    // faults of EARLIER steps whose counters have been zeroed since (steps run without LM_STEP_STATS in between)
    if (sticky & (1u | 2u | 4u | 8u)) return LM_ENOSPC;
    if (sticky & 16u) return LM_ESTATE;
    if (sticky & 32u) return LM_ESTATE;                                            // a neighbour strip's message never came (peer_wait_kernel gave up)
    return LM_OK;
}

int64_t lm_launch_count(lm_handle h) { return h ? h->launches : 0; }

int lm_join(lm_handle h, void *stream)
{
    if (!h) return LM_EINVAL;
    LM_CUDA(cudaSetDevice(h->device));
    return join_side(h, as_stream(stream));
}

int lm_set_option(lm_handle h, int32_t option, int64_t value)
{
    if (!h) return LM_EINVAL;
    switch (option) {
        case LM_OPT_RESOLVE_UPL:
            if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return LM_EINVAL;
            h->resolve_upl = (int)value;
            return LM_OK;
        case LM_OPT_OVERLAP:
            if (value < 0 || value > 1) return LM_EINVAL;
            h->overlap = (int)value;
            return LM_OK;
        case LM_OPT_FIND_PATH:
            if (value < 0 || value > 1) return LM_EINVAL;
            h->find_path = (int)value;
            return LM_OK;
        case LM_OPT_RESOLVE_HEAVY_MIN:
            if (value < 0 || value > 0xffff) return LM_EINVAL;
            h->resolve_heavy_min = (int)value;
            return LM_OK;
        case LM_OPT_RESOLVE_BATCH:
            if (value != 1 && value != 4 && value != 8) return LM_EINVAL;
            h->resolve_batch = (int)value;
            return LM_OK;
        case LM_OPT_NORM:
            if (value != LM_NORM_1 && value != LM_NORM_2 && value != LM_NORM_INF) return LM_EINVAL;
            h->norm = (int)value;
            return LM_OK;
        case LM_OPT_RESOLVE_MODE: {
            if (value < 0 || value > 1) return LM_EINVAL;
            if (h->stage != 0) return LM_ESTATE;                 // not between the stages of a step
            if (value == 1 && !h->sp_snap) {
                LM_CUDA(cudaSetDevice(h->device));
                bool ok = dev_alloc(&h->sp_snap, h->max_particles);
                ok = ok && dev_alloc(&h->tile_scratch, 4 * h->max_particles + h->max_cells + 64);
                ok = ok && dev_alloc(&h->tile_scratch_used, 1);
                if (!ok) {
                    cudaFree(h->sp_snap); cudaFree(h->tile_scratch); cudaFree(h->tile_scratch_used);
                    h->sp_snap = nullptr; h->tile_scratch = nullptr; h->tile_scratch_used = nullptr;
                    return LM_ENOMEM;
                }
            }
            h->resolve_mode = (int)value;
            return LM_OK;
        }
        case LM_OPT_INTERACT_MODE:
            if (value < 0 || value > 2) return LM_EINVAL;
            if (h->stage != 0) return LM_ESTATE;                 // not between the stages of a step
            h->interact_mode = (int)value;
            return LM_OK;
        case LM_OPT_DRAW_BATCH:
            if (value < 0 || value > 32) return LM_EINVAL;
            h->draw_batch = (int)value;
            return LM_OK;
        case LM_OPT_TILE_CAP:
            if (value < 0 || value > 6144) return LM_EINVAL;
            h->tile_cap = (int)value;
            return LM_OK;
        case LM_OPT_TILE_REC_CAP:
            if (value < 0 || value > 16384) return LM_EINVAL;
            h->tile_rec_cap = (int)value;
            return LM_OK;
        case LM_OPT_HEAVY_MIN:
            if (value < 0 || value > (1ll << 40)) return LM_EINVAL;
            h->heavy_min = value;
            return LM_OK;
        case LM_OPT_PEER_WAIT_CYCLES:
            if (value < 1) return LM_EINVAL;
            h->peer_wait_cycles = (long long)value;
            return LM_OK;
        case LM_OPT_RECORD_DEBUG:
            if (value < 0 || value > 3) return LM_EINVAL;
            h->record_debug = (int)value;
            return LM_OK;
        case LM_OPT_SCATTER_PASSES:
            if (value < 0 || value > 64) return LM_EINVAL;
            h->scatter_passes = (int)value;
            return LM_OK;
        case LM_OPT_TILE_PATH:
            if (value < 0 || value > 1) return LM_EINVAL;
            h->tile_path = (int)value;
            return LM_OK;
        case LM_OPT_ADVECT_MODE:
            if (value < 0 || value > 1) return LM_EINVAL;
            h->advect_mode = (int)value;
            h->uv4_stale = true;
            return LM_OK;
        case LM_OPT_RESOLVE_TILE_SHAPE:
            if (value < 0 || value > 3) return LM_EINVAL;
            h->resolve_tile_shape = (int)value;
            return LM_OK;
        case LM_OPT_RESOLVE_MEGA_MIN:
            if (value < 0 || value > (1ll << 30)) return LM_EINVAL;
            h->resolve_mega_min = (int)value;
            return LM_OK;
        case LM_OPT_RESOLVE_TILE_SMEM:
            if (value < 1024 || value > 200 * 1024) return LM_EINVAL;
            h->resolve_tile_smem = (int)value;
            return LM_OK;
        default: return LM_EINVAL;
    }
}

// ---- analysis reductions on a snapshot (csrc/analysis.cu); handle-free like lm_pair_uniforms -----------------
int lm_pair_distance_hist(const float *lat, const float *lon, int64_t n, float radius_m, int32_t bins,
                          uint64_t *hist_out, void *stream)
{
    if (n < 0 || n >= (1ll << 24) || bins < 1 || bins > LM_PDH_MAX_BINS || !hist_out || !(radius_m > 0.f)) return LM_EINVAL;
    if (n > 0 && (!lat || !lon)) return LM_EINVAL;
    LM_CUDA(launch_pair_distance_hist(lat, lon, n, radius_m, bins, reinterpret_cast<unsigned long long *>(hist_out),
                                      as_stream(stream), nullptr));
    return LM_OK;
}

int lm_rasterize(const float *lon, const float *lat, const int8_t *species, int64_t n, double lon_min, double lon_max,
                 double lat_min, double lat_max, int32_t width, int32_t height, uint32_t *counts_out, int32_t *top_out,
                 void *stream)
{
    if (n < 0 || n >= (1ll << 31) || width < 1 || height < 1 || (int64_t)width * height >= (1ll << 30)) return LM_EINVAL;
    if (!(lon_max > lon_min) || !(lat_max > lat_min) || !counts_out || !top_out) return LM_EINVAL;
    if (n > 0 && (!lon || !lat)) return LM_EINVAL;
    LM_CUDA(launch_raster(lon, lat, species, n, lon_min, lon_max, lat_min, lat_max, width, height, counts_out, top_out,
                          as_stream(stream), nullptr));
    return LM_OK;
}

int lm_compose_frame(const uint32_t *counts, const int32_t *top, const int8_t *species, int32_t width, int32_t height,
                     int32_t mode, const uint8_t *palette_rgb, uint8_t *rgb_out, void *stream)
{
    if (width < 1 || height < 1 || (int64_t)width * height >= (1ll << 30) || !palette_rgb || !rgb_out) return LM_EINVAL;
    if ((mode != LM_FRAME_LAST_DRAWN && mode != LM_FRAME_PLURALITY) || (mode == LM_FRAME_LAST_DRAWN ? !top : !counts)) return LM_EINVAL;
    LM_CUDA(launch_compose(counts, top, species, width, height, mode, palette_rgb, rgb_out, as_stream(stream), nullptr));
    return LM_OK;
}

// ---- lossless delta packing of the position record (csrc/record.cu); handle-free ----------------------------
int lm_record_delta_pack(const float *prev_lon, const float *prev_lat, const float *lon, const float *lat, int64_t n,
                         int16_t *dlon_out, int16_t *dlat_out, uint32_t *esc_out, int64_t esc_cap, uint32_t *esc_count,
                         void *stream)
{
    if (n < 0 || n >= (1ll << 31) || esc_cap < 0 || !esc_count || (esc_cap > 0 && !esc_out)) return LM_EINVAL;
    if (n > 0 && (!prev_lon || !prev_lat || !lon || !lat || !dlon_out || !dlat_out)) return LM_EINVAL;
    LM_CUDA(launch_record_delta_pack(prev_lon, prev_lat, lon, lat, n, dlon_out, dlat_out, esc_out, esc_cap, esc_count,
                                     as_stream(stream)));
    return LM_OK;
}

int lm_record_delta_unpack_host(const float *prev_lon, const float *prev_lat, const int16_t *dlon, const int16_t *dlat,
                                const uint32_t *esc, int64_t n_esc, int64_t n, float *lon_out, float *lat_out,
                                int32_t n_threads)
{
    if (n < 0 || n >= (1ll << 31) || n_esc < 0 || (n_esc > 0 && !esc) || n_threads < 1 || n_threads > 1024) return LM_EINVAL;
    if (n > 0 && (!prev_lon || !prev_lat || !dlon || !dlat || !lon_out || !lat_out)) return LM_EINVAL;
    const int64_t marked = record_delta_unpack_host(prev_lon, prev_lat, dlon, dlat, esc, n_esc, n, lon_out, lat_out, n_threads);
    return marked == n_esc ? LM_OK : LM_EINVAL;
}

}  // extern "C"
