// Latitude-strip decomposition: the device side of migration and of the one-row halo.
//
// The reference's only data-parallel split is contiguous particle tiles for ADVECTION
// (particle_advecter.py:38-66,143-148); its interaction phase is serial (interaction_simulator.py:82-117).
// Here both phases run on every GPU of a box: GPU g owns the particles of global cell rows
// [row0_g, row0_{g+1}) of ONE shared cell grid.  Per step and per strip boundary:
//
//   migration   particles whose new row left the strip are packed by bin_count (csrc/bin.cu) into
//               mig_send[dir]; the neighbour's arrivals are unpacked behind the local particles and
//               binned with them.
//   ghost row   the pair search uses a half stencil that looks north, so a strip needs the FIRST row
//               of the strip to its north: positions + ids + that row's cell table (ghost_send ->
//               ghost_recv), appended behind the owned particles as local row `rows_owned`.
//   species     row0 is even, so the strip boundary is crossed only by phases 6-8 of the canonical
//               cell-phase order (anchor row odd).  The northern strip sends the species of its first
//               row after ITS phases 0-5 (gsp), the southern strip resolves phases 6-8 against them
//               and sends them back (gret).  The northern strip does not touch that row in phases
//               6-8, so the result is exactly the single-GPU sequential order.
//
// All messages have a fixed capacity with the live count in a header, so no size negotiation (and no
// host synchronisation) is needed to post the transfers; the transport itself (NCCL send/recv between
// neighbours through torch.distributed, or device copies when several strips share a GPU) lives on the
// host side: lagrangian_microbes_b200/strips.py.
#include "lm_internal.cuh"

namespace lm {

__global__ void __launch_bounds__(256) unpack_arrivals_kernel(const int4 *__restrict__ rec, int n, int first,
                                                              float *__restrict__ lon, float *__restrict__ lat,
                                                              int8_t *__restrict__ sp, int32_t *__restrict__ id)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int4 r = rec[1 + k];
    lon[first + k] = __int_as_float(r.x);
    lat[first + k] = __int_as_float(r.y);
    id[first + k] = r.z;
    sp[first + k] = (int8_t)r.w;
}

cudaError_t launch_unpack_arrivals(lm_handle_s *h, int dir, int n_arrive, int first, float *lon, float *lat, int8_t *sp,
                                   int32_t *id, cudaStream_t s)
{
    if (n_arrive <= 0) return cudaSuccess;
    unpack_arrivals_kernel<<<(n_arrive + 255) / 256, 256, 0, s>>>(h->mig_recv[dir], n_arrive, first, lon, lat, sp, id);
    ++h->launches;
    return cudaGetLastError();
}

// ---- ghost row: positions, ids and cell table of the first owned row -> the strip to the south ----
__global__ void __launch_bounds__(256) ghost_pack_kernel(const float *__restrict__ lon, const float *__restrict__ lat,
                                                         const int32_t *__restrict__ id,
                                                         const int32_t *__restrict__ cell_start, int ncx, int row_cap,
                                                         int ghost_cap, int32_t *__restrict__ msg, Counters *ctr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_row = cell_start[ncx];
    const int n_send = min(n_row, ghost_cap);
    int32_t *cs_row = msg + GHOST_HDR;
    int32_t *m_lon = cs_row + row_cap + 1, *m_lat = m_lon + ghost_cap, *m_id = m_lat + ghost_cap;
    if (k == 0) {
        msg[0] = n_send;
        msg[1] = ncx;
        msg[2] = msg[3] = 0;
        if (n_row > ghost_cap) atomicAdd(&ctr->n_xfer_overflow, (unsigned int)(n_row - ghost_cap));
    }
    if (k <= ncx) cs_row[k] = min(cell_start[k], n_send);
    if (k < n_send) {
        m_lon[k] = __float_as_int(lon[k]);
        m_lat[k] = __float_as_int(lat[k]);
        m_id[k] = id[k];
    }
}

cudaError_t launch_ghost_pack(lm_handle_s *h, const float *lon, const float *lat, const int32_t *id, cudaStream_t s)
{
    const int work = (int)((h->ghost_cap > h->grid.ncx + 1) ? h->ghost_cap : h->grid.ncx + 1);
    // peer-memory exchange: packed straight into the southern neighbour's receive buffer (peer stores over NVLink)
    int32_t *msg = h->peer[0].connected ? h->peer[0].ghost_recv : h->ghost_send;
    ghost_pack_kernel<<<(work + 255) / 256, 256, 0, s>>>(lon, lat, id, h->cell_start, h->grid.ncx, (int)h->row_cap,
                                                          (int)h->ghost_cap, msg, h->ctr);
    ++h->launches;
    return cudaGetLastError();
}

// appended behind the owned particles; the ghost cells become local row `rows_owned` of the cell table
__global__ void __launch_bounds__(256) ghost_unpack_kernel(const int32_t *__restrict__ msg, int ncx, int row_cap,
                                                           int ghost_cap, int n_owned, int room,
                                                           float *__restrict__ lon, float *__restrict__ lat,
                                                           int32_t *__restrict__ id, int32_t *__restrict__ cs_ghost_row,
                                                           Counters *ctr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_msg = min(msg[0], ghost_cap);
    const int n_g = min(n_msg, room);
    const int32_t *cs_row = msg + GHOST_HDR;
    const int32_t *m_lon = cs_row + row_cap + 1, *m_lat = m_lon + ghost_cap, *m_id = m_lat + ghost_cap;
    if (k == 0 && (n_g < n_msg || msg[1] != ncx)) atomicAdd(&ctr->n_xfer_overflow, (unsigned int)max(n_msg - n_g, 1));
    if (k <= ncx) cs_ghost_row[k] = n_owned + min(cs_row[k], n_g);
    if (k < n_g) {
        lon[n_owned + k] = __int_as_float(m_lon[k]);
        lat[n_owned + k] = __int_as_float(m_lat[k]);
        id[n_owned + k] = m_id[k];
    }
}

cudaError_t launch_ghost_unpack(lm_handle_s *h, float *lon, float *lat, int32_t *id, int n_owned, cudaStream_t s)
{
    const int ncx = h->grid.ncx;
    const int work = (int)((h->ghost_cap > ncx + 1) ? h->ghost_cap : ncx + 1);
    const int64_t room = h->max_particles - n_owned;
    ghost_unpack_kernel<<<(work + 255) / 256, 256, 0, s>>>(h->ghost_recv, ncx, (int)h->row_cap, (int)h->ghost_cap,
                                                            n_owned, (int)(room < 0 ? 0 : room), lon, lat, id,
                                                            h->cell_start + (size_t)h->strip.rows_owned * ncx, h->ctr);
    ++h->launches;
    return cudaGetLastError();
}

// ---- species of the boundary row: north -> south after phase 5, south -> north after phase 8 ----
// count_at: cell_start entry holding the END of the range; base_at: entry holding its START (or null = 0)
__global__ void __launch_bounds__(256) species_copy_kernel(const int8_t *__restrict__ src, int8_t *__restrict__ dst,
                                                           const int32_t *__restrict__ end_at,
                                                           const int32_t *__restrict__ beg_at, int src_rel, int dst_rel,
                                                           int cap)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int beg = beg_at ? *beg_at : 0;
    const int n = min(*end_at - beg, cap);
    if (k >= n) return;
    dst[(dst_rel ? beg : 0) + k] = src[(src_rel ? beg : 0) + k];
}

static cudaError_t species_copy(lm_handle_s *h, const int8_t *src, int8_t *dst, const int32_t *end_at,
                                const int32_t *beg_at, bool src_rel, bool dst_rel, cudaStream_t s)
{
    const int cap = (int)h->ghost_cap;
    species_copy_kernel<<<(cap + 255) / 256, 256, 0, s>>>(src, dst, end_at, beg_at, src_rel ? 1 : 0, dst_rel ? 1 : 0, cap);
    ++h->launches;
    return cudaGetLastError();
}

cudaError_t launch_row0_species_pack(lm_handle_s *h, const int8_t *sp, cudaStream_t s)
{
    return species_copy(h, sp, h->peer[0].connected ? h->peer[0].gsp_recv : h->gsp_send, h->cell_start + h->grid.ncx, nullptr, false,
                        false, s);
}

cudaError_t launch_row0_species_unpack(lm_handle_s *h, int8_t *sp, cudaStream_t s)
{
    return species_copy(h, h->gret_recv, sp, h->cell_start + h->grid.ncx, nullptr, false, false, s);
}

cudaError_t launch_ghost_species_unpack(lm_handle_s *h, int8_t *sp, int /*n_owned*/, cudaStream_t s)
{
    const int32_t *beg = h->cell_start + (size_t)h->strip.rows_owned * h->grid.ncx;
    return species_copy(h, h->gsp_recv, sp, beg + h->grid.ncx, beg, false, true, s);
}

cudaError_t launch_ghost_species_pack(lm_handle_s *h, const int8_t *sp, int /*n_owned*/, cudaStream_t s)
{
    const int32_t *beg = h->cell_start + (size_t)h->strip.rows_owned * h->grid.ncx;
    return species_copy(h, sp, h->peer[1].connected ? h->peer[1].gret_recv : h->gret_send, beg + h->grid.ncx, beg, true, false, s);
}

// ---- peer-memory exchange: flags and the migrants' copy ---------------------------------------------------------------------
// A message is complete at the neighbour when its flag holds the step's sequence number: the flag store follows the kernel that
// wrote the message on the same stream, behind a system-scope fence; the consumer spins on its own flag word (written by the
// peer over NVLink), then fences.
__global__ void peer_signal_kernel(unsigned int *flag, unsigned int seq)
{
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int *>(flag) = seq;
}

// The spin is bounded: a neighbour that died (or raised and left) would otherwise leave this GPU spinning for ever -- after
// about a minute of SM clocks the wait gives up, latches fault bit 5 (lm_sync_stats: LM_ESTATE) and every later wait of the
// handle returns at once, so the run ends with an error instead of a hung device.
__global__ void peer_wait_kernel(const unsigned int *flag, unsigned int seq, unsigned int *sticky, long long limit_cycles)
{
    if (*reinterpret_cast<volatile unsigned int *>(sticky) & 32u) return;
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<const volatile unsigned int *>(flag) - seq) < 0) {     // wrap-safe: flag >= seq
        if (clock64() - t0 > limit_cycles) {
            atomicOr(sticky, 32u);
            break;
        }
    }
    __threadfence_system();
}

// the live part of a migration message: [0].x records behind the header
__global__ void __launch_bounds__(256) peer_push_mig_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, int send_cap)
{
    const int n = 1 + min(max(src[0].x, 0), send_cap);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst[k] = src[k];
}

// The migration counts of a step, stored by the device straight into mapped pinned host memory.  A cudaMemcpyAsync of
// the same 16 bytes would queue on the D2H copy engine BEHIND the bulk copies of the per-step record (2.7 ms for a 12.5 M
// microbe strip) and with it stall the whole step: a store from an SM is a posted write over PCIe and passes them.
__global__ void xfer_counts_kernel(const int32_t *send0, const int32_t *send1, const int32_t *recv0, const int32_t *recv1,
                                   volatile int32_t *host_counts)
{
    if (threadIdx.x == 0) {
        host_counts[0] = send0 ? *send0 : 0;
        host_counts[1] = send1 ? *send1 : 0;
        host_counts[2] = recv0 ? *recv0 : 0;
        host_counts[3] = recv1 ? *recv1 : 0;
        __threadfence_system();
    }
}

__global__ void words_to_host_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ last, volatile uint32_t *host, int words)
{
    for (int k = threadIdx.x; k < words; k += blockDim.x) host[k] = src[k];
    if (threadIdx.x == 0 && last) host[words] = *last;
    __threadfence_system();
}

// ``words`` 32-bit words of src, then one word of ``last`` (optional), into mapped pinned host memory
cudaError_t launch_words_to_host(const uint32_t *src, const uint32_t *last, uint32_t *host_dev, int words, cudaStream_t s, int64_t *launches)
{
    words_to_host_kernel<<<1, 64, 0, s>>>(src, last, host_dev, words);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_xfer_counts(const void *send0, const void *send1, const void *recv0, const void *recv1, int32_t *host_counts_dev,
                               cudaStream_t s, int64_t *launches)
{
    xfer_counts_kernel<<<1, 32, 0, s>>>(static_cast<const int32_t *>(send0), static_cast<const int32_t *>(send1),
                                        static_cast<const int32_t *>(recv0), static_cast<const int32_t *>(recv1), host_counts_dev);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_peer_signal(unsigned int *flag, unsigned int seq, cudaStream_t s, int64_t *launches)
{
    peer_signal_kernel<<<1, 1, 0, s>>>(flag, seq);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_peer_wait(const unsigned int *flag, unsigned int seq, unsigned int *sticky, long long limit_cycles, cudaStream_t s,
                             int64_t *launches)
{
    peer_wait_kernel<<<1, 1, 0, s>>>(flag, seq, sticky, limit_cycles);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_peer_push_mig(const int4 *src, int4 *dst, int64_t send_cap, cudaStream_t s, int64_t *launches)
{
    peer_push_mig_kernel<<<64, 256, 0, s>>>(src, dst, (int)send_cap);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace lm
