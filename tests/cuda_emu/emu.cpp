// TEST INFRASTRUCTURE ONLY -- the launcher of the CPU stand-in for the CUDA execution model (see cuda_runtime.h).
#include <cuda_runtime.h>

thread_local uint3 threadIdx, blockIdx, blockDim, gridDim;
namespace emu {
thread_local Warp *t_warp = nullptr;
thread_local Cta *t_cta = nullptr;
thread_local int t_lane = 0;
void *dyn_smem = nullptr;

void launch(unsigned int grid, unsigned int block, size_t smem, const std::function<void()> &body)
{
    if (grid == 0 || block == 0) return;
    void *raw = nullptr;
    if (posix_memalign(&raw, 128, smem ? smem : 128) != 0) abort();
    dyn_smem = raw;
    const unsigned int n_warps = (block + 31) / 32;
    // one set of `block` threads per launch; they run the CTAs one after the other (function-local static storage
    // stands in for shared memory, so two CTAs must never be in flight together), a barrier between CTAs
    Cta cta;
    pthread_barrier_init(&cta.bar, nullptr, block);
    cta.warps.resize(n_warps);
    for (unsigned int w = 0; w < n_warps; ++w)
        pthread_barrier_init(&cta.warps[w].bar, nullptr, std::min(32u, block - 32 * w));
    std::vector<std::thread> threads;
    threads.reserve(block);
    for (unsigned int t = 0; t < block; ++t) {
        threads.emplace_back([&, t]() {
            threadIdx = uint3{t, 0, 0};
            blockDim = uint3{block, 1, 1};
            gridDim = uint3{grid, 1, 1};
            t_cta = &cta;
            t_warp = &cta.warps[t / 32];
            t_lane = (int)(t % 32);
            for (unsigned int b = 0; b < grid; ++b) {
                blockIdx = uint3{b, 0, 0};
                body();
                pthread_barrier_wait(&cta.bar);            // every thread of CTA b is done before CTA b + 1 starts
            }
        });
    }
    for (auto &th : threads) th.join();
    for (unsigned int w = 0; w < n_warps; ++w) pthread_barrier_destroy(&cta.warps[w].bar);
    pthread_barrier_destroy(&cta.bar);
    free(raw);
    dyn_smem = nullptr;
}
}  // namespace emu
