// TEST INFRASTRUCTURE ONLY -- a race detector for the kernels: the emulated library (one std::thread per CUDA thread,
// see cuda_runtime.h) built with -fsanitize=thread and driven through the product's own C ABI.  ThreadSanitizer
// understands the pthread barriers behind __syncthreads / __syncwarp / the warp collectives and the __atomic builtins
// behind atomicAdd & co, so what it reports is exactly the class of bug a missing barrier is on a GPU: two CUDA threads
// touching the same shared / global location without a barrier or an atomic in between.
//
//   python tests/cuda_emu/emu_build.py --tsan && tests/cuda_emu/_build/tsan_driver        (a few minutes)
//
// Workload: a random cloud with knots (so the whole-warp and whole-CTA resolver paths run), pair search + RPS through
// lm_interact_rps with the nine-phase and the tiled resolver (shared-memory and scratch tiles; the two must agree),
// three fused lm_step calls on a synthetic velocity field, a two-strip staged step, the analysis kernels and the delta-packed record.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/lm_b200.h"

#define CHECK(x) do { int rc__ = (x); if (rc__ != 0) { fprintf(stderr, "%s -> %d\n", #x, rc__); return 1; } } while (0)

int main(int argc, char **argv)
{
    const bool quick = argc > 1 && strcmp(argv[1], "quick") == 0;      // the subset the CPU test suite runs (about a minute)
    std::mt19937 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const int n = quick ? 320 : 900;
    const double r = 0.01, h = r * (1.0 + 1.0 / 1048576.0);
    const int ncx = 70, ncy = 19;
    std::vector<float> lon(n), lat(n);
    std::vector<int8_t> sp0(n);
    for (int i = 0; i < n; ++i) { lon[i] = (float)(201.0 + ncx * h * U(rng)); lat[i] = (float)(32.0 + ncy * h * U(rng)); sp0[i] = (int8_t)(1 + (int)(3 * U(rng))); }
    int k = 0;
    for (int m : {quick ? 100 : 140, 60, 25, 25}) {         // knots: whole-CTA, whole-warp and long lane-walked units
        const double cx = (int)(ncx * U(rng)), cy = (int)(ncy * U(rng));
        for (int j = 0; j < m; ++j, ++k) { lon[k] = (float)(201.0 + h * (cx + U(rng))); lat[k] = (float)(32.0 + h * (cy + U(rng))); }
    }
    lm_grid grid = {201.0, 32.0, 1.0 / h, ncx, ncy};
    lm_rps_params prm = {0.55, 0.6, 0.9, 5, 17};
    std::vector<int8_t> ref;
    // mode, tile smem, mega_min, heavy_min, resolver batch, units per lane, find path
    const long long opts[8][7] = {{0, 32768, 0, 0, 4, 0, 0}, {1, 32768, 0, 0, 4, 0, 0}, {1, 1024, 0, 0, 4, 0, 0}, {1, 32768, 48, 0, 4, 0, 0},
                                  {0, 32768, 0, 8, 1, 2, 1}, {0, 32768, 0, 24, 8, 8, 0}, {1, 32768, 0, 8, 4, 0, 1}, {0, 32768, 0, 1, 4, 1, 0}};
    int n_opts = 0;
    for (auto &o : opts) {
        if (quick && (++n_opts == 3 || n_opts > 4)) continue;      // quick: nine phases, tiled, tiled with the whole-CTA path
        lm_handle hd = nullptr;
        CHECK(lm_create(&hd, 0, n, 1 << 14, 60 * n));
        CHECK(lm_set_grid(hd, &grid));
        CHECK(lm_set_option(hd, LM_OPT_INTERACT_MODE, 0));             // the round-1 pipeline and its resolvers
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_MODE, o[0]));
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_TILE_SMEM, o[1]));
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_MEGA_MIN, o[2]));
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_HEAVY_MIN, o[3]));
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_BATCH, o[4]));
        CHECK(lm_set_option(hd, LM_OPT_RESOLVE_UPL, o[5]));
        CHECK(lm_set_option(hd, LM_OPT_FIND_PATH, o[6]));
        std::vector<int8_t> sp(sp0);
        std::vector<int32_t> pairs(2 * 60 * n);
        CHECK(lm_interact_rps(hd, lon.data(), lat.data(), sp.data(), n, r, &prm, pairs.data(), 60 * n, nullptr, nullptr));
        lm_stats st;
        CHECK(lm_sync_stats(hd, &st, nullptr));
        printf("mode %lld smem %lld mega %lld: %lld pairs\n", o[0], o[1], o[2], (long long)st.n_pairs);
        if (ref.empty()) ref = sp;
        else if (memcmp(ref.data(), sp.data(), n) != 0) { fprintf(stderr, "resolvers disagree\n"); return 2; }
        CHECK(lm_destroy(hd));
    }

    // the fused tile kernel (LM_OPT_INTERACT_MODE = 1, the default): shared-memory tiles with the whole-warp and whole-CTA
    // paths, draws one lane at a time, tiles too full for shared memory (global-memory path); all must agree
    {
        std::vector<int8_t> ref1;
        const long long topts[3][2] = {{0, 0}, {0, 1}, {256, 32}};       // LM_OPT_TILE_CAP, LM_OPT_DRAW_BATCH
        int k_opt = 0;
        for (auto &o : topts) {
            if (quick && ++k_opt == 2) continue;
            lm_handle hd = nullptr;
            CHECK(lm_create(&hd, 0, n, 1 << 14, 60 * n));
            CHECK(lm_set_grid(hd, &grid));
            CHECK(lm_set_option(hd, LM_OPT_INTERACT_MODE, 1));
            CHECK(lm_set_option(hd, LM_OPT_TILE_CAP, o[0]));
            CHECK(lm_set_option(hd, LM_OPT_DRAW_BATCH, o[1]));
            std::vector<int8_t> sp(sp0);
            std::vector<int32_t> pairs(2 * 60 * n);
            CHECK(lm_interact_rps(hd, lon.data(), lat.data(), sp.data(), n, r, &prm, pairs.data(), 60 * n, nullptr, nullptr));
            lm_stats st;
            CHECK(lm_sync_stats(hd, &st, nullptr));
            printf("fused tile kernel, tile cap %lld, draw batch %lld: %lld pairs\n", o[0], o[1], (long long)st.n_pairs);
            if (ref1.empty()) ref1 = sp;
            else if (memcmp(ref1.data(), sp.data(), n) != 0) { fprintf(stderr, "fused tile kernel: settings disagree\n"); return 2; }
            CHECK(lm_destroy(hd));
        }
    }

    // the hybrid path (LM_OPT_INTERACT_MODE = 2, the default): round-1 pipeline for the light units, heavy units queued per phase
    // and resolved in rounds of matchings from shared memory by whole warps / the whole CTA; a low LM_OPT_HEAVY_MIN queues many
    {
        for (long long hmin : {0ll, 40ll}) {
            if (quick && hmin) continue;
            lm_handle hd = nullptr;
            CHECK(lm_create(&hd, 0, n, 1 << 14, 60 * n));
            CHECK(lm_set_grid(hd, &grid));
            CHECK(lm_set_option(hd, LM_OPT_HEAVY_MIN, hmin));
            std::vector<int8_t> sp(sp0);
            std::vector<int32_t> pairs(2 * 60 * n);
            CHECK(lm_interact_rps(hd, lon.data(), lat.data(), sp.data(), n, r, &prm, pairs.data(), 60 * n, nullptr, nullptr));
            lm_stats st;
            CHECK(lm_sync_stats(hd, &st, nullptr));
            printf("hybrid path, heavy_min %lld: %lld pairs\n", hmin, (long long)st.n_pairs);
            CHECK(lm_destroy(hd));
        }
    }

    // explicit-order resolver (mark / fire rounds with 64-bit atomicMin) on the pair list of a search, other norms
    if (!quick) {
        lm_handle hd = nullptr;
        CHECK(lm_create(&hd, 0, n, 1 << 14, 60 * n));
        CHECK(lm_set_grid(hd, &grid));
        std::vector<int32_t> pairs(2 * 60 * n);
        std::vector<long long> found(1);
        CHECK(lm_find_pairs(hd, lon.data(), lat.data(), n, r, pairs.data(), 60 * n, (int64_t *)found.data(), nullptr));
        lm_stats st;
        CHECK(lm_sync_stats(hd, &st, nullptr));
        std::vector<double> u(st.n_pairs);
        CHECK(lm_pair_uniforms(pairs.data(), st.n_pairs, 5, 17, u.data(), nullptr));
        std::vector<int8_t> sp(sp0);
        int32_t rounds = 0;
        CHECK(lm_resolve_rps(hd, pairs.data(), u.data(), st.n_pairs, sp.data(), n, 0.55, 0.6, 0.9, &rounds, nullptr));
        printf("explicit resolver: %lld pairs, %d rounds\n", (long long)st.n_pairs, rounds);
        for (long long norm : {LM_NORM_1, LM_NORM_INF}) {
            CHECK(lm_set_option(hd, LM_OPT_NORM, norm));
            CHECK(lm_find_pairs(hd, lon.data(), lat.data(), n, r, pairs.data(), 60 * n, nullptr, nullptr));
            CHECK(lm_sync_stats(hd, &st, nullptr));
        }
        CHECK(lm_destroy(hd));
    }

    // fused steps on a synthetic field (solid-body-like flow on a 12 x 10 grid, 3 time levels)
    const int T = 3, Y = 10, X = 12;
    std::vector<float> Uf(T * Y * X), Vf(T * Y * X), glon(X), glat(Y);
    for (int x = 0; x < X; ++x) glon[x] = 200.5f + 0.2f * x;
    for (int y = 0; y < Y; ++y) glat[y] = 31.5f + 0.2f * y;
    for (int t = 0; t < T; ++t) for (int y = 0; y < Y; ++y) for (int x = 0; x < X; ++x) {
        Uf[(t * Y + y) * X + x] = 0.2f * (float)(y - Y / 2) / Y + 0.02f * t;
        Vf[(t * Y + y) * X + x] = -0.2f * (float)(x - X / 2) / X;
    }
    for (int mode = 0; mode < 4; ++mode) {                                 // 0, 1: round-1 pipeline, nine phases / tiled; 2: fused tile kernel; 3: hybrid
        lm_handle hd = nullptr;
        CHECK(lm_create(&hd, 0, n, 1 << 14, 60 * n));
        CHECK(lm_set_field(hd, Uf.data(), Vf.data(), glon.data(), glat.data(), T, Y, X));
        lm_grid g2 = {200.9, 31.9, 1.0 / h, ncx + 20, ncy + 20};
        CHECK(lm_set_grid(hd, &g2));
        CHECK(lm_set_option(hd, LM_OPT_INTERACT_MODE, mode == 2 ? 1 : (mode == 3 ? 2 : 0)));
        CHECK(lm_set_option(hd, LM_OPT_ADVECT_MODE, mode >= 2));
        if (mode < 2) CHECK(lm_set_option(hd, LM_OPT_RESOLVE_MODE, mode));
        CHECK(lm_state_set(hd, lon.data(), lat.data(), sp0.data(), nullptr, n, nullptr));
        std::vector<int32_t> pairs(2 * 60 * n);
        for (int step = 0; step < (quick ? 1 : 3); ++step) {
            lm_stage_times stt = {{0, 0, 0, 0}, {1, 1, 1, 1}, {0.01f * step, 0.01f * step + 0.005f, 0.01f * step + 0.005f, 0.01f * step + 0.01f}};
            lm_rps_params p2 = {0.55, 0.55, 0.55, 3, (uint64_t)step};
            CHECK(lm_step(hd, LM_STEP_ADVECT | LM_STEP_INTERACT | LM_STEP_EMIT_PAIRS | LM_STEP_STATS | (step ? LM_STEP_DIFFUSE : 0), &stt, 3600.f,
                          0.001, r, &p2, pairs.data(), 60 * n, nullptr));
            lm_stats st;
            CHECK(lm_sync_stats(hd, &st, nullptr));
            printf("lm_step mode %d step %d: %lld pairs, %lld out of bounds\n", mode, step, (long long)st.n_pairs, (long long)st.n_out_of_bounds);
        }
        std::vector<float> a(n), b(n);
        std::vector<int8_t> c(n);
        CHECK(lm_state_get(hd, a.data(), b.data(), c.data(), nullptr));
        // the per-step record pipeline: next step scatters + copies its record, then the host-copy variant
        CHECK(lm_record_next_step(hd, a.data(), b.data(), c.data()));
        {
            lm_stage_times stt = {{0, 0, 0, 0}, {1, 1, 1, 1}, {0.03f, 0.035f, 0.035f, 0.04f}};
            lm_rps_params p2 = {0.55, 0.55, 0.55, 3, 3};
            CHECK(lm_step(hd, LM_STEP_ADVECT | LM_STEP_INTERACT, &stt, 3600.f, 0.0, r, &p2, nullptr, 0, nullptr));
        }
        CHECK(lm_host_copies_sync(hd));
        CHECK(lm_state_get_host(hd, a.data(), b.data(), c.data(), nullptr));
        CHECK(lm_host_copies_sync(hd));
        CHECK(lm_destroy(hd));
    }

    // two latitude strips: routing passes, then two staged steps with the exchange buffers copied between the neighbours
    if (!quick) {
        lm_grid g2 = {200.9, 31.9, 1.0 / h, ncx + 20, 40};
        lm_handle hs[2] = {nullptr, nullptr};
        lm_strip_buffers bf[2];
        std::vector<int32_t> ids(n);
        for (int i = 0; i < n; ++i) ids[i] = i;
        std::vector<int32_t> pairs[2] = {std::vector<int32_t>(2 * 60 * n), std::vector<int32_t>(2 * 60 * n)};
        for (int s = 0; s < 2; ++s) {
            CHECK(lm_create(&hs[s], 0, n + 512, 1 << 14, 60 * n));
            CHECK(lm_set_field(hs[s], Uf.data(), Vf.data(), glon.data(), glat.data(), T, Y, X));
            CHECK(lm_strip_alloc(hs[s], 1024, 1024, g2.ncx + 8));
            CHECK(lm_set_grid(hs[s], &g2));
            // (fused tile kernel on both strips: the default)
            lm_strip st = {s ? 16 : 0, s ? 24 : 16, s, 1 - s};
            CHECK(lm_set_strip(hs[s], &st));
            CHECK(lm_strip_buffers_get(hs[s], &bf[s]));
            const int first = s ? n / 2 : 0, cnt = s ? n - n / 2 : n / 2;          // contiguous tiles, wherever the microbes are
            CHECK(lm_state_set(hs[s], lon.data() + first, lat.data() + first, sp0.data() + first, ids.data() + first, cnt, nullptr));
        }
        auto staged = [&](int flags, const lm_stage_times *stt, const lm_rps_params *p2) -> int {
            for (int s = 0; s < 2; ++s) CHECK(lm_step_move(hs[s], flags, stt, 3600.f, 0.0, p2, nullptr));
            memcpy(bf[0].mig_recv[1], bf[1].mig_send[0], (size_t)bf[1].mig_bytes);
            memcpy(bf[1].mig_recv[0], bf[0].mig_send[1], (size_t)bf[0].mig_bytes);
            for (int s = 0; s < 2; ++s) CHECK(lm_step_bin(hs[s], nullptr));
            if (flags & LM_STEP_INTERACT) memcpy(bf[0].ghost_recv, bf[1].ghost_send, (size_t)bf[1].ghost_bytes);
            for (int s = 0; s < 2; ++s) CHECK(lm_step_interact_begin(hs[s], r, pairs[s].data(), 60 * n, nullptr));
            if (flags & LM_STEP_INTERACT) memcpy(bf[0].gsp_recv, bf[1].gsp_send, (size_t)bf[1].species_bytes);
            for (int s = 0; s < 2; ++s) CHECK(lm_step_interact_end(hs[s], nullptr));
            if (flags & LM_STEP_INTERACT) memcpy(bf[1].gret_recv, bf[0].gret_send, (size_t)bf[0].species_bytes);
            for (int s = 0; s < 2; ++s) CHECK(lm_step_finish(hs[s], nullptr));
            return 0;
        };
        lm_rps_params p0 = {0.55, 0.55, 0.55, 3, 0};
        for (int pass = 0; pass < 40; ++pass) {
            for (int s = 0; s < 2; ++s) { lm_strip st = {s ? 16 : 0, s ? 24 : 16, s, 1 - s}; CHECK(lm_set_strip(hs[s], &st)); }
            if (staged(0, nullptr, &p0)) return 1;
            lm_stats a, b;
            lm_sync_stats(hs[0], &a, nullptr); lm_sync_stats(hs[1], &b, nullptr);
            if (a.n_misrouted + b.n_misrouted == 0) break;
        }
        for (int step = 0; step < 2; ++step) {
            lm_stage_times stt = {{0, 0, 0, 0}, {1, 1, 1, 1}, {0.01f * step, 0.01f * step + 0.005f, 0.01f * step + 0.005f, 0.01f * step + 0.01f}};
            lm_rps_params p2 = {0.55, 0.55, 0.55, 3, (uint64_t)step};
            if (staged(LM_STEP_ADVECT | LM_STEP_INTERACT | LM_STEP_EMIT_PAIRS | LM_STEP_STATS, &stt, &p2)) return 1;
            lm_stats a, b;
            CHECK(lm_sync_stats(hs[0], &a, nullptr)); CHECK(lm_sync_stats(hs[1], &b, nullptr));
            printf("strips step %d: %lld + %lld pairs, %lld + %lld microbes, %lld moved\n", step, (long long)a.n_pairs, (long long)b.n_pairs,
                   (long long)a.n_particles, (long long)b.n_particles, (long long)(a.n_moved_in + b.n_moved_in));
        }
        for (int s = 0; s < 2; ++s) CHECK(lm_destroy(hs[s]));
    }

    // analysis kernels
    std::vector<uint64_t> hist(72);
    CHECK(lm_pair_distance_hist(lat.data(), lon.data(), n, 6371.228e3f, 70, hist.data(), nullptr));
    unsigned long long tot = 0;
    for (auto v : hist) tot += v;
    if (tot != (unsigned long long)n * (n - 1) / 2) { fprintf(stderr, "histogram total\n"); return 3; }
    std::vector<uint32_t> counts(3 * 40 * 24);
    std::vector<int32_t> top(40 * 24);
    std::vector<uint8_t> rgb(3 * 40 * 24);
    const uint8_t pal[12] = {255, 255, 255, 255, 0, 0, 50, 205, 50, 0, 0, 255};
    CHECK(lm_rasterize(lon.data(), lat.data(), sp0.data(), n, 201.0, 201.7, 32.0, 32.2, 40, 24, counts.data(), top.data(), nullptr));
    CHECK(lm_compose_frame(counts.data(), top.data(), sp0.data(), 40, 24, LM_FRAME_LAST_DRAWN, pal, rgb.data(), nullptr));
    // the delta-packed record: pack (vector and scalar paths, a small escape list that overflows), decode, compare
    {
        std::vector<float> cl(lon), ca(lat);
        for (int i = 0; i < n; ++i) { cl[i] += (float)(0.01 * (U(rng) - 0.5)); ca[i] += (float)(0.01 * (U(rng) - 0.5)); }
        for (int i = 0; i < 12; ++i) cl[7 * i] += 3.0f;                    // far jumps: escapes from several warps at once
        std::vector<int16_t> dl(n + 1), da(n + 1);
        for (int off = 0; off < 2; ++off) {                                // off = 1: int16 outputs not 8-byte aligned -> scalar kernel
            std::vector<uint32_t> esc(2 * 64);
            uint32_t n_esc = 0;
            CHECK(lm_record_delta_pack(lon.data(), lat.data(), cl.data(), ca.data(), n, dl.data() + off, da.data() + off, esc.data(), 64, &n_esc, nullptr));
            if (n_esc < 12 || n_esc > 64) { fprintf(stderr, "escapes %u\n", n_esc); return 4; }
            std::vector<float> ol(n), oa(n);
            CHECK(lm_record_delta_unpack_host(lon.data(), lat.data(), dl.data() + off, da.data() + off, esc.data(), n_esc, n, ol.data(), oa.data(), 2));
            if (memcmp(ol.data(), cl.data(), 4 * n) || memcmp(oa.data(), ca.data(), 4 * n)) { fprintf(stderr, "delta record round trip\n"); return 4; }
            uint32_t over = 0;
            CHECK(lm_record_delta_pack(lon.data(), lat.data(), cl.data(), ca.data(), n, dl.data() + off, da.data() + off, esc.data(), 4, &over, nullptr));
            if (over != n_esc) { fprintf(stderr, "overflow count %u vs %u\n", over, n_esc); return 4; }
        }
        printf("delta record: round trip exact\n");
    }
    printf("done\n");
    return 0;
}
