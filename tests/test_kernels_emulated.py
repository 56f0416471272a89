"""The CUDA kernels of csrc/pairs.cu and csrc/analysis.cu EXECUTED on the CPU (tests/cuda_emu: one std::thread per CUDA
thread, barriers for __syncthreads / __syncwarp / the warp collectives) and held against the oracle.

Why: kernels written when no GPU was at hand (the tiled resolver, the snapshot analyses) could otherwise only be
compiled.  The emulator is itself checked by running the kernels that ARE verified on hardware -- the pair search and the
nine-phase resolver -- through it: they must reproduce cKDTree's pair set and the reference's sequential rule
(interactions.py:13-40 in canonical cell-phase order) here exactly as they do on the B200.  Sizes are tiny: a warp
shuffle costs two pthread barriers."""
import ctypes
import os
import sys

import numpy as np
import pytest

from oracle import analysis as oa
from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps
from oracle.pairs import cell_index

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))

P = (0.55, 0.6, 0.9)
R = 0.01
H = R * (1 + 2.0 ** -20)


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = ctypes.CDLL(emu_build.build())
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_double
    L.emu_interact.restype = i64
    L.emu_interact.argtypes = [vp, vp, vp, vp, vp, i32, i32, dbl, dbl, dbl, i32, i32, i32, i32, i32, dbl, dbl, dbl, dbl, u64, u64,
                               i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, i64, vp, vp, vp]
    L.emu_pair_distance_hist.restype = i32
    L.emu_pair_distance_hist.argtypes = [vp, vp, i64, ctypes.c_float, i32, vp]
    L.emu_raster.restype = i32
    L.emu_raster.argtypes = [vp, vp, vp, i64, dbl, dbl, dbl, dbl, i32, i32, vp, vp, i32, vp, vp]
    return L


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


class Cloud:
    def __init__(self, seed, ncx, ncy, n, knots=0, knot_size=(10, 31)):
        rng = np.random.default_rng(seed)
        self.grid = dict(x0=200.0, y0=30.0, inv_h=1.0 / H, ncx=ncx, ncy=ncy)
        lon = 200.0 + ncx * H * rng.random(n)
        lat = 30.0 + ncy * H * rng.random(n)
        k = 0
        for _ in range(knots):
            m = int(rng.integers(*knot_size))
            lon[k:k + m] = 200.0 + H * (int(rng.integers(0, ncx)) + rng.random(m))
            lat[k:k + m] = 30.0 + H * (int(rng.integers(0, ncy)) + rng.random(m))
            k += m
        self.lon, self.lat = lon.astype(np.float32), lat.astype(np.float32)
        self.sp0 = rng.integers(1, 4, n).astype(np.int8)
        self.sp0[::37] = 0                                  # species outside {1, 2, 3}
        self.n = n
        cx = cell_index(self.lon, 200.0, 1.0 / H, ncx)
        cy = cell_index(self.lat, 30.0, 1.0 / H, ncy)
        key = cy.astype(np.int64) * ncx + cx
        self.ids = np.lexsort((np.arange(n), key)).astype(np.int32)            # storage slot -> id: (cell, id) order
        self.cell_start = np.zeros(ncx * ncy + 1, dtype=np.int32)
        np.cumsum(np.bincount(key, minlength=ncx * ncy), out=self.cell_start[1:])
        self.pairs = opairs.query_pairs_reference_array(self.lon, self.lat, R)
        self.order, self.phase = orps.cell_phase_order(self.pairs, self.lon, self.lat, self.grid)
        self.u = philox.pair_uniforms(self.order[:, 0], self.order[:, 1], 17, 5)      # step 17, seed 5

    def oracle(self, sp, first, last):
        sel = (self.phase >= first) & (self.phase <= last)
        out, _ = orps.rps_sequential_c(sp.copy(), self.order[sel], self.u[sel], *P)
        return out

    def run(self, emu, sp, mode, first=0, last=8, tile_smem=32768, heavy_min=0, batch=4, upl=0, find_path=0, want_pairs=False,
            mega_min=0):
        g = self.grid
        lon_s, lat_s = np.ascontiguousarray(self.lon[self.ids]), np.ascontiguousarray(self.lat[self.ids])
        sp_s = np.ascontiguousarray(sp[self.ids])
        cap = self.pairs.shape[0] + 8
        pairs_out = np.full((cap, 2), -1, dtype=np.int32) if want_pairs else None
        ret = emu.emu_interact(_ptr(lon_s), _ptr(lat_s), _ptr(self.ids), _ptr(self.cell_start), _ptr(sp_s), self.n, self.n,
                               g["x0"], g["y0"], g["inv_h"], g["ncx"], g["ncy"], 0, g["ncy"], g["ncy"], R, *P, 5, 17,
                               mode, first, last, tile_smem, heavy_min, batch, upl, find_path, mega_min, _ptr(pairs_out), cap,
                               cap + 64,
                               None, None, None)
        assert ret >= 0
        found, launches = ret & ((1 << 48) - 1), ret >> 48
        assert found == self.pairs.shape[0], "pairs found"
        out = np.empty_like(sp)
        out[self.ids] = sp_s
        if want_pairs:
            return out, launches, opairs.sort_pairs(pairs_out[:found])
        return out, launches


@pytest.fixture(scope="module")
def cloud():
    return Cloud(1, 74, 19, 1800, knots=10)                  # two tiles across, two up, ragged; knots of 10-30 microbes


def test_emulator_reproduces_the_hardware_verified_kernels(emu, cloud):
    """Control: pair search + nine-phase resolver (verified on the B200) through the emulator."""
    got, launches, pairs = cloud.run(emu, cloud.sp0, mode=0, want_pairs=True)
    assert np.array_equal(pairs, cloud.pairs), "pair set differs from cKDTree.query_pairs"
    want = cloud.oracle(cloud.sp0, 0, 8)
    assert int((want != cloud.sp0).sum()) > 100
    assert np.array_equal(got, want), "%d species differ" % int((got != want).sum())
    assert launches == 10                                     # 1 search + 9 phases
    got1, _ = cloud.run(emu, cloud.sp0, mode=0, batch=1, heavy_min=8, upl=2, find_path=1)     # other code paths of the same kernels
    assert np.array_equal(got1, want)


@pytest.mark.parametrize("tile_smem,heavy_min", [(32768, 0), (1024, 0), (32768, 8)])
def test_tiled_resolver_executed(emu, cloud, tile_smem, heavy_min):
    """resolve_tiled_kernel (LM_OPT_RESOLVE_MODE = 1): shared-memory tiles, scratch tiles (1 KB limit), whole-warp path."""
    want = cloud.oracle(cloud.sp0, 0, 8)
    got, launches = cloud.run(emu, cloud.sp0, mode=1, tile_smem=tile_smem, heavy_min=heavy_min)
    assert np.array_equal(got, want), "%d species differ" % int((got != want).sum())
    assert launches == 2                                      # 1 search + 1 tiled launch


def test_tiled_resolver_phase_ranges(emu, cloud):
    """The two ranges a strip launches around its halo exchange: 0-5, then 6-8 on the result."""
    mid_want = cloud.oracle(cloud.sp0, 0, 5)
    mid, _ = cloud.run(emu, cloud.sp0, mode=1, first=0, last=5)
    assert np.array_equal(mid, mid_want)
    end, _ = cloud.run(emu, mid, mode=1, first=6, last=8)
    assert np.array_equal(end, cloud.oracle(cloud.sp0, 0, 8))


def test_tiled_resolver_crowded_cells(emu):
    """~12 microbes per cell: every unit is long, cells continue across 32-particle chunks, the relative dense limit
    keeps the lane walk, a low absolute limit forces whole warps."""
    c = Cloud(2, 20, 6, 1400, knots=3, knot_size=(40, 60))
    want = c.oracle(c.sp0, 0, 8)
    assert (np.diff(c.cell_start) > 32).any()
    for heavy_min in (0, 16):
        got, _ = c.run(emu, c.sp0, mode=1, heavy_min=heavy_min)
        assert np.array_equal(got, want), "heavy_min %d: %d species differ" % (heavy_min, int((got != want).sum()))


def test_tiled_resolver_knots_on_the_whole_cta(emu):
    """Knots of 150-250 microbes in single cells (what a long run produces): units of 10^4 pairs and more go to the
    whole-CTA path (256 partners of an anchor per scan); with a low limit every unit above 48 pairs does."""
    c = Cloud(5, 24, 10, 900, knots=2, knot_size=(150, 251))
    want = c.oracle(c.sp0, 0, 8)
    assert np.diff(c.cell_start).max() >= 150 and c.pairs.shape[0] > 20000
    for mega_min, smem in ((0, 32768), (48, 1024)):
        got, _ = c.run(emu, c.sp0, mode=1, mega_min=mega_min, tile_smem=smem)
        assert np.array_equal(got, want), "mega_min %d: %d species differ" % (mega_min, int((got != want).sum()))
    got0, _ = c.run(emu, c.sp0, mode=0)                     # the nine-phase resolver on the same knots (control)
    assert np.array_equal(got0, want)


def test_pair_distance_histogram_executed(emu):
    rng = np.random.default_rng(3)
    for kind, n in (("patch", 700), ("clustered", 600), ("global", 1300)):
        if kind == "patch":
            lat, lon = 25 + 10 * rng.random(n), 205 + 10 * rng.random(n)
        elif kind == "clustered":
            lat = 30 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
            lon = 210 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
            lat[:20] = lat[20:40]; lon[:20] = lon[20:40]
        else:
            lat, lon = -80 + 160 * rng.random(n), 360 * rng.random(n)
        lat, lon = lat.astype(np.float32), lon.astype(np.float32)
        hist = np.full(72, 99, dtype=np.uint64)                  # the call zeroes its output
        assert emu.emu_pair_distance_hist(_ptr(lat), _ptr(lon), n, 6371.228e3, 70, _ptr(hist)) == 0
        h = hist.astype(np.int64)
        lo, up = oa.pdh_bounds(lat, lon, bins=70)
        assert h.sum() == n * (n - 1) // 2
        assert np.all(lo <= h) and np.all(h <= up), (kind, h - lo, up - h)
    hist = np.zeros(4, dtype=np.uint64)                          # bins = 2: nearly everything beyond the last bin
    assert emu.emu_pair_distance_hist(_ptr(lat), _ptr(lon), n, 6371.228e3, 2, _ptr(hist)) == 0
    assert hist.sum() == n * (n - 1) // 2 and hist[3] > 0


@pytest.mark.parametrize("mode", [0, 1])
def test_rasteriser_executed(emu, mode):
    rng = np.random.default_rng(4)
    n = 5000
    lon = (204.0 + 12.0 * rng.random(n)).astype(np.float32)
    lat = (24.0 + 12.0 * rng.random(n)).astype(np.float32)
    lon[:5] = [205.0, 215.0, np.nextafter(np.float32(215.0), np.float32(0)), 210.0, np.nan]
    lat[:5] = [25.0, 30.0, 30.0, 35.0, 30.0]
    sp = rng.integers(0, 5, n).astype(np.int8)
    w, h = 40, 24
    pal = np.array([[255, 255, 255], [255, 0, 0], [50, 205, 50], [0, 0, 255]], dtype=np.uint8)
    counts = np.full((3, h, w), 7, dtype=np.uint32)
    top = np.full((h, w), 7, dtype=np.int32)
    rgb = np.zeros((h, w, 3), dtype=np.uint8)
    assert emu.emu_raster(_ptr(lon), _ptr(lat), _ptr(sp), n, 205.0, 215.0, 25.0, 35.0, w, h, _ptr(counts), _ptr(top), mode,
                          _ptr(pal), _ptr(rgb)) == 0
    want_counts, want_top = oa.raster_reference(lon, lat, sp, 205.0, 215.0, 25.0, 35.0, w, h)
    assert np.array_equal(counts, want_counts) and np.array_equal(top, want_top)
    assert np.array_equal(rgb, oa.compose_reference(want_counts, want_top, sp, pal, mode))


@pytest.mark.parametrize("seed,ncx,ncy,n", [(11, 5, 3, 120), (12, 64, 16, 900), (13, 65, 17, 900), (14, 141, 5, 1200),
                                             (15, 7, 40, 700)])
def test_tiled_resolver_other_grids(emu, seed, ncx, ncy, n):
    """Grids smaller than a tile, exactly one tile, one cell more than a tile, three tiles across, three up."""
    c = Cloud(seed, ncx, ncy, n, knots=4)
    want = c.oracle(c.sp0, 0, 8)
    got, _ = c.run(emu, c.sp0, mode=1)
    assert np.array_equal(got, want), "%d species differ" % int((got != want).sum())


@pytest.mark.parametrize("mode", [0, 1])
def test_strip_geometry(emu, mode):
    """One latitude strip of a larger domain as the library holds it: rows [4, 12) owned, row 12 appended
    as a ghost row whose microbes are partners only; phases 0-5, then 6-8, against the sequential order restricted to
    the pairs this strip owns (anchor in an owned row).  mode 0 is the hardware-verified control."""
    c = Cloud(21, 70, 19, 1700, knots=6)
    g = c.grid
    row0, rows_owned = 4, 8
    cy = cell_index(c.lat, g["y0"], g["inv_h"], g["ncy"])
    cx = cell_index(c.lon, g["x0"], g["inv_h"], g["ncx"])
    local = (cy >= row0) & (cy <= row0 + rows_owned)                    # owned rows + the ghost row
    idx = np.nonzero(local)[0]
    key = (cy[idx] - row0).astype(np.int64) * g["ncx"] + cx[idx]
    o = np.lexsort((idx, key))
    ids = idx[o].astype(np.int32)                                       # storage slot -> global id; ghosts come last
    n_all = ids.size
    n_owned = int((cy[ids] < row0 + rows_owned).sum())
    rows_local = rows_owned + 1
    cell_start = np.zeros(g["ncx"] * rows_local + 1, dtype=np.int32)
    np.cumsum(np.bincount(key, minlength=g["ncx"] * rows_local), out=cell_start[1:])
    anchor_row = np.minimum(cy[c.order[:, 0]], cy[c.order[:, 1]])
    mine = (anchor_row >= row0) & (anchor_row < row0 + rows_owned)
    order, phase, u = c.order[mine], c.phase[mine], c.u[mine]
    sp = c.sp0.copy()
    lon_s, lat_s = np.ascontiguousarray(c.lon[ids]), np.ascontiguousarray(c.lat[ids])
    for first, last in ((0, 5), (6, 8)):
        sel = (phase >= first) & (phase <= last)
        want, _ = orps.rps_sequential_c(sp.copy(), order[sel], u[sel], *P)
        sp_s = np.ascontiguousarray(sp[ids])
        cap = order.shape[0] + 64
        ret = emu.emu_interact(_ptr(lon_s), _ptr(lat_s), _ptr(ids), _ptr(cell_start), _ptr(sp_s), n_owned, n_all,
                               g["x0"], g["y0"], g["inv_h"], g["ncx"], g["ncy"], row0, rows_owned, rows_local, R, *P, 5, 17,
                               mode, first, last, 32768, 0, 4, 0, 0, 0, None, 0, cap, None, None, None)
        assert ret >= 0 and (ret & ((1 << 48) - 1)) == order.shape[0]
        got = sp.copy()
        got[ids] = sp_s
        assert np.array_equal(got, want), "phases %d-%d: %d species differ" % (first, last, int((got != want).sum()))
        sp = want
    assert (sp[ids[n_owned:]] != c.sp0[ids[n_owned:]]).any()            # the ghost row's microbes took part


# ----------------------------------------------------------------------------------------------------------------------
# The whole C ABI (csrc/api.cu on top of every kernel), executed: lm_create .. lm_step .. lm_state_get on host arrays.
@pytest.fixture(scope="module")
def abi():
    import emu_build
    from lagrangian_microbes_b200 import _lib
    return _lib.declare(ctypes.CDLL(emu_build.build()))


def _small_field():
    from conftest import golden
    from oracle import rk4 as ork4
    g = golden("rk4_small.npz")
    return ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"])


@pytest.mark.parametrize("interact_mode,resolve_mode,stats_every_step", [(2, 0, True), (2, 0, False), (1, 0, True),
                                                                         (0, 0, True), (0, 1, False)])
def test_fused_steps_through_the_c_abi(abi, interact_mode, resolve_mode, stats_every_step):
    """Four lm_step calls (RK4 advection, binning, pair search, RPS, stats) on 1,200 microbes, step by step against the
    oracle: positions vs the RK4 restatement from identical inputs, pairs vs cKDTree on the library's positions,
    species vs the reference rule in canonical order -- tests/test_gpu_parity.py::test_fused_simulation_matches_oracle_loop
    in miniature, on the emulator, with the fused tile kernel (the default) and with the round-1 pipeline (nine-phase and
    tiled resolver)."""
    from lagrangian_microbes_b200 import _lib
    from lagrangian_microbes_b200.engine import make_grid
    from lagrangian_microbes_b200.particle_advecter import StageClock
    from oracle import rk4 as ork4
    L = abi
    fs = _small_field()
    n, r, p, seed, dt, n_steps = 1200, 0.01, (0.55, 0.6, 0.9), 5, 3600.0, 4
    rng = np.random.default_rng(8)
    lon = (201.5 + 0.25 * rng.random(n)).astype(np.float32)
    lat = (32.5 + 0.16 * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    h = ctypes.c_void_p()
    max_cells, cap = 1 << 16, 40 * n
    assert L.lm_create(ctypes.byref(h), 0, n, max_cells, cap) == 0
    try:
        u, v = np.ascontiguousarray(fs.u), np.ascontiguousarray(fs.v)
        glon, glat = np.ascontiguousarray(fs.lon), np.ascontiguousarray(fs.lat)
        assert L.lm_set_field(h, _ptr(u), _ptr(v), _ptr(glon), _ptr(glat), *u.shape) == 0
        grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, n, max_cells, margin=0.1)
        assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_INTERACT_MODE, interact_mode) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_MODE, resolve_mode) == 0
        assert L.lm_state_set(h, _ptr(lon), _ptr(lat), _ptr(sp0), None, n, None) == 0
        clock = StageClock(fs.time)
        pairs = np.zeros((cap, 2), dtype=np.int32)
        lon_prev, lat_prev, sp_ref = lon.copy(), lat.copy(), sp0.copy()
        gl, ga, gs = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int8)
        t, ti, total, launches0 = 0.0, 0, 0, L.lm_launch_count(h)
        flags = _lib.LM_STEP_ADVECT | _lib.LM_STEP_INTERACT | _lib.LM_STEP_EMIT_PAIRS
        for step in range(n_steps):
            st_times = clock.next_step(dt)
            prm = _lib.RpsParams(*p, seed, step)
            f = flags | (_lib.LM_STEP_STATS if stats_every_step else 0)
            assert L.lm_step(h, f, ctypes.byref(st_times), dt, 0.0, r, ctypes.byref(prm), _ptr(pairs), cap, None) == 0
            stats = _lib.Stats()
            assert L.lm_sync_stats(h, ctypes.byref(stats), None) == 0
            assert L.lm_state_get(h, _ptr(gl), _ptr(ga), _ptr(gs), None) == 0
            a32, b32, ti_new, oob = ork4.rk4_step_f32(fs, lon_prev, lat_prev, t, dt, ti)
            assert oob == 0 and stats.n_out_of_bounds == 0
            assert np.max(np.abs(gl - a32) / np.abs(a32)) < 1e-6 and np.max(np.abs(ga - b32) / np.abs(b32)) < 1e-6
            want_pairs = opairs.query_pairs_reference_array(gl, ga, r)
            assert stats.n_pairs == want_pairs.shape[0]
            assert np.array_equal(opairs.sort_pairs(pairs[:stats.n_pairs]), want_pairs)
            order, _ = orps.canonical_order(want_pairs, gl, ga, grid.as_dict(), mode=interact_mode)
            uu = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
            sp_ref, _ = orps.rps_sequential_c(sp_ref, order, uu, *p)
            assert np.array_equal(gs, sp_ref), "step %d: %d species differ" % (step, int((gs != sp_ref).sum()))
            if stats_every_step:
                assert list(stats.species_count)[1:] == [int((sp_ref == s).sum()) for s in (1, 2, 3)]
            lon_prev, lat_prev, t, ti = gl.copy(), ga.copy(), t + dt, ti_new
            total += stats.n_pairs
        assert total > 1500 and int((sp_ref != sp0).sum()) > 100
        per_step = (L.lm_launch_count(h) - launches0) / float(n_steps)
        # the scatter to id order in id windows (LM_OPT_SCATTER_PASSES; auto only splits above 8 M microbes): same arrays
        for passes in (2, 3, 7):
            assert L.lm_set_option(h, _lib.LM_OPT_SCATTER_PASSES, passes) == 0
            wl, wa, ws = np.full(n, np.nan, np.float32), np.full(n, np.nan, np.float32), np.full(n, -1, np.int8)
            assert L.lm_state_get(h, _ptr(wl), _ptr(wa), _ptr(ws), None) == 0
            assert np.array_equal(wl, gl) and np.array_equal(wa, ga) and np.array_equal(ws, gs)
        assert L.lm_set_option(h, _lib.LM_OPT_SCATTER_PASSES, 65) == _lib.LM_EINVAL
        per_step -= 1                             # lm_sync_stats after every step: one small kernel stores the counters into mapped host memory
        if interact_mode == 2:
            assert 25 <= per_step <= 28           # advection, binning (6), pair search, nine phase launches + nine heavy-unit launches
        elif interact_mode == 1:
            assert per_step <= 17                 # advection, binning (6), one tile launch + at most six boundary launches
        else:
            assert per_step < 12 if resolve_mode == 1 else per_step >= 16      # one resolver launch instead of nine
    finally:
        assert L.lm_destroy(h) == 0


# (the GPU suite runs every mode x transport; here: the default mode on both transports, one case each of the other modes)
@pytest.mark.parametrize("n_strips,interact_mode,resolve_mode,peer", [(2, 2, 0, False), (3, 2, 0, True), (3, 1, 0, True),
                                                                      (2, 0, 1, False)])
def test_latitude_strips_through_the_c_abi(abi, n_strips, interact_mode, resolve_mode, peer):
    """The whole strip protocol executed: G handles, particles handed out in contiguous tiles, routing
    passes until every microbe sits in its strip, then fused steps in the five stages of include/lm_b200.h with the
    exchange buffers copied between neighbours -- against ONE handle stepping all microbes on the same grid.
    Positions, pair set and species must agree bit for bit (strips with the tiled resolver vs a single handle with the
    nine phases, and the other way round)."""
    from lagrangian_microbes_b200 import _lib
    from lagrangian_microbes_b200.engine import make_grid
    from lagrangian_microbes_b200.particle_advecter import StageClock
    from lagrangian_microbes_b200.strips import cell_rows, strip_edges
    L, G = abi, n_strips
    fs = _small_field()
    n, r, p, seed, dt, n_steps = 700, 0.01, (0.55, 0.6, 0.9), 3, 3600.0, 3
    rng = np.random.default_rng(20 + G)
    lon = (201.5 + 0.2 * rng.random(n)).astype(np.float32)
    lat = (32.5 + 0.2 * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    ids = np.arange(n, dtype=np.int32)
    u, v = np.ascontiguousarray(fs.u), np.ascontiguousarray(fs.v)
    glon, glat = np.ascontiguousarray(fs.lon), np.ascontiguousarray(fs.lat)
    max_cells, cap = 1 << 14, 40 * n
    grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, n, max_cells, margin=0.1)
    edges = strip_edges(np.bincount(cell_rows(lat, grid), minlength=grid.ncy), G)
    flags_full = _lib.LM_STEP_ADVECT | _lib.LM_STEP_INTERACT | _lib.LM_STEP_EMIT_PAIRS | _lib.LM_STEP_STATS

    def make(mode):
        h = ctypes.c_void_p()
        assert L.lm_create(ctypes.byref(h), 0, n + 512, max_cells, cap) == 0
        assert L.lm_set_field(h, _ptr(u), _ptr(v), _ptr(glon), _ptr(glat), *u.shape) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_INTERACT_MODE, interact_mode) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_MODE, mode) == 0
        return h

    single = make(1 - resolve_mode)
    strips = [make(resolve_mode) for _ in range(G)]
    try:
        assert L.lm_set_grid(single, ctypes.byref(grid)) == 0
        assert L.lm_state_set(single, _ptr(lon), _ptr(lat), _ptr(sp0), None, n, None) == 0
        bufs, pair_bufs = [], []
        per = n // G
        for k, h in enumerate(strips):
            assert L.lm_strip_alloc(h, 512, 512, grid.ncx + 8) == 0
            assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
            b = _lib.StripBuffers()
            assert L.lm_strip_buffers_get(h, ctypes.byref(b)) == 0
            bufs.append(b)
            pair_bufs.append(np.zeros((cap, 2), dtype=np.int32))
            sl = slice(k * per, (k + 1) * per if k < G - 1 else n)          # the reference's contiguous tiles
            st = _lib.Strip(edges[k], edges[k + 1] - edges[k], int(k > 0), int(k < G - 1))
            assert L.lm_set_strip(h, ctypes.byref(st)) == 0
            a, b_, c_, d_ = (np.ascontiguousarray(x[sl]) for x in (lon, lat, sp0, ids))
            assert L.lm_state_set(h, _ptr(a), _ptr(b_), _ptr(c_), _ptr(d_), a.size, None) == 0
        if peer:
            # the peer-memory exchange of the multi-GPU run: every strip connects to its neighbours' receive buffers (plain
            # pointers: one process), the packing kernels write there and lm_step_push posts the flags -- no copies by the caller
            exports = []
            for h in strips:
                e = _lib.PeerExport()
                assert L.lm_strip_peer_export(h, ctypes.byref(e)) == 0
                exports.append(e)
            for k, h in enumerate(strips):
                if k > 0:
                    assert L.lm_strip_peer_connect(h, 0, ctypes.byref(exports[k - 1]), 0) == 0
                if k < G - 1:
                    assert L.lm_strip_peer_connect(h, 1, ctypes.byref(exports[k + 1]), 0) == 0

        def copy(dst, src, nbytes):
            if not peer:
                ctypes.memmove(dst, src, nbytes)

        def push(kind):
            if peer:
                for h in strips:
                    assert L.lm_step_push(h, kind, None) == 0

        def staged(flags, st_times, prm):
            for h in strips:
                assert L.lm_step_move(h, flags, ctypes.byref(st_times) if st_times is not None else None, dt, 0.0,
                                      ctypes.byref(prm), None) == 0
            for k in range(G):                                            # index 0 = south side, 1 = north side
                if k > 0:
                    copy(bufs[k - 1].mig_recv[1], bufs[k].mig_send[0], bufs[k].mig_bytes)
                if k < G - 1:
                    copy(bufs[k + 1].mig_recv[0], bufs[k].mig_send[1], bufs[k].mig_bytes)
            push(_lib.LM_XCHG_MIG)
            for h in strips:
                assert L.lm_step_bin(h, None) == 0
            halo = bool(flags & _lib.LM_STEP_INTERACT)
            if halo:
                for k in range(1, G):
                    copy(bufs[k - 1].ghost_recv, bufs[k].ghost_send, bufs[k].ghost_bytes)
            push(_lib.LM_XCHG_GHOST)
            for k, h in enumerate(strips):
                assert L.lm_step_interact_begin(h, r, _ptr(pair_bufs[k]), cap, None) == 0
            if halo:
                for k in range(1, G):
                    copy(bufs[k - 1].gsp_recv, bufs[k].gsp_send, bufs[k].species_bytes)
            push(_lib.LM_XCHG_GSP)
            for h in strips:
                assert L.lm_step_interact_end(h, None) == 0
            if halo:
                for k in range(G - 1):
                    copy(bufs[k + 1].gret_recv, bufs[k].gret_send, bufs[k].species_bytes)
            push(_lib.LM_XCHG_GRET)
            for h in strips:
                assert L.lm_step_finish(h, None) == 0

        def strip_stats(h):
            s = _lib.Stats()
            rc = L.lm_sync_stats(h, ctypes.byref(s), None)
            assert rc in (0, _lib.LM_ESTATE), rc                          # LM_ESTATE: misrouted microbes are still on their way
            return s

        prm0 = _lib.RpsParams(*p, seed, 0)
        for _ in range(50):                                               # settle: one hop per pass
            for k, h in enumerate(strips):
                st = _lib.Strip(edges[k], edges[k + 1] - edges[k], int(k > 0), int(k < G - 1))
                assert L.lm_set_strip(h, ctypes.byref(st)) == 0
            staged(0, None, prm0)
            if sum(strip_stats(h).n_misrouted for h in strips) == 0:
                break
        else:
            raise AssertionError("routing did not converge")
        assert sum(L.lm_state_size(h) for h in strips) == n

        clock_a, clock_b = StageClock(fs.time), StageClock(fs.time)
        pairs_single = np.zeros((cap, 2), dtype=np.int32)
        rec_bufs = [(np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int8)) for _ in range(2)]
        moved = 0
        for step in range(n_steps):
            prm = _lib.RpsParams(*p, seed, step)
            assert L.lm_step(single, flags_full, ctypes.byref(clock_a.next_step(dt)), dt, 0.0, r, ctypes.byref(prm),
                             _ptr(pairs_single), cap, None) == 0
            s1 = _lib.Stats()
            assert L.lm_sync_stats(single, ctypes.byref(s1), None) == 0
            rec_h = strips[min(1, G - 1)]                                  # the in-step record by ids of one strip, on two steps of three
            if step % 3 != 1:
                assert L.lm_record_next_step_ids(rec_h, *(_ptr(a) for a in rec_bufs[step & 1])) == 0
            staged(flags_full, clock_b.next_step(dt), prm)
            if step % 3 != 1:
                assert L.lm_host_copies_sync(rec_h) == 0
                m = L.lm_record_count(rec_h)
                assert m == L.lm_state_size(rec_h) and m > 0
                ri, rl, ra, rs = (a[:m] for a in rec_bufs[step & 1])
                fl, fa, fsp = np.full(n, np.nan, np.float32), np.full(n, np.nan, np.float32), np.full(n, -1, np.int8)
                assert L.lm_state_get(rec_h, _ptr(fl), _ptr(fa), _ptr(fsp), None) == 0
                assert np.unique(ri).size == m and np.array_equal(fl[ri], rl) and np.array_equal(fa[ri], ra) and np.array_equal(fsp[ri], rs)
                assert np.isnan(np.delete(fl, ri)).all()                   # exactly the strip's microbes
            st = [strip_stats(h) for h in strips]
            assert sum(s.n_misrouted for s in st) == 0 and sum(s.n_particles for s in st) == n
            assert sum(s.n_pairs for s in st) == s1.n_pairs
            moved += sum(s.n_moved_in for s in st)
            wl, wa, ws = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int8)
            assert L.lm_state_get(single, _ptr(wl), _ptr(wa), _ptr(ws), None) == 0
            gl, ga, gs = np.full(n, np.nan, np.float32), np.full(n, np.nan, np.float32), np.full(n, -1, np.int8)
            for h in strips:                                              # global ids: every strip fills its own microbes
                assert L.lm_state_get(h, _ptr(gl), _ptr(ga), _ptr(gs), None) == 0
            assert np.array_equal(gl, wl) and np.array_equal(ga, wa), "positions, step %d" % step
            got = opairs.sort_pairs(np.concatenate([pb[:s.n_pairs] for pb, s in zip(pair_bufs, st)]))
            assert np.array_equal(got, opairs.sort_pairs(pairs_single[:s1.n_pairs])), "pair set, step %d" % step
            assert np.array_equal(gs, ws), "species, step %d: %d differ" % (step, int((gs != ws).sum()))
        assert moved > 0                                                  # microbes did cross strip boundaries
        if peer:
            # a neighbour that never sends (it died, or raised and left): the stage that waits for its message gives up
            # after LM_OPT_PEER_WAIT_CYCLES instead of spinning for ever, the fault is reported once as LM_ESTATE, and
            # further waits of the handle return at once
            lone = strips[0]
            assert L.lm_set_option(lone, _lib.LM_OPT_PEER_WAIT_CYCLES, 0) == _lib.LM_EINVAL
            assert L.lm_set_option(lone, _lib.LM_OPT_PEER_WAIT_CYCLES, 2000000) == 0       # emulator clock: 2 ms
            prm = _lib.RpsParams(*p, seed, n_steps)
            import time as _time
            t0 = _time.time()
            assert L.lm_step_move(lone, flags_full, ctypes.byref(clock_b.next_step(dt)), dt, 0.0, ctypes.byref(prm), None) == 0
            assert L.lm_step_push(lone, _lib.LM_XCHG_MIG, None) == 0
            assert L.lm_step_bin(lone, None) == 0                         # its northern neighbour has not stepped
            assert L.lm_sync_stats(lone, ctypes.byref(_lib.Stats()), None) == _lib.LM_ESTATE
            assert _time.time() - t0 < 30.0
    finally:
        for h in [single] + strips:
            L.lm_destroy(h)


# ----------------------------------------------------------------------------------------------------------------------
# The committed golden vectors (tests/golden/, made with the unmodified reference function and with SciPy) through the
# emulated C ABI: the CPU suite pins the kernels' logic to the reference's own outputs, not only the GPU suite.
@pytest.mark.parametrize("name,interact_mode,resolve_mode", [("rps_oddspecies", 2, 0), ("rps_clustered", 2, 0),
                                                              ("rps_knots", 2, 0),
                                                              ("rps_oddspecies", 1, 0), ("rps_clustered", 1, 0),
                                                              ("rps_oddspecies", 0, 0), ("rps_oddspecies", 0, 1),
                                                              ("rps_clustered", 0, 1)])
def test_golden_species_through_the_c_abi(abi, name, interact_mode, resolve_mode):
    from conftest import golden
    from lagrangian_microbes_b200 import _lib
    L = abi
    g = golden(name + ".npz")
    n, n_pairs = g["lon"].size, g["pairs_ref_order"].shape[0]
    h = ctypes.c_void_p()
    assert L.lm_create(ctypes.byref(h), 0, n, 1 << 16, n_pairs + 64) == 0
    try:
        grid = _lib.Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1]))
        assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_INTERACT_MODE, interact_mode) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_MODE, resolve_mode) == 0
        lon, lat = np.ascontiguousarray(g["lon"]), np.ascontiguousarray(g["lat"])
        # fused path, canonical order (tile-round for the fused tile kernel, cell-phase for the round-1 pipeline):
        # species made by the unmodified reference function fed that order
        species = g["species0"].copy()
        prm = _lib.RpsParams(float(g["pRS"]), float(g["pPR"]), float(g["pSP"]), int(g["seed"]), int(g["step"]))
        found = np.zeros(1, dtype=np.int64)
        assert L.lm_interact_rps(h, _ptr(lon), _ptr(lat), _ptr(species), n, float(g["r"]), ctypes.byref(prm), None, 0,
                                 _ptr(found), None) == 0
        stats = _lib.Stats()
        assert L.lm_sync_stats(h, ctypes.byref(stats), None) == 0
        assert stats.n_pairs == n_pairs == found[0]
        assert np.array_equal(species, g[{2: "species_round", 1: "species_tile", 0: "species_cell"}[interact_mode]])
        if resolve_mode == 0:
            # explicit-order resolver (M-ref): the reference's own set-iteration order and draws
            species = g["species0"].copy()
            pairs = np.ascontiguousarray(g["pairs_ref_order"], dtype=np.int32)
            u = np.ascontiguousarray(g["u_ref"], dtype=np.float64)
            rounds = ctypes.c_int32(0)
            assert L.lm_resolve_rps(h, _ptr(pairs), _ptr(u), n_pairs, _ptr(species), n, float(g["pRS"]), float(g["pPR"]),
                                    float(g["pSP"]), ctypes.byref(rounds), None) == 0
            assert np.array_equal(species, g["species_ref"]) and rounds.value > 1
    finally:
        L.lm_destroy(h)


def test_golden_pair_sets_through_the_c_abi(abi):
    """SciPy's pair sets (p = 2, 1, inf) incl. points exactly r apart, duplicates and collinear points."""
    from conftest import golden
    from lagrangian_microbes_b200 import _lib
    from lagrangian_microbes_b200.engine import make_grid
    L = abi
    cases = []
    g = golden("pairs_cases.npz")
    for name in ("exact345", "dups_collinear", "lattice"):
        cases.append((g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"]), g[name + "_pairs"], _lib.LM_NORM_2))
    g = golden("pairs_norms.npz")
    for tag, code in (("p1", _lib.LM_NORM_1), ("pinf", _lib.LM_NORM_INF)):
        cases.append((g["exact_lon"], g["exact_lat"], float(g["exact_r"]), g["exact_pairs_" + tag], code))
        cases.append((g["blob_lon"][:300], g["blob_lat"][:300], float(g["blob_r"]), None, code))
    for lon, lat, r, want, code in cases:
        lon, lat = np.ascontiguousarray(lon, dtype=np.float32), np.ascontiguousarray(lat, dtype=np.float32)
        if want is None:
            want = opairs.query_pairs_reference_array(lon, lat, r, p={_lib.LM_NORM_1: 1, _lib.LM_NORM_INF: np.inf}[code])
        want = opairs.sort_pairs(np.asarray(want).reshape(-1, 2))
        n, cap = lon.size, want.shape[0] + 16
        h = ctypes.c_void_p()
        assert L.lm_create(ctypes.byref(h), 0, n, 1 << 16, 0) == 0
        try:
            grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, n, 1 << 16, margin=0.0)
            assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
            assert L.lm_set_option(h, _lib.LM_OPT_NORM, code) == 0
            pairs = np.zeros((cap, 2), dtype=np.int32)
            found = np.zeros(1, dtype=np.int64)
            assert L.lm_find_pairs(h, _ptr(lon), _ptr(lat), n, r, _ptr(pairs), cap, _ptr(found), None) == 0
            assert found[0] == want.shape[0]
            assert np.array_equal(opairs.sort_pairs(pairs[:found[0]]), want)
        finally:
            L.lm_destroy(h)


@pytest.mark.parametrize("resolve_mode", [0, 1])
def test_hand_off_overflow_is_reported_not_resolved(abi, resolve_mode):
    """A hand-off buffer smaller than the pairs found: LM_ENOSPC from lm_sync_stats, the species left alone (both
    resolvers give up before touching them), the true pair count reported so that the caller can size it and repeat."""
    from conftest import golden
    from lagrangian_microbes_b200 import _lib
    L = abi
    g = golden("rps_oddspecies.npz")
    n, n_pairs = g["lon"].size, g["pairs_ref_order"].shape[0]
    h = ctypes.c_void_p()
    assert L.lm_create(ctypes.byref(h), 0, n, 1 << 16, n_pairs // 2) == 0
    try:
        grid = _lib.Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1]))
        assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
        assert L.lm_set_option(h, _lib.LM_OPT_INTERACT_MODE, 0) == 0     # the round-1 pipeline; the fused tile kernel has no hand-off
        assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_MODE, resolve_mode) == 0
        species = g["species0"].copy()
        prm = _lib.RpsParams(float(g["pRS"]), float(g["pPR"]), float(g["pSP"]), int(g["seed"]), int(g["step"]))
        assert L.lm_interact_rps(h, _ptr(np.ascontiguousarray(g["lon"])), _ptr(np.ascontiguousarray(g["lat"])), _ptr(species), n,
                                 float(g["r"]), ctypes.byref(prm), None, 0, None, None) == 0
        stats = _lib.Stats()
        assert L.lm_sync_stats(h, ctypes.byref(stats), None) == _lib.LM_ENOSPC
        assert stats.n_pairs == n_pairs
        assert np.array_equal(species, g["species0"])
    finally:
        L.lm_destroy(h)


def test_kernels_are_race_free_under_thread_sanitizer():
    """The emulated library under -fsanitize=thread (tests/cuda_emu/tsan_driver.cpp, quick subset): both resolvers with
    shared-memory / scratch tiles and the whole-warp / whole-CTA paths, fused steps with diffusion, the record pipeline,
    the analysis kernels.  A missing __syncthreads() / __syncwarp() / atomic in a kernel is a data-race report here.
    (The full driver -- all resolver options, explicit-order resolver, two strips -- is run by hand:
    profiles/r1z_tsan_emulated.txt.)"""
    import subprocess
    import emu_build
    try:
        exe = emu_build.build_tsan()
    except subprocess.CalledProcessError:
        pytest.skip("g++ -fsanitize=thread is not available here")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 history_size=4 exitcode=66")
    res = subprocess.run([exe, "quick"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1500)
    err = res.stderr.decode(errors="replace")
    if "FATAL: ThreadSanitizer" in err:                      # the sanitizer runtime cannot start on this kernel (ASLR layout)
        pytest.skip("ThreadSanitizer cannot run here: " + err.strip().splitlines()[0])
    assert "ThreadSanitizer" not in err, err[:4000]
    assert res.returncode == 0, (res.returncode, err[:2000])
    assert b"done" in res.stdout


@pytest.mark.parametrize("shape", [1, 2, 3])
def test_tiled_resolver_other_tile_shapes(abi, shape):
    """LM_OPT_RESOLVE_TILE_SHAPE: 32 x 16, 128 x 16 and 64 x 32 cells per tile (same halo) on the clustered golden case and
    on a grid with several tiles of every shape."""
    from conftest import golden
    from lagrangian_microbes_b200 import _lib
    L = abi
    g = golden("rps_clustered.npz")
    c = Cloud(31 + shape, 150, 37, 2200, knots=6)
    cases = [(np.ascontiguousarray(g["lon"]), np.ascontiguousarray(g["lat"]), g["species0"], float(g["r"]),
              (float(g["pRS"]), float(g["pPR"]), float(g["pSP"])), int(g["seed"]), int(g["step"]), g["species_cell"],
              _lib.Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1]))),
             (c.lon, c.lat, c.sp0, R, P, 5, 17, c.oracle(c.sp0, 0, 8),
              _lib.Grid(c.grid["x0"], c.grid["y0"], c.grid["inv_h"], c.grid["ncx"], c.grid["ncy"]))]
    for lon, lat, sp0, r, p, seed, step, want, grid in cases:
        n = lon.size
        h = ctypes.c_void_p()
        assert L.lm_create(ctypes.byref(h), 0, n, 1 << 16, 40 * n) == 0
        try:
            assert L.lm_set_grid(h, ctypes.byref(grid)) == 0
            assert L.lm_set_option(h, _lib.LM_OPT_INTERACT_MODE, 0) == 0        # the round-1 pipeline
            assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_MODE, 1) == 0
            assert L.lm_set_option(h, _lib.LM_OPT_RESOLVE_TILE_SHAPE, shape) == 0
            species = sp0.copy()
            prm = _lib.RpsParams(*p, seed, step)
            assert L.lm_interact_rps(h, _ptr(lon), _ptr(lat), _ptr(species), n, r, ctypes.byref(prm), None, 0, None, None) == 0
            assert L.lm_sync_stats(h, None, None) == 0
            assert np.array_equal(species, want), "shape %d: %d species differ" % (shape, int((species != want).sum()))
        finally:
            L.lm_destroy(h)


# ----------------------------------------------------------------------------------------------------------------------
# The fused tile kernel (csrc/interact.cu, LM_OPT_INTERACT_MODE = 1, the default device path): pair search + RPS in one
# pass, canonical tile-round order (oracle/rps.py::tile_round_order).  Pair set vs cKDTree, species vs the reference rule.
@pytest.fixture(scope="module")
def emu_tile():
    import emu_build
    L = ctypes.CDLL(emu_build.build())
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_double
    L.emu_interact_tile.restype = i64
    L.emu_interact_tile.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, dbl, i32, dbl, dbl, dbl, u64, u64, i32, i32,
                                    i32, i32, i32, i32, vp, i64]
    return L


class TileCloud(Cloud):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.order, self.phase = orps.tile_round_order(self.pairs, self.lon, self.lat, self.grid)
        self.u = philox.pair_uniforms(self.order[:, 0], self.order[:, 1], 17, 5)

    def run_tile(self, L, sp, first=0, last=14, tile_cap=0, draw_batch=0, rps=True, rec_cap=0, path=0):
        g = self.grid
        lon_s, lat_s = np.ascontiguousarray(self.lon[self.ids]), np.ascontiguousarray(self.lat[self.ids])
        sp_s = np.ascontiguousarray(sp[self.ids])
        cap = self.pairs.shape[0] + 8
        pairs_out = np.full((cap, 2), -1, dtype=np.int32)
        ret = L.emu_interact_tile(_ptr(lon_s), _ptr(lat_s), _ptr(self.ids), _ptr(self.cell_start), _ptr(sp_s) if rps else None,
                                  self.n, g["ncx"], g["ncy"], 0, g["ncy"], g["ncy"], R, 2, *P, 5, 17, first, last, tile_cap,
                                  draw_batch, rec_cap, path, _ptr(pairs_out), cap)
        assert ret >= 0
        found, launches = ret & ((1 << 48) - 1), ret >> 48
        out = np.empty_like(sp)
        out[self.ids] = sp_s
        return out, launches, opairs.sort_pairs(pairs_out[:found])


@pytest.mark.parametrize("seed,ncx,ncy,n,knots,knot_size,tile_cap,draw_batch,rec_cap,path", [
    (1, 74, 19, 1800, 10, (10, 31), 0, 0, 0, 0),      # three tiles across, two up, ragged; records in shared memory; knots on the whole-warp path
    (1, 74, 19, 1800, 10, (10, 31), 0, 0, 512, 0),    # ... record buffer too small in some directions of some tiles: both paths side by side
    (2, 74, 37, 3000, 6, (40, 60), 0, 1, 0, 1),       # lane walk, draws taken one lane at a time
    (3, 40, 20, 1500, 2, (150, 200), 256, 32, 0, 0),  # 18,000-slot cells on the whole-CTA path; tiles too full for shared memory
])
def test_fused_tile_kernel_executed(emu_tile, seed, ncx, ncy, n, knots, knot_size, tile_cap, draw_batch, rec_cap, path):
    c = TileCloud(seed, ncx, ncy, n, knots=knots, knot_size=knot_size)
    want = c.oracle(c.sp0, 0, 14)
    assert int((want != c.sp0).sum()) > 100 and set(np.unique(c.phase)) >= set(range(15))
    got, launches, pairs = c.run_tile(emu_tile, c.sp0, tile_cap=tile_cap, draw_batch=draw_batch, rec_cap=rec_cap, path=path)
    assert np.array_equal(pairs, c.pairs), "pair set differs from cKDTree.query_pairs"
    assert np.array_equal(got, want), "%d species differ" % int((got != want).sum())
    assert launches == 7                                       # one tile launch + six boundary phases


def test_fused_tile_kernel_phase_ranges_and_pair_search_alone(emu_tile):
    """The two ranges a strip launches around its halo exchange (0-11, then 12-14 on the result) and the pair search
    without species (lm_find_pairs)."""
    c = TileCloud(4, 70, 33, 2500, knots=4, knot_size=(30, 50))
    mid, _, p_lo = c.run_tile(emu_tile, c.sp0, first=0, last=11)
    assert np.array_equal(mid, c.oracle(c.sp0, 0, 11))
    end, _, p_hi = c.run_tile(emu_tile, mid, first=12, last=14)
    assert np.array_equal(end, c.oracle(c.sp0, 0, 14))
    assert np.array_equal(opairs.sort_pairs(np.concatenate((p_lo, p_hi))), c.pairs)      # every pair exactly once
    same, _, pairs = c.run_tile(emu_tile, c.sp0, rps=False)
    assert np.array_equal(pairs, c.pairs) and np.array_equal(same, c.sp0)
