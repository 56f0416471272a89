"""The tiled RPS resolver (LM_OPT_RESOLVE_MODE = 1: one launch per phase range, species of a tile + halo in shared
memory, csrc/pairs.cu::resolve_tiled_kernel) must give exactly what the nine phase launches give: the reference rule
(interactions.py:13-40) applied sequentially in the canonical cell-phase order (oracle/rps.py).  Covered: the golden
species made with the unmodified reference function, live clouds (short units, crowded cells, knots that go to the
whole-warp path), tiles that do not fit in shared memory (global scratch path), the fused multi-step loop, and
latitude strips (phase ranges 0-5 / 6-8 around the halo exchange) against a single handle."""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps

# (File name: sorts after every verified GPU test, so that a fault in a kernel that has never run cannot poison the
# CUDA context of the tests before it.)
# First run on hardware: the round-1 driver run (GPUTEST_r01.json), all green.
pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _round1_pipeline(monkeypatch):
    """These tests cover the round-1 pipeline (pair search -> hand-off -> nine phase launches / tiled resolver,
    csrc/pairs.cu, cell-phase order), still shipped as LM_OPT_INTERACT_MODE = 0."""
    from lagrangian_microbes_b200.engine import Engine
    monkeypatch.setattr(Engine, "DEFAULT_INTERACT_MODE", 0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def tiled(eng, smem=32768, heavy_min=0, shape=0):
    from lagrangian_microbes_b200._lib import (LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_MODE, LM_OPT_RESOLVE_TILE_SHAPE,
                                               LM_OPT_RESOLVE_TILE_SMEM)
    eng.set_option(LM_OPT_RESOLVE_MODE, 1)
    eng.set_option(LM_OPT_RESOLVE_TILE_SMEM, smem)
    eng.set_option(LM_OPT_RESOLVE_HEAVY_MIN, heavy_min)
    eng.set_option(LM_OPT_RESOLVE_TILE_SHAPE, shape)


@pytest.mark.parametrize("name", ["rps_uniform", "rps_clustered", "rps_oddspecies"])
@pytest.mark.parametrize("smem", [32768, 1024])
def test_golden_species(engine_factory, name, smem):
    from lagrangian_microbes_b200._lib import Grid
    g = golden(name + ".npz")
    n = g["lon"].size
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=g["pairs_ref_order"].shape[0] + 64)
    eng.set_grid(Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1])))
    tiled(eng, smem)
    species = dev(g["species0"].copy())
    eng.interact_rps(dev(g["lon"]), dev(g["lat"]), species, float(g["r"]), float(g["pRS"]), float(g["pPR"]), float(g["pSP"]),
                     int(g["seed"]), int(g["step"]))
    assert eng.sync_stats().n_pairs == g["pairs_ref_order"].shape[0]
    assert np.array_equal(species.cpu().numpy(), g["species_cell"])


def _cloud(kind, rng):
    if kind == "uniform":
        n = 150000
        side = np.sqrt(n / 4900.0)
        return 205 + side * rng.random(n), 25 + side * rng.random(n), 0.02
    if kind == "crowded":                                   # ~40 microbes per cell
        n = 40000
        return 205 + 0.3 * rng.random(n), 25 + 0.3 * rng.random(n), 0.01
    n = 60000                                               # knots of 10-25 microbes in a sparse background
    lon, lat = 205 + 3.0 * rng.random(n), 25 + 3.0 * rng.random(n)
    k = 0
    for c in range(300):
        m = int(rng.integers(10, 26))
        lon[k:k + m] = 205.005 + 0.01 * int(rng.integers(0, 290)) + 0.004 * rng.random(m)
        lat[k:k + m] = 25.005 + 0.01 * int(rng.integers(0, 290)) + 0.004 * rng.random(m)
        k += m
    return lon, lat, 0.01


@pytest.mark.parametrize("kind", ["uniform", "crowded", "knots"])
def test_live_species(engine_factory, kind):
    from lagrangian_microbes_b200.engine import make_grid
    rng = np.random.default_rng(11)
    lon, lat, r = _cloud(kind, rng)
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    n = lon.size
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    sp0[::53] = 0                                           # species outside {1, 2, 3}: draw, no winner
    p = (0.55, 0.6, 0.9)
    want_pairs = opairs.query_pairs_reference_array(lon, lat, r)
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=want_pairs.shape[0] + 64)
    grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, n, eng.max_cells, margin=0.1)
    eng.set_grid(grid)
    order, _ = orps.cell_phase_order(want_pairs, lon, lat, grid.as_dict())
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 17, 5)
    want_sp, draws = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    assert draws > 0
    lon_d, lat_d = dev(lon), dev(lat)
    # (shared memory per tile, whole-warp limit): default; every tile in the global scratch; every unit with more than
    # 8 pairs on the whole-warp path; no whole-warp path at all
    for smem, heavy_min, shape in [(32768, 0, 0), (1024, 0, 0), (32768, 8, 0), (65536, 0xffff, 0), (32768, 0, 1), (32768, 0, 2),
                                   (1024, 8, 3)]:
        tiled(eng, smem, heavy_min, shape)
        species = dev(sp0.copy())
        eng.interact_rps(lon_d, lat_d, species, r, *p, 5, 17)
        assert eng.sync_stats().n_pairs == want_pairs.shape[0]
        bad = int((species.cpu().numpy() != want_sp).sum())
        assert bad == 0, "smem %d heavy_min %d shape %d: %d species differ" % (smem, heavy_min, shape, bad)


def test_knots_on_the_whole_cta(engine_factory):
    """Knots of several hundred microbes in one cell (config 2 run past step 3,000): units of 10^4..10^5 pairs, resolved
    by the whole CTA.  Default limit, a low limit (every unit above 48 pairs), and the global-scratch variant."""
    from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_MEGA_MIN
    from lagrangian_microbes_b200.engine import make_grid
    rng = np.random.default_rng(23)
    n, r = 30000, 0.01
    lon, lat = 205 + 2.0 * rng.random(n), 25 + 2.0 * rng.random(n)
    k = 0
    for c in range(12):
        m = int(rng.integers(150, 600))
        lon[k:k + m] = 205.005 + 0.01 * int(rng.integers(0, 190)) + 0.006 * rng.random(m)
        lat[k:k + m] = 25.005 + 0.01 * int(rng.integers(0, 190)) + 0.006 * rng.random(m)
        k += m
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    p = (0.55, 0.6, 0.9)
    want_pairs = opairs.query_pairs_reference_array(lon, lat, r)
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=want_pairs.shape[0] + 64)
    grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, n, eng.max_cells, margin=0.1)
    eng.set_grid(grid)
    order, _ = orps.cell_phase_order(want_pairs, lon, lat, grid.as_dict())
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 17, 5)
    want_sp, _ = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    lon_d, lat_d = dev(lon), dev(lat)
    for mega_min, smem in [(0, 32768), (48, 32768), (48, 1024), (1 << 30, 32768)]:
        tiled(eng, smem)
        eng.set_option(LM_OPT_RESOLVE_MEGA_MIN, mega_min)
        species = dev(sp0.copy())
        eng.interact_rps(lon_d, lat_d, species, r, *p, 5, 17)
        assert eng.sync_stats().n_pairs == want_pairs.shape[0]
        bad = int((species.cpu().numpy() != want_sp).sum())
        assert bad == 0, "mega_min %d smem %d: %d species differ" % (mega_min, smem, bad)


def test_fused_loop_equals_the_phase_launches():
    """Twelve fused steps (advection, re-binning, regridding) with the tiled resolver against the default resolver."""
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from test_gpu_strips import P, R, particles, small_fs
    fs = small_fs()
    lon, lat, sp = particles(60000, 5, clustered=True)
    sims = [FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, seed=3, emit_pairs=True, pair_capacity=80 * lon.size,
                            regrid_every=4, grid_margin=0.25) for _ in range(2)]
    tiled(sims[1].engine)
    for step in range(12):
        a, b = sims[0].step(check=True), sims[1].step(check=True)
        assert a.n_pairs == b.n_pairs and list(a.species_count) == list(b.species_count), "step %d" % step
    for x, y in zip(sims[0].download(), sims[1].download()):
        assert np.array_equal(x, y)
    assert sims[1].engine.launch_count() < sims[0].engine.launch_count() - 12 * 7     # one launch instead of nine


@pytest.mark.parametrize("G,clustered", [(2, False), (3, True)])
def test_strips_with_the_tiled_resolver_equal_single_handle(G, clustered):
    from lagrangian_microbes_b200.strips import LocalTransport, StripSet
    from test_gpu_strips import P, R, compare_step, particles, single, small_fs
    n, seed = 40000, 21 + G
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered)
    ids = np.arange(n, dtype=np.int32)
    per = n // G
    cut = [slice(g * per, (g + 1) * per if g < G - 1 else n) for g in range(G)]
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=40 * G, grid_margin=0.25, regrid_every=0)
    for strip in ss.strips:
        tiled(strip.engine)
    sim = single(lon, lat, sp, ss.grid, fs, seed)          # default resolver on the single handle
    total = 0
    for step in range(5):
        total += compare_step(ss, sim, step)
    assert total > 1000
    ss.close()


def test_mode_option_validation(engine_factory):
    from lagrangian_microbes_b200._lib import LM_EINVAL, LM_OPT_RESOLVE_MODE, LM_OPT_RESOLVE_TILE_SMEM, LmError
    eng = engine_factory(max_particles=64, max_cells=1024)
    for opt, bad in ((LM_OPT_RESOLVE_MODE, 2), (LM_OPT_RESOLVE_MODE, -1), (LM_OPT_RESOLVE_TILE_SMEM, 100),
                     (LM_OPT_RESOLVE_TILE_SMEM, 1 << 20)):
        with pytest.raises(LmError) as ei:
            eng.set_option(opt, bad)
        assert ei.value.code == LM_EINVAL
    eng.set_option(LM_OPT_RESOLVE_MODE, 1)
    eng.set_option(LM_OPT_RESOLVE_MODE, 0)
