"""GPU parity tests: the CUDA path, called through the C ABI (Engine -> liblm_b200.so), against the
oracle on the same seeded inputs and against the committed golden fixtures.

Bars (north_star): pair set bit-exact vs cKDTree.query_pairs; species bit-exact vs the reference
interaction function fed the same per-pair stream in the same pair order; positions within 1e-6
relative of the RK4 oracle (float32-faithful AND float64).
"""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import philox
from oracle import rk4 as ork4
from oracle import rps as orps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PAIR_CASES = ["exact345", "dups_collinear", "lattice", "uniform", "tiny_lat", "empty", "single"]
RPS_CASES = ["rps_uniform", "rps_clustered", "rps_oddspecies", "rps_knots"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def grid_from_golden(g):
    from lagrangian_microbes_b200._lib import Grid
    return Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1]))


def auto_grid(eng, lon, lat, r, **kw):
    from lagrangian_microbes_b200.engine import make_grid
    if lon.size == 0:
        g = make_grid(0.0, 1.0, 0.0, 1.0, r, 1, eng.max_cells, **kw)
    else:
        g = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, lon.size,
                      eng.max_cells, **kw)
    eng.set_grid(g)
    return g


def gpu_pairs(eng, lon, lat, r, cap=None):
    n = lon.size
    cap = int(cap if cap is not None else max(1024, 40 * n))
    out = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    if n == 0:
        return np.zeros((0, 2), dtype=np.int64)
    npairs = eng.find_pairs(dev(lon.astype(np.float32)), dev(lat.astype(np.float32)), r, out)
    return opairs.sort_pairs(out[:npairs].cpu().numpy())


# ------------------------------------------------------------------------------------------------------
def test_pair_uniforms_bit_exact(engine_factory):
    eng = engine_factory(max_particles=1024, max_cells=1024)
    rng = np.random.default_rng(0)
    i = rng.integers(0, 2**31 - 2, 50000)
    j = i + rng.integers(1, 1000, 50000)
    j = np.minimum(j, 2**31 - 1)
    pairs = np.stack((i, j), -1).astype(np.int32)
    for seed, step in ((0, 0), (42, 7), (2**40 + 5, 2**33 + 1)):
        u = eng.pair_uniforms(dev(pairs), seed, step).cpu().numpy()
        assert np.array_equal(u, philox.pair_uniforms(pairs[:, 0], pairs[:, 1], step, seed))


@pytest.mark.parametrize("name", PAIR_CASES)
@pytest.mark.parametrize("coarse", [1.0, 0.05])
def test_find_pairs_golden(engine_factory, name, coarse):
    g = golden("pairs_cases.npz")
    lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
    eng = engine_factory(max_particles=max(lon.size, 16), max_cells=1 << 22)
    # coarse < 1 shrinks the cell budget -> cell edge k*r with k > 1
    auto_grid(eng, lon, lat, r, margin=0.0, cells_per_particle=2.0 * coarse)
    got = gpu_pairs(eng, lon, lat, r)
    assert np.array_equal(got, g[name + "_pairs"].astype(np.int64))


@pytest.mark.parametrize("n,r,kind", [(1000, 0.3, "uniform"), (50000, 0.02, "uniform"), (200000, 0.01, "uniform"),
                                      (100000, 0.005, "clustered"), (30000, 0.05, "line"), (4096, 0.0, "dups")])
def test_find_pairs_vs_live_ckdtree(engine_factory, n, r, kind):
    rng = np.random.default_rng(n)
    if kind == "uniform":
        side = np.sqrt(n / 4900.0)
        lon, lat = 205 + side * rng.random(n), 25 + side * rng.random(n)
    elif kind == "clustered":
        c = rng.random((40, 2)) * 4 + np.array([208.0, 28.0])
        w = rng.integers(0, 40, n)
        pts = c[w] + rng.normal(0, 0.03, (n, 2))
        lon, lat = pts[:, 0], pts[:, 1]
    elif kind == "line":
        lon = 200 + 3 * rng.random(n)
        lat = np.full(n, 12.5) + 1e-4 * rng.random(n)
    else:
        base = rng.random((n // 4, 2)) + np.array([200.0, 0.0])
        pts = np.repeat(base, 4, axis=0)                       # every point 4 times: 6 pairs each at r = 0
        lon, lat = pts[:, 0], pts[:, 1]
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1 << 22)
    auto_grid(eng, lon, lat, r, margin=0.1)
    want = opairs.query_pairs_reference_array(lon, lat, r)
    got = gpu_pairs(eng, lon, lat, r, cap=want.shape[0] + 1024)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_find_pairs_property_based(engine_factory):
    """Hypothesis: random clouds (uniform / clustered / duplicated / collinear, any radius) -- the pair set equals
    cKDTree.query_pairs, for pair search alone and fused with RPS (whose species then equal the oracle's)."""
    from hypothesis import HealthCheck, given, settings, strategies as st_

    eng = engine_factory(max_particles=4096, max_cells=1 << 20, max_pairs=4_000_000)
    out = torch.empty((4_000_000, 2), dtype=torch.int32, device="cuda")

    # deterministic by default; LM_HYPOTHESIS_EXAMPLES=N runs a randomised campaign of N examples instead
    import os
    campaign = int(os.environ.get("LM_HYPOTHESIS_EXAMPLES", "0"))

    @settings(max_examples=campaign or 40, deadline=None, derandomize=not campaign, database=None,
              suppress_health_check=list(HealthCheck))
    @given(n=st_.integers(0, 3000), kind=st_.sampled_from(["uniform", "clustered", "dups", "line"]),
           r=st_.sampled_from([0.003, 0.01, 0.05, 0.3]), seed=st_.integers(0, 2**31 - 1), coarse=st_.sampled_from([1.0, 0.02]))
    def check(n, kind, r, seed, coarse):
        rng = np.random.default_rng(seed)
        if kind == "uniform":
            lon, lat = 205 + rng.random(n), 30 + rng.random(n)
        elif kind == "clustered":
            c = rng.integers(0, 5, n)
            lon, lat = 205 + 0.2 * c + rng.normal(0, 0.004, n), 30 + 0.1 * c + rng.normal(0, 0.004, n)
        elif kind == "dups":
            base = rng.integers(0, max(n // 3, 1), n)
            lon, lat = 205 + (base % 37) * 0.007, 30 + (base // 37) * 0.007
        else:
            lon, lat = 205 + np.arange(n) * (r / 3.0), np.full(n, 30.25)
        lon, lat = lon.astype(np.float32), lat.astype(np.float32)
        want = opairs.query_pairs_reference_array(lon, lat, r)
        if want.shape[0] > 3_900_000 or n == 0:
            return
        grid = auto_grid(eng, lon, lat, r, margin=0.0, cells_per_particle=2.0 * coarse)
        assert np.array_equal(gpu_pairs(eng, lon, lat, r, cap=want.shape[0] + 8), want)
        sp0 = rng.integers(0, 4, n).astype(np.int8)
        species = dev(sp0.copy())
        eng.interact_rps(dev(lon), dev(lat), species, r, 0.3, 0.6, 0.9, seed % 1000, seed % 7, pairs_out=out)
        st = eng.sync_stats()
        assert st.n_pairs == want.shape[0]
        assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), want)
        order, _ = orps.canonical_order(want, lon, lat, grid.as_dict())
        u = philox.pair_uniforms(order[:, 0], order[:, 1], seed % 7, seed % 1000)
        want_sp, _ = orps.rps_sequential_c(sp0.copy(), order, u, 0.3, 0.6, 0.9)
        assert np.array_equal(species.cpu().numpy(), want_sp)

    check()


def test_dense_uniform_cloud_fills_both_mask_words_and_the_fill_buffer(engine_factory):
    """~28 microbes per cell: lanes with 65-128 candidates (second mask word), warps with more than 1024 hits
    (fill buffer overflow) and warps with more than 128 candidates per lane (two-pass path) side by side."""
    rng = np.random.default_rng(31)
    n, r = 60000, 0.05
    side = np.sqrt(n / 11200.0)
    lon = (205 + side * rng.random(n)).astype(np.float32)
    lat = (30 + side * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    want = opairs.query_pairs_reference_array(lon, lat, r)
    assert want.shape[0] > 2_000_000
    eng = engine_factory(max_particles=n, max_cells=1 << 20, max_pairs=want.shape[0] + 64)
    grid = auto_grid(eng, lon, lat, r, margin=0.1)
    out = torch.empty((want.shape[0] + 64, 2), dtype=torch.int32, device="cuda")
    species = dev(sp0.copy())
    eng.interact_rps(dev(lon), dev(lat), species, r, 0.55, 0.55, 0.55, 4, 2, pairs_out=out)
    st = eng.sync_stats()
    assert st.n_pairs == want.shape[0]
    assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), want)
    order, _ = orps.canonical_order(want, lon, lat, grid.as_dict())
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 2, 4)
    want_sp, _ = orps.rps_sequential_c(sp0.copy(), order, u, 0.55, 0.55, 0.55)
    assert np.array_equal(species.cpu().numpy(), want_sp)


def test_pairs_outside_grid_are_clamped_not_lost(engine_factory):
    rng = np.random.default_rng(5)
    n = 20000
    lon = (205 + 2 * rng.random(n)).astype(np.float32)
    lat = (25 + 2 * rng.random(n)).astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1 << 20)
    from lagrangian_microbes_b200.engine import make_grid
    eng.set_grid(make_grid(205.5, 206.5, 25.5, 26.5, 0.02, n, eng.max_cells, margin=0.0))   # grid covers 1/4 of them
    want = opairs.query_pairs_reference_array(lon, lat, 0.02)
    got = gpu_pairs(eng, lon, lat, 0.02, cap=want.shape[0] + 10)
    assert np.array_equal(got, want)
    assert eng.sync_stats().n_clamped > 0


def test_pair_capacity_overflow_is_reported(engine_factory):
    from lagrangian_microbes_b200._lib import LmError, LM_ENOSPC
    rng = np.random.default_rng(6)
    n = 5000
    lon = (205 + 0.5 * rng.random(n)).astype(np.float32)
    lat = (25 + 0.5 * rng.random(n)).astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1 << 20)
    auto_grid(eng, lon, lat, 0.02)
    want = opairs.query_pairs_reference_array(lon, lat, 0.02)
    out = torch.empty((100, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(LmError) as ei:
        eng.find_pairs(dev(lon), dev(lat), 0.02, out)
    assert ei.value.code == LM_ENOSPC
    assert eng.sync_stats(raise_on_overflow=False).n_pairs == want.shape[0]     # the count is still exact
    wantset = {tuple(p) for p in want}
    assert all(tuple(sorted(p)) in wantset for p in out.cpu().numpy().tolist())   # what was written is valid


# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RPS_CASES)
def test_resolve_rps_in_the_references_own_pair_order(engine_factory, name):
    """Explicit-order resolver fed the CPython-set order + per-pair stream the golden was made with
    by the unmodified reference function."""
    g = golden(name + ".npz")
    eng = engine_factory(max_particles=g["lon"].size, max_cells=1024, max_pairs=g["pairs_ref_order"].shape[0])
    species = dev(g["species0"].copy())
    rounds = eng.resolve_rps(dev(g["pairs_ref_order"]), dev(g["u_ref"]), species, float(g["pRS"]), float(g["pPR"]),
                             float(g["pSP"]))
    assert np.array_equal(species.cpu().numpy(), g["species_ref"])
    assert 1 <= rounds < (2000 if name == "rps_knots" else 200)      # at least the largest degree: a knot of 90 has chains of hundreds


@pytest.mark.parametrize("name", RPS_CASES)
def test_resolve_rps_lexicographic_and_reversed_orders(engine_factory, name):
    g = golden(name + ".npz")
    pairs = opairs.sort_pairs(g["pairs_ref_order"])
    eng = engine_factory(max_particles=g["lon"].size, max_cells=1024, max_pairs=pairs.shape[0])
    prm = (float(g["pRS"]), float(g["pPR"]), float(g["pSP"]))
    for order in (pairs, pairs[::-1].copy()):
        u = philox.pair_uniforms(order[:, 0], order[:, 1], 5, 11)
        want, _ = orps.rps_sequential_c(g["species0"].copy(), order, u, *prm)
        species = dev(g["species0"].copy())
        eng.resolve_rps(dev(order.astype(np.int32)), dev(u), species, *prm)
        assert np.array_equal(species.cpu().numpy(), want)


# (LM_OPT_INTERACT_MODE, LM_OPT_FIND_PATH): the hybrid path (default: round-1 pipeline for the light units + rounds of
# matchings for the queued heavy units), auto / every warp of its pair search on the two-pass (dense cluster) path | the
# fused tile kernel | the round-1 pipeline alone, auto / two-pass
RESOLVE_MODES = [(2, 0), (2, 1), (1, 0), (0, 0), (0, 1)]
GOLDEN_SPECIES = {2: "species_round", 1: "species_tile", 0: "species_cell"}


@pytest.mark.parametrize("imode,mode", RESOLVE_MODES)
@pytest.mark.parametrize("name", RPS_CASES)
def test_interact_rps_canonical_order_golden(engine_factory, name, imode, mode):
    """Fused pair search + RPS on the grid the golden was made for; golden species come from the
    unmodified reference function run in the device's canonical order (cell-round order for the hybrid path, tile-round
    order for the fused tile kernel, cell-phase order for the round-1 pipeline)."""
    from lagrangian_microbes_b200._lib import LM_OPT_FIND_PATH, LM_OPT_INTERACT_MODE
    g = golden(name + ".npz")
    g = dict(g, species_cell=g[GOLDEN_SPECIES[imode]])
    n = g["lon"].size
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=g["pairs_ref_order"].shape[0] + 64)
    eng.set_option(LM_OPT_INTERACT_MODE, imode)
    eng.set_option(LM_OPT_FIND_PATH, mode)
    eng.set_grid(grid_from_golden(g))
    species = dev(g["species0"].copy())
    out = torch.empty((g["pairs_ref_order"].shape[0] + 64, 2), dtype=torch.int32, device="cuda")
    eng.interact_rps(dev(g["lon"]), dev(g["lat"]), species, float(g["r"]), float(g["pRS"]), float(g["pPR"]),
                     float(g["pSP"]), int(g["seed"]), int(g["step"]), pairs_out=out)
    st = eng.sync_stats()
    assert st.n_pairs == g["pairs_ref_order"].shape[0]
    assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), opairs.sort_pairs(g["pairs_ref_order"]))
    assert np.array_equal(species.cpu().numpy(), g["species_cell"])
    # without pair emission the species result is the same and the count is still reported
    species2 = dev(g["species0"].copy())
    eng.interact_rps(dev(g["lon"]), dev(g["lat"]), species2, float(g["r"]), float(g["pRS"]), float(g["pPR"]),
                     float(g["pSP"]), int(g["seed"]), int(g["step"]))
    assert eng.sync_stats().n_pairs == st.n_pairs
    assert np.array_equal(species2.cpu().numpy(), g["species_cell"])


@pytest.mark.parametrize("imode,mode", RESOLVE_MODES)
@pytest.mark.parametrize("n,r,p", [(100000, 0.01, (0.55, 0.55, 0.55)), (150000, 0.02, (0.5, 0.6, 0.9)),
                                   (40000, 0.05, (0.9, 0.9, 0.9)), (600000, 0.004, (0.55, 0.55, 0.55))])
def test_interact_rps_vs_oracle_live(engine_factory, n, r, p, imode, mode):
    from lagrangian_microbes_b200._lib import LM_OPT_FIND_PATH, LM_OPT_INTERACT_MODE
    rng = np.random.default_rng(n + 1)
    side = np.sqrt(n / 4900.0)
    lon = (205 + side * rng.random(n)).astype(np.float32)
    lat = (25 + side * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    want_pairs = opairs.query_pairs_reference_array(lon, lat, r)
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=want_pairs.shape[0] + 64)
    eng.set_option(LM_OPT_INTERACT_MODE, imode)
    eng.set_option(LM_OPT_FIND_PATH, mode)
    grid = auto_grid(eng, lon, lat, r, margin=0.25)       # the 600k case: 2770 x 2770 cells, several warps per row
    out = torch.empty((want_pairs.shape[0] + 64, 2), dtype=torch.int32, device="cuda")
    species = dev(sp0.copy())
    eng.interact_rps(dev(lon), dev(lat), species, r, *p, 77, 1234, pairs_out=out)
    st = eng.sync_stats()
    assert st.n_pairs == want_pairs.shape[0]
    assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), want_pairs)
    order, _ = orps.canonical_order(want_pairs, lon, lat, grid.as_dict(), mode=imode)
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 1234, 77)
    want_sp, draws = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    assert draws > 0 and np.array_equal(species.cpu().numpy(), want_sp)


# ------------------------------------------------------------------------------------------------------
def test_binning_puts_state_in_cell_id_order(engine_factory):
    rng = np.random.default_rng(9)
    n = 70000
    lon = (205 + 3 * rng.random(n)).astype(np.float32)
    lat = (25 + 2 * rng.random(n)).astype(np.float32)
    sp = rng.integers(1, 4, n).astype(np.int8)
    eng = engine_factory(max_particles=n, max_cells=1 << 22)
    grid = auto_grid(eng, lon, lat, 0.01)
    eng.state_set(dev(lon), dev(lat), dev(sp))
    blon, blat, bsp, bid, cs = [t.cpu().numpy() for t in eng.state_view()]
    cx = opairs.cell_index(blon, grid.x0, grid.inv_h, grid.ncx)
    cy = opairs.cell_index(blat, grid.y0, grid.inv_h, grid.ncy)
    key = cy * grid.ncx + cx
    assert np.all(np.diff(key) >= 0)                                    # sorted by cell
    same = np.diff(key) == 0
    assert np.all(np.diff(bid)[same] > 0)                               # by id inside a cell
    assert np.array_equal(np.sort(bid), np.arange(n))                   # a permutation
    assert np.array_equal(blon, lon[bid]) and np.array_equal(blat, lat[bid]) and np.array_equal(bsp, sp[bid])
    assert cs[0] == 0 and cs[-1] == n and np.array_equal(np.diff(cs), np.bincount(key, minlength=grid.ncx * grid.ncy))
    lon_o = torch.empty(n, dtype=torch.float32, device="cuda")
    lat_o = torch.empty(n, dtype=torch.float32, device="cuda")
    sp_o = torch.empty(n, dtype=torch.int8, device="cuda")
    eng.state_get(lon_o, lat_o, sp_o)
    assert np.array_equal(lon_o.cpu().numpy(), lon) and np.array_equal(lat_o.cpu().numpy(), lat)
    assert np.array_equal(sp_o.cpu().numpy(), sp)


def test_diffusion_kick_matches_the_philox_stream(engine_factory):
    rng = np.random.default_rng(10)
    n = 100000
    lon = (205 + 3 * rng.random(n)).astype(np.float32)
    lat = (25 + 2 * rng.random(n)).astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1024)
    amp = float(np.sqrt(6 * 3600.0 * 100 / 1e10))                        # Kh = 100 m^2/s -> 0.0147 deg
    dl, da = dev(lon.copy()), dev(lat.copy())
    eng.diffuse(dl, da, amp, 99, 17)
    ua, ub = philox.particle_uniforms(np.arange(n), 17, 99, 0)
    want_lat = (lat.astype(np.float64) + (-1.0 + 2.0 * ua) * amp).astype(np.float32)   # lat first, then lon
    want_lon = (lon.astype(np.float64) + (-1.0 + 2.0 * ub) * amp).astype(np.float32)
    assert np.array_equal(da.cpu().numpy(), want_lat) and np.array_equal(dl.cpu().numpy(), want_lon)
    kick = da.cpu().numpy().astype(np.float64) - lat
    assert abs(kick.std() - amp / np.sqrt(3)) < 0.01 * amp and np.abs(kick).max() <= amp * 1.0001


# ------------------------------------------------------------------------------------------------------
def _small_fs():
    g = golden("rk4_small.npz")
    return g, ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"])


def _set_field(eng, fs):
    eng.set_field(dev(fs.u), dev(fs.v), dev(fs.lon), dev(fs.lat))


def test_advect_rk4_matches_oracle_over_130_steps(engine_factory):
    from lagrangian_microbes_b200.particle_advecter import StageClock
    g, fs = _small_fs()
    n = g["lon0"].size
    eng = engine_factory(max_particles=n, max_cells=1024)
    _set_field(eng, fs)
    lon, lat = dev(g["lon0"].copy()), dev(g["lat0"].copy())
    clock = StageClock(fs.time)
    ref_lon, ref_lat, t, ti = g["lon0"].copy(), g["lat0"].copy(), 0.0, 0
    worst32 = worst64 = 0.0
    mismatches = 0
    eng.reset_stats()
    for step in range(int(g["steps"])):
        # single step from IDENTICAL inputs (the GPU's own previous positions)
        prev_lon, prev_lat = lon.cpu().numpy(), lat.cpu().numpy()
        eng.advect_rk4(lon, lat, clock.next_step(3600.0), 3600.0)
        a32, b32, ti_new, _ = ork4.rk4_step_f32(fs, prev_lon, prev_lat, t, 3600.0, ti)
        a64, b64, _, _ = ork4.rk4_step_f64(fs, prev_lon, prev_lat, t, 3600.0, ti)
        gl, ga = lon.cpu().numpy(), lat.cpu().numpy()
        worst32 = max(worst32, np.max(np.abs(gl - a32) / np.abs(a32)), np.max(np.abs(ga - b32) / np.abs(b32)))
        # a particle within one step of the grid edge can be out of bounds in one precision and not in
        # the other (it is then left where it was): compare with float64 where both paths moved it
        ok = ((a32 != prev_lon) | (b32 != prev_lat)) & ((a64 != prev_lon) | (b64 != prev_lat))
        worst64 = max(worst64, np.max(np.abs(gl - a64)[ok] / np.abs(a64)[ok]), np.max(np.abs(ga - b64)[ok] / np.abs(b64)[ok]))
        mismatches += int((gl != a32).sum() + (ga != b32).sum())
        t, ti = t + 3600.0, ti_new
    print("advect: worst rel vs f32-faithful %.3g, vs f64 %.3g, bitwise mismatches %d of %d"
          % (worst32, worst64, mismatches, 2 * n * int(g["steps"])))
    assert worst32 < 1e-6 and worst64 < 1e-6            # north_star tolerance
    assert mismatches <= 1e-4 * 2 * n * int(g["steps"])  # measured on B200: 0 of 1,040,000 (bit-identical)
    st = eng.sync_stats()
    assert st.n_out_of_bounds >= 5 * int(g["steps"])     # the 5 particles east of the grid, every step
    # trajectories also end where the golden (oracle-only) trajectories end, to chaotic-growth tolerance
    assert np.max(np.abs(lon.cpu().numpy() - g["lon_f64"]) / g["lon_f64"]) < 1e-4


def test_advect_known_answers(engine_factory):
    from lagrangian_microbes_b200._lib import StageTimes
    X, Y = 31, 25
    glon = (200.0 + np.arange(X) / 3.0).astype(np.float32)
    glat = (20.0 + np.arange(Y) / 3.0).astype(np.float32)
    rng = np.random.default_rng(2)
    n = 5000
    lon = (202 + 6 * rng.random(n)).astype(np.float32)
    lat = (22 + 4 * rng.random(n)).astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1024)
    st = StageTimes()
    # zero field: fixed point
    z = np.zeros((2, Y, X), dtype=np.float32)
    eng.set_field(dev(z), dev(z), dev(glon), dev(glat))
    dl, da = dev(lon.copy()), dev(lat.copy())
    eng.advect_rk4(dl, da, st, 3600.0)
    assert np.array_equal(dl.cpu().numpy(), lon) and np.array_equal(da.cpu().numpy(), lat)
    # uniform zonal flow: dlon = U0 dt / (111120 cos lat) exactly, lat untouched
    u = np.full((2, Y, X), 0.25, dtype=np.float32)
    eng.set_field(dev(u), dev(z), dev(glon), dev(glat))
    dl, da = dev(lon.copy()), dev(lat.copy())
    eng.advect_rk4(dl, da, st, 3600.0)
    want = lon.astype(np.float64) + 0.25 * 3600.0 / (111120.0 * np.cos(lat.astype(np.float64) * np.pi / 180))
    assert np.array_equal(da.cpu().numpy(), lat)
    assert np.max(np.abs(dl.cpu().numpy() - want) / want) < 1e-7
    # uniform meridional flow, time-interpolated between 0.25 and 0.75 with fraction 0.5 -> 0.5 m/s
    v = np.stack([np.full((Y, X), 0.25), np.full((Y, X), 0.75)]).astype(np.float32)
    eng.set_field(dev(z), dev(v), dev(glon), dev(glat))
    st2 = StageTimes()
    for k in range(4):
        st2.ti[k], st2.interp[k], st2.frac[k] = 0, 1, 0.5
    dl, da = dev(lon.copy()), dev(lat.copy())
    eng.advect_rk4(dl, da, st2, 3600.0)
    want = lat.astype(np.float64) + 0.5 * 3600.0 / 111120.0
    assert np.array_equal(dl.cpu().numpy(), lon)
    assert np.max(np.abs(da.cpu().numpy() - want) / want) < 1e-7


# ------------------------------------------------------------------------------------------------------
def test_fused_simulation_matches_oracle_loop(engine_factory):
    """24 fused steps (advect + interact, config-1 style: lattice patch, p = 0.55, r = 0.01 deg) checked
    step by step: positions vs the RK4 oracle from identical inputs, pairs vs cKDTree on the GPU's
    positions, species vs the reference rule in the canonical order."""
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from lagrangian_microbes_b200.particle_advecter import uniform_particle_locations

    class HostFS:
        def __init__(self, fs):
            self.u, self.v, self.lon, self.lat, self.time = fs.u, fs.v, fs.lon, fs.lat, fs.time

        def to_device(self, device):
            return tuple(torch.from_numpy(a).to(device) for a in (self.u, self.v, self.lon, self.lat))

    g, fs = _small_fs()
    n = 160 * 160
    # half a 160x80 lattice (spacing 0.0101 deg > r: it pairs up only once the flow strains it, like
    # BASELINE config 1) and half uniform-random microbes over the same box (pairs from the first step)
    rng = np.random.default_rng(4)
    l1, a1 = uniform_particle_locations(n // 2, 32.0, 33.6059, 205.0, 205.0 + 79 * 0.0101)
    lons = np.concatenate([l1, 205.0 + 1.6 * rng.random(n // 2)])
    lats = np.concatenate([a1, 32.0 + 1.6 * rng.random(n // 2)])
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    p = (0.55, 0.55, 0.55)
    sim = FusedSimulation(lons, lats, sp0, 0.01, *p, HostFS(fs), dt_seconds=3600.0, seed=5, emit_pairs=True,
                          regrid_every=4, grid_margin=0.25)
    lon_prev, lat_prev = lons.astype(np.float32), lats.astype(np.float32)
    sp_ref = sp0.copy()
    t, ti, total_pairs = 0.0, 0, 0
    for step in range(24):
        grid = sim.grid.as_dict()                    # the grid this step will bin with
        sim.step(check=True)
        npairs = sim.last_stats.n_pairs
        gl, ga, gs = sim.download()
        a32, b32, ti_new, oob = ork4.rk4_step_f32(fs, lon_prev, lat_prev, t, 3600.0, ti)
        a64, b64, _, _ = ork4.rk4_step_f64(fs, lon_prev, lat_prev, t, 3600.0, ti)
        assert oob == 0
        assert np.max(np.abs(gl - a64) / np.abs(a64)) < 1e-6 and np.max(np.abs(ga - b64) / np.abs(b64)) < 1e-6
        assert np.max(np.abs(gl - a32) / np.abs(a32)) < 1e-6 and np.max(np.abs(ga - b32) / np.abs(b32)) < 1e-6
        want_pairs = opairs.query_pairs_reference_array(gl, ga, 0.01)
        assert npairs == want_pairs.shape[0]
        got_pairs = opairs.sort_pairs(sim.pairs[:npairs].cpu().numpy())
        assert np.array_equal(got_pairs, want_pairs)
        order, _ = orps.canonical_order(want_pairs, gl, ga, grid)
        u = philox.pair_uniforms(order[:, 0], order[:, 1], step, 5)
        sp_ref, _ = orps.rps_sequential_c(sp_ref, order, u, *p)
        assert np.array_equal(gs, sp_ref)
        lon_prev, lat_prev, t, ti = gl, ga, t + 3600.0, ti_new
        total_pairs += npairs
    print("fused: %d pairs over 24 steps, %d species changed" % (total_pairs, int((sp_ref != sp0).sum())))
    assert total_pairs > 0
    assert np.bincount(sp_ref, minlength=4)[1:].sum() == n


def test_overlapped_record_and_side_stream_phases_match_the_serial_path(engine_factory):
    """lm_record_next_step (record scattered + copied under the step) and LM_OPT_OVERLAP (RPS phases of step k
    under the advection of step k+1) against the same run with everything on one stream."""
    from lagrangian_microbes_b200._lib import LM_OPT_OVERLAP
    from lagrangian_microbes_b200.simulation import FusedSimulation

    class HostFS:
        def __init__(self, fs):
            self.u, self.v, self.lon, self.lat, self.time = fs.u, fs.v, fs.lon, fs.lat, fs.time

        def to_device(self, device):
            return tuple(torch.from_numpy(a).to(device) for a in (self.u, self.v, self.lon, self.lat))

    g, fs = _small_fs()
    n = 120000
    rng = np.random.default_rng(12)
    lons = (201.0 + 2.0 * rng.random(n)).astype(np.float32)
    lats = (31.0 + 2.0 * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    p = (0.55, 0.6, 0.9)
    sims = []
    for overlap in (1, 0):
        sim = FusedSimulation(lons, lats, sp0, 0.01, *p, HostFS(fs), dt_seconds=3600.0, seed=9, emit_pairs=True,
                              regrid_every=4, grid_margin=0.25)
        sim.engine.set_option(LM_OPT_OVERLAP, overlap)
        sims.append(sim)
    rec = [tuple(torch.empty(n, dtype=dt).pin_memory() for dt in (torch.float32, torch.float32, torch.int8)) for _ in range(2)]
    for step in range(10):
        sims[0].step(record=rec[step & 1])
        sims[1].step()
        wl, wa, ws = sims[1].download()
        sims[0].engine.host_copies_sync()
        r = rec[step & 1]
        assert np.array_equal(r[0].numpy(), wl) and np.array_equal(r[1].numpy(), wa), "recorded positions, step %d" % step
        assert np.array_equal(r[2].numpy(), ws), "recorded species, step %d" % step
    gl, ga, gs = sims[0].download()
    assert np.array_equal(gl, wl) and np.array_equal(ga, wa) and np.array_equal(gs, ws)
    assert int((ws != sp0).sum()) > 1000


@pytest.mark.parametrize("imode,mode", RESOLVE_MODES)
def test_interact_rps_dense_clusters_take_the_warp_cooperative_path(engine_factory, imode, mode):
    """Clusters of ~1500 microbes inside one or two cells: candidate pairs per unit >> HEAVY_TESTS, so the
    units are resolved by the whole warp (prefix scan over 3->3 species maps) and the pair search takes
    its direct (unstaged) path.  Same bit-exact bar."""
    rng = np.random.default_rng(21)
    centres = np.array([[207.003, 30.004], [207.5, 30.5], [208.0099, 31.0001]])      # the last straddles cell edges
    pts = [c + rng.normal(0, 0.002, (1500, 2)) for c in centres]
    pts.append(np.array([206.5, 29.5]) + 2.0 * rng.random((20000, 2)))              # background
    pts = np.concatenate(pts)
    perm = rng.permutation(pts.shape[0])
    lon, lat = pts[perm, 0].astype(np.float32), pts[perm, 1].astype(np.float32)
    n = lon.size
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    sp0[::97] = 0                                                                       # a few non-RPS species
    r, p = 0.01, (0.55, 0.6, 0.5)
    want_pairs = opairs.query_pairs_reference_array(lon, lat, r)
    assert want_pairs.shape[0] > 2_000_000
    from lagrangian_microbes_b200._lib import LM_OPT_FIND_PATH, LM_OPT_INTERACT_MODE
    eng = engine_factory(max_particles=n, max_cells=1 << 20, max_pairs=want_pairs.shape[0] + 64)
    eng.set_option(LM_OPT_INTERACT_MODE, imode)
    eng.set_option(LM_OPT_FIND_PATH, mode)
    grid = auto_grid(eng, lon, lat, r, margin=0.1)
    out = torch.empty((want_pairs.shape[0] + 64, 2), dtype=torch.int32, device="cuda")
    species = dev(sp0.copy())
    eng.interact_rps(dev(lon), dev(lat), species, r, *p, 3, 8, pairs_out=out)
    st = eng.sync_stats()
    assert st.n_pairs == want_pairs.shape[0]
    assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), want_pairs)
    order, _ = orps.canonical_order(want_pairs, lon, lat, grid.as_dict(), mode=imode)
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 8, 3)
    want_sp, _ = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    assert np.array_equal(species.cpu().numpy(), want_sp)


def test_rps_hand_off_capacity_overflow_is_reported(engine_factory):
    from lagrangian_microbes_b200._lib import LmError, LM_ENOSPC
    rng = np.random.default_rng(22)
    n = 5000
    lon = (205 + 0.5 * rng.random(n)).astype(np.float32)
    lat = (25 + 0.5 * rng.random(n)).astype(np.float32)
    eng = engine_factory(max_particles=n, max_cells=1 << 20, max_pairs=100)           # far too small
    from lagrangian_microbes_b200._lib import LM_OPT_INTERACT_MODE
    eng.set_option(LM_OPT_INTERACT_MODE, 0)      # only the round-1 pipeline has a hand-off; the fused tile kernel has none
    auto_grid(eng, lon, lat, 0.02)
    species = dev(rng.integers(1, 4, n).astype(np.int8))
    eng.interact_rps(dev(lon), dev(lat), species, 0.02, 0.5, 0.5, 0.5, 0, 0)
    with pytest.raises(LmError) as ei:
        eng.sync_stats()
    assert ei.value.code == LM_ENOSPC
