"""The RPS resolver's tuning knobs must never change a result: LM_OPT_RESOLVE_BATCH (pairs a lane loads ahead per
iteration of its stream walk, with in-batch forwarding of rewritten species), LM_OPT_RESOLVE_HEAVY_MIN (when a unit
goes to the whole warp) and LM_OPT_RESOLVE_UPL, each against the reference rule run in the canonical cell-phase order
(oracle/rps.py restating interactions.py:13-40 / interaction_simulator.py:104-105)."""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _round1_pipeline(monkeypatch):
    """These tests cover the round-1 pipeline (pair search -> hand-off -> nine phase launches / tiled resolver,
    csrc/pairs.cu, cell-phase order), still shipped as LM_OPT_INTERACT_MODE = 0."""
    from lagrangian_microbes_b200.engine import Engine
    monkeypatch.setattr(Engine, "DEFAULT_INTERACT_MODE", 0)

# (LM_OPT_RESOLVE_BATCH, LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_UPL)
SETTINGS = [(4, 0, 0), (8, 0, 0), (4, 24, 1), (8, 0xffff, 2), (1, 24, 4), (8, 8, 8)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def configure(eng, batch, heavy_min, upl):
    from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_BATCH, LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_UPL
    eng.set_option(LM_OPT_RESOLVE_BATCH, batch)
    eng.set_option(LM_OPT_RESOLVE_HEAVY_MIN, heavy_min)
    eng.set_option(LM_OPT_RESOLVE_UPL, upl)


@pytest.mark.parametrize("name", ["rps_uniform", "rps_clustered", "rps_oddspecies"])
def test_golden_species_under_every_setting(engine_factory, name):
    from lagrangian_microbes_b200._lib import Grid
    g = golden(name + ".npz")
    n = g["lon"].size
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=g["pairs_ref_order"].shape[0] + 64)
    eng.set_grid(Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1])))
    lon, lat = dev(g["lon"]), dev(g["lat"])
    for setting in SETTINGS:
        configure(eng, *setting)
        species = dev(g["species0"].copy())
        eng.interact_rps(lon, lat, species, float(g["r"]), float(g["pRS"]), float(g["pPR"]), float(g["pSP"]),
                         int(g["seed"]), int(g["step"]))
        assert eng.sync_stats().n_pairs == g["pairs_ref_order"].shape[0]
        assert np.array_equal(species.cpu().numpy(), g["species_cell"]), "setting %r" % (setting,)


def _cloud(kind, rng):
    if kind == "uniform":                                   # ~6 pairs per microbe, short units
        n = 150000
        side = np.sqrt(n / 4900.0)
        return 205 + side * rng.random(n), 25 + side * rng.random(n), 0.02
    if kind == "crowded":                                   # ~40 microbes per cell: every unit has hundreds of pairs
        n = 40000
        return 205 + 0.3 * rng.random(n), 25 + 0.3 * rng.random(n), 0.01
    # a few cells with 10-25 microbes in a sparse background: the units that sit just under the warp limit
    n = 60000
    lon, lat = 205 + 3.0 * rng.random(n), 25 + 3.0 * rng.random(n)
    k = 0
    for c in range(300):
        m = int(rng.integers(10, 26))
        lon[k:k + m] = 205.005 + 0.01 * int(rng.integers(0, 290)) + 0.004 * rng.random(m)
        lat[k:k + m] = 25.005 + 0.01 * int(rng.integers(0, 290)) + 0.004 * rng.random(m)
        k += m
    return lon, lat, 0.01


@pytest.mark.parametrize("kind", ["uniform", "crowded", "knots"])
def test_live_species_under_every_setting(engine_factory, kind):
    from lagrangian_microbes_b200.engine import make_grid
    rng = np.random.default_rng(11)
    lon, lat, r = _cloud(kind, rng)
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    n = lon.size
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    sp0[::53] = 0                                           # a few species outside {1, 2, 3}: draw, no winner
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
    for setting in [(1, 0, 0)] + SETTINGS:
        configure(eng, *setting)
        species = dev(sp0.copy())
        eng.interact_rps(lon_d, lat_d, species, r, *p, 5, 17)
        assert eng.sync_stats().n_pairs == want_pairs.shape[0]
        bad = int((species.cpu().numpy() != want_sp).sum())
        assert bad == 0, "setting %r: %d species differ" % (setting, bad)


def test_resolver_options_reject_other_values(engine_factory):
    from lagrangian_microbes_b200._lib import LM_EINVAL, LM_OPT_RESOLVE_BATCH, LM_OPT_RESOLVE_HEAVY_MIN, LmError
    eng = engine_factory(max_particles=64, max_cells=1024)
    for opt, bad in ((LM_OPT_RESOLVE_BATCH, 0), (LM_OPT_RESOLVE_BATCH, 2), (LM_OPT_RESOLVE_BATCH, 16),
                     (LM_OPT_RESOLVE_HEAVY_MIN, -1), (LM_OPT_RESOLVE_HEAVY_MIN, 1 << 16)):
        with pytest.raises(LmError) as ei:
            eng.set_option(opt, bad)
        assert ei.value.code == LM_EINVAL
