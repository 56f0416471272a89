"""GPU parity of the radius query under the other Minkowski norms SciPy evaluates without pow():
``query_pairs(r, p=interaction_norm)`` with p = 1 and p = inf (interaction_simulator.py:27,98; LM_OPT_NORM).
Bar: the pair set is bit-exact vs cKDTree (golden fixtures made with SciPy + the live call), and the species of
the fused path are bit-exact vs the reference rule run over those pairs in the canonical cell-phase order."""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

NORM_CASES = ["exact", "uniform", "blob", "tiny_lat", "dups_collinear"]
NORMS = [("p1", 1), ("pinf", np.inf)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fit_grid(eng, lon, lat, r, **kw):
    from lagrangian_microbes_b200.engine import make_grid
    g = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), r, lon.size, eng.max_cells, **kw)
    eng.set_grid(g)
    return g


def gpu_pairs(eng, lon, lat, r, cap):
    out = torch.empty((int(cap), 2), dtype=torch.int32, device="cuda")
    n_pairs = eng.find_pairs(dev(lon), dev(lat), r, out)
    return opairs.sort_pairs(out[:n_pairs].cpu().numpy())


# (LM_OPT_INTERACT_MODE, LM_OPT_FIND_PATH): fused tile kernel | round-1 pair search, auto / every warp on the two-pass path
@pytest.mark.parametrize("imode,mode", [(2, 0), (2, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("tag,p", NORMS)
@pytest.mark.parametrize("name", NORM_CASES)
def test_find_pairs_other_norms_golden(engine_factory, name, tag, p, imode, mode):
    from lagrangian_microbes_b200._lib import LM_OPT_FIND_PATH, LM_OPT_INTERACT_MODE
    g = golden("pairs_norms.npz")
    lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
    want = g["%s_pairs_%s" % (name, tag)].astype(np.int64)
    eng = engine_factory(max_particles=max(lon.size, 16), max_cells=1 << 20)
    eng.set_norm(p)
    eng.set_option(LM_OPT_INTERACT_MODE, imode)
    eng.set_option(LM_OPT_FIND_PATH, mode)
    fit_grid(eng, lon, lat, r, margin=0.0)
    assert np.array_equal(gpu_pairs(eng, lon, lat, r, want.shape[0] + 64), want)
    # the same handle switched back: the Euclidean set lies between the two
    eng.set_norm(2)
    p2 = gpu_pairs(eng, lon, lat, r, g["%s_pairs_pinf" % name].shape[0] + 64)
    assert np.array_equal(p2, opairs.query_pairs_reference_array(lon, lat, r) if lon.size > 1 else p2)
    assert g["%s_pairs_p1" % name].shape[0] <= p2.shape[0] <= g["%s_pairs_pinf" % name].shape[0]


@pytest.mark.parametrize("p", [1, np.inf])
@pytest.mark.parametrize("n,r,kind", [(200000, 0.01, "uniform"), (100000, 0.005, "clustered"), (30000, 0.05, "line")])
def test_find_pairs_other_norms_vs_live_ckdtree(engine_factory, n, r, kind, p):
    rng = np.random.default_rng(n + 3)
    if kind == "uniform":
        side = np.sqrt(n / 4900.0)
        lon, lat = 205 + side * rng.random(n), -0.5 * side + side * rng.random(n)          # straddles the equator
    elif kind == "clustered":
        c = rng.random((40, 2)) * 4 + np.array([208.0, 28.0])
        pts = c[rng.integers(0, 40, n)] + rng.normal(0, 0.03, (n, 2))
        lon, lat = pts[:, 0], pts[:, 1]
    else:
        lon, lat = 200 + 3 * rng.random(n), np.full(n, 12.5) + 1e-4 * rng.random(n)
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    want = opairs.query_pairs_reference_array(lon, lat, r, p=p)
    eng = engine_factory(max_particles=n, max_cells=1 << 22)
    eng.set_norm(p)
    fit_grid(eng, lon, lat, r, margin=0.1)
    got = gpu_pairs(eng, lon, lat, r, want.shape[0] + 1024)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("p", [1, np.inf])
def test_interact_rps_other_norms_vs_oracle(engine_factory, p):
    """Fused search + RPS: the pairs of that norm, resolved in the canonical order, equal the reference rule."""
    n, r, prob = 120000, 0.012, (0.5, 0.6, 0.9)
    rng = np.random.default_rng(5)
    side = np.sqrt(n / 4900.0)
    lon = (205 + side * rng.random(n)).astype(np.float32)
    lat = (25 + side * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    want = opairs.query_pairs_reference_array(lon, lat, r, p=p)
    eng = engine_factory(max_particles=n, max_cells=1 << 22, max_pairs=want.shape[0] + 64)
    eng.set_norm(p)
    grid = fit_grid(eng, lon, lat, r, margin=0.25)
    out = torch.empty((want.shape[0] + 64, 2), dtype=torch.int32, device="cuda")
    species = dev(sp0.copy())
    eng.interact_rps(dev(lon), dev(lat), species, r, *prob, 9, 31, pairs_out=out)
    st = eng.sync_stats()
    assert st.n_pairs == want.shape[0]
    assert np.array_equal(opairs.sort_pairs(out[:st.n_pairs].cpu().numpy()), want)
    order, _ = orps.canonical_order(want, lon, lat, grid.as_dict())
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 31, 9)
    want_sp, draws = orps.rps_sequential_c(sp0.copy(), order, u, *prob)
    assert draws > 0 and np.array_equal(species.cpu().numpy(), want_sp)


def test_fused_simulation_and_strips_with_the_max_norm():
    """interaction_norm reaches lm_step and the staged (strip) step: 3 strips on one device == a single handle,
    and the pairs of every step are cKDTree's for p = inf on the positions of that step."""
    from test_gpu_strips import P, R, compare_step, particles, single, small_fs
    from lagrangian_microbes_b200.strips import LocalTransport, StripSet
    from lagrangian_microbes_b200.simulation import FusedSimulation
    G, n, seed = 3, 20000, 4
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered=True)
    ids = np.arange(n, dtype=np.int32)
    cut = [slice(g, n, G) for g in range(G)]
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=60 * G, grid_margin=0.25, regrid_every=0, interaction_norm=np.inf)
    sim = single(lon, lat, sp, ss.grid, fs, seed)
    sim.engine.set_norm(np.inf)
    euclid = FusedSimulation(lon, lat, sp, R, *P, fs, seed=seed, pair_capacity=60 * n, regrid_every=0, grid_margin=0.25)
    for step in range(3):
        compare_step(ss, sim, step)
        wl, wa, _ = sim.download()
        want = opairs.query_pairs_reference_array(wl, wa, R, p=np.inf)
        n_pairs = sim.last_stats.n_pairs
        assert np.array_equal(opairs.sort_pairs(sim.pairs[:n_pairs].cpu().numpy()), want)
        assert euclid.step(check=True).n_pairs < n_pairs             # the max-norm ball is the larger one
    ss.close()


def test_norm_option_rejects_other_values(engine_factory):
    from lagrangian_microbes_b200._lib import LM_EINVAL, LM_OPT_NORM, LmError
    eng = engine_factory(max_particles=64, max_cells=1024)
    for bad in (3, -1, 7):
        with pytest.raises(LmError) as ei:
            eng.set_option(LM_OPT_NORM, bad)
        assert ei.value.code == LM_EINVAL
    with pytest.raises(NotImplementedError):
        eng.set_norm(3)
