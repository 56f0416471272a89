"""Generate the golden fixtures in tests/golden/ FROM THE UNMODIFIED REFERENCE.

Run in the authoring container (needs /root/reference and SciPy):

    python tests/golden/make_golden.py

What is pinned by what:
  rps_*.npz    species after the reference's own sequential loop
               (interaction_simulator.py:104-105) over the reference's own function
               (interactions.py:13-40, imported from /root/reference, NOT restated), with
               ``np.random.rand`` patched to return the per-pair uniform u[k] so that the stream is
               per pair rather than per draw.  Two orders per case:
                 *_ref   the CPython-set iteration order of ``cKDTree.query_pairs`` -- literally
                         what the reference iterates over;
                 *_cell  the canonical cell-phase order of the round-1 device pipeline (LM_OPT_INTERACT_MODE = 0)
                         for a fixed grid;
                 *_round the canonical cell-round order of the hybrid device path (LM_OPT_INTERACT_MODE = 2, the default:
                         oracle/rps.py::cell_round_order) for the same grid;
                 *_tile  the canonical tile-round order of the fused tile kernel (LM_OPT_INTERACT_MODE = 1,
                         oracle/rps.py::tile_round_order) for the same grid.
  pairs_*.npz  pair sets from ``cKDTree(...).query_pairs(r, p=2)`` -- the library call the
               reference makes (interaction_simulator.py:93,98).
  rk4_*.npz    outputs of oracle/rk4.py (regression only -- parity unpinned, see that file).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import interactions as ref_interactions  # noqa: E402  (the reference module, unmodified)

from oracle import pairs as opairs  # noqa: E402
from oracle import philox  # noqa: E402
from oracle import rk4 as ork4  # noqa: E402
from oracle import rps as orps  # noqa: E402


def run_reference_rps(species0, pair_list, u, params):
    """The reference loop, verbatim, on a copy of species0."""
    props = {"species": species0.copy()}
    cursor = {"k": 0}
    real_rand = ref_interactions.np.random.rand
    ref_interactions.np.random.rand = lambda: float(u[cursor["k"]])
    try:
        for k, pair in enumerate(pair_list):
            cursor["k"] = k
            ref_interactions.rock_paper_scissors_interaction(params, props, pair[0], pair[1])
    finally:
        ref_interactions.np.random.rand = real_rand
    return props["species"]


def fixed_grid(lon, lat, r, k=1):
    h = r * (1.0 + 2.0 ** -20) * k
    x0, y0 = float(np.floor(lon.min())), float(np.floor(lat.min()))
    ncx = int((float(lon.max()) - x0) / h) + 2
    ncy = int((float(lat.max()) - y0) / h) + 2
    return dict(x0=x0, y0=y0, inv_h=1.0 / h, ncx=ncx, ncy=ncy)


def rps_case(name, lon, lat, species0, r, params, seed, step, grid_k=1):
    lon = lon.astype(np.float32)
    lat = lat.astype(np.float32)
    pair_set = opairs.query_pairs_reference(lon, lat, r)            # the reference's set
    ref_order = np.array(list(pair_set), dtype=np.int64).reshape(-1, 2)   # CPython set iteration order
    u_ref = philox.pair_uniforms(ref_order[:, 0], ref_order[:, 1], step, seed)
    species_ref = run_reference_rps(species0, [tuple(map(int, p)) for p in ref_order], u_ref, params)

    grid = fixed_grid(lon, lat, r, grid_k)
    cell_order, phases = orps.cell_phase_order(opairs.pairs_from_set(pair_set), lon, lat, grid)
    u_cell = philox.pair_uniforms(cell_order[:, 0], cell_order[:, 1], step, seed)
    species_cell = run_reference_rps(species0, [tuple(map(int, p)) for p in cell_order], u_cell, params)

    round_order, _ = orps.cell_round_order(opairs.pairs_from_set(pair_set), lon, lat, grid)
    u_round = philox.pair_uniforms(round_order[:, 0], round_order[:, 1], step, seed)
    species_round = run_reference_rps(species0, [tuple(map(int, p)) for p in round_order], u_round, params)

    tile_order, _ = orps.tile_round_order(opairs.pairs_from_set(pair_set), lon, lat, grid)
    u_tile = philox.pair_uniforms(tile_order[:, 0], tile_order[:, 1], step, seed)
    species_tile = run_reference_rps(species0, [tuple(map(int, p)) for p in tile_order], u_tile, params)

    np.savez_compressed(os.path.join(HERE, name + ".npz"), lon=lon, lat=lat, r=np.float64(r), species0=species0,
                        pairs_tile_order=tile_order.astype(np.int32), species_tile=species_tile,
                        pairs_round_order=round_order.astype(np.int32), species_round=species_round,
                        pRS=params["pRS"], pPR=params["pPR"], pSP=params["pSP"], seed=np.int64(seed),
                        step=np.int64(step), pairs_ref_order=ref_order.astype(np.int32), u_ref=u_ref,
                        species_ref=species_ref, grid=np.array([grid["x0"], grid["y0"], grid["inv_h"]]),
                        grid_n=np.array([grid["ncx"], grid["ncy"]], dtype=np.int32),
                        pairs_cell_order=cell_order.astype(np.int32), species_cell=species_cell)
    print("%-16s N=%d P=%d changed(ref)=%d changed(cell)=%d changed(tile)=%d" % (
        name, lon.size, ref_order.shape[0], int((species_ref != species0).sum()), int((species_cell != species0).sum()),
        int((species_tile != species0).sum())))


def make_rps():
    rng = np.random.default_rng(1)
    n = 3000
    lon = 205.0 + rng.random(n)
    lat = 25.0 + rng.random(n)
    np.random.seed(0)
    _, params, props = ref_interactions.rock_paper_scissors(n, 0.55, 0.55, 0.55)      # the reference factory
    rps_case("rps_uniform", lon, lat, props["species"], 0.03, params, seed=42, step=7)

    # clustered with exact duplicates, asymmetric probabilities, coarser grid (k=2)
    n = 2000
    centres = rng.random((12, 2)) * np.array([2.0, 1.0]) + np.array([210.0, 30.0])
    which = rng.integers(0, 12, n)
    pts = centres[which] + rng.normal(0.0, 0.02, (n, 2))
    pts[::50] = pts[1::50][: pts[::50].shape[0]]                      # duplicates
    np.random.seed(1)
    _, params, props = ref_interactions.rock_paper_scissors(n, 0.5, 0.6, 0.9)
    rps_case("rps_clustered", pts[:, 0], pts[:, 1], props["species"], 0.01, params, seed=2**40 + 5, step=2**33 + 1, grid_k=2)

    # species outside {1,2,3}: the reference draws, finds no winner, changes nothing
    n = 600
    lon = 200.0 + 0.3 * rng.random(n)
    lat = 10.0 + 0.3 * rng.random(n)
    sp = rng.integers(0, 5, n).astype(np.int8)
    rps_case("rps_oddspecies", lon, lat, sp, 0.02, {"pRS": 0.3, "pPR": 0.7, "pSP": 0.5}, seed=3, step=0)

    # knots: 50-90 microbes inside one or two cells on a sparse background -- HEAVY units of the hybrid path (more than 1,024
    # candidate pairs: rounds of matchings) next to light ones
    n = 800
    lon = 207.0 + 0.4 * rng.random(n)
    lat = 31.0 + 0.4 * rng.random(n)
    k = 0
    for m, (cxk, cyk, sig) in zip((90, 70, 50), ((207.105, 31.105, 0.0015), (207.2999, 31.2501, 0.003), (207.25, 31.1, 0.002))):
        lon[k:k + m] = cxk + sig * rng.standard_normal(m)
        lat[k:k + m] = cyk + sig * rng.standard_normal(m)
        k += m
    perm = rng.permutation(n)
    np.random.seed(2)
    _, params, props = ref_interactions.rock_paper_scissors(n, 0.55, 0.6, 0.5)
    rps_case("rps_knots", lon[perm], lat[perm], props["species"], 0.01, params, seed=11, step=3)


def make_pairs():
    rng = np.random.default_rng(2)
    cases = {}
    # 3-4-5 triangles in float32-exact coordinates: distance exactly r (inclusive predicate)
    base = np.array([[0.0, 0.0], [0.1875, 0.25], [0.375, 0.5], [0.1875, 0.0], [0.375, 0.25]]) + np.array([208.0, 30.0])
    cases["exact345"] = (base[:, 0], base[:, 1], 0.3125)
    # duplicates + collinear
    x = np.concatenate([np.full(20, 210.5), np.linspace(210.0, 210.2, 41)])
    y = np.concatenate([np.full(20, 31.25), np.full(41, 31.0)])
    cases["dups_collinear"] = (x, y, 0.01)
    # a piece of the config-1 lattice (spacing 10/699 > r: no pairs)
    g = np.linspace(205, 215, 700)[:40]
    gx, gy = np.meshgrid(g, np.linspace(25, 35, 700)[:40])
    cases["lattice"] = (gx.ravel(), gy.ravel(), 0.01)
    # uniform random at config-1 density
    n = 20000
    side = np.sqrt(n / 4900.0)
    cases["uniform"] = (205 + side * rng.random(n), 25 + side * rng.random(n), 0.01)
    # near-zero latitudes: squares are not exactly representable -> exercises the rounding order
    cases["tiny_lat"] = (180 + 0.2 * rng.random(4000), 1e-3 * rng.random(4000), 2e-5)
    cases["empty"] = (np.zeros(0), np.zeros(0), 0.01)
    cases["single"] = (np.array([210.0]), np.array([30.0]), 0.01)
    out = {}
    for name, (lon, lat, r) in cases.items():
        lon = lon.astype(np.float32)
        lat = lat.astype(np.float32)
        if lon.size >= 2:
            pr = opairs.pairs_from_set(opairs.query_pairs_reference(lon, lat, r))
        else:
            pr = np.zeros((0, 2), dtype=np.int64)
        out[name + "_lon"], out[name + "_lat"], out[name + "_r"] = lon, lat, np.float64(r)
        out[name + "_pairs"] = pr.astype(np.int32)
        print("pairs %-16s N=%d P=%d" % (name, lon.size, pr.shape[0]))
    np.savez_compressed(os.path.join(HERE, "pairs_cases.npz"), **out)


def make_pairs_norms():
    """pairs_norms.npz: ``cKDTree(...).query_pairs(r, p)`` for the other norms SciPy evaluates without pow():
    p = 1 and p = inf (interaction_norm, interaction_simulator.py:27,98)."""
    rng = np.random.default_rng(12)
    cases = {}
    # float32-exact coordinates at distance EXACTLY r in the 1-norm (0.125 + 0.25) / in the max norm (0.375)
    base = np.array([[0.0, 0.0], [0.125, 0.25], [0.25, 0.5], [0.375, 0.0], [0.375, 0.375], [0.25, 0.125],
                     [0.75, 0.375], [0.375, -0.375]]) + np.array([208.0, 30.0])
    cases["exact"] = (base[:, 0], base[:, 1], 0.375)
    n = 20000
    side = np.sqrt(n / 4900.0)
    cases["uniform"] = (205 + side * rng.random(n), 25 + side * rng.random(n), 0.01)
    # a dense blob: cells with hundreds of particles (the pair search's two-pass path)
    cases["blob"] = (210 + 0.006 * rng.standard_normal(700), 30 + 0.006 * rng.standard_normal(700), 0.01)
    # near-zero latitudes, both signs: differences are not Sterbenz-exact in float32
    cases["tiny_lat"] = (180 + 0.2 * rng.random(4000), 1e-3 * (rng.random(4000) - 0.5), 2e-5)
    x = np.concatenate([np.full(20, 210.5), np.linspace(210.0, 210.2, 41)])
    y = np.concatenate([np.full(20, 31.25), np.full(41, 31.0)])
    cases["dups_collinear"] = (x, y, 0.01)
    out = {}
    for name, (lon, lat, r) in cases.items():
        lon = lon.astype(np.float32)
        lat = lat.astype(np.float32)
        out[name + "_lon"], out[name + "_lat"], out[name + "_r"] = lon, lat, np.float64(r)
        for tag, p in (("p1", 1), ("pinf", np.inf)):
            pr = opairs.pairs_from_set(opairs.query_pairs_reference(lon, lat, r, p=p))
            out["%s_pairs_%s" % (name, tag)] = pr.astype(np.int32)
            print("pairs %-16s p=%-4s N=%d P=%d" % (name, p, lon.size, pr.shape[0]))
    np.savez_compressed(os.path.join(HERE, "pairs_norms.npz"), **out)


def small_fieldset(seed=0, T=5, Y=31, X=46, land=True):
    from lagrangian_microbes_b200.velocity_fields import synthetic_uv
    lon = (200.0 + np.arange(X) / 3.0).astype(np.float32)
    lat = (40.0 - np.arange(Y) / 3.0).astype(np.float32)            # descending, like OSCAR
    time = np.array([0.0, 400000.0, 432000.0 * 2, 432000.0 * 3.5, 432000.0 * 5])[:T]   # irregular
    u, v = synthetic_uv(lon, lat, time, kind="random_fourier", seed=seed, n_modes=12, rms_speed=0.3, land=land)
    return lon, lat, time, u, v


def make_rk4():
    lon, lat, time, u, v = small_fieldset()
    fs = ork4.FieldSet(lon, lat, time, u, v)
    rng = np.random.default_rng(3)
    n = 4000
    plon = (201.0 + 13.0 * rng.random(n)).astype(np.float32)
    plat = (31.0 + 8.0 * rng.random(n)).astype(np.float32)
    plon[:5] = np.float32(215.5)              # beyond the grid's east edge (215.0): out of bounds
    plon[5:10] = fs.lon[7:12]                 # exactly on grid lines
    plat[5:10] = fs.lat[3:8]
    ti = 0
    t = 0.0
    l32, a32 = plon.copy(), plat.copy()
    l64, a64 = plon.astype(np.float64), plat.astype(np.float64)
    steps = 130                               # crosses the 400000 s snapshot (step 112)
    for _ in range(steps):
        l32, a32, ti_new, _ = ork4.rk4_step_f32(fs, l32, a32, t, 3600.0, ti)
        l64, a64, _, _ = ork4.rk4_step_f64(fs, l64, a64, t, 3600.0, ti)
        ti, t = ti_new, t + 3600.0
    np.savez_compressed(os.path.join(HERE, "rk4_small.npz"), grid_lon=lon, grid_lat=lat, grid_time=time, u=u, v=v,
                        lon0=plon, lat0=plat, steps=np.int32(steps), lon_f32=l32, lat_f32=a32, lon_f64=l64, lat_f64=a64)
    rel = np.max(np.abs(l32 - l64) / np.abs(l64))
    print("rk4_small: %d steps, max rel |f32-f64| lon = %.3g, moved max %.3f deg" % (steps, rel, np.abs(l64 - plon).max()))


if __name__ == "__main__":
    only = sys.argv[1:]                       # e.g. ``make_golden.py pairs_norms``; default: everything
    for name, fn in (("rps", make_rps), ("pairs", make_pairs), ("pairs_norms", make_pairs_norms), ("rk4", make_rk4)):
        if not only or name in only:
            fn()
