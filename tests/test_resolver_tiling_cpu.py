"""The tiled resolver (csrc/pairs.cu::resolve_tiled_kernel, LM_OPT_RESOLVE_MODE = 1) rests on one claim: a tile of cells
plus a halo of 6 columns and 2 rows, taken from the species BEFORE the first phase and run through every unit that lies
completely inside the loaded region in canonical order, ends with exactly the global sequential result on the tile's
interior (tile origins on even rows and columns).  This is that scheme in NumPy + the C restatement of the reference
loop (oracle/rps.py; interactions.py:13-40, interaction_simulator.py:104-105), for the three phase ranges the library
launches: 0-8 (single handle), 0-5 and 6-8 (either side of a strip's halo exchange)."""
import numpy as np
import pytest

from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps
from oracle.pairs import cell_index

TILE_X, TILE_Y, HX, HY = 64, 16, 6, 2          # csrc/pairs.cu


def _case(seed, ncx, ncy, per_cell, clustered):
    rng = np.random.default_rng(seed)
    r = 0.01
    h = r * (1 + 2.0 ** -20)
    grid = dict(x0=200.0, y0=30.0, inv_h=1.0 / h, ncx=ncx, ncy=ncy)
    n = int(ncx * ncy * per_cell)
    lon = 200.0 + ncx * h * rng.random(n)
    lat = 30.0 + ncy * h * rng.random(n)
    if clustered:                                  # knots of 10-30 microbes: long units next to tile borders
        k = 0
        for _ in range(40):
            m = int(rng.integers(10, 31))
            lon[k:k + m] = 200.0 + h * (int(rng.integers(0, ncx)) + rng.random(m))
            lat[k:k + m] = 30.0 + h * (int(rng.integers(0, ncy)) + rng.random(m))
            k += m
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    pairs = opairs.query_pairs_reference_array(lon, lat, r)
    order, phase = orps.cell_phase_order(pairs, lon, lat, grid)
    cx = cell_index(lon, grid["x0"], grid["inv_h"], ncx)
    cy = cell_index(lat, grid["y0"], grid["inv_h"], ncy)
    return sp0, order, phase, cx, cy


def _tiled(sp_before, order, u, cx, cy, ncx, ncy, tile_x, tile_y, hx, hy, p):
    out = sp_before.copy()
    for ty in range(0, ncy, tile_y):
        for tx in range(0, ncx, tile_x):
            loaded = (cx >= tx - hx) & (cx < tx + tile_x + hx) & (cy >= ty - hy) & (cy < ty + tile_y + hy)
            sel = loaded[order[:, 0]] & loaded[order[:, 1]]         # units completely inside the loaded region
            got, _ = orps.rps_sequential_c(sp_before.copy(), order[sel], u[sel], *p)
            interior = (cx >= tx) & (cx < tx + tile_x) & (cy >= ty) & (cy < ty + tile_y)
            out[interior] = got[interior]
    return out


@pytest.mark.parametrize("first,last", [(0, 8), (0, 5), (6, 8)])
@pytest.mark.parametrize("clustered", [False, True])
def test_tile_plus_halo_reproduces_the_sequential_order(first, last, clustered):
    p = (0.55, 0.6, 0.9)
    # small tiles (same halo) so that a modest grid has many tile borders; then the kernel's own tile size
    for seed, (tile_x, tile_y, ncx, ncy) in enumerate([(8, 4, 42, 23), (16, 6, 50, 20), (TILE_X, TILE_Y, 150, 37)]):
        sp0, order, phase, cx, cy = _case(100 + seed, ncx, ncy, 2.5, clustered)
        u = philox.pair_uniforms(order[:, 0], order[:, 1], 3, 9)
        before = order[phase < first]
        sp_before, _ = orps.rps_sequential_c(sp0.copy(), before, u[phase < first], *p)
        rng_sel = (phase >= first) & (phase <= last)
        want, draws = orps.rps_sequential_c(sp_before.copy(), order[rng_sel], u[rng_sel], *p)
        assert draws > 100
        got = _tiled(sp_before, order[rng_sel], u[rng_sel], cx, cy, ncx, ncy, tile_x, tile_y, HX, HY, p)
        assert np.array_equal(got, want), "%d species differ" % int((got != want).sum())


def test_a_thinner_halo_is_not_enough():
    """The halo is needed: with one row, or two columns, the scheme goes wrong somewhere."""
    p = (0.55, 0.55, 0.55)
    bad_rows = bad_cols = 0
    for seed in range(3):
        sp0, order, phase, cx, cy = _case(seed, 40, 22, 3.0, False)
        u = philox.pair_uniforms(order[:, 0], order[:, 1], 0, 1)
        want, _ = orps.rps_sequential_c(sp0.copy(), order, u, *p)
        bad_rows += int((_tiled(sp0, order, u, cx, cy, 40, 22, 8, 4, HX, 1, p) != want).sum())
        bad_cols += int((_tiled(sp0, order, u, cx, cy, 40, 22, 8, 4, 2, HY, p) != want).sum())
    assert bad_rows > 0 and bad_cols > 0


@pytest.mark.parametrize("first,last", [(0, 8), (0, 5), (6, 8)])
@pytest.mark.parametrize("clustered", [False, True])
def test_transliterated_kernels_on_the_emulated_hand_off(first, last, clustered):
    """The index arithmetic of the CUDA resolvers (records, continuation segments, anchor / partner offsets, the tiled
    kernel's row deltas and interior write-back), transliterated in tests/handoff_emulator.py and run on an emulated
    hand-off: the nine-phase resolver validates the emulated layout against the oracle, the tiled one must agree."""
    import handoff_emulator as he
    p = (0.55, 0.6, 0.9)
    ncx, ncy = 90, 21                                  # two tiles across, two up, ragged on both sides
    rng = np.random.default_rng(7 + first)
    r = 0.01
    h = r * (1 + 2.0 ** -20)
    grid = dict(x0=200.0, y0=30.0, inv_h=1.0 / h, ncx=ncx, ncy=ncy)
    n = 5000
    lon = 200.0 + ncx * h * rng.random(n)
    lat = 30.0 + ncy * h * rng.random(n)
    if clustered:                                      # cells with 40-70 microbes: segments continue across chunks
        k = 0
        for _ in range(12):
            m = int(rng.integers(40, 71))
            lon[k:k + m] = 200.0 + h * (int(rng.integers(0, ncx)) + rng.random(m))
            lat[k:k + m] = 30.0 + h * (int(rng.integers(0, ncy)) + rng.random(m))
            k += m
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    sp0[::41] = 0
    pairs = opairs.query_pairs_reference_array(lon, lat, r)
    order, phase = orps.cell_phase_order(pairs, lon, lat, grid)
    u = philox.pair_uniforms(order[:, 0], order[:, 1], 2, 4)
    H = he.build(lon, lat, order, u, p, grid)
    if clustered:
        assert (np.diff(H["cell_start"]) > 32).any()
    sel_before = phase < first
    sp_before, _ = orps.rps_sequential_c(sp0.copy(), order[sel_before], u[sel_before], *p)
    sel = (phase >= first) & (phase <= last)
    want, draws = orps.rps_sequential_c(sp_before.copy(), order[sel], u[sel], *p)
    assert draws > 100
    stored = sp_before[H["ids"]].astype(np.int64)       # storage order
    got_phases = he.resolve_phases(H, stored.copy(), first, last)
    assert np.array_equal(got_phases, want[H["ids"]]), "the emulated hand-off / phase walk disagrees with the oracle"
    got_tiled = he.resolve_tiled(H, stored.copy(), first, last)
    assert np.array_equal(got_tiled, want[H["ids"]]), "%d species differ" % int((got_tiled != want[H["ids"]]).sum())
    small = he.resolve_tiled(H, stored.copy(), first, last, tile_x=16, tile_y=4)
    assert np.array_equal(small, want[H["ids"]])
