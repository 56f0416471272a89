"""CPU tests of the analysis oracle (oracle/analysis.py) and of the host side of the analysis / plotter mirrors.

The reference holds no golden output for these (its histogram code is Julia, sandbox/pairwise_distance_histogram_*.jl;
parity unpinned, see the oracle's header), so the restatement is checked against closed-form distances and against
its own float64 bounds."""
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import analysis as oa

R = float(oa.R32)


def test_haversine_known_answers():
    f = np.float32
    # one degree of longitude on the equator, one degree of latitude anywhere: R * pi / 180
    assert oa.haversine_distance32(f(0), f(10), f(0), f(11)) == pytest.approx(R * np.pi / 180, rel=3e-7)
    assert oa.haversine_distance32(f(40), f(10), f(41), f(10)) == pytest.approx(R * np.pi / 180, rel=3e-6)
    # same point, antipodes, a quarter of a great circle
    assert oa.haversine_distance32(f(33), f(210), f(33), f(210)) == 0.0
    assert oa.haversine_distance32(f(0), f(0), f(0), f(180)) == pytest.approx(R * np.pi, rel=3e-7)
    assert oa.haversine_distance32(f(0), f(0), f(90), f(77)) == pytest.approx(R * np.pi / 2, rel=3e-7)
    # the two marks the reference draws on its plot (pairwise_distance_histogram_distributed.jl:140-141)
    assert oa.haversine_distance32(f(25), f(-145), f(35), f(-155)) == pytest.approx(1.4675e6, rel=2e-3)
    assert oa.haversine_distance32(f(30), f(-150), f(30), f(-150.01)) == pytest.approx(963.0, rel=2e-3)


def test_bin_rule_on_constructed_distances():
    # points on a meridian at chosen distances from the first one: bin = round(10 log10 d)
    want_m = np.array([0.5, 1.0, 1.1, 1.2, 10.0, 31.0, 33.0, 1000.0, 1.0e5, 5.0e6])
    lat = np.concatenate(([0.0], np.degrees(want_m / R))).astype(np.float32)
    lon = np.full(lat.size, 200.0, dtype=np.float32)
    h = np.zeros(72, dtype=np.int64)
    for d in oa.haversine_distance32(lat[0], lon[0], lat[1:], lon[1:]):
        h[int(np.rint(10 * np.log10(max(1.0, float(d)))))] += 1
    only_first = np.zeros(72, dtype=np.int64)
    for b in (0, 0, 0, 1, 10, 15, 15, 30, 50, 67):
        only_first[b] += 1
    assert np.array_equal(h, only_first)


@pytest.mark.parametrize("kind", ["patch", "clustered", "global"])
def test_reference_histogram_is_inside_its_float64_bounds(kind):
    rng = np.random.default_rng(5)
    n = 700
    if kind == "patch":                   # the reference's 10 x 10 degree patch
        lat = 25 + 10 * rng.random(n); lon = 205 + 10 * rng.random(n)
    elif kind == "clustered":             # metres to kilometres apart, with exact duplicates
        lat = 30 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        lon = 210 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        lat[:20] = lat[20:40]; lon[:20] = lon[20:40]
    else:
        lat = -80 + 160 * rng.random(n); lon = 360 * rng.random(n)
    lat, lon = lat.astype(np.float32), lon.astype(np.float32)
    h = oa.pair_distance_hist_reference(lat, lon, bins=70)
    lo, up = oa.pdh_bounds(lat, lon, bins=70)
    assert h.sum() == n * (n - 1) // 2 and h.shape == (72,)
    assert np.all(lo <= h) and np.all(h <= up)
    assert (up - lo).sum() <= 2e-3 * h.sum() + 4          # the band is thin: bounds that pin almost every pair
    if kind == "clustered":
        assert h[0] >= 20                                  # coincident microbes: max(1, d) -> bin 0
    if kind == "global":
        assert h[71] > 0                                   # beyond bin 70 (10^7.05 m): the reference would index out of range


def test_raster_and_compose_reference_by_hand():
    # 4 x 2 pixels over lon [0, 4) x lat [0, 2); row 0 is the northern one
    lon = np.array([0.5, 0.6, 3.9, 4.0, -0.1, 2.0, 2.5], dtype=np.float32)
    lat = np.array([0.5, 0.4, 1.9, 1.0, 1.0, 2.0, 1.5], dtype=np.float32)
    sp = np.array([1, 2, 3, 1, 1, 2, 7], dtype=np.int8)
    counts, top = oa.raster_reference(lon, lat, sp, 0.0, 4.0, 0.0, 2.0, 4, 2)
    assert counts.shape == (3, 2, 4) and top.shape == (2, 4)
    assert counts[0, 1, 0] == 1 and counts[1, 1, 0] == 1 and counts[2, 0, 3] == 1 and counts.sum() == 3
    assert top[1, 0] == 1 and top[0, 3] == 2 and top[0, 2] == 6 and (top == -1).sum() == 5
    pal = np.array([[9, 9, 9], [255, 0, 0], [50, 205, 50], [0, 0, 255]], dtype=np.uint8)
    last = oa.compose_reference(counts, top, sp, pal, 0)
    assert tuple(last[1, 0]) == (50, 205, 50) and tuple(last[0, 3]) == (0, 0, 255)
    assert tuple(last[0, 2]) == (9, 9, 9)                  # species 7 has no colour
    plur = oa.compose_reference(counts, top, sp, pal, 1)
    assert tuple(plur[1, 0]) == (255, 0, 0) and tuple(plur[0, 2]) == (9, 9, 9)


def test_png_writer_round_trips(tmp_path):
    from lagrangian_microbes_b200 import io as lmio
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    path = lmio.write_png(str(tmp_path / "f.png"), img)
    PIL = pytest.importorskip("PIL.Image")
    back = np.asarray(PIL.open(path).convert("RGB"))
    assert np.array_equal(back, img)


def test_species_count_series_and_figure(tmp_path):
    from lagrangian_microbes_b200 import analysis, io as lmio
    rng = np.random.default_rng(1)
    N, Nt = 500, 6
    sp = rng.integers(1, 4, (N, Nt)).astype(np.int8)
    t0, dt = datetime(2018, 1, 1), timedelta(hours=1)
    times = [t0 + k * dt for k in range(Nt)]
    pos = rng.random((N, Nt)).astype(np.float32)
    lmio.write_particle_file(str(tmp_path / "microbe_data.nc"), {"longitude": pos, "latitude": pos, "species": sp}, times)
    ts, r, p, s = analysis.species_count_series(str(tmp_path), t0, t0 + Nt * dt, dt)
    assert ts == times
    for k, got in enumerate((r, p, s)):
        assert np.array_equal(got, (sp == k + 1).sum(axis=0))           # analysis.py:33-35
    png = analysis.species_count_figure(str(tmp_path), t0, t0 + Nt * dt, dt)
    assert os.path.basename(png) == "species_count.png" and os.path.getsize(png) > 100
    assert open(png + ".csv").read().count("\n") == Nt + 1
    assert np.allclose(analysis.bin_lengths_m(70)[[0, 10, 69]], [1.0, 10.0, 10 ** 6.9])


def test_analysis_and_plotter_have_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from lagrangian_microbes_b200 import analysis
    from lagrangian_microbes_b200.microbe_plotter import MicrobePlotter
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        analysis.pairwise_distance_histogram(np.zeros(4), np.zeros(4))
    mp = MicrobePlotter(microbe_marker_size=10, dark_theme=True)
    assert mp.marker_px == 1 and tuple(mp.palette[0]) == (0, 0, 0) and tuple(mp.palette[2]) == (50, 205, 50)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mp.render(np.zeros(4), np.zeros(4), np.ones(4, dtype=np.int8))


def _kernel_edges(bins, radius):
    """csrc/analysis.cu::launch_pair_distance_hist: a-values at the bin edges, rounded UP to float32."""
    slots = bins + 2
    e = np.full(slots + 1, np.inf, dtype=np.float32)
    e[0] = -np.inf
    for k in range(1, slots):
        half = 10.0 ** ((k - 0.5) / 10.0) / (2.0 * radius)
        if half < np.pi / 2:
            x = np.sin(half) ** 2
            f = np.float32(x)
            e[k] = f if float(f) >= x else np.nextafter(f, np.float32(np.inf))
    return e


@pytest.mark.parametrize("kind", ["patch", "clustered", "global"])
def test_threshold_binning_of_the_kernel_is_inside_the_bounds(kind):
    """The kernel does not evaluate sqrt / asin / log10 per pair: it compares the float32 haversine argument `a` with
    bin edges mapped into a-space.  Emulated here in NumPy (same float32 operation order, same edge table, the
    guess + fix-up loop replaced by the search it converges to) and held against the float64 bounds."""
    rng = np.random.default_rng(11)
    n = 600
    if kind == "patch":
        lat = 25 + 10 * rng.random(n); lon = 205 + 10 * rng.random(n)
    elif kind == "clustered":
        lat = 30 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        lon = 210 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        lat[:20] = lat[20:40]; lon[:20] = lon[20:40]
    else:
        lat = -80 + 160 * rng.random(n); lon = 360 * rng.random(n)
    lat, lon = lat.astype(np.float32), lon.astype(np.float32)
    f = np.float32
    edges = _kernel_edges(70, R)
    assert np.all(np.diff(edges[1:].astype(np.float64)) >= 0) and edges[72] == np.inf
    c = oa._cospi32(lat * f(1.0 / 180.0))
    h = np.zeros(72, dtype=np.int64)
    for i in range(n - 1):
        d1 = oa._sinpi32((lat[i + 1:] - lat[i]) * f(1.0 / 360.0))
        d2 = oa._sinpi32((lon[i + 1:] - lon[i]) * f(1.0 / 360.0))
        a = d1 * d1 + ((d2 * d2) * c[i]) * c[i + 1:]
        assert a.dtype == np.float32
        slot = np.searchsorted(edges[1:72], a, side="right")       # slot k: edges[k] <= a < edges[k + 1]
        h += np.bincount(slot, minlength=72)
    lo, up = oa.pdh_bounds(lat, lon, bins=70)
    assert h.sum() == n * (n - 1) // 2
    assert np.all(lo <= h) and np.all(h <= up)


def test_first_guess_and_fix_up_converge():
    """The kernel's slot search: guess = rint(c0 + c1 log2 a) clamped, then walk up / down against the edge table."""
    edges = _kernel_edges(70, R)
    c0, c1 = np.float32(10 * np.log10(2 * R)), np.float32(5 * np.log10(2.0))
    rng = np.random.default_rng(3)
    a_all = np.concatenate(([0.0, 1e-45, 1e-30, 1.0, 1.5, 0.999999], 10.0 ** rng.uniform(-16, 0, 4000))).astype(np.float32)
    worst = 0
    for a in a_all:
        with np.errstate(divide="ignore"):
            g = c0 + c1 * np.log2(a)
        g = 0 if not np.isfinite(g) or g < 0 else int(min(np.rint(g), 71))
        steps = 0
        while a >= edges[g + 1]:
            g += 1; steps += 1
        while a < edges[g]:
            g -= 1; steps += 1
        assert g == int(np.searchsorted(edges[1:72], a, side="right"))
        worst = max(worst, steps)
    assert worst <= 3                                              # asin(x) / x <= pi / 2: at most two bins off at the far end
