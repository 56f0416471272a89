"""GPU parity of the snapshot analyses (csrc/analysis.cu) against oracle/analysis.py, through the C ABI.

Tolerances, stated: the rasteriser is integer work and must be bit-exact.  The pair-distance histogram bins a float32
quantity; a pair whose 10 log10(d) lies within 5e-5 of a bin edge (float32 rounding of the haversine argument, see
oracle/analysis.py::pdh_bounds) may fall on either side, every other pair must be in its bin, and the entries must sum
to N (N - 1) / 2 exactly.
"""
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import analysis as oa

# (File name: sorts after every verified GPU test, so that a fault in a kernel that has never run cannot poison the
# CUDA context of the tests before it.)
# First run on hardware: the round-1 driver run (GPUTEST_r01.json), all green.
pytestmark = pytest.mark.gpu


def _cloud(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "patch":
        lat = 25 + 10 * rng.random(n); lon = 205 + 10 * rng.random(n)
    elif kind == "clustered":
        lat = 30 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        lon = 210 + 1e-3 * rng.standard_normal(n) * rng.random(n) ** 4
        if n >= 40:
            lat[:20] = lat[20:40]; lon[:20] = lon[20:40]
    else:
        lat = -80 + 160 * rng.random(n); lon = 360 * rng.random(n)
    return lat.astype(np.float32), lon.astype(np.float32)


@pytest.mark.parametrize("kind", ["patch", "clustered", "global"])
@pytest.mark.parametrize("n", [2, 257, 1025, 3000])
def test_pair_distance_histogram_inside_the_oracle_bounds(kind, n):
    from lagrangian_microbes_b200 import analysis
    lat, lon = _cloud(kind, n, 7 + n)
    h = analysis.pairwise_distance_histogram(lat, lon, bins=70)
    assert h.shape == (72,) and h.dtype == np.int64
    assert h.sum() == n * (n - 1) // 2
    lo, up = oa.pdh_bounds(lat, lon, bins=70)
    assert np.all(lo <= h) and np.all(h <= up), (h - lo, up - h)
    # and the float32 restatement of the reference's code may differ from the kernel only by edge pairs
    ref = oa.pair_distance_hist_reference(lat, lon, bins=70)
    assert np.abs(ref - h).sum() <= 2 * (up - lo).sum()


def test_pair_distance_histogram_edge_cases():
    from lagrangian_microbes_b200 import analysis
    for n in (0, 1):
        h = analysis.pairwise_distance_histogram(np.zeros(n, np.float32), np.zeros(n, np.float32), bins=70)
        assert h.shape == (72,) and not h.any()
    # two microbes 1 km apart on a meridian: bin 30; coincident: bin 0; antipodal: beyond bin 70
    R = float(oa.R32)
    lat = np.array([0.0, np.degrees(1000.0 / R)], dtype=np.float32)
    h = analysis.pairwise_distance_histogram(lat, np.array([200.0, 200.0], np.float32))
    assert h[30] == 1 and h.sum() == 1
    h = analysis.pairwise_distance_histogram(np.array([10.0, 10.0], np.float32), np.array([5.0, 5.0], np.float32))
    assert h[0] == 1 and h.sum() == 1
    h = analysis.pairwise_distance_histogram(np.array([0.0, 0.0], np.float32), np.array([0.0, 180.0], np.float32))
    assert h[71] == 1 and h.sum() == 1
    # other bin counts: the smallest and the largest the ABI takes
    lat, lon = _cloud("global", 900, 1)
    for bins in (1, 40, 126):
        h = analysis.pairwise_distance_histogram(lat, lon, bins=bins)
        lo, up = oa.pdh_bounds(lat, lon, bins=bins)
        assert h.shape == (bins + 2,) and h.sum() == 900 * 899 // 2
        assert np.all(lo <= h) and np.all(h <= up)
    with pytest.raises(ValueError):
        analysis.pairwise_distance_histogram(lat, lon, bins=127)


def test_pair_distance_histogram_at_scale_and_per_species():
    """Size-independent properties at a size the oracle cannot reach: exact total, (near) invariance under a
    permutation of the microbes, per-species totals."""
    from lagrangian_microbes_b200 import analysis
    n = 120_000
    lat, lon = _cloud("patch", n, 3)
    h = analysis.pairwise_distance_histogram(lat, lon, bins=70)
    assert h.sum() == n * (n - 1) // 2
    perm = np.random.default_rng(0).permutation(n)
    h2 = analysis.pairwise_distance_histogram(lat[perm], lon[perm], bins=70)
    assert h2.sum() == h.sum()
    assert np.abs(h2 - h).sum() <= 1e-4 * h.sum()          # c_i * c_j is rounded in the other order: edge pairs only
    sp = np.random.default_rng(1).integers(1, 4, n).astype(np.int8)
    per = analysis.species_pair_distance_histograms(lon, lat, sp, bins=70)
    for s in (1, 2, 3):
        m = int((sp == s).sum())
        assert per[s].sum() == m * (m - 1) // 2
    small = sp[:2000]
    per = analysis.species_pair_distance_histograms(lon[:2000], lat[:2000], small, bins=70)
    for s in (1, 2, 3):
        lo, up = oa.pdh_bounds(lat[:2000][small == s], lon[:2000][small == s], bins=70)
        assert np.all(lo <= per[s]) and np.all(per[s] <= up)


def _raster_gpu(lon, lat, sp, extent, w, h, mode, palette):
    import ctypes
    import torch
    from lagrangian_microbes_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    p = ctypes.c_void_p
    lo, la = torch.from_numpy(lon).to(dev), torch.from_numpy(lat).to(dev)
    s = torch.from_numpy(sp).to(dev) if sp is not None else None
    counts = torch.full((3, h, w), 77, dtype=torch.int32, device=dev)           # the call initialises its outputs
    top = torch.full((h, w), 77, dtype=torch.int32, device=dev)
    rgb = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)
    stream = p(torch.cuda.current_stream().cuda_stream)
    sp_ptr = p(s.data_ptr()) if s is not None else p(0)
    _lib.check(L.lm_rasterize(p(lo.data_ptr()), p(la.data_ptr()), sp_ptr, lon.size, *extent, w, h, p(counts.data_ptr()),
                              p(top.data_ptr()), stream), "lm_rasterize")
    _lib.check(L.lm_compose_frame(p(counts.data_ptr()), p(top.data_ptr()), sp_ptr, w, h, mode, palette.tobytes(),
                                  p(rgb.data_ptr()), stream), "lm_compose_frame")
    torch.cuda.synchronize()
    return counts.cpu().numpy().view(np.uint32), top.cpu().numpy(), rgb.cpu().numpy()


@pytest.mark.parametrize("with_species", [True, False])
@pytest.mark.parametrize("mode", [0, 1])
def test_rasteriser_is_bit_exact(mode, with_species):
    rng = np.random.default_rng(4)
    n = 300_000
    lon = (204.0 + 12.0 * rng.random(n)).astype(np.float32)       # some microbes outside the extent on every side
    lat = (24.0 + 12.0 * rng.random(n)).astype(np.float32)
    lon[:5] = [205.0, 215.0, np.nextafter(np.float32(215.0), np.float32(0)), 210.0, np.nan]   # edges: in, out, in, in, skipped
    lat[:5] = [25.0, 30.0, 30.0, 35.0, 30.0]
    sp = rng.integers(0, 5, n).astype(np.int8) if with_species else None   # 0 and 4: counted in no species, no colour
    extent = (205.0, 215.0, 25.0, 35.0)
    w, h = 640, 360
    pal = np.array([[255, 255, 255], [255, 0, 0], [50, 205, 50], [0, 0, 255]], dtype=np.uint8)
    counts, top, rgb = _raster_gpu(lon, lat, sp, extent, w, h, mode, pal)
    want_counts, want_top = oa.raster_reference(lon, lat, sp, *extent, w, h)
    assert np.array_equal(counts, want_counts)
    assert np.array_equal(top, want_top)
    assert np.array_equal(rgb, oa.compose_reference(want_counts, want_top, sp, pal, mode))


def test_microbe_plotter_writes_the_reference_s_frames(tmp_path):
    from lagrangian_microbes_b200 import io as lmio
    from lagrangian_microbes_b200.microbe_plotter import MicrobePlotter
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    N, Nt = 20_000, 3
    lon = (205 + 10 * rng.random((N, Nt))).astype(np.float32)
    lat = (25 + 10 * rng.random((N, Nt))).astype(np.float32)
    sp = rng.integers(1, 4, (N, Nt)).astype(np.int8)
    t0, dt = datetime(2018, 1, 1), timedelta(hours=1)
    lmio.write_particle_file(str(tmp_path / "microbe_data.nc"), {"longitude": lon, "latitude": lat, "species": sp},
                             [t0 + k * dt for k in range(Nt)])
    mp = MicrobePlotter(N_procs=4, dark_theme=True, microbe_marker_size=36, output_dir=str(tmp_path),
                        extent=(205.0, 215.0, 25.0, 35.0), width=400, height=300)
    assert mp.marker_px == 2
    mp.plot_frames(t0, t0 + Nt * dt, dt)
    for i in range(Nt):
        path = os.path.join(str(tmp_path), "lagrangian_microbes_%05d.png" % i)       # microbe_plotter.py:149
        img = np.asarray(PIL.open(path).convert("RGB"))
        assert img.shape == (300, 400, 3)
        counts, top = oa.raster_reference(lon[:, i], lat[:, i], sp[:, i], 205.0, 215.0, 25.0, 35.0, 200, 150)
        want = oa.compose_reference(counts, top, sp[:, i], mp.palette, 0)
        assert np.array_equal(img, np.repeat(np.repeat(want, 2, axis=0), 2, axis=1))
