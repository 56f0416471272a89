"""Snapshot analyses of the microbe population on the B200.

Mirrors /root/reference/analysis.py (``species_count_figure``, :20-58) and the reference's pair-distance histogram
(sandbox/pairwise_distance_histogram_distributed.jl:33-64, :93-138; its unfinished CUDA kernel:
sandbox/pairwise_distance_histogram_gpu.jl:14-43):

    pairwise_distance_histogram(lats, lons, bins, R)     one point set, all N (N - 1) / 2 pairs   -> lm_pair_distance_hist
    species_pair_distance_histograms(lons, lats, species, bins)    the per-species call pattern of :127-138
    species_count_series / species_count_figure          the census of analysis.py:31-35

The histogram is a CUDA kernel (csrc/analysis.cu); there is no CPU fallback.  In the fused loop the per-step census
comes from the device for free (``lm_stats.species_count``, ``FusedSimulation.step(check=True)``); the census functions
here serve the reference's file-based workflow, where the data are already on the host.
"""
import ctypes
import logging
import os

import numpy as np

from . import io as lmio
from .interactions import PAPER, ROCK, SCISSORS

logger = logging.getLogger(__name__)

EARTH_RADIUS_M = 6371.228e3          # pairwise_distance_histogram_distributed.jl:103 (R32)


def _device_f32(x, device):
    import torch
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)


def pairwise_distance_histogram(lats, lons, bins=70, R=EARTH_RADIUS_M, device=None):
    """All-pairs haversine distance histogram of one point set, ``bin = round(10 log10(max(1, d[m])))``
    (pairwise_distance_histogram_distributed.jl:54-64 summed over i).

    lats, lons: degrees, NumPy arrays or CUDA tensors (float32 is what the reference converts to, :110-111).
    Returns int64[bins + 2] on the host: [b] = pairs in bin b for b = 0..bins, [bins + 1] = pairs beyond the last bin;
    entry k of the reference's 1-based ``sub_hist`` is entry k here.  The entries sum to N (N - 1) / 2."""
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("pairwise_distance_histogram needs a CUDA device: there is no CPU fallback")
    L = _lib.lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if not 1 <= int(bins) <= _lib.LM_PDH_MAX_BINS:
        raise ValueError("bins must be in 1..%d" % _lib.LM_PDH_MAX_BINS)
    with torch.cuda.device(device):
        la, lo = _device_f32(lats, device), _device_f32(lons, device)
        assert la.ndim == 1 and la.shape == lo.shape
        hist = torch.empty(int(bins) + 2, dtype=torch.int64, device=device)
        _lib.check(L.lm_pair_distance_hist(ctypes.c_void_p(la.data_ptr()), ctypes.c_void_p(lo.data_ptr()), la.numel(),
                                           float(R), int(bins), ctypes.c_void_p(hist.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "lm_pair_distance_hist")
        return hist.cpu().numpy()


def species_pair_distance_histograms(lons, lats, species, bins=70, R=EARTH_RADIUS_M, device=None):
    """{ROCK: hist, PAPER: hist, SCISSORS: hist}: one histogram per species over that species' microbes, as
    plot_pairwise_histogram does (pairwise_distance_histogram_distributed.jl:113-138)."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lo, la = _device_f32(lons, device), _device_f32(lats, device)
    sp = species if isinstance(species, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(species, dtype=np.int8))
    sp = sp.to(device)
    out = {}
    for s in (ROCK, PAPER, SCISSORS):
        sel = sp == s                                   # order-preserving selection, like mlons[species .== ROCK]
        out[int(s)] = pairwise_distance_histogram(la[sel], lo[sel], bins=bins, R=R, device=device)
    return out


def bin_lengths_m(bins=70):
    """Abscissa the reference plots the histogram against (:149): 10^(k / 10) metres for k = 0..bins-1."""
    return 10.0 ** (np.arange(bins) / 10.0)


def species_count_series(output_dir, start_time, end_time, dt):
    """(times, n_rocks, n_papers, n_scissors) from ``microbe_data.nc`` -- analysis.py:20-35."""
    iters = (end_time - start_time) // dt
    times = [start_time + n * dt for n in range(iters)]
    microbe_data = lmio.read_particle_file(os.path.join(output_dir, "microbe_data.nc"))
    species = microbe_data["species"]
    counts = np.zeros((3, iters), dtype=np.int64)
    logger.info("Calculating species count time series...")
    for i in range(iters):
        col = np.asarray(species[:, i])
        for k, s in enumerate((ROCK, PAPER, SCISSORS)):
            counts[k, i] = int(np.sum(col == s))
    return times, counts[0], counts[1], counts[2]


_NAMED_COLORS = {"red": (255, 0, 0), "limegreen": (50, 205, 50), "blue": (0, 0, 255), "green": (0, 128, 0),
                 "white": (255, 255, 255), "black": (0, 0, 0), "dimgray": (105, 105, 105)}


def color_rgb(name):
    """RGB bytes of the matplotlib colour names the reference uses (interactions.py:8-10, analysis.py:45-47)."""
    return _NAMED_COLORS[name]


def species_count_figure(output_dir, start_time, end_time, dt, png_filename="species_count.png", width=1000, height=600):
    """analysis.py:20-58: the three census curves (red / green / blue, :45-47) as a PNG in ``output_dir``; the series
    itself is also saved next to it as ``<png_filename>.csv``.  No axes text: matplotlib is not a dependency."""
    times, n_rocks, n_papers, n_scissors = species_count_series(output_dir, start_time, end_time, dt)
    img = np.full((height, width, 3), 255, dtype=np.uint8)
    series = ((n_rocks, color_rgb("red")), (n_papers, color_rgb("green")), (n_scissors, color_rgb("blue")))
    top = max(1, int(max(int(s.max()) if s.size else 0 for s, _ in series)))
    iters = len(times)
    for s, c in series:
        if iters == 0:
            break
        x = np.round(np.linspace(0, width - 1, 4 * width)).astype(np.int64)
        v = np.interp(np.linspace(0, iters - 1, 4 * width), np.arange(iters), s.astype(np.float64))
        y = (height - 1 - np.round(v / top * (height - 1))).astype(np.int64)
        img[np.clip(y, 0, height - 1), x] = c
    png_filepath = os.path.join(output_dir, png_filename)
    logger.info("Saving species count time series figure: {:s}".format(png_filepath))
    lmio.write_png(png_filepath, img)
    with open(png_filepath + ".csv", "w") as f:
        f.write("time,rocks,papers,scissors\n")
        for t, a, b, c in zip(times, n_rocks, n_papers, n_scissors):
            f.write("%s,%d,%d,%d\n" % (t.isoformat(), a, b, c))
    return png_filepath
