"""Timing of the snapshot-analysis kernels (csrc/analysis.cu) with CUDA events: the pair-distance histogram at the size
of the reference's own benchmark (10^4 points, 70 bins: 4.035 s on one Julia process, 0.27 s on 40 --
sandbox/pairwise_distance_histogram_distributed.jl:11-20) and at BASELINE config 1's size (490,000 microbes, three
species), and one frame of the rasteriser.  One JSON line per measurement."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from lagrangian_microbes_b200 import _lib  # noqa: E402
from lagrangian_microbes_b200.microbe_plotter import MicrobePlotter  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


for n, reps in ((10_000, 20), (163_333, 5), (490_000, 2)):
    lat = torch.from_numpy((20 + 20 * rng.random(n)).astype(np.float32)).to(dev)       # random_points_distributed, :84-88
    lon = torch.from_numpy((20 + 20 * rng.random(n)).astype(np.float32)).to(dev)
    hist = torch.empty(72, dtype=torch.int64, device=dev)
    p = ctypes.c_void_p

    def run():
        _lib.check(L.lm_pair_distance_hist(p(lat.data_ptr()), p(lon.data_ptr()), n, 6371.228e3, 70, p(hist.data_ptr()),
                                           p(torch.cuda.current_stream().cuda_stream)), "lm_pair_distance_hist")
    ms = time_ms(run, reps)
    pairs = n * (n - 1) // 2
    assert int(hist.sum().item()) == pairs
    print(json.dumps({"kernel": "pair_distance_hist", "n": n, "pairs": pairs, "ms": round(ms, 4),
                      "pairs_per_s": pairs / (ms * 1e-3), "reference_julia_1proc_s_at_1e4": 4.035}), flush=True)

n = 12_500_000
lon = torch.from_numpy((180 + 60 * rng.random(n)).astype(np.float32)).to(dev)
lat = torch.from_numpy((26.25 + 7.5 * rng.random(n)).astype(np.float32)).to(dev)
sp = torch.from_numpy(rng.integers(1, 4, n).astype(np.int8)).to(dev)
mp = MicrobePlotter(microbe_marker_size=9, width=1600, height=900)
ms = time_ms(lambda: mp.render(lon, lat, sp), 5)
print(json.dumps({"kernel": "rasterize + compose + D2H of the frame", "n": n, "ms": round(ms, 4), "frame": [900, 1600]}), flush=True)
