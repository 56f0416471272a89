"""GPU probe (not a bench line): lm_record_delta_pack at the default shard size -- device time per launch with CUDA
events on the launching stream, algorithmic bytes (R 16 + W 4 per microbe) against MEASURED_PEAKS.json's HBM number,
and the pinned D2H time of the packed record (4 B) beside the plain one (8 B).  One JSON line.

    python tools/record_pack_probe.py [n_microbes]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lagrangian_microbes_b200 import _lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
    assert torch.cuda.is_available(), "needs a CUDA device"
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    prev_lon = 205 + 10 * torch.rand(n, device="cuda", generator=g); prev_lat = 25 + 10 * torch.rand(n, device="cuda", generator=g)
    lon = prev_lon + 0.02 * (torch.rand(n, device="cuda", generator=g) - 0.5); lat = prev_lat + 0.02 * (torch.rand(n, device="cuda", generator=g) - 0.5)
    dl = torch.empty(n, dtype=torch.int16, device="cuda"); da = torch.empty_like(dl)
    cap = n // 16
    esc = torch.empty((cap, 2), dtype=torch.int32, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")           # > L2 (126 MB)

    def launch():
        _lib.check(L.lm_record_delta_pack(p(prev_lon), p(prev_lat), p(lon), p(lat), n, p(dl), p(da), p(esc), cap, p(cnt), s), "pack")

    for _ in range(3):
        launch()
    times = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); launch(); b.record(); b.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.median(times))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 0) or 0)
    host4 = torch.empty(n, dtype=torch.int32).pin_memory(); host8 = torch.empty(2 * n, dtype=torch.int32).pin_memory()
    dev4 = torch.empty(n, dtype=torch.int32, device="cuda"); dev8 = torch.empty(2 * n, dtype=torch.int32, device="cuda")

    def d2h(dst, src):
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dst.copy_(src, non_blocking=True); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    # host decode of the same record (lm_record_delta_unpack_host) on the box's cores
    import time
    from lagrangian_microbes_b200.record import unpack_delta_record_native
    m = int(cnt.item())
    h = [t.cpu().numpy() for t in (prev_lon, prev_lat, dl, da)]
    esc_h = esc.cpu().numpy().view(np.uint32)[:m]
    out = (np.empty(n, np.float32), np.empty(n, np.float32))
    decode_ms = {}
    for th in sorted({1, 8, min(32, os.cpu_count() or 1)}):
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); unpack_delta_record_native(h[0], h[1], h[2], h[3], esc_h, n_threads=th, out=out); ts.append(time.perf_counter() - t0)
        decode_ms[str(th)] = 1e3 * min(ts)
    exact = bool(np.array_equal(out[0].view(np.uint32), lon.cpu().numpy().view(np.uint32)) and
                 np.array_equal(out[1].view(np.uint32), lat.cpu().numpy().view(np.uint32)))
    print(json.dumps({"kernel": "record_delta_pack_kernel<vec>", "round_trip_bit_exact": exact, "host_decode_ms_by_threads": decode_ms, "microbes": n, "launch_ms": ms, "escapes": int(cnt.item()),
                      "algorithmic_bytes": 20 * n, "achieved_gbs": 20 * n / ms / 1e6, "peak_gbs": peak or None,
                      "frac": (20 * n / ms / 1e6 / peak) if peak else None,
                      "d2h_ms_packed_4B": d2h(host4, dev4), "d2h_ms_plain_8B": d2h(host8, dev8), "l2": "flushed between launches"}))


if __name__ == "__main__":
    main()
