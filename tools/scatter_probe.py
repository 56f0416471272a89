"""The record scatter (cell order -> particle-id order) alone and inside the end-to-end loop, per LM_OPT_SCATTER_PASSES.
python tools/scatter_probe.py [n] -> one JSON line per setting (GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                      # noqa: E402
from lagrangian_microbes_b200 import _lib                          # noqa: E402
from lagrangian_microbes_b200.simulation import FusedSimulation    # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
hfs = bench.make_fieldset(64)
lon, lat, sp, _ = bench.workload_particles("shard", n, 0, 1)


def new_sim():
    return FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                           pair_capacity=8 * n, regrid_every=16, grid_margin=0.5, stream_field=True)


def alone(sim, passes, what, reps=20):
    sim.engine.set_option(_lib.LM_OPT_SCATTER_PASSES, passes)
    dev = sim.engine.device
    lo = torch.empty(n, dtype=torch.float32, device=dev)
    la = torch.empty(n, dtype=torch.float32, device=dev)
    s8 = torch.empty(n, dtype=torch.int8, device=dev)
    args = {"lon+lat": (lo, la, None), "lon": (lo, None, None), "species": (None, None, s8), "all": (lo, la, s8)}[what]
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    tot = 0.0
    for k in range(reps + 2):
        flush.zero_()                                             # 256 MB: L2 holds nothing of the previous repetition
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.engine.state_get(*args)
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            tot += e0.elapsed_time(e1)
    return tot / reps


def e2e(passes, mode, steps=60):
    sim = new_sim()
    sim.engine.set_option(_lib.LM_OPT_SCATTER_PASSES, passes)
    rec = [(torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
            torch.empty(n, dtype=torch.int8).pin_memory()) for _ in range(2)]

    def one(k):
        r = rec[k & 1]
        sim.step(record={"full": r, "pos": (r[0], r[1], None), "none": None}[mode])

    for k in range(6):
        one(k)
    sim.engine.host_copies_sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    sim.engine.host_copies_sync(); sim.engine.join(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    sim.engine.close()
    return ms


def e2e_debug(debug, mode="full", steps=60):
    """The e2e loop with parts of the record left out (LM_OPT_RECORD_DEBUG): 1 = no D2H copies, 2 = no scatter."""
    sim = new_sim()
    sim.engine.set_option(_lib.LM_OPT_RECORD_DEBUG, debug)
    rec = [(torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
            torch.empty(n, dtype=torch.int8).pin_memory()) for _ in range(2)]
    for k in range(6):
        sim.step(record=rec[k & 1])
    sim.engine.host_copies_sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        sim.step(record=rec[k & 1])
    sim.engine.host_copies_sync(); sim.engine.join(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    sim.engine.close()
    return ms


if len(sys.argv) > 2 and sys.argv[2] == "decompose":
    print(json.dumps({"probe": "e2e loop", "record": "none", "ms_per_step": round(e2e(0, "none"), 4)}), flush=True)
    for debug, what in ((0, "scatter + D2H (the record)"), (1, "scatter only, no D2H"), (2, "D2H only, no scatter"), (3, "neither (events only)")):
        print(json.dumps({"probe": "e2e loop, parts of the record", "what": what, "ms_per_step": round(e2e_debug(debug), 4)}), flush=True)
    sys.exit(0)

sim = new_sim()
for _ in range(4):
    sim.step()
for what in ("lon+lat", "lon", "species", "all"):
    for passes in (1, 2, 4, 8, 16):
        print(json.dumps({"probe": "scatter alone", "n": n, "arrays": what, "passes": passes,
                          "ms": round(alone(sim, passes, what), 4)}), flush=True)
sim.engine.close()
del sim
print(json.dumps({"probe": "e2e loop", "record": "none", "ms_per_step": round(e2e(1, "none"), 4)}), flush=True)
for passes in (0, 1, 2):
    for mode in ("pos", "full"):
        print(json.dumps({"probe": "e2e loop", "record": mode, "passes": passes, "ms_per_step": round(e2e(passes, mode), 4)}), flush=True)
