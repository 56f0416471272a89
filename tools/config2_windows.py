"""BASELINE config 2 (490,000 microbes, 7,670 hourly steps) run from scratch under ONE interaction setting; prints the step time
(CUDA events) over windows of 100 steps along the run -- positions do not depend on the setting, so the windows of different
settings see the same cell occupancies.   python tools/config2_windows.py <interact_mode> <heavy_min> [steps]"""
import json
import sys

import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from lagrangian_microbes_b200._lib import LM_OPT_ADVECT_MODE, LM_OPT_HEAVY_MIN, LM_OPT_INTERACT_MODE  # noqa: E402
from lagrangian_microbes_b200.simulation import FusedSimulation  # noqa: E402

mode, hmin = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 7670
n = 490_000
hfs = bench.make_fieldset(64)
lon, lat, sp, _ = bench.workload_particles("config2", n, 0, 1)
sim = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True, pair_capacity=64 * n,
                      regrid_every=16, grid_margin=0.5)
sim.engine.set_option(LM_OPT_ADVECT_MODE, 1)
sim.engine.set_option(LM_OPT_INTERACT_MODE, mode)
sim.engine.set_option(LM_OPT_HEAVY_MIN, hmin)
out, worst = [], 0.0
while sim.iteration < steps:
    k = min(100, steps - sim.iteration)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.engine.join()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(k):
        sim.step()
    sim.engine.join()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / k
    worst = max(worst, ms)
    out.append(round(ms, 3))
sim.check_faults()
tot = sum(a * 100 for a in out[:-1]) + out[-1] * (steps - 100 * (len(out) - 1))
print(json.dumps({"interact_mode": mode, "heavy_min": hmin, "steps": steps, "total_device_s": round(tot / 1e3, 3), "worst_window_ms": worst,
                  "first_window_ms": out[0], "ms_per_step_by_100_steps": out}))
