"""A/B of LM_OPT_HEAVY_MIN (hybrid mode: candidate pairs above which a unit goes to the rounds-of-matchings queue) and of the
round-1 pipeline alone (mode 0), on fresh and stirred states, inside one run per workload (every setting is a valid
sequential order, so they can be switched mid-run).

    python tools/heavy_sweep.py config2:3000:100 config2:6000:100 config3:0:10 config3:300:10 shard:0:20 shard:1000:20
"""
import json
import sys

import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from lagrangian_microbes_b200._lib import LM_OPT_ADVECT_MODE, LM_OPT_HEAVY_MIN, LM_OPT_INTERACT_MODE  # noqa: E402
from lagrangian_microbes_b200.simulation import FusedSimulation  # noqa: E402

SETTINGS = [(2, 1024), (2, 2048), (2, 4096), (2, 8192), (2, 16384), (2, 65536), (0, 0), (2, 1024)]

hfs = bench.make_fieldset(64)
sims = {}
for arg in sys.argv[1:]:
    workload, before, timed = arg.split(":")
    before, timed = int(before), int(timed)
    if workload not in sims:
        n = bench.default_n(workload)
        lon, lat, sp, _ = bench.workload_particles(workload, n, 0, 1)
        sims[workload] = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                                         pair_capacity=(36 if workload == "config3" else 64 if workload == "config2" else 14) * n,
                                         regrid_every=16, grid_margin=0.5)
        sims[workload].engine.set_option(LM_OPT_ADVECT_MODE, 1)
    sim = sims[workload]
    while sim.iteration < before:
        sim.step()
    for mode, hmin in SETTINGS:
        sim.engine.join()
        torch.cuda.synchronize()
        sim.engine.set_option(LM_OPT_INTERACT_MODE, mode)
        sim.engine.set_option(LM_OPT_HEAVY_MIN, hmin)
        for _ in range(3):
            sim.step()
        sim.engine.join()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(timed):
            sim.step()
        sim.engine.join()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / timed
        sim.step(timing=True)
        ph = sim.engine.phase_times()
        st = sim.stats()
        print(json.dumps({"workload": workload, "step": sim.iteration, "mode": mode, "heavy_min": hmin, "ms_per_step": round(ms, 4),
                          "find_ms": round(ph[2], 4), "rps_ms": round(ph[3], 4), "pairs": int(st.n_pairs)}), flush=True)
    sim.engine.set_option(LM_OPT_INTERACT_MODE, 2)
    sim.engine.set_option(LM_OPT_HEAVY_MIN, 0)
