"""A/B of the RPS resolver's knobs (LM_OPT_RESOLVE_BATCH, LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_UPL) on fresh and
stirred states.

    [LM_SWEEP="batch,heavy_min,upl;..."] python tools/resolve_sweep.py config2:1500:200 shard:0:20 shard:1000:20 \
        config3:0:10 config3:400:10

Each argument is workload:steps_before:timed_steps; consecutive arguments of one workload continue the same run.
Prints one JSON line per (state, setting): ms per step (CUDA events over the timed steps, side stream joined) and the
event-timed RPS phase of one step.  Every setting produces identical species (tests/test_gpu_resolver_options.py), so
the settings can be switched inside one run."""
import json
import os
import sys

import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_BATCH, LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_UPL  # noqa: E402
from lagrangian_microbes_b200.simulation import FusedSimulation  # noqa: E402

SETTINGS = [(1, 0, 0), (4, 0, 0), (8, 0, 0), (8, 64, 0), (8, 32, 0), (4, 64, 0), (1, 64, 0), (1, 0, 0)]
if os.environ.get("LM_SWEEP"):
    SETTINGS = [tuple(int(x) for x in t.split(",")) for t in os.environ["LM_SWEEP"].split(";")]

hfs = bench.make_fieldset(64)
sims = {}
for arg in sys.argv[1:]:
    workload, before, timed = arg.split(":")
    before, timed = int(before), int(timed)
    if workload not in sims:
        n = bench.default_n(workload)
        lon, lat, sp, _ = bench.workload_particles(workload, n, 0, 1)
        sims[workload] = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                                         pair_capacity=(36 if workload == "config3" else 14) * n, regrid_every=16,
                                         grid_margin=0.5)
    sim = sims[workload]
    while sim.iteration < before:
        sim.step()
    for batch, heavy_min, upl in SETTINGS:
        sim.engine.set_option(LM_OPT_RESOLVE_BATCH, batch)
        sim.engine.set_option(LM_OPT_RESOLVE_HEAVY_MIN, heavy_min)
        sim.engine.set_option(LM_OPT_RESOLVE_UPL, upl)
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
        print(json.dumps({"workload": workload, "step": sim.iteration, "batch": batch, "heavy_min": heavy_min or 160, "upl": upl,
                          "ms_per_step": round(ms, 4), "rps_ms": round(ph[3], 4), "find_ms": round(ph[2], 4),
                          "pairs": int(st.n_pairs)}), flush=True)
    sim.engine.set_option(LM_OPT_RESOLVE_BATCH, 4)
    sim.engine.set_option(LM_OPT_RESOLVE_HEAVY_MIN, 0)
    sim.engine.set_option(LM_OPT_RESOLVE_UPL, 0)
