"""A/B of the tiled RPS resolver (LM_OPT_RESOLVE_MODE = 1) against the nine phase launches, on fresh and stirred states.

    python tools/tiled_sweep.py config2:1500:200 shard:0:20 shard:1000:20 config3:0:10 config3:400:10

Each argument is workload:steps_before:timed_steps; consecutive arguments of one workload continue the same run.
Both modes produce identical species (tests/test_gpu_zz_resolver_tiled.py), so they can be switched inside one run.
Prints one JSON line per (state, setting): ms per step (CUDA events, side stream joined) and the event-timed phases."""
import json
import sys

import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_MODE, LM_OPT_RESOLVE_TILE_SHAPE, LM_OPT_RESOLVE_TILE_SMEM  # noqa: E402
from lagrangian_microbes_b200.simulation import FusedSimulation  # noqa: E402

# (mode, shared memory per tile, tile shape: 0 = 64 x 16 cells, 1 = 32 x 16, 2 = 128 x 16, 3 = 64 x 32)
SETTINGS = [(0, 32768, 0), (1, 16384, 0), (1, 32768, 0), (1, 65536, 0), (1, 131072, 0), (1, 32768, 1), (1, 32768, 2), (1, 65536, 2),
            (1, 32768, 3), (1, 65536, 3), (0, 32768, 0)]

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
    for mode, smem, shape in SETTINGS:
        sim.engine.join()
        torch.cuda.synchronize()
        sim.engine.set_option(LM_OPT_RESOLVE_MODE, mode)
        sim.engine.set_option(LM_OPT_RESOLVE_TILE_SMEM, smem)
        sim.engine.set_option(LM_OPT_RESOLVE_TILE_SHAPE, shape)
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
        print(json.dumps({"workload": workload, "step": sim.iteration, "mode": mode, "tile_smem": smem, "tile_shape": shape,
                          "ms_per_step": round(ms, 4), "rps_ms": round(ph[3], 4), "find_ms": round(ph[2], 4),
                          "pairs": int(st.n_pairs), "species": [int(c) for c in st.species_count]}), flush=True)
    sim.engine.set_option(LM_OPT_RESOLVE_MODE, 0)
