"""How the step time evolves as the flow stirs the microbes into filaments: phase times every `every` steps."""
import sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from lagrangian_microbes_b200.simulation import FusedSimulation
workload, total, every = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
hfs = bench.make_fieldset(64)
n = bench.default_n(workload)
lon, lat, sp, _ = bench.workload_particles(workload, n, 0, 1)
sim = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                      pair_capacity=(20 if workload == "config3" else 64 if workload == "config2" else 8) * n, regrid_every=16, grid_margin=0.5)
import os
if os.environ.get('LM_RESOLVE_UPL'):
    from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_UPL
    sim.engine.set_option(LM_OPT_RESOLVE_UPL, int(os.environ['LM_RESOLVE_UPL']))
if os.environ.get('LM_RESOLVE_MODE'):
    from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_MODE
    sim.engine.set_option(LM_OPT_RESOLVE_MODE, int(os.environ['LM_RESOLVE_MODE']))
done = 0
import time
t_wall = time.time()
while done < total:
    for _ in range(every - 1):
        sim.step()
    sim.step(timing=True)
    ph = sim.engine.phase_times()
    st = sim.stats()
    done += every
    lo, la, _ = (None, None, None)
    print(json.dumps({"step": done, "wall_s": round(time.time() - t_wall, 2), "pairs": int(st.n_pairs), "rho": st.n_pairs / n, "phases_ms": [round(x, 3) for x in ph],
                      "species": list(st.species_count)[1:]}), flush=True)
# occupancy of the cell grid at the end
g = sim.grid
cs = sim.engine.state_view(rows=g.ncy)[4].to(torch.int64)
occ = (cs[1:] - cs[:-1]).cpu().numpy()
hist = np.bincount(np.minimum(occ, 100000))
edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 10**9]
out = {}
for lo, hi in zip(edges[:-1], edges[1:]):
    sel = (occ >= lo) & (occ < hi)
    out["%d-%d" % (lo, hi - 1)] = [int(sel.sum()), int(occ[sel].sum()), float((occ[sel].astype(np.float64) ** 2).sum() / 2)]
print(json.dumps({"cells": int(occ.size), "max_occupancy": int(occ.max()), "by_occupancy [cells, particles, m^2/2]": out}), flush=True)
