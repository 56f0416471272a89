import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from lagrangian_microbes_b200.simulation import FusedSimulation
from lagrangian_microbes_b200._lib import LM_OPT_OVERLAP
hfs = bench.make_fieldset(64)
n = bench.default_n("shard")
lon, lat, sp, _ = bench.workload_particles("shard", n, 0, 1)
sim = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                      pair_capacity=8 * n, regrid_every=16, grid_margin=0.5)
sim.engine.set_option(LM_OPT_OVERLAP, 0)
for _ in range(1000):
    sim.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sim.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
