import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from lagrangian_microbes_b200.simulation import FusedSimulation
hfs = bench.make_fieldset(64)
n = 12_500_000
lon, lat, sp, _ = bench.workload_particles("shard", n, 0, 1)
def run(stream_field, rec_mode, steps=60):
    sim = FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.55, 0.55, hfs, dt_seconds=3600.0, seed=0, emit_pairs=True,
                          pair_capacity=8 * n, regrid_every=16, grid_margin=0.5, stream_field=stream_field)
    rec = [(torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
            torch.empty(n, dtype=torch.int8).pin_memory()) for _ in range(2)]
    def one(k):
        if rec_mode == "instep": sim.step(record=rec[k & 1])
        elif rec_mode == "sp_only": sim.step(record=(None, None, rec[k & 1][2]))
        elif rec_mode == "pos_only": sim.step(record=(rec[k & 1][0], rec[k & 1][1], None))
        elif rec_mode == "after":
            sim.step(); sim.record_to_host(*rec[k & 1])
        else: sim.step()
    for k in range(5): one(k)
    sim.engine.host_copies_sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps): one(k)
    sim.engine.host_copies_sync(); sim.engine.join(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print("stream_field=%s record=%s: %.3f ms/step" % (stream_field, rec_mode, ms), flush=True)
    sim.engine.close()
for sf in (False, True):
    for rm in ("none", "sp_only", "pos_only", "instep", "after"):
        run(sf, rm)
