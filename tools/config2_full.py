"""BASELINE config 2 to its full length: 490,000 microbes (700 x 700 lattice, 25-35N 205-215E), 7,670 hourly steps of
the time-varying synthetic OSCAR-grid field on one B200 (the reference's documented run: README.md:7,12 of the
reference), with PARITY CHECKPOINTS on the way: at each checkpoint step the state before the step is downloaded, the
step is run with counters and pair emission, and

  * positions are held against the float64 RK4 restatement from the same inputs (1e-6 relative, north_star),
  * the emitted pair set against cKDTree.query_pairs on the device's positions (exact),
  * the species against the reference rule run sequentially over those pairs in the device's canonical order (exact).

Between checkpoints the step time is measured with CUDA events over windows of 50 steps.  One JSON line per checkpoint
and a summary line; exit status 1 if a checkpoint fails.

    python tools/config2_full.py [--steps 7670] [--checkpoints 100,1000,2000,3000,4000,5000,6000,7000,7670]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(steps=7670, checkpoints=(100, 1000, 2000, 3000, 4000, 5000, 6000, 7000, 7670), window=50, out=None, interact_mode=None,
        advect_mode=None, n=490_000):
    import torch
    import bench
    from lagrangian_microbes_b200._lib import LM_OPT_ADVECT_MODE, LM_OPT_INTERACT_MODE
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from oracle import pairs as opairs, philox, rk4 as ork4, rps as orps
    hfs = bench.make_fieldset(64)
    fs = ork4.FieldSet(hfs.lon, hfs.lat, hfs.time, hfs.u, hfs.v)
    lon, lat, sp, _ = bench.workload_particles("config2", n, 0, 1)
    p, r, seed = bench.P_RPS, bench.RADIUS, 0
    sim = FusedSimulation(lon, lat, sp, r, *p, hfs, dt_seconds=3600.0, seed=seed, emit_pairs=True, pair_capacity=64 * n,
                          regrid_every=16, grid_margin=0.5)
    if interact_mode is not None:
        sim.engine.set_option(LM_OPT_INTERACT_MODE, interact_mode)
    if advect_mode is not None:
        sim.engine.set_option(LM_OPT_ADVECT_MODE, advect_mode)
    checkpoints = sorted(c for c in set(checkpoints) if 1 <= c <= steps)
    lines, ok_all, ms_fresh, ms_max = [], True, None, 0.0
    t_wall = time.time()

    def emit(d):
        lines.append(d)
        s = json.dumps(d)
        print(s, flush=True)
        if out:
            out.write(s + "\n")
            out.flush()

    def timed_window(k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.engine.join()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(k):
            sim.step()
        sim.engine.join()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / k

    while sim.iteration < steps:
        nxt = next((c for c in checkpoints if c > sim.iteration), steps + 1)
        todo = min(nxt - 1, steps) - sim.iteration            # plain steps before the checkpoint step
        ms = None
        if todo >= window:
            # untimed steps up to the window, then the `window` steps before the checkpoint timed with CUDA events
            for _ in range(todo - window):
                sim.step()
            ms = timed_window(window)
        else:
            for _ in range(todo):
                sim.step()
        if sim.iteration >= steps:
            break
        # ---- the checkpoint step itself (step index sim.iteration, 0-based), checked against the oracles
        step = sim.iteration
        lon0, lat0, sp0 = sim.download()
        t0, ti0 = sim.clock.t, sim.clock.ti                       # the particle clock and Parcels' cached time index before the step
        grid = sim.grid.as_dict()
        st = sim.step(check=True)
        gl, ga, gs = sim.download()
        grid_after = sim.grid.as_dict()
        # positions: float64 RK4 restatement from the same inputs
        a64, b64, _, _ = ork4.rk4_step_f64(fs, lon0, lat0, t0, 3600.0, ti0)
        rel = float(max(np.max(np.abs(gl - a64) / np.abs(a64)), np.max(np.abs(ga - b64) / np.abs(b64))))
        want_pairs = opairs.query_pairs_reference_array(gl, ga, r)
        got_pairs = opairs.sort_pairs(sim.pairs[:st.n_pairs].cpu().numpy())
        pairs_ok = bool(st.n_pairs == want_pairs.shape[0] and np.array_equal(got_pairs, want_pairs))
        order, _ = orps.canonical_order(want_pairs, gl, ga, grid, mode=sim.engine.interact_mode)
        u = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
        want_sp, draws = orps.rps_sequential_c(sp0.copy(), order, u, *p)
        species_ok = bool(np.array_equal(gs, want_sp))
        occ = orps.cell_ranks(gl, ga, grid)[3]
        ok = pairs_ok and species_ok and rel < 1e-6
        ok_all = ok_all and ok
        if ms is not None:
            ms_fresh = ms if ms_fresh is None else ms_fresh
            ms_max = max(ms_max, ms)
        emit({"checkpoint_step": step + 1, "ms_per_step_before": None if ms is None else round(ms, 4), "pairs": int(st.n_pairs),
              "rho": round(st.n_pairs / float(n), 4), "max_cell_occupancy": int(occ.max()), "draws": int(draws),
              "positions_rel_err_vs_f64": rel, "pairs_exact": pairs_ok, "species_exact": species_ok,
              "species_count": [int(c) for c in st.species_count[1:]], "cells": [grid["ncx"], grid["ncy"]],
              "regridded": grid_after != grid, "ok": ok, "wall_s": round(time.time() - t_wall, 1)})
    sim.check_faults()
    emit({"summary": "config2", "steps": int(sim.iteration), "microbes": n, "checkpoints": len(checkpoints), "all_ok": ok_all,
          "ms_per_step_first_window": ms_fresh, "ms_per_step_max_window": ms_max,
          "interact_mode": sim.engine.interact_mode, "wall_s": round(time.time() - t_wall, 1)})
    return ok_all, lines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=7670)
    ap.add_argument("--checkpoints", default="100,1000,2000,3000,4000,5000,6000,7000,7670")
    ap.add_argument("--interact-mode", type=int, default=None)
    ap.add_argument("--advect-mode", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    fh = open(a.out, "w") if a.out else None
    ok, _ = run(a.steps, [int(c) for c in a.checkpoints.split(",")], out=fh, interact_mode=a.interact_mode, advect_mode=a.advect_mode)
    sys.exit(0 if ok else 1)
