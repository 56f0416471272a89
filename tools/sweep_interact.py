#!/usr/bin/env python
"""BASELINE config 5: interaction-only sweep (pair search + RPS, no advection) on one B200.

N microbes uniform-random in a square patch at the config-1 areal density (4,900 / deg^2), radius 0.5-5 km
(0.005-0.05 degrees), p in {0.5, 0.6, 0.9}.  Prints one JSON line per case: pairs, rho, ms per step (CUDA events),
pairs/s, microbes/s.  Usage:  python tools/sweep_interact.py [--max-n 50000000] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=50_000_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from lagrangian_microbes_b200.simulation import FusedSimulation
    cases = [(n, r, 0.55) for n in (1_000_000, 5_000_000, 20_000_000, 50_000_000) for r in (0.005, 0.01, 0.02, 0.05)]
    cases += [(5_000_000, 0.01, p) for p in (0.5, 0.6, 0.9)]
    for n, r, p in cases:
        if n > args.max_n:
            continue
        rho_est = 0.5 * np.pi * r * r * 4900.0
        if n * rho_est > 1.5e9:                      # keep the pair list under 2^31 entries / 20 GB
            continue
        rng = np.random.default_rng(n % 1000 + int(r * 1e4))
        side = np.sqrt(n / 4900.0)
        lon = (205.0 + side * rng.random(n)).astype(np.float32)
        lat = (10.0 + side * rng.random(n)).astype(np.float32)
        sp = rng.integers(1, 4, n).astype(np.int8)
        cap = int(max(1 << 20, 1.3 * rho_est * n + 4 * np.sqrt(rho_est * n) + 1e5))
        sim = FusedSimulation(lon, lat, sp, r, p, p, p, None, emit_pairs=True, pair_capacity=cap, regrid_every=0,
                              grid_margin=0.1, advect=False)
        for _ in range(3):
            sim.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sim.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        st = sim.stats()
        sim.step(timing=True)
        ph = sim.engine.phase_times()
        counts = list(st.species_count)
        print(json.dumps({"n": n, "r_deg": r, "p": p, "pairs": int(st.n_pairs), "rho": st.n_pairs / n, "ms_per_step": ms,
                          "pairs_per_s": st.n_pairs / (ms * 1e-3), "microbes_per_s": n / (ms * 1e-3),
                          "pair_search_ms": ph[2], "rps_resolve_ms": ph[3],
                          "grid": [sim.grid.ncx, sim.grid.ncy], "species_after": counts[1:]}), flush=True)
        sim.engine.close()
        del sim
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
