"""Randomised campaign for the DEFAULT interaction path (LM_OPT_INTERACT_MODE = 2, hybrid: pair search -> hand-off -> nine
phase launches for light units + the queue of heavy units resolved in rounds of matchings by warps / CTAs) on the CPU
emulator (tests/cuda_emu): random grids, densities, knots of up to several hundred microbes per cell (units that a warp
stages in shared memory, units a whole CTA stages, units too big to stage), heavy thresholds, odd species values,
probabilities 0 / 1 -- each against the reference rule applied sequentially in the cell-round order (oracle/rps.py).
Also checks the pair set against cKDTree.  Not part of the test suite (minutes); one JSON line per case.

    python tools/emulated_campaign_hybrid.py [n_cases] [seed]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))

import emu_build  # noqa: E402
from lagrangian_microbes_b200 import _lib  # noqa: E402
from oracle import pairs as opairs, philox, rps as orps  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
L = _lib.declare(ctypes.CDLL(emu_build.build()))
R = 0.01
H = R * (1 + 2.0 ** -20)


def ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


bad = 0
t_start = time.time()
for case in range(n_cases):
    ncx, ncy = int(rng.integers(3, 120)), int(rng.integers(2, 50))
    dens = float(rng.choice([0.3, 1.0, 2.5, 5.0]))
    n = max(8, min(3500, int(ncx * ncy * dens)))
    lon = 200.0 + ncx * H * rng.random(n)
    lat = 30.0 + ncy * H * rng.random(n)
    k = 0
    knots = []
    for _ in range(int(rng.integers(0, 7))):
        m = int(rng.choice([12, 40, 90, 200, 330, 600]))
        if k + m > n:
            break
        kx, ky = int(rng.integers(0, ncx)), int(rng.integers(0, ncy))
        spread = float(rng.choice([1.0, 1.0, 2.0]))              # 2.0: the knot spills over into the neighbouring cells (heavy cross units)
        lon[k:k + m] = np.clip(200.0 + H * (kx + spread * rng.random(m)), 200.0, 200.0 + ncx * H * 0.999999)
        lat[k:k + m] = np.clip(30.0 + H * (ky + spread * rng.random(m)), 30.0, 30.0 + ncy * H * 0.999999)
        knots.append(m)
        k += m
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    sp0 = rng.integers(0, 5, n).astype(np.int8)               # 0 and 4: not rock / paper / scissors
    p = tuple(float(x) for x in rng.choice([0.0, 0.3, 0.55, 0.9, 1.0], 3))
    seed, step = int(rng.integers(0, 1 << 40)), int(rng.integers(0, 1 << 33))
    heavy_min = int(rng.choice([0, 1, 16, 256, 1024, 4096]))   # 0 = the default (1024)
    find_path = int(rng.choice([0, 0, 1]))
    grid = dict(x0=200.0, y0=30.0, inv_h=1.0 / H, ncx=ncx, ncy=ncy)
    pairs = opairs.query_pairs_reference_array(lon, lat, R)
    order, _ = orps.cell_round_order(pairs, lon, lat, grid, heavy_min=heavy_min if heavy_min else 1024)
    u = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
    want, _ = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    h = ctypes.c_void_p()
    cap = pairs.shape[0] + 64
    assert L.lm_create(ctypes.byref(h), 0, n, max(1 << 12, ncx * ncy), cap) == 0
    g = _lib.Grid(200.0, 30.0, 1.0 / H, ncx, ncy)
    assert L.lm_set_grid(h, ctypes.byref(g)) == 0
    for o, v in ((_lib.LM_OPT_INTERACT_MODE, 2), (_lib.LM_OPT_HEAVY_MIN, heavy_min), (_lib.LM_OPT_FIND_PATH, find_path)):
        assert L.lm_set_option(h, o, v) == 0
    species = sp0.copy()
    out = np.zeros((cap, 2), dtype=np.int32)
    prm = _lib.RpsParams(*p, seed, step)
    assert L.lm_interact_rps(h, ptr(lon), ptr(lat), ptr(species), n, R, ctypes.byref(prm), ptr(out), cap, None, None) == 0
    st = _lib.Stats()
    rc = L.lm_sync_stats(h, ctypes.byref(st), None)
    L.lm_destroy(h)
    pairs_ok = st.n_pairs == pairs.shape[0] and np.array_equal(opairs.sort_pairs(out[:st.n_pairs]), pairs)
    ok = rc == 0 and pairs_ok and np.array_equal(species, want)
    bad += 0 if ok else 1
    cell = (np.floor((lat.astype(np.float64) - 30.0) / H).clip(0, ncy - 1) * ncx + np.floor((lon.astype(np.float64) - 200.0) / H).clip(0, ncx - 1)).astype(np.int64)
    print(json.dumps({"case": case, "ncx": ncx, "ncy": ncy, "n": n, "pairs": int(pairs.shape[0]), "max_cell": int(np.bincount(cell).max()), "knots": knots,
                      "p": p, "heavy_min": heavy_min, "find_path": find_path, "changed": int((want != sp0).sum()), "rc": rc, "pairs_ok": bool(pairs_ok),
                      "ok": bool(ok)}), flush=True)
print(json.dumps({"cases": n_cases, "failed": bad, "seconds": round(time.time() - t_start, 1)}))
