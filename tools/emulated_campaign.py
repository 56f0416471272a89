"""Randomised campaign for the tiled resolver on the CPU emulator (tests/cuda_emu): random grids, densities, knots,
tile shapes, shared-memory limits and dense / whole-CTA limits, each against the reference rule applied
sequentially in canonical order (oracle/rps.py).  Not part of the test suite (minutes); one JSON line per case.

    python tools/emulated_campaign.py [n_cases] [seed]
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
    ncx, ncy = int(rng.integers(3, 200)), int(rng.integers(2, 60))
    dens = float(rng.choice([0.3, 1.0, 2.5, 6.0]))
    n = max(8, min(4000, int(ncx * ncy * dens)))
    lon = 200.0 + ncx * H * rng.random(n)
    lat = 30.0 + ncy * H * rng.random(n)
    k = 0
    for _ in range(int(rng.integers(0, 9))):
        m = int(rng.integers(5, 121))
        if k + m > n:
            break
        lon[k:k + m] = 200.0 + H * (int(rng.integers(0, ncx)) + rng.random(m))
        lat[k:k + m] = 30.0 + H * (int(rng.integers(0, ncy)) + rng.random(m))
        k += m
    lon, lat = lon.astype(np.float32), lat.astype(np.float32)
    sp0 = rng.integers(0, 5, n).astype(np.int8)               # 0 and 4: not rock / paper / scissors
    p = tuple(float(x) for x in rng.choice([0.0, 0.3, 0.55, 0.9, 1.0], 3))
    seed, step = int(rng.integers(0, 1 << 40)), int(rng.integers(0, 1 << 33))
    opts = dict(shape=int(rng.integers(0, 4)), smem=int(rng.choice([1024, 4096, 32768])), heavy=int(rng.choice([0, 4, 64])),
                mega=int(rng.choice([0, 32, 1 << 20])))
    grid = dict(x0=200.0, y0=30.0, inv_h=1.0 / H, ncx=ncx, ncy=ncy)
    pairs = opairs.query_pairs_reference_array(lon, lat, R)
    order, phase = orps.cell_phase_order(pairs, lon, lat, grid)
    u = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
    want, _ = orps.rps_sequential_c(sp0.copy(), order, u, *p)
    h = ctypes.c_void_p()
    assert L.lm_create(ctypes.byref(h), 0, n, max(1 << 12, ncx * ncy), pairs.shape[0] + 64) == 0
    g = _lib.Grid(200.0, 30.0, 1.0 / H, ncx, ncy)
    assert L.lm_set_grid(h, ctypes.byref(g)) == 0
    for o, v in ((_lib.LM_OPT_RESOLVE_MODE, 1), (_lib.LM_OPT_RESOLVE_TILE_SHAPE, opts["shape"]), (_lib.LM_OPT_RESOLVE_TILE_SMEM, opts["smem"]),
                 (_lib.LM_OPT_RESOLVE_HEAVY_MIN, opts["heavy"]), (_lib.LM_OPT_RESOLVE_MEGA_MIN, opts["mega"])):
        assert L.lm_set_option(h, o, v) == 0
    species = sp0.copy()
    prm = _lib.RpsParams(*p, seed, step)
    assert L.lm_interact_rps(h, ptr(lon), ptr(lat), ptr(species), n, R, ctypes.byref(prm), None, 0, None, None) == 0
    st = _lib.Stats()
    rc = L.lm_sync_stats(h, ctypes.byref(st), None)
    L.lm_destroy(h)
    ok = rc == 0 and st.n_pairs == pairs.shape[0] and np.array_equal(species, want)
    bad += 0 if ok else 1
    print(json.dumps({"case": case, "ncx": ncx, "ncy": ncy, "n": n, "pairs": int(pairs.shape[0]), "max_cell": int(np.bincount(
        (np.floor((lat.astype(np.float64) - 30.0) / H).clip(0, ncy - 1) * ncx + np.floor((lon.astype(np.float64) - 200.0) / H).clip(0, ncx - 1)).astype(np.int64)).max()),
        "p": p, **opts, "changed": int((want != sp0).sum()), "ok": bool(ok)}), flush=True)
print(json.dumps({"cases": n_cases, "failed": bad, "seconds": round(time.time() - t_start, 1)}))
sys.exit(1 if bad else 0)
