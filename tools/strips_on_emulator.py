"""Developer tool, NOT part of the product: the strip host layer (StripSet, transports, settle / regrid / rebalance) against
the CPU emulator of the kernels, with the configuration of tests/test_gpu_strips.py::test_strips_over_nccl_equal_single_handle
(contiguous tiles, tight grid margin, regridding every 4 steps) on G strips in one process.

    python -O tools/strips_on_emulator.py [G] [n] [steps] [peer]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
prelude = open(os.path.join(HERE, "host_layer_on_emulator.py")).read().split("from lagrangian_microbes_b200.simulation import FusedSimulation")[0]
exec(compile(prelude, "host_layer_on_emulator.py(prelude)", "exec"))

import numpy as np                                                          # noqa: E402

_real_as_tensor = torch.as_tensor                                          # noqa: F821


def _as_tensor(obj, *a, **k):
    """Raw "device" pointers of the emulator are host pointers: wrap them through ctypes."""
    if isinstance(k.get("device"), _Dev):                                  # noqa: F821
        k.pop("device")
    iface = getattr(obj, "__cuda_array_interface__", None)
    if iface is not None:
        dt = np.dtype(iface["typestr"])
        count = int(np.prod(iface["shape"]))
        buf = (ctypes.c_ubyte * (count * dt.itemsize)).from_address(iface["data"][0])   # noqa: F821
        return torch.from_numpy(np.frombuffer(buf, dtype=dt).reshape(iface["shape"]))   # noqa: F821
    return _real_as_tensor(obj, *a, **k)


torch.as_tensor = _as_tensor                                               # noqa: F821
torch.Tensor.cuda = lambda self, *a, **k: self                             # noqa: F821
import lagrangian_microbes_b200.strips as strips                           # noqa: E402
strips.torch.device = fake_device                                          # noqa: F821
import test_gpu_strips as T                                                 # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
peer = len(sys.argv) > 4 and sys.argv[4] == "peer"
seed = 21
if len(sys.argv) > 5 and sys.argv[5] == "twin":
    # bench.py::parity_twin at N = G ranks: contiguous id tiles of a uniform cloud, settle() has to route 7/8 of the microbes
    import bench
    hfs = bench.make_fieldset(8)
    per_rank = n // G
    rng = np.random.default_rng(12345)
    side = float(np.sqrt(n / 27_800.0))
    lon = (200.0 + side * rng.random(n)).astype(np.float32)
    lat = (30.0 + side * rng.random(n)).astype(np.float32)
    sp = rng.integers(1, 4, n).astype(np.int8)
    ids = np.arange(n, dtype=np.int32)
    cut = [slice(r * per_rank, (r + 1) * per_rank) for r in range(G)]
    for slack in ((max(1.6, float(G)),) if peer else (1.6, max(1.6, float(G)))):
        try:
            ss = strips.StripSet((strips.LocalPeerTransport if peer else strips.LocalTransport)(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                                 [ids[c] for c in cut], n, bench.RADIUS, *bench.P_RPS, hfs, dt_seconds=bench.DT, seed=7, local_strips=list(range(G)),
                                 emit_pairs=False, slack=slack, grid_margin=0.5, regrid_every=2)
            print("twin, slack %.1f: settled, edges %s, sizes %s, capacity %d" % (slack, ss.edges, [s.engine.state_size() for s in ss.strips],
                                                                                  ss.strips[0].engine.max_particles), flush=True)
            for k in range(steps):
                ss.step()
            ss.stats()
            print("twin, slack %.1f: %d steps ok" % (slack, steps), flush=True)
            ss.close()
        except Exception as e:
            print("twin, slack %.1f: FAILED: %s" % (slack, e), flush=True)
    sys.exit(0)
fs = T.small_fs()
lon, lat, sp = T.particles(n, seed, clustered=True)
ids = np.arange(n, dtype=np.int32)
per = n // G
cut = [slice(r * per, (r + 1) * per if r < G - 1 else n) for r in range(G)]
ss = strips.StripSet((strips.LocalPeerTransport if peer else strips.LocalTransport)(G), [lon[c] for c in cut], [lat[c] for c in cut],
                     [sp[c] for c in cut], [ids[c] for c in cut], n, T.R, *T.P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                     pairs_per_particle=40 * G, grid_margin=0.05, regrid_every=4, cells_headroom=3.0)
print("settled: edges", ss.edges, "grid", ss.grid.ncx, ss.grid.ncy, "sizes", [s.engine.state_size() for s in ss.strips], flush=True)
sim = T.single(lon, lat, sp, ss.grid, fs, seed, regrid_every=4, grid_margin=0.05)
for k in range(steps):
    try:
        pairs = T.compare_step(ss, sim, k)
    except Exception:
        for st in ss.strips:
            c = _lib.Stats()                                                # noqa: F821
            rc = st.engine.L.lm_sync_stats(st.engine.h, ctypes.byref(c), None)   # noqa: F821
            print("strip", st.index, "rc", rc, "pairs", c.n_pairs, "pair cap", None if st.pairs is None else st.pairs.shape[0],
                  "particles", c.n_particles, "max", st.engine.max_particles, "moved in/out", c.n_moved_in, c.n_moved_out)
        raise
    print("step %d: %d pairs, edges %s, sizes %s" % (k, pairs, ss.edges, [s.engine.state_size() for s in ss.strips]), flush=True)
# the in-step record of the first strip (StripSet.step(record=slot)) against the state read back after the step
for k in range(3):
    ss.step(record=k & 1)
    sim.step()
    want = ss.local_state()[0]
    ss.host_copies_sync()
    got = ss.record_view(k & 1)
    assert all(np.array_equal(a, b) for a, b in zip(got, want)) and got[0].size == want[0].size > 0
print("in-step strip record ok")
print("ok")
