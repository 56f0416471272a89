"""GPU parity of the latitude-strip decomposition: G strips -- several handles on one device
with device-copy exchanges, or one process per GPU with NCCL send/recv -- must reproduce the single-handle
fused step BIT FOR BIT (positions, pair set, species), which in turn is checked against the oracle in
tests/test_gpu_parity.py::test_fused_simulation_matches_oracle_loop."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import rk4 as ork4

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

P = (0.55, 0.6, 0.9)
R = 0.01


class HostFS:
    def __init__(self, fs):
        self.u, self.v, self.lon, self.lat, self.time = fs.u, fs.v, fs.lon, fs.lat, fs.time

    def to_device(self, device):
        return tuple(torch.from_numpy(a).to(device) for a in (self.u, self.v, self.lon, self.lat))


def small_fs():
    g = golden("rk4_small.npz")
    return HostFS(ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"]))


def particles(n, seed, clustered=False):
    rng = np.random.default_rng(seed)
    lon = 201.0 + 1.2 * rng.random(n)          # open water in the golden field (its land block starts at 204.67E)
    lat = 32.0 + 1.2 * rng.random(n)
    if clustered:                       # a dense blob: unbalanced rows, heavy cells next to a strip boundary
        k = n // 3
        lon[:k] = 201.6 + 0.05 * rng.standard_normal(k)
        lat[:k] = 32.55 + 0.05 * rng.standard_normal(k)
    sp = rng.integers(1, 4, n).astype(np.int8)
    return lon.astype(np.float32), lat.astype(np.float32), sp


def single(lon, lat, sp, grid, fs, seed, Kh=0.0, regrid_every=0, grid_margin=0.5):
    """The single-handle reference run on the SAME grid as the strips."""
    from lagrangian_microbes_b200.simulation import FusedSimulation
    sim = FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, Kh=Kh, seed=seed, emit_pairs=True,
                          pair_capacity=40 * lon.size, regrid_every=regrid_every, grid_margin=grid_margin,
                          max_cells=max(1 << 20, 4 * grid.ncx * grid.ncy))
    sim.engine.set_grid(grid)
    sim.grid = grid
    sim.engine.state_set(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda(), torch.from_numpy(sp).cuda())
    return sim


def compare_step(ss, sim, step):
    st = sim.step(check=True)
    ss.step(check=True)
    n_pairs, n_part, counts = ss.totals()
    assert n_part == sim.n
    lon, lat, sp = ss.gather()
    wl, wa, ws = sim.download()
    assert np.array_equal(lon, wl) and np.array_equal(lat, wa), "positions differ at step %d" % step
    assert n_pairs == st.n_pairs, "pair count %d != %d at step %d" % (n_pairs, st.n_pairs, step)
    got = opairs.sort_pairs(np.concatenate([p.reshape(-1, 2) for p in ss.local_pairs()]))
    want = opairs.sort_pairs(sim.pairs[:st.n_pairs].cpu().numpy())
    assert np.array_equal(got, want), "pair set differs at step %d" % step
    assert np.array_equal(sp, ws), "species differ at step %d (%d microbes)" % (step, int((sp != ws).sum()))
    assert counts == list(st.species_count)
    return st.n_pairs


# peer = True: the strips exchange through each other's device pointers -- the peer-memory path of the multi-GPU run (packing
# kernels writing into the neighbour's receive buffer, flag stores, spinning consumers) on one device
@pytest.mark.parametrize("G,clustered,peer", [(2, False, False), (3, True, False), (4, False, False), (7, True, False),
                                              (2, False, True), (3, True, True), (5, False, True)])
def test_strips_on_one_device_equal_single_handle(G, clustered, peer):
    from lagrangian_microbes_b200.strips import LocalPeerTransport, StripSet
    from lagrangian_microbes_b200.strips import LocalTransport as _Copy
    LocalTransport = LocalPeerTransport if peer else _Copy
    n, seed = 40000, 11 + G
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered)
    ids = np.arange(n, dtype=np.int32)
    # hand every strip a contiguous TILE of the particles (the reference's split): settle() must route them
    per = n // G
    cut = [slice(g * per, (g + 1) * per if g < G - 1 else n) for g in range(G)]
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=40 * G, grid_margin=0.25, regrid_every=0)     # a dense blob puts most pairs into one strip
    assert all(e % 16 == 0 for e in ss.edges[:-1]) and ss.edges[-1] == ss.grid.ncy
    sim = single(lon, lat, sp, ss.grid, fs, seed)
    total, moved = 0, 0
    for step in range(6):
        total += compare_step(ss, sim, step)
        moved += sum(st.n_moved_in for st in ss.last_stats)
    print("G=%d: %d pairs, %d migrations, edges %s" % (G, total, moved, ss.edges))
    assert total > 1000 and moved > 0
    ss.close()


@pytest.mark.parametrize("peer", [False, True])
def test_in_step_record_of_a_strip(peer):
    """StripSet.step(record=slot): (ids, lon, lat, species) of the first strip's owned microbes, copied inside the step on
    the library's copy stream (lm_record_next_step_ids) -- equal to the state read back after the step, on steps with
    and without a record in between (the re-binning two steps on must wait for a lagging species copy)."""
    from lagrangian_microbes_b200.strips import LocalPeerTransport, LocalTransport, StripSet
    G, n, seed = 3, 50000, 5
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered=True)
    ids = np.arange(n, dtype=np.int32)
    cut = [slice(g, n, G) for g in range(G)]
    ss = StripSet((LocalPeerTransport if peer else LocalTransport)(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=40 * G, grid_margin=0.25, regrid_every=4)
    sim = single(lon, lat, sp, ss.grid, fs, seed, regrid_every=4, grid_margin=0.25)
    kept = []
    for step in range(9):
        slot = None if step % 3 == 1 else (len(kept) & 1)
        ss.step(record=slot)
        sim.step()
        if slot is None:
            continue
        want = ss.local_state()[0]                              # read back after the step (synchronises)
        ss.host_copies_sync()
        got = ss.record_view(slot)
        assert got[0].size == want[0].size > 0
        for a, b, what in zip(got, want, ("ids", "lon", "lat", "species")):
            assert np.array_equal(a, b), "%s of the record differ at step %d" % (what, step)
        wl, wa, ws = sim.download()                             # and they are the single handle's values of those microbes
        assert np.array_equal(wl[got[0]], got[1]) and np.array_equal(wa[got[0]], got[2]) and np.array_equal(ws[got[0]], got[3])
        kept.append(step)
    assert len(kept) == 6
    ss.close()


def test_record_by_ids_on_a_single_handle():
    """lm_record_next_step_ids without strips: the whole state in storage order, every id once."""
    from lagrangian_microbes_b200.simulation import FusedSimulation
    n, seed = 30000, 8
    fs = small_fs()
    lon, lat, sp = particles(n, seed)
    sim = FusedSimulation(lon, lat, sp, R, *P, fs, seed=seed, pair_capacity=40 * n, regrid_every=0, grid_margin=0.25)
    pin = lambda dt: torch.empty(n, dtype=dt).pin_memory()
    bufs = [(pin(torch.int32), pin(torch.float32), pin(torch.float32), pin(torch.int8)) for _ in range(2)]
    for step in range(5):
        sim.engine.record_next_step_ids(*bufs[step & 1])
        sim.step()
        sim.engine.host_copies_sync()
        assert sim.engine.record_count() == n
        ri, rl, ra, rs = (b.numpy() for b in bufs[step & 1])
        wl, wa, ws = sim.download()
        assert np.array_equal(np.sort(ri), np.arange(n))
        assert np.array_equal(wl[ri], rl) and np.array_equal(wa[ri], ra) and np.array_equal(ws[ri], rs)
    sim.engine.close()


def test_strips_with_diffusion_and_rebalancing():
    from lagrangian_microbes_b200.strips import LocalTransport, StripSet
    G, n, seed = 3, 30000, 3
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered=True)
    ids = np.arange(n, dtype=np.int32)
    row_of = lambda ss_: None
    cut = [slice(g, n, G) for g in range(G)]                # round-robin: almost everybody starts on the wrong strip
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, Kh=20.0, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=40, grid_margin=0.25, rebalance_every=2)
    sim = single(lon, lat, sp, ss.grid, fs, seed, Kh=20.0)
    edges0 = list(ss.edges)
    for step in range(6):
        compare_step(ss, sim, step)
    sizes = [s.engine.state_size() for s in ss.strips]
    print("edges %s -> %s, strip sizes %s" % (edges0, ss.edges, sizes))
    assert max(sizes) < 1.5 * n / G                          # rebalancing keeps the strips level
    ss.close()


@pytest.mark.parametrize("peer", [False, True])
def test_strips_regrid_with_the_single_handle(peer):
    """A tight grid (margin 0.03 degrees) that has to be re-fitted as the cloud moves: strips and single handle
    follow the same policy, must pick the same grids at the same steps and stay bit-identical."""
    from lagrangian_microbes_b200.strips import LocalPeerTransport, StripSet
    from lagrangian_microbes_b200.strips import LocalTransport as _Copy
    LocalTransport = LocalPeerTransport if peer else _Copy
    G, n, seed = 3, 30000, 8
    fs = small_fs()
    lon, lat, sp = particles(n, seed)
    ids = np.arange(n, dtype=np.int32)
    cut = [slice(g, n, G) for g in range(G)]
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, seed=seed, local_strips=list(range(G)), slack=3.0,
                  pairs_per_particle=40, grid_margin=0.03, regrid_every=2, cells_headroom=3.0)
    sim = single(lon, lat, sp, ss.grid, fs, seed, regrid_every=2, grid_margin=0.03)
    grids = set()
    for step in range(12):
        assert ss.grid.as_dict() == sim.grid.as_dict(), "grids diverged before step %d" % step
        grids.add(tuple(sorted(ss.grid.as_dict().items())))
        compare_step(ss, sim, step)
    print("%d different grids in 12 steps" % len(grids))
    assert len(grids) >= 2
    ss.close()


def test_exchange_buffer_overflow_is_reported():
    from lagrangian_microbes_b200._lib import LmError, LM_ENOSPC, LM_ESTATE
    from lagrangian_microbes_b200.strips import LocalTransport, StripSet
    G, n = 2, 20000
    fs = small_fs()
    lon, lat, sp = particles(n, 0)
    ids = np.arange(n, dtype=np.int32)
    cut = [slice(g, n, G) for g in range(G)]
    # 64-record messages: the initial routing (half of each tile) takes many passes but loses nobody ...
    ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                  [ids[c] for c in cut], n, R, *P, fs, local_strips=list(range(G)), slack=3.0, send_cap=64,
                  pairs_per_particle=40)
    got = np.sort(np.concatenate([p[0] for p in ss.local_state()]))
    assert np.array_equal(got, ids)
    # ... while a real step whose migrants do not fit must not pass silently
    ss.edges = [0, ss.edges[1] + 8, ss.grid.ncy]           # move the boundary by 8 rows: > 64 particles change hands
    for s in ss.strips:
        s.rows = (ss.edges[s.index], ss.edges[s.index + 1])
        s.engine.set_strip(s.rows[0], s.rows[1] - s.rows[0], s.index > 0, s.index < G - 1)
    with pytest.raises(LmError) as ei:
        ss.step(check=True)
    assert ei.value.code == LM_ESTATE
    ss.close()
    with pytest.raises(LmError) as ei:                      # the ghost row does not fit 8 records
        ss = StripSet(LocalTransport(G), [lon[c] for c in cut], [lat[c] for c in cut], [sp[c] for c in cut],
                      [ids[c] for c in cut], n, R, *P, fs, local_strips=list(range(G)), slack=3.0, ghost_cap=8)
        ss.step(check=True)
    assert ei.value.code == LM_ENOSPC


# ------------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, n, seed, steps, out_dir, kind="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lagrangian_microbes_b200.strips import DistTransport, PeerTransport, StripSet
        fs = small_fs()
        lon, lat, sp = particles(n, seed, clustered=True)
        ids = np.arange(n, dtype=np.int32)
        per = n // world
        c = slice(rank * per, (rank + 1) * per if rank < world - 1 else n)
        # "nccl": send / recv between neighbours; "peer": the neighbours' buffers mapped through CUDA IPC, peer stores + flags
        # (pairs_per_particle: strips are cut on 16-row boundaries, so the dense blob of the test lies in one or two of them)
        ss = StripSet(PeerTransport() if kind == "peer" else DistTransport(), lon[c], lat[c], sp[c], ids[c], n, R, *P, fs, seed=seed, slack=3.0,
                      pairs_per_particle=40 * world, grid_margin=0.05, regrid_every=4, cells_headroom=3.0)
        grid0 = (ss.grid.x0, ss.grid.y0, ss.grid.inv_h, ss.grid.ncx, ss.grid.ncy)      # the grid of step 0 (re-fitted later)
        rec = []
        for step in range(steps):
            ss.step(check=True)
            n_pairs, n_part, counts = ss.totals()
            lo, la, s_ = ss.gather()
            prs = ss.local_pairs()[0]
            rec.append((n_pairs, lo, la, s_, prs))
        if True:
            np.savez(os.path.join(out_dir, "rank%d.npz" % rank), grid=np.array(grid0[:3]),
                     grid_n=np.array(grid0[3:]), n_pairs=np.array([r[0] for r in rec]),
                     **{"lon%d" % k: r[1] for k, r in enumerate(rec)}, **{"lat%d" % k: r[2] for k, r in enumerate(rec)},
                     **{"sp%d" % k: r[3] for k, r in enumerate(rec)}, **{"pairs%d" % k: r[4] for k, r in enumerate(rec)})
        ss.close()
    except BaseException:
        # a rank that fails must not wait for the others (they are inside a collective or spin on a flag this rank would
        # have raised): report and leave at once, so that mp.spawn ends the remaining ranks instead of hanging the test
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,kind", [(2, "nccl"), (2, "peer"), (4, "nccl"), (4, "peer")])
def test_strips_over_nccl_equal_single_handle(tmp_path, world, kind):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from lagrangian_microbes_b200._lib import Grid
    n, seed, steps = 60000, 21, 16          # tight grid margin: the grid is re-fitted several times on the way
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, n, seed, steps, str(tmp_path), kind), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    g = parts[0]
    grid = Grid(float(g["grid"][0]), float(g["grid"][1]), float(g["grid"][2]), int(g["grid_n"][0]), int(g["grid_n"][1]))
    fs = small_fs()
    lon, lat, sp = particles(n, seed, clustered=True)
    sim = single(lon, lat, sp, grid, fs, seed, regrid_every=4, grid_margin=0.05)
    for k in range(steps):
        st = sim.step(check=True)
        wl, wa, ws = sim.download()
        assert int(g["n_pairs"][k]) == st.n_pairs
        for p in parts:                                      # every rank gathered the same global record
            assert np.array_equal(p["lon%d" % k], wl) and np.array_equal(p["lat%d" % k], wa)
            assert np.array_equal(p["sp%d" % k], ws)
        got = opairs.sort_pairs(np.concatenate([p["pairs%d" % k].reshape(-1, 2) for p in parts]))
        assert np.array_equal(got, opairs.sort_pairs(sim.pairs[:st.n_pairs].cpu().numpy()))
