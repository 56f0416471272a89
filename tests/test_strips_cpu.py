"""Host-side logic of the latitude-strip decomposition, on CPU: the partitioner, and the exchange protocol
driven through the product's transports (``DistTransport`` over gloo with world_size 2 and 3,
``LocalTransport`` in-process) by a NumPy emulation of the strip stages (tests/strip_emulator.py), checked
against the single-domain oracle loop."""
import os
import socket

import numpy as np
import pytest
import torch

import strip_emulator as emu
from oracle import pairs as opairs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_strip_edges_are_tile_aligned_balanced_and_cover_all_rows():
    from lagrangian_microbes_b200.strips import ROW_ALIGN, strip_edges
    assert ROW_ALIGN == 16                                       # LM_TILE_H: the tile height of the fused interaction pass
    rng = np.random.default_rng(0)
    for ncy, G in ((40, 2), (200, 8), (64, 3), (1000, 8), (130, 8), (18, 2), (7, 1)):
        for kind in ("uniform", "blob", "random"):
            if kind == "uniform":
                h = np.full(ncy, 100)
            elif kind == "blob":
                h = np.zeros(ncy, dtype=int)
                h[ncy // 3: ncy // 3 + max(2, ncy // 10)] = 1000
            else:
                h = rng.integers(0, 500, ncy)
            e = strip_edges(h, G)
            assert e[0] == 0 and e[-1] == ncy and len(e) == G + 1
            assert all(b - a >= 16 for a, b in zip(e[:-2], e[1:-1])) and e[-1] - e[-2] >= 2
            assert all(x % 16 == 0 for x in e[:-1])
            if kind == "uniform" and ncy >= 64 * G:
                share = np.add.reduceat(h, e[:-1])
                assert share.max() - share.min() <= 2 * 100 * 16     # within two aligned rows of each other
    with pytest.raises(ValueError):
        strip_edges(np.ones(7), 4)
    with pytest.raises(ValueError):
        strip_edges(np.ones(17), 2)
    assert strip_edges(np.ones(101), 8, align=2)[1] % 2 == 0     # the round-1 pipeline only needs even rows
    e = strip_edges(np.r_[np.zeros(30), np.full(10, 50)], 2, max_rows=24)
    assert max(b - a for a, b in zip(e[:-1], e[1:])) <= 24


def test_remap_edges_keeps_latitudes_aligned_rows_and_room_for_every_strip():
    from lagrangian_microbes_b200.engine import make_grid
    from lagrangian_microbes_b200.strips import remap_edges, strip_edges
    old = make_grid(200.0, 210.0, 20.0, 40.0, 0.01, 4_000_000, 1 << 24)
    edges = strip_edges(np.full(old.ncy, 10), 8)
    for box in ((200.5, 211.0, 21.0, 41.5), (199.0, 209.0, 18.5, 38.0), (200.0, 210.0, 20.0, 40.0)):
        for r in (0.01, 0.02):
            new = make_grid(*box, r, 4_000_000, 1 << 24)
            e = remap_edges(old, edges, new)
            assert e[0] == 0 and e[-1] == new.ncy and len(e) == 9
            assert all(x % 16 == 0 for x in e[:-1]) and all(b - a >= 16 for a, b in zip(e[:-2], e[1:-1])) and e[-1] - e[-2] >= 2
            lat_old = [old.y0 + k / old.inv_h for k in edges[1:-1]]
            lat_new = [new.y0 + k / new.inv_h for k in e[1:-1]]
            inside = [(a, b) for a, b in zip(lat_old, lat_new) if new.y0 + 4 / new.inv_h < a < new.y0 + (new.ncy - 4) / new.inv_h]
            assert inside and all(abs(a - b) <= 8.01 / new.inv_h for a, b in inside)       # within half a tile height
    tiny = make_grid(200.0, 200.01, 20.0, 20.01, 0.01, 100, 1 << 16, margin=0.0)
    with pytest.raises(ValueError):
        remap_edges(old, edges, tiny)


def test_cell_rows_matches_the_oracle_cell_index():
    from lagrangian_microbes_b200.engine import make_grid
    from lagrangian_microbes_b200.strips import cell_rows
    rng = np.random.default_rng(1)
    lat = np.concatenate((rng.uniform(20, 40, 100000), [-1e9, 1e9, 25.0])).astype(np.float32)
    g = make_grid(200.0, 210.0, 25.0, 35.0, 0.01, 1000000, 1 << 22)
    assert np.array_equal(cell_rows(lat, g), opairs.cell_index(lat, g.y0, g.inv_h, g.ncy))
    assert cell_rows(np.array([np.nan], dtype=np.float32), g)[0] == 0          # the device bins NaN into row 0


def _check_against_single(n, seed, n_steps, parts):
    """parts: list of dicts(ids, lon, lat, sp, pairs<k>) over strips."""
    grid, lon, lat, sp, ids = emu.make_case(n, seed)
    want = emu.run_single(grid, lon, lat, sp, ids, n_steps, seed)
    all_ids = np.concatenate([p["ids"] for p in parts])
    assert np.array_equal(np.sort(all_ids), ids), "particles lost or duplicated by migration"
    got_lon, got_lat, got_sp = np.empty_like(lon), np.empty_like(lat), np.empty_like(sp)
    for p in parts:
        got_lon[p["ids"]], got_lat[p["ids"]], got_sp[p["ids"]] = p["lon"], p["lat"], p["sp"]
    w_lon, w_lat, w_sp, _ = want[-1]
    assert np.array_equal(got_lon, w_lon) and np.array_equal(got_lat, w_lat)
    for k in range(n_steps):
        prs = np.concatenate([p["pairs%d" % k].reshape(-1, 2) for p in parts])
        assert np.array_equal(opairs.sort_pairs(prs), want[k][3]), "pair set differs at step %d" % k
    assert want[-1][3].shape[0] > 100                      # the case does interact
    assert np.array_equal(got_sp, w_sp), "species differ from the single-domain sequential loop"


@pytest.mark.parametrize("G", [2, 3, 5])
def test_protocol_with_local_transport_equals_single_domain(G):
    from lagrangian_microbes_b200.strips import LocalTransport, cell_rows, strip_edges
    n, seed, n_steps = 6000, 3, 4
    grid, lon, lat, sp, ids = emu.make_case(n, seed)
    edges = strip_edges(np.bincount(cell_rows(lat, grid), minlength=grid.ncy), G)
    row = cell_rows(lat, grid)
    strips = []
    for g in range(G):
        m = (row >= edges[g]) & (row < edges[g + 1])
        strips.append(emu.NumpyStrip(g, G, grid, (edges[g], edges[g + 1]), lon[m], lat[m], sp[m], ids[m]))
    pairs = emu.run_strips(LocalTransport(G), strips, n_steps, seed)
    parts = []
    for g, s in enumerate(strips):
        d = dict(ids=s.ids, lon=s.lon, lat=s.lat, sp=s.sp)
        d.update({"pairs%d" % k: pairs[k][g] for k in range(n_steps)})
        parts.append(d)
    _check_against_single(n, seed, n_steps, parts)
    moved = sum(int(np.sum((cell_rows(s.lat, grid) < s.rows[0]) | (cell_rows(s.lat, grid) >= s.rows[1]))) for s in strips)
    assert moved == 0


@pytest.mark.parametrize("world", [2, 3])
def test_protocol_over_gloo_equals_single_domain(tmp_path, world):
    import torch.multiprocessing as mp
    n, seed, n_steps = 6000, 5, 3
    mp.spawn(emu.gloo_worker, args=(world, _free_port(), n, n_steps, seed, str(tmp_path)), nprocs=world, join=True)
    parts = [dict(np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))) for r in range(world)]
    assert all(np.array_equal(p["edges"], parts[0]["edges"]) for p in parts)
    _check_against_single(n, seed, n_steps, parts)
