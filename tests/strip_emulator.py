"""NumPy emulation of the latitude-strip protocol (TEST INFRASTRUCTURE ONLY).

Each emulated strip runs the five stages of include/lm_b200.h's staged step with the ORACLE doing the
compute (cKDTree pair query, the restated RPS loop) and the PRODUCT's host logic doing everything else:
``strips.strip_edges`` / ``strips.cell_rows`` for the partition, ``strips.EXCHANGES`` + a product transport
(``DistTransport`` over gloo, or ``LocalTransport``) for the messages.  The claim under test is the one the
CUDA path relies on: with strip boundaries on multiples of the tile height, "phases 0-11 (inside tiles
and across vertical tile boundaries), hand the first row's species south, phases 12-14 (across horizontal tile
boundaries), hand them back" equals the single-domain sequential loop in the canonical tile-round order -- for any
number of strips.
"""
import numpy as np
import torch

from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps

CAP = 4096          # records per message (fixed capacity, live count in the header)
P_RPS = (0.55, 0.6, 0.9)
RADIUS = 0.01


def displacement(ids, step):
    """Deterministic per-particle drift (float32), large enough to cross strip boundaries."""
    a = (ids.astype(np.float64) * 0.7548776662466927 + 0.31 * step) % 1.0
    b = (ids.astype(np.float64) * 0.5698402909980532 + 0.17 * step) % 1.0
    return ((a - 0.5) * 0.03).astype(np.float32), ((b - 0.5) * 0.03).astype(np.float32)


def grid_dict(g):
    return dict(x0=g.x0, y0=g.y0, inv_h=g.inv_h, ncx=g.ncx, ncy=g.ncy)


class NumpyStrip:
    def __init__(self, index, n_strips, grid, rows, lon, lat, sp, ids):
        self.index, self.n_strips, self.grid, self.rows = index, n_strips, grid, rows
        self.lon, self.lat, self.sp, self.ids = (np.array(a) for a in (lon, lat, sp, ids))
        z = lambda: torch.zeros(4 * (CAP + 1), dtype=torch.int32)
        self.buffers = {"mig_send": [z(), z()], "mig_recv": [z(), z()], "ghost_send": z(), "ghost_recv": z(),
                        "gsp_send": z(), "gsp_recv": z(), "gret_send": z(), "gret_recv": z()}
        self.n_ghost = 0
        self.pairs = np.zeros((0, 2), dtype=np.int64)

    def _rows_of(self, lat):
        from lagrangian_microbes_b200.strips import cell_rows
        return cell_rows(lat, self.grid)

    @staticmethod
    def _pack(buf, *cols):
        a = buf.numpy().reshape(-1, 4)
        n = cols[0].size
        assert n <= CAP
        a[0, 0] = n
        for k, c in enumerate(cols):
            a[1:1 + n, k] = c.view(np.int32) if c.dtype == np.float32 else c.astype(np.int32)

    @staticmethod
    def _unpack(buf, dtypes):
        a = buf.numpy().reshape(-1, 4)
        n = int(a[0, 0])
        out = []
        for k, dt in enumerate(dtypes):
            col = a[1:1 + n, k].copy()
            out.append(col.view(np.float32) if dt == np.float32 else col.astype(dt))
        return out

    # ---- the five stages -----------------------------------------------------------------------------
    def move(self, step, advect=True):
        if advect:
            dx, dy = displacement(self.ids, step)
            self.lon = (self.lon + dx).astype(np.float32)
            self.lat = (self.lat + dy).astype(np.float32)
        row = self._rows_of(self.lat)
        south = row < self.rows[0] if self.index > 0 else np.zeros(row.size, bool)
        north = row >= self.rows[1] if self.index < self.n_strips - 1 else np.zeros(row.size, bool)
        for side, m in ((0, south), (1, north)):
            self._pack(self.buffers["mig_send"][side], self.lon[m], self.lat[m], self.ids[m], self.sp[m])
        keep = ~(south | north)
        self.lon, self.lat, self.sp, self.ids = self.lon[keep], self.lat[keep], self.sp[keep], self.ids[keep]

    def bin(self):
        for side in (0, 1):
            if (side == 0 and self.index == 0) or (side == 1 and self.index == self.n_strips - 1):
                continue
            lon, lat, ids, sp = self._unpack(self.buffers["mig_recv"][side], (np.float32, np.float32, np.int32, np.int8))
            self.lon, self.lat = np.concatenate((self.lon, lon)), np.concatenate((self.lat, lat))
            self.ids, self.sp = np.concatenate((self.ids, ids)), np.concatenate((self.sp, sp))
        g = self.grid
        cx = opairs.cell_index(self.lon, g.x0, g.inv_h, g.ncx)
        cy = self._rows_of(self.lat)
        self.misrouted = int(np.sum((cy < self.rows[0]) | (cy >= self.rows[1])))
        order = np.lexsort((self.ids, cy * g.ncx + cx))
        self.lon, self.lat, self.sp, self.ids = self.lon[order], self.lat[order], self.sp[order], self.ids[order]
        self.row = cy[order]
        first = self.row == self.rows[0]
        self.n_row0 = int(first.sum())
        if self.index > 0:
            self._pack(self.buffers["ghost_send"], self.lon[first], self.lat[first], self.ids[first])

    def interact_begin(self, step, seed):
        n = self.lon.size
        lon, lat, ids = self.lon, self.lat, self.ids
        if self.index < self.n_strips - 1:
            glon, glat, gid = self._unpack(self.buffers["ghost_recv"], (np.float32, np.float32, np.int32))
            lon, lat, ids = np.concatenate((lon, glon)), np.concatenate((lat, glat)), np.concatenate((ids, gid))
        self.n_ghost = lon.size - n
        self.sp_all = np.concatenate((self.sp, np.zeros(self.n_ghost, np.int8)))
        loc = opairs.query_pairs_reference_array(lon, lat, RADIUS).reshape(-1, 2)
        loc = loc[(loc[:, 0] < n) | (loc[:, 1] < n)]              # ghost-ghost pairs belong to the strip to the north
        # the tile-round order ranks the microbes of a cell by storage index; storage is in (cell, id) order on both
        # sides of the boundary (the ghost row arrives sorted), so that is the rank by particle id the device uses
        self.order, self.phase = orps.tile_round_order(loc, lon, lat, grid_dict(self.grid))
        gi, gj = ids[self.order[:, 0]].astype(np.int64), ids[self.order[:, 1]].astype(np.int64)
        self.u = philox.pair_uniforms(np.minimum(gi, gj), np.maximum(gi, gj), step, seed)
        self.pairs = np.stack((np.minimum(gi, gj), np.maximum(gi, gj)), -1)
        lo = self.phase <= 11             # the tile phases and the vertical-boundary phases
        self.sp_all, _ = orps.rps_sequential_c(self.sp_all, self.order[lo], self.u[lo], *P_RPS)
        if self.index > 0:
            self._pack(self.buffers["gsp_send"], self.sp_all[:self.n_row0])

    def interact_end(self):
        n = self.lon.size
        if self.index < self.n_strips - 1:
            (gsp,) = self._unpack(self.buffers["gsp_recv"], (np.int8,))
            assert gsp.size == self.n_ghost
            self.sp_all[n:] = gsp
        hi = self.phase >= 12             # the horizontal-boundary phases: the only ones that cross a strip boundary
        self.sp_all, _ = orps.rps_sequential_c(self.sp_all, self.order[hi], self.u[hi], *P_RPS)
        if self.index < self.n_strips - 1:
            self._pack(self.buffers["gret_send"], self.sp_all[n:])
        self.sp = self.sp_all[:n].copy()

    def finish(self):
        if self.index > 0:
            (back,) = self._unpack(self.buffers["gret_recv"], (np.int8,))
            assert back.size == self.n_row0
            self.sp[:self.n_row0] = back


def run_strips(transport, strips, n_steps, seed, advect=True):
    """Drive emulated strips through n_steps; returns per-step list of per-strip pair arrays."""
    all_pairs = []
    for step in range(n_steps):
        for s in strips:
            s.move(step, advect)
        transport.exchange("mig", strips)
        for s in strips:
            s.bin()
        transport.exchange("ghost", strips)
        for s in strips:
            s.interact_begin(step, seed)
        transport.exchange("gsp", strips)
        for s in strips:
            s.interact_end()
        transport.exchange("gret", strips)
        for s in strips:
            s.finish()
        all_pairs.append([s.pairs for s in strips])
    return all_pairs


def run_single(grid, lon, lat, sp, ids, n_steps, seed):
    """The single-domain oracle loop on the same inputs: per step (lon, lat, species by id, sorted pairs)."""
    lon, lat, sp = lon.copy(), lat.copy(), sp.copy()
    out = []
    for step in range(n_steps):
        dx, dy = displacement(ids, step)
        lon, lat = (lon + dx).astype(np.float32), (lat + dy).astype(np.float32)
        prs = opairs.query_pairs_reference_array(lon, lat, RADIUS)
        order, _ = orps.tile_round_order(prs, lon, lat, grid_dict(grid))
        u = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
        sp, _ = orps.rps_sequential_c(sp, order, u, *P_RPS)
        out.append((lon.copy(), lat.copy(), sp.copy(), opairs.sort_pairs(prs)))
    return out


def make_case(n, seed):
    """Particles in a ~0.6 x 0.6 degree patch (ids = index), the global grid and the reference solution inputs."""
    from lagrangian_microbes_b200.engine import make_grid
    rng = np.random.default_rng(seed)
    lon = (205.0 + 0.6 * rng.random(n)).astype(np.float32)
    lat = (30.0 + 0.6 * rng.random(n)).astype(np.float32)
    sp = rng.integers(1, 4, n).astype(np.int8)
    ids = np.arange(n, dtype=np.int32)
    grid = make_grid(float(lon.min()), float(lon.max()), float(lat.min()), float(lat.max()), RADIUS, n, 1 << 20,
                     margin=0.1, cells_per_particle=2.0)
    return grid, lon, lat, sp, ids


def gloo_worker(rank, world, port, n, n_steps, seed, result_dir):
    """One rank of the gloo test: emulated strip `rank`, product DistTransport."""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lagrangian_microbes_b200.strips import DistTransport, cell_rows, strip_edges
        grid, lon, lat, sp, ids = make_case(n, seed)
        # every rank starts from a contiguous TILE of the particles (the reference's split, particle_advecter.py:38-66)
        per = n // world
        mine = slice(rank * per, (rank + 1) * per if rank < world - 1 else n)
        T = DistTransport()
        hist = np.bincount(cell_rows(lat[mine], grid), minlength=grid.ncy).astype(np.float64)
        edges = strip_edges(T.all_sum([hist]), world)
        s = NumpyStrip(rank, world, grid, (edges[rank], edges[rank + 1]), lon[mine], lat[mine], sp[mine], ids[mine])
        # settle: route the tile's particles to their strips (one hop per pass)
        for _ in range(world + 1):
            s.move(0, advect=False)
            T.exchange("mig", [s])
            s.bin()
            if T.all_sum([[s.misrouted]])[0] == 0:
                break
        else:
            raise AssertionError("settle did not converge")
        pairs = run_strips(T, [s], n_steps, seed)
        np.savez(os.path.join(result_dir, "rank%d.npz" % rank), ids=s.ids, lon=s.lon, lat=s.lat, sp=s.sp,
                 edges=np.array(edges), **{"pairs%d" % k: p[0] for k, p in enumerate(pairs)})
    finally:
        dist.destroy_process_group()
