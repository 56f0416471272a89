"""Latitude-strip sharding of the fused step over the GPUs of one box.

The reference parallelises only advection, over contiguous particle tiles with no communication
(/root/reference/particle_advecter.py:38-66,143-148); its interaction phase is one serial process
(/root/reference/interaction_simulator.py:82-117).  Here every GPU runs the whole step -- advection,
binning, pair search, rock-paper-scissors -- on the particles of one latitude strip of a cell grid shared
by all GPUs, and exchanges with its two neighbours only:

    stage (C ABI)            then exchange                        direction
    lm_step_move             "mig"   particles that left the strip   both ways
    lm_step_bin              "ghost" first owned row (pos, id, cells) north -> south
    lm_step_interact_begin   "gsp"   that row's species after phase 5 north -> south
    lm_step_interact_end     "gret"  the same species after phase 8  south -> north
    lm_step_finish

Strip boundaries sit on multiples of 16 cell rows (the tile height of the fused interaction pass), which makes the pair set, the species and the positions
bit-identical to a single GPU running the same grid (tests/test_gpu_strips.py).

``StripSet`` drives any number of strips held by THIS process: one per process under torchrun
(``DistTransport``: NCCL send/recv between neighbour ranks through torch.distributed), or several on one
device in one process (``LocalTransport``: device copies; used by the single-GPU tests).
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import RpsParams
from .engine import Engine, make_grid

# (kind, send buffer, side it leaves through, receive buffer at the neighbour)
#   side 0 = south, 1 = north;  a message leaving through side s arrives at the neighbour's side 1 - s
EXCHANGES = {
    "mig": [("mig_send", 0, "mig_recv"), ("mig_send", 1, "mig_recv")],
    "ghost": [("ghost_send", 0, "ghost_recv")],
    "gsp": [("gsp_send", 0, "gsp_recv")],
    "gret": [("gret_send", 1, "gret_recv")],
}


def cell_rows(lat, grid):
    """Global cell row of float32 latitudes, bit for bit what csrc/bin.cu::cell_coord computes."""
    q = np.floor((np.asarray(lat, dtype=np.float32).astype(np.float64) - grid.y0) * grid.inv_h)
    q = np.where(q >= 0.0, q, 0.0)                       # also NaN -> 0
    return np.minimum(q, grid.ncy - 1).astype(np.int64)


ROW_ALIGN = _lib.LM_TILE_H      # strip boundaries sit on multiples of the tile height of the fused interaction pass


def _snap(e, align):
    return int(e) - int(e) % align


def strip_edges(row_counts, n_strips, max_rows=None, align=ROW_ALIGN):
    """Row boundaries e[0] = 0 < e[1] < ... < e[G] = ncy, all interior ones multiples of ``align`` (the tile height of
    the fused interaction pass; 16 is also even, which is all the round-1 pipeline needs), splitting the particles of
    ``row_counts`` (particles per global cell row) as evenly as the rows allow.  Every strip gets at least
    ``align`` rows and (optionally) at most ``max_rows``."""
    row_counts = np.asarray(row_counts, dtype=np.int64)
    ncy, G = int(row_counts.size), int(n_strips)
    if G < 1 or (G > 1 and ncy < align * (G - 1) + 2):
        raise ValueError("need at least %d cell rows per strip (ncy=%d, strips=%d)" % (align, ncy, G))
    cum = np.concatenate(([0], np.cumsum(row_counts)))   # cum[e] = particles in rows < e
    total = int(cum[-1])
    edges = [0]
    for g in range(1, G):
        target = total * g / float(G)
        e = int(np.searchsorted(cum, target, side="left"))
        # nearest aligned row to the ideal cut
        lo_e = _snap(e, align)
        hi_e = lo_e if lo_e == e else lo_e + align
        e = lo_e if abs(cum[min(lo_e, ncy)] - target) <= abs(cum[min(hi_e, ncy)] - target) else hi_e
        lo = edges[-1] + align
        hi = _snap(ncy - 2 - align * (G - g - 1), align)          # the strips above keep >= align rows, the last >= 2
        if max_rows is not None:
            hi = min(hi, edges[-1] + _snap(max_rows, align))
            # leave the remaining strips enough room under max_rows as well
            need = ncy - max_rows * (G - g)
            lo = max(lo, need + (-need) % align)
        e = max(lo, min(e, hi))
        edges.append(e)
    edges.append(ncy)
    if any(b <= a for a, b in zip(edges[:-1], edges[1:])):
        raise ValueError("strips do not fit: %s (ncy=%d, align=%d)" % (edges, ncy, align))
    if max_rows is not None and max(b - a for a, b in zip(edges[:-1], edges[1:])) > max_rows:
        raise ValueError("strips do not fit max_rows=%d: %s" % (max_rows, edges))
    return edges


def remap_edges(old_grid, old_edges, new_grid, align=ROW_ALIGN):
    """Strip boundaries on a re-fitted grid: every interior boundary keeps (nearly) its latitude, snapped to a
    multiple of ``align`` rows of the new grid, every strip keeps at least ``align`` rows."""
    G = len(old_edges) - 1
    if G > 1 and new_grid.ncy < align * (G - 1) + 2:
        raise ValueError("the cell grid has %d rows: too few for %d strips" % (new_grid.ncy, G))
    edges = [0]
    for k in range(1, G):
        lat_edge = old_grid.y0 + old_edges[k] / old_grid.inv_h
        e = align * int(round((lat_edge - new_grid.y0) * new_grid.inv_h / float(align)))
        hi = _snap(new_grid.ncy - 2 - align * (G - k - 1), align)
        e = max(edges[-1] + align, min(e, hi))
        edges.append(e)
    edges.append(new_grid.ncy)
    return edges


# ------------------------------------------------------------------------------------------------------
# transports
# ------------------------------------------------------------------------------------------------------
class LocalTransport:
    """All strips live in this process (same device): the exchange is a device copy on the current stream."""

    def __init__(self, n_strips):
        self.n_strips = n_strips

    def exchange(self, kind, strips):
        by_index = {s.index: s for s in strips}
        for name, side, rname in EXCHANGES[kind]:
            for s in strips:
                nb = by_index.get(s.index + (1 if side else -1))
                if nb is None:
                    continue
                src = s.buffers[name]
                dst = nb.buffers[rname]
                src = src[side] if isinstance(src, list) else src
                dst = dst[1 - side] if isinstance(dst, list) else dst
                dst.copy_(src, non_blocking=True)

    def all_sum(self, values):
        return np.sum(np.asarray(values, dtype=np.float64), axis=0)

    def all_max(self, values):
        return np.max(np.asarray(values, dtype=np.float64), axis=0)

    def all_gather_rows(self, per_strip):
        return list(per_strip)


class DistTransport:
    """One strip per rank of a torch.distributed group; neighbours are rank - 1 (south) and rank + 1 (north).

    NCCL point-to-point over NVLink for CUDA buffers (the transfers are ordered after the kernels already
    queued on the current stream and the following kernels wait for them -- no host synchronisation);
    gloo for CPU tensors (tests/test_strips_gloo.py)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.n_strips = dist.get_world_size(group)

    def exchange(self, kind, strips):
        dist = self.dist
        (s,) = strips
        assert s.index == self.rank
        ops = []
        for name, side, rname in EXCHANGES[kind]:
            peer = self.rank + (1 if side else -1)
            if 0 <= peer < self.n_strips:
                src = s.buffers[name]
                ops.append(dist.P2POp(dist.isend, src[side] if isinstance(src, list) else src, peer, self.group))
            # the matching receive: the message that leaves the OTHER neighbour through its side `side`
            peer = self.rank - (1 if side else -1)
            if 0 <= peer < self.n_strips:
                dst = s.buffers[rname]
                ops.append(dist.P2POp(dist.irecv, dst[1 - side] if isinstance(dst, list) else dst, peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _reduce(self, values, op):
        (v,) = values
        t = torch.as_tensor(np.asarray(v, dtype=np.float64))
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def all_sum(self, values):
        return self._reduce(values, self.dist.ReduceOp.SUM)

    def all_max(self, values):
        return self._reduce(values, self.dist.ReduceOp.MAX)

    def all_gather_rows(self, per_strip):
        (v,) = per_strip
        out = [None] * self.n_strips
        self.dist.all_gather_object(out, v, group=self.group)
        return out


class _PeerMixin:
    """Peer-memory exchange (include/lm_b200.h: lm_peer_export): the packing kernels of a strip write straight into the
    neighbour's receive buffers over NVLink / NVSwitch, a flag store follows each message and the consuming stage spins on its
    flag -- stream-ordered on both sides, no NCCL send / recv on the data path.  ``connect`` is called once by StripSet."""

    XCHG = {"mig": _lib.LM_XCHG_MIG, "ghost": _lib.LM_XCHG_GHOST, "gsp": _lib.LM_XCHG_GSP, "gret": _lib.LM_XCHG_GRET}
    peer = True

    def exchange(self, kind, strips):
        for s in strips:
            s.engine.step_push(self.XCHG[kind])


class LocalPeerTransport(_PeerMixin, LocalTransport):
    """Several strips in this process, exchanging through each other's device pointers (single-GPU tests of the peer path)."""

    def connect(self, strips):
        exports = {s.index: s.engine.peer_export() for s in strips}
        for s in strips:
            for side, nb in ((0, s.index - 1), (1, s.index + 1)):
                if nb in exports:
                    s.engine.peer_connect(side, exports[nb], use_ipc=False)


class PeerTransport(_PeerMixin, DistTransport):
    """One strip per rank; the neighbours' buffers are mapped through CUDA IPC.  torch.distributed carries the handles once
    (and the small host-side reductions: counts, bounding boxes); the per-step messages never touch it."""

    def connect(self, strips):
        (s,) = strips
        exports = [None] * self.n_strips
        self.dist.all_gather_object(exports, s.engine.peer_export(), group=self.group)
        for side, nb in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= nb < self.n_strips:
                s.engine.peer_connect(side, exports[nb], use_ipc=True)
        self.dist.barrier(group=self.group)            # nobody steps before everybody is connected (flags zeroed)


# ------------------------------------------------------------------------------------------------------
class Strip:
    """One latitude strip: an Engine (one lm_handle) + its exchange buffers + its pair list."""

    def __init__(self, index, engine, pairs):
        self.index = index
        self.engine = engine
        self.pairs = pairs
        self.buffers = engine.strip_buffers()
        self.rows = (0, 0)


class StripSet:
    """The fused advect -> bin -> pair search -> RPS loop over latitude strips.

    Every process passes the particles IT holds (any subset, e.g. the reference's contiguous tiles) with
    their GLOBAL ids; the constructor fits one global cell grid, cuts it into strips of (nearly) equal
    particle counts on even rows and routes every particle to its strip.
    """

    def __init__(self, transport, lons, lats, species, ids, n_total, radius, pRS, pPR, pSP, fieldset, dt_seconds=3600.0,
                 Kh=0.0, seed=0, emit_pairs=True, pairs_per_particle=8, slack=1.3, send_cap=None, ghost_cap=None,
                 grid_margin=0.5, cells_per_particle=2.0, local_strips=None, device=None, interact=True, advect=True,
                 rebalance_every=0, stream_field=False, regrid_every=16, cells_headroom=1.5, interaction_norm=2):
        from .particle_advecter import StageClock
        self.transport = transport
        G = transport.n_strips
        self.n_strips = G
        self.n_total = int(n_total)
        self.radius = float(radius)
        self.rps = RpsParams(float(pRS), float(pPR), float(pSP), int(seed), 0)
        self.dt = float(dt_seconds)
        self.Kh_deg2 = float(Kh) / 1e10                          # particle_advecter.py:121
        self.diffuse_amp = float(np.sqrt(6 * np.fabs(np.float32(self.dt)) * self.Kh_deg2))
        self.iteration = 0
        self.interact = bool(interact)
        self.advect = bool(advect)
        self.emit_pairs = bool(emit_pairs) and self.interact
        self.rebalance_every = int(rebalance_every)
        self.regrid_every = int(regrid_every)
        self.grid_margin = float(grid_margin)
        self.cells_per_particle = float(cells_per_particle)
        # which strips this process holds; lons/lats/... are then lists, one entry per local strip
        if local_strips is None:
            local_strips = [transport.rank]
            lons, lats, species, ids = [lons], [lats], [species], [ids]
        self.local = list(local_strips)
        lons = [np.ascontiguousarray(a, dtype=np.float32) for a in lons]
        lats = [np.ascontiguousarray(a, dtype=np.float32) for a in lats]
        species = [np.ascontiguousarray(a, dtype=np.int8) for a in species]
        ids = [np.ascontiguousarray(a, dtype=np.int32) for a in ids]

        # ---- one global grid, fitted to the global bounding box
        big = 1e30
        box = [[-(a.min() if a.size else big) for a in lons], [(a.max() if a.size else -big) for a in lons],
               [-(a.min() if a.size else big) for a in lats], [(a.max() if a.size else -big) for a in lats]]
        m = transport.all_max([[box[0][k], box[1][k], box[2][k], box[3][k]] for k in range(len(self.local))])
        x0, x1, y0, y1 = -m[0], m[1], -m[2], m[3]
        self.per_strip = int(math.ceil(self.n_total / float(G)))
        self.max_particles = int(slack * self.per_strip) + 1024
        cells_budget = int(max(4 * cells_per_particle * self.n_total, 1 << 20))
        self.grid = make_grid(x0, x1, y0, y1, self.radius, self.n_total, cells_budget, margin=grid_margin,
                              cells_per_particle=cells_per_particle)
        g = self.grid
        if G > 1 and g.ncy < ROW_ALIGN * (G - 1) + 2:
            raise ValueError("the cell grid has %d rows: too few for %d strips (boundaries sit on multiples of %d rows)"
                             % (g.ncy, G, ROW_ALIGN))
        # rows a strip may own: room for the balanced share with the same slack as the particles
        self.max_rows = min(g.ncy, max(2 * ROW_ALIGN, int(slack * math.ceil(g.ncy / float(G))) + ROW_ALIGN))
        self.max_cells = int(cells_headroom * (self.max_rows + 1) * g.ncx)      # room for re-fitted grids
        self._bbox = (x0, x1, y0, y1)
        row_density = self.n_total / float(g.ncy)
        self.ghost_cap = int(ghost_cap if ghost_cap is not None else max(4096, 4 * row_density))
        self.send_cap = int(send_cap if send_cap is not None else max(4096, 8 * row_density))

        # ---- strip boundaries from the global row histogram
        hist = [np.bincount(cell_rows(a, g), minlength=g.ncy).astype(np.float64) for a in lats]
        self.row_counts = transport.all_sum(hist)
        self.edges = strip_edges(self.row_counts, G, self.max_rows)

        self.fieldset = fieldset
        self.clock = StageClock(fieldset.time) if fieldset is not None else None
        assert fieldset is not None or not self.advect, "advection needs a fieldset"
        self.strips = []
        pair_cap = int(max(1 << 16, pairs_per_particle * self.max_particles))
        for k, idx in enumerate(self.local):
            eng = Engine(max_particles=self.max_particles + self.ghost_cap, max_cells=self.max_cells,
                         max_pairs=pair_cap if self.interact else 0, device=device)
            eng.strip_alloc(self.send_cap, self.ghost_cap, int(cells_headroom * g.ncx) + 8)
            eng.set_norm(interaction_norm)
            eng.set_grid(g)
            if fieldset is not None and not stream_field:
                eng.set_field(*fieldset.to_device(eng.device))
            pairs = torch.empty((pair_cap, 2), dtype=torch.int32, device=eng.device) if self.emit_pairs else None
            self.strips.append(Strip(idx, eng, pairs))
        self.streamer = None
        if fieldset is not None and stream_field:
            from .simulation import FieldWindowStreamer
            self.streamer = FieldWindowStreamer(fieldset, [s.engine for s in self.strips])
        self._record = None
        self._record_pinned = None
        self._record_slot_count = [None, None]
        self._record_last = None
        if getattr(transport, "peer", False):
            transport.connect(self.strips)
        self._apply_edges()
        for s, lo, la, sp, i in zip(self.strips, lons, lats, species, ids):
            if lo.size > s.engine.max_particles:
                raise ValueError("strip %d was handed %d particles, capacity %d" % (s.index, lo.size, s.engine.max_particles))
            dev = s.engine.device
            s.engine.state_set(torch.from_numpy(lo).to(dev), torch.from_numpy(la).to(dev), torch.from_numpy(sp).to(dev),
                               torch.from_numpy(i).to(dev))
        self.settle()
        self.last_stats = None

    # ---- strip geometry -----------------------------------------------------------------------------
    def _apply_edges(self):
        G = self.n_strips
        for s in self.strips:
            r0, r1 = self.edges[s.index], self.edges[s.index + 1]
            s.engine.set_strip(r0, r1 - r0, s.index > 0, s.index < G - 1)
            s.rows = (r0, r1)

    def _staged(self, flags, st_times=None):
        T, S = self.transport, self.strips
        for s in S:
            s.engine.step_move(flags, st_times, self.dt, self.diffuse_amp, self.rps)
        halo = bool(flags & _lib.LM_STEP_INTERACT)        # routing passes move particles only
        T.exchange("mig", S)
        for s in S:
            s.engine.step_bin()
        if halo:
            T.exchange("ghost", S)
        for s in S:
            s.engine.step_interact_begin(self.radius, s.pairs)
        if halo:
            T.exchange("gsp", S)
        for s in S:
            s.engine.step_interact_end()
        if halo:
            T.exchange("gret", S)
        for s in S:
            s.engine.step_finish()

    def settle(self, max_passes=100000):
        """Route every particle to the strip that owns its row: empty steps (no advection, no interaction)
        until no particle is held by a strip it does not belong to.  One hop and at most ``send_cap``
        particles per boundary and pass."""
        passes, best, stalled = 0, None, 0
        while True:
            for s in self.strips:       # (re)declare the strip: marks the state as not binned, so the pass re-bins
                s.engine.set_strip(s.rows[0], s.rows[1] - s.rows[0], s.index > 0, s.index < self.n_strips - 1)
            self._staged(0)
            mis = [[s.engine.sync_stats(allow_misrouted=True).n_misrouted] for s in self.strips]
            left = int(self.transport.all_sum(mis)[0])
            if left == 0:
                return passes
            passes += 1
            stalled = 0 if (best is None or left < best) else stalled + 1
            best = left if best is None else min(best, left)
            if passes > max_passes or stalled > self.n_strips + 2:
                raise RuntimeError("particles could not be routed to their strips (%d left)" % left)

    def regrid(self, x0, x1, y0, y1):
        """Re-fit the global cell grid to the bounding box (x0, x1, y0, y1) of all particles, keep every strip
        boundary at (nearly) the latitude it had, and route the particles of the rows that changed hands."""
        old, old_edges = self.grid, self.edges
        g = make_grid(x0, x1, y0, y1, self.radius, self.n_total, int(max(4 * self.cells_per_particle * self.n_total, 1 << 20)),
                      margin=self.grid_margin, cells_per_particle=self.cells_per_particle)
        G = self.n_strips
        edges = remap_edges(old, old_edges, g)
        rows = max(b - a for a, b in zip(edges[:-1], edges[1:]))
        if (rows + 1) * g.ncx > self.max_cells or g.ncy < 2 * G:
            raise RuntimeError("the re-fitted cell grid (%d x %d) does not fit the strips' cell tables" % (g.ncx, g.ncy))
        self.grid, self.edges, self._bbox = g, edges, (x0, x1, y0, y1)
        for s in self.strips:
            s.engine.set_grid(g)
            s.rows = (edges[s.index], edges[s.index + 1])
        self.settle()
        return g

    def _maybe_regrid(self, stats):
        """Same policy as FusedSimulation._maybe_regrid, on the bounding box of ALL strips."""
        big = 1e30
        box = [[-st.bbox[0] if st.n_particles else -big, st.bbox[1] if st.n_particles else -big,
                -st.bbox[2] if st.n_particles else -big, st.bbox[3] if st.n_particles else -big] for st in stats]
        m = self.transport.all_max(box)
        x0, x1, y0, y1 = -m[0], m[1], -m[2], m[3]
        bx0, bx1, by0, by1 = self._bbox
        mg = self.grid_margin
        guard = 0.25 * mg
        grew = (x0 < bx0 - mg + guard) or (x1 > bx1 + mg - guard) or (y0 < by0 - mg + guard) or (y1 > by1 + mg - guard)
        shrank = (x1 - x0 + 2 * mg) * (y1 - y0 + 2 * mg) < 0.5 * (bx1 - bx0 + 2 * mg) * (by1 - by0 + 2 * mg)
        if grew or shrank:
            self.regrid(x0, x1, y0, y1)
            return True
        return False

    def rebalance(self):
        """Re-cut the strips from the current per-row particle counts (read from the device cell tables) and
        route the particles of the rows that changed hands."""
        g = self.grid
        hist = []
        for s in self.strips:
            rows = s.rows[1] - s.rows[0]
            cs = s.engine.state_view(rows=rows)[4]
            row_start = cs[::g.ncx][:rows + 1].to(torch.int64).cpu().numpy()
            h = np.zeros(g.ncy, dtype=np.float64)
            h[s.rows[0]:s.rows[1]] = np.diff(row_start)
            hist.append(h)
        self.row_counts = self.transport.all_sum(hist)
        new = strip_edges(self.row_counts, self.n_strips, self.max_rows)
        if new != self.edges:
            self.edges = new
            for s in self.strips:
                s.rows = (new[s.index], new[s.index + 1])
            self.settle()
        return self.edges

    # ---- stepping -----------------------------------------------------------------------------------
    def _record_buffers(self):
        if self._record_pinned is None:
            cap = self.strips[0].engine.max_particles
            pin = lambda dt: [torch.empty(cap, dtype=dt).pin_memory() for _ in range(2)]
            self._record_pinned = [pin(torch.int32), pin(torch.float32), pin(torch.float32), pin(torch.int8)]
        return self._record_pinned

    def step(self, check=False, timing=False, record=None):
        """One step of every local strip.  ``record`` = 0 / 1: the FIRST local strip writes this step's record -- (ids, lon,
        lat, species) of its owned particles in storage order, what a per-strip output file holds, like the reference's
        per-tile chunk pickles (particle_advecter.py:201-214) -- into slot ``record`` of two sets of pinned host buffers,
        inside the step: positions and ids are copied under the pair search, species after the RPS phases, on the
        library's copy stream (lm_record_next_step_ids).  ``record_view(slot)`` after ``host_copies_sync()``."""
        if record is not None:
            bufs = self._record_buffers()
            self.strips[0].engine.record_next_step_ids(*(b[record] for b in bufs))
            self._record_slot_count[record] = None
            self._record_last = record
        flags = 0
        st_times = None
        win = None
        if self.advect:
            flags |= _lib.LM_STEP_ADVECT
            st_times = self.clock.next_step(self.dt)
            if self.streamer is not None:
                win = self.streamer.upload(st_times)
        if timing:
            flags |= _lib.LM_STEP_TIMING
        self.rps.step = self.iteration
        if self.Kh_deg2 > 0 and self.iteration > 0:
            flags |= _lib.LM_STEP_DIFFUSE
        self.iteration += 1
        if self.interact:
            flags |= _lib.LM_STEP_INTERACT
        if self.emit_pairs:
            flags |= _lib.LM_STEP_EMIT_PAIRS
        regrid_now = self.regrid_every > 0 and self.iteration % self.regrid_every == 0
        if check or regrid_now:
            flags |= _lib.LM_STEP_STATS
        self._staged(flags, st_times)
        if record is not None:
            self._record_slot_count[record] = self.strips[0].engine.record_count()
        if win is not None:
            self.streamer.release(win)
        out = None
        if check or regrid_now:
            out = self.stats()
            if regrid_now:
                self._maybe_regrid(out)
        if self.rebalance_every > 0 and self.iteration % self.rebalance_every == 0:
            self.rebalance()
        return out

    def stats(self):
        """Per-strip counters of the last step (synchronises); raises on overflow / misrouted particles."""
        self.last_stats = [s.engine.sync_stats() for s in self.strips]
        oob = sum(st.n_out_of_bounds for st in self.last_stats)
        if oob:
            from .particle_advecter import OutOfBoundsError
            raise OutOfBoundsError("%d particle(s) left the velocity grid at iteration %d" % (oob, self.iteration))
        return self.last_stats

    def totals(self):
        """Global (pairs, particles, species counts[4]) of the last step, summed over all strips of all processes."""
        st = self.last_stats if self.last_stats is not None else self.stats()
        v = self.transport.all_sum([[s.n_pairs, s.n_particles] + list(s.species_count) for s in st])
        return int(v[0]), int(v[1]), [int(x) for x in v[2:6]]

    # ---- outputs ------------------------------------------------------------------------------------
    def local_state(self):
        """Per local strip: (ids, lon, lat, species) of the owned particles as NumPy arrays (storage order)."""
        out = []
        for s in self.strips:
            n = s.engine.state_size()
            lon, lat, sp, ids, _ = s.engine.state_view(rows=1)
            out.append((ids[:n].cpu().numpy(), lon[:n].cpu().numpy(), lat[:n].cpu().numpy(), sp[:n].cpu().numpy()))
        return out

    def local_pairs(self):
        """Per local strip: the (P, 2) global-id pairs of the last step (needs emit_pairs and a stats() sync)."""
        st = self.last_stats if self.last_stats is not None else self.stats()
        return [s.pairs[:x.n_pairs].cpu().numpy() for s, x in zip(self.strips, st)]

    def record_to_host(self, slot):
        """(Kept for callers that decide AFTER a step that they want its record; ``step(record=slot)`` copies the same
        record inside the step and is what overlaps with the pair search.)
        Asynchronous per-step record of the FIRST local strip into pinned host buffers: (ids, lon, lat,
        species) of its owned particles in storage order -- what a per-strip output file holds, like the
        reference's per-tile chunk pickles (particle_advecter.py:201-214).  Two slots alternate; returns the
        number of particles recorded.  Call ``host_copies_sync()`` before reading the buffers."""
        s = self.strips[0]
        dev = s.engine.device
        if self._record is None:
            cap = s.engine.max_particles
            mk = lambda dt: [torch.empty(cap, dtype=dt, device=dev) for _ in range(2)]
            pin = lambda dt: [torch.empty(cap, dtype=dt).pin_memory() for _ in range(2)]
            self._record = dict(stage=[mk(torch.int32), mk(torch.float32), mk(torch.float32), mk(torch.int8)],
                                host=[pin(torch.int32), pin(torch.float32), pin(torch.float32), pin(torch.int8)],
                                stream=torch.cuda.Stream(device=dev), copied=[torch.cuda.Event() for _ in range(2)])
            for e in self._record["copied"]:
                e.record()
        R = self._record
        n = s.engine.state_size()
        lon, lat, sp, ids, _ = s.engine.state_view(rows=1)
        cur = torch.cuda.current_stream()
        cur.wait_event(R["copied"][slot])                    # the staging slot was drained two records ago
        for k, src in enumerate((ids, lon, lat, sp)):
            R["stage"][k][slot][:n].copy_(src[:n], non_blocking=True)
        staged = torch.cuda.Event()
        staged.record(cur)
        with torch.cuda.stream(R["stream"]):
            R["stream"].wait_event(staged)
            for k in range(4):
                R["host"][k][slot][:n].copy_(R["stage"][k][slot][:n], non_blocking=True)
            R["copied"][slot].record(R["stream"])
        return n

    def record_view(self, slot):
        """(ids, lon, lat, species) NumPy views of record slot ``slot`` (valid until that slot is armed again)."""
        n = self._record_slot_count[slot]
        return tuple(b[slot][:n].numpy() for b in self._record_pinned)

    def host_copies_sync(self):
        if self._record is not None:
            self._record["stream"].synchronize()
        if self._record_pinned is not None:
            self.strips[0].engine.host_copies_sync()

    def gather(self):
        """Global (lon, lat, species) in particle-id order on every process (the reference's per-step record,
        interaction_simulator.py:108-110) -- for tests and small runs; big runs write per-strip files."""
        parts = self.transport.all_gather_rows(self.local_state())
        lon = np.full(self.n_total, np.nan, dtype=np.float32)
        lat = np.full(self.n_total, np.nan, dtype=np.float32)
        sp = np.zeros(self.n_total, dtype=np.int8)
        for ids, lo, la, s_ in parts:
            lon[ids], lat[ids], sp[ids] = lo, la, s_
        return lon, lat, sp

    def close(self):
        for s in self.strips:
            s.engine.close()
