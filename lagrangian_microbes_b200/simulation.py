"""FusedSimulation: advect -> (diffuse) -> bin -> pair search + RPS in one device-resident loop.

The reference couples its two phases through disk: ``ParticleAdvecter`` writes every position to
``particle_data.nc`` and ``InteractionSimulator`` reads it back step by step
(/root/reference/rock_paper_scissors_example.py:25-36).  Species never feed back into advection,
so running both per step on resident state is semantically identical and removes
the disk round trip.  This class is that fused driver loop; it produces the same per-step record
the two reference classes produce between them (positions after step n, species after step n's
interactions: particle_advecter.py:233-235, interaction_simulator.py:108-110).

One ``step()`` is one ``lm_step`` call (include/lm_b200.h): ~17 kernel launches, no host
synchronisation.  Every ``regrid_every`` steps the particles' bounding box is read back (one
small sync) and the binning grid is re-fitted if particles approach its edge.
"""
import numpy as np
import torch

from . import _lib
from ._lib import RpsParams
from .engine import Engine, make_grid
from .particle_advecter import OutOfBoundsError, StageClock


class FieldWindowStreamer:
    """Velocity snapshots stay in pinned host memory; each step uploads the (2 or 3) time levels its four RK4
    stages bracket into one of two device windows on a side stream, overlapped with the previous step's
    kernels (velocity input path; the reference re-opens the year's NetCDF file per
    ``time_step`` call instead: particle_advecter.py:160-183)."""

    def __init__(self, fieldset, engines):
        self.engines = list(engines)
        dev = self.engines[0].device
        self._u_host = torch.from_numpy(fieldset.u).pin_memory()
        self._v_host = torch.from_numpy(fieldset.v).pin_memory()
        _, Y, X = fieldset.u.shape
        self._win_u = [torch.zeros((3, Y, X), dtype=torch.float32, device=dev) for _ in range(2)]
        self._win_v = [torch.zeros((3, Y, X), dtype=torch.float32, device=dev) for _ in range(2)]
        self._grid_lon = torch.from_numpy(fieldset.lon).to(dev)
        self._grid_lat = torch.from_numpy(fieldset.lat).to(dev)
        for e in self.engines:
            e.set_field(self._win_u[0], self._win_v[0], self._grid_lon, self._grid_lat)
        self._h2d_stream = torch.cuda.Stream(device=dev)
        self._win_ready = [torch.cuda.Event() for _ in range(2)]
        self._win_free = [torch.cuda.Event() for _ in range(2)]
        for e in self._win_free:
            e.record()
        self._win_idx = 0
        self.h2d_bytes_last_step = 0

    def upload(self, st_times):
        """Upload the time levels this step samples, point the engines at them and rebase ``st_times.ti``.
        Returns the window index to hand to ``release`` once the step's kernels are queued."""
        lo = min(st_times.ti[k] for k in range(4))
        hi = max(st_times.ti[k] + (1 if st_times.interp[k] else 0) for k in range(4))
        cnt = hi - lo + 1
        assert cnt <= 3, "one RK4 step spans more than three velocity snapshots"
        k = self._win_idx
        self._win_idx ^= 1
        with torch.cuda.stream(self._h2d_stream):
            self._h2d_stream.wait_event(self._win_free[k])        # the step that last read window k is done
            self._win_u[k][:cnt].copy_(self._u_host[lo:hi + 1], non_blocking=True)
            self._win_v[k][:cnt].copy_(self._v_host[lo:hi + 1], non_blocking=True)
            self._win_ready[k].record(self._h2d_stream)
        torch.cuda.current_stream().wait_event(self._win_ready[k])
        for e in self.engines:
            e.update_field_data(self._win_u[k], self._win_v[k])
        for j in range(4):
            st_times.ti[j] -= lo
        self.h2d_bytes_last_step = 2 * cnt * self._win_u[k][0].numel() * 4
        return k

    def release(self, k):
        self._win_free[k].record()


class FusedSimulation:
    def __init__(self, lons, lats, species, radius, pRS, pPR, pSP, fieldset, dt_seconds=3600.0, Kh=0.0, seed=0,
                 emit_pairs=True, pair_capacity=None, regrid_every=16, grid_margin=0.5, cells_per_particle=2.0,
                 max_cells=None, device=None, interact=True, advect=True, stream_field=False, interaction_norm=2):
        lons = np.ascontiguousarray(lons, dtype=np.float32)      # Parcels keeps float32 positions
        lats = np.ascontiguousarray(lats, dtype=np.float32)
        species = np.ascontiguousarray(species, dtype=np.int8)
        n = lons.size
        assert lats.size == n and species.size == n
        self.n = n
        self.radius = float(radius)
        self.rps = RpsParams(float(pRS), float(pPR), float(pSP), int(seed), 0)
        self.dt = float(dt_seconds)
        self.Kh_deg2 = float(Kh) / 1e10                          # particle_advecter.py:121
        self.diffuse_amp = float(np.sqrt(6 * np.fabs(np.float32(self.dt)) * self.Kh_deg2))
        self.seed = int(seed)
        self.iteration = 0
        self.regrid_every = int(regrid_every)
        self.grid_margin = float(grid_margin)
        self.cells_per_particle = float(cells_per_particle)
        self.interact = bool(interact)
        self.advect = bool(advect)

        if max_cells is None:
            max_cells = int(max(4 * cells_per_particle * n, 1 << 20))
        if pair_capacity is None:
            pair_capacity = max(1 << 20, 24 * n)                 # BASELINE config 3 reaches 19 pairs per microbe
        # max_pairs sizes the pair-search -> resolver hand-off buffer of the round-1 pipeline (4 B per pair per step)
        self.engine = Engine(max_particles=n, max_cells=max_cells, max_pairs=int(pair_capacity) if interact else 0,
                             device=device)
        self.engine.set_norm(interaction_norm)                   # query_pairs(r, p=interaction_norm)
        dev = self.engine.device
        self.fieldset = fieldset
        self.stream_field = bool(stream_field) and fieldset is not None
        self.h2d_bytes_last_step = 0
        self.streamer = None
        if fieldset is not None and self.stream_field:
            self.streamer = FieldWindowStreamer(fieldset, [self.engine])
            self.clock = StageClock(fieldset.time)
        elif fieldset is not None:
            self.engine.set_field(*fieldset.to_device(dev))
            self.clock = StageClock(fieldset.time)
        else:
            assert not advect, "advection needs a fieldset"
            self.clock = None

        self.emit_pairs = bool(emit_pairs) and self.interact
        if self.emit_pairs:
            self.pairs = torch.empty((int(pair_capacity), 2), dtype=torch.int32, device=dev)
        else:
            self.pairs = None

        self._fit_grid(float(lons.min()), float(lons.max()), float(lats.min()), float(lats.max()))
        self.engine.state_set(torch.from_numpy(lons).to(dev), torch.from_numpy(lats).to(dev),
                              torch.from_numpy(species).to(dev))
        self.total_pairs = 0
        self.last_stats = None

    # ---- grid policy -------------------------------------------------------------------------------
    def _fit_grid(self, x0, x1, y0, y1):
        self._bbox = (x0, x1, y0, y1)
        g = make_grid(x0, x1, y0, y1, self.radius, self.n, self.engine.max_cells, margin=self.grid_margin,
                      cells_per_particle=self.cells_per_particle)
        self.engine.set_grid(g)
        self.grid = g

    def _maybe_regrid(self, st):
        x0, x1, y0, y1 = st.bbox[0], st.bbox[1], st.bbox[2], st.bbox[3]
        bx0, bx1, by0, by1 = self._bbox
        m = self.grid_margin
        guard = 0.25 * m
        grew = (x0 < bx0 - m + guard) or (x1 > bx1 + m - guard) or (y0 < by0 - m + guard) or (y1 > by1 + m - guard)
        shrank = (x1 - x0 + 2 * m) * (y1 - y0 + 2 * m) < 0.5 * (bx1 - bx0 + 2 * m) * (by1 - by0 + 2 * m)
        if grew or shrank:
            self._fit_grid(x0, x1, y0, y1)

    # ---- stepping ----------------------------------------------------------------------------------
    def step(self, check=False, timing=False, record=None):
        """One fused step.  ``record`` = (lon, lat, species) pinned CPU tensors: the step's record is written there
        asynchronously, overlapped with the step (engine.host_copies_sync() before reading)."""
        if record is not None:
            self.engine.record_next_step(*record)
        flags = 0
        st_times = None
        win = None
        if self.advect:
            flags |= _lib.LM_STEP_ADVECT
            st_times = self.clock.next_step(self.dt)
            if self.streamer is not None:
                win = self.streamer.upload(st_times)
                self.h2d_bytes_last_step = self.streamer.h2d_bytes_last_step
        if timing:
            flags |= _lib.LM_STEP_TIMING
        self.rps.step = self.iteration          # 0-based step index = InteractionSimulator's ``i``
        if self.Kh_deg2 > 0 and self.iteration > 0:
            flags |= _lib.LM_STEP_DIFFUSE       # kick of the previous iteration (particle_advecter.py:240-242)
        self.iteration += 1
        if self.interact:
            flags |= _lib.LM_STEP_INTERACT
        if self.emit_pairs:
            flags |= _lib.LM_STEP_EMIT_PAIRS
        # the grid fixes the canonical pair order, so it is re-fitted on a fixed schedule: ``check`` must not change results
        regrid_now = self.regrid_every > 0 and self.iteration % self.regrid_every == 0
        want_stats = check or regrid_now
        if want_stats:
            flags |= _lib.LM_STEP_STATS
        self.engine.step(flags, st_times, self.dt, self.diffuse_amp, self.radius, self.rps, self.pairs)
        if win is not None:
            self.streamer.release(win)
        if want_stats:
            st = self.engine.sync_stats()
            self.last_stats = st
            if st.n_out_of_bounds:
                raise OutOfBoundsError("%d particle(s) left the velocity grid at iteration %d"
                                       % (st.n_out_of_bounds, self.iteration))
            if regrid_now:
                self._maybe_regrid(st)
            return st
        return None

    def check_faults(self):
        """Raise if any step since the last check overflowed a capacity (pair list truncated, hand-off or exchange buffer
        too small) -- the device latches such faults across steps (lm_sync_stats), so steps run without ``check=True``
        cannot lose them.  Synchronises."""
        self.engine.sync_stats()

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
        self.check_faults()

    def run_to_file(self, output_dir, start_time, end_time, dt, stride=1, filename="microbe_data.nc", packed=False):
        """The reference's end product in one pass: ``(end_time - start_time) // dt`` fused steps, their per-step
        record written as ``microbe_data.nc`` -- what rock_paper_scissors_example.py:25-36 produces through
        ParticleAdvecter.time_step + create_netcdf_file + InteractionSimulator.time_step, without the round trip
        through ``particle_data.nc``.  Column k holds the positions after step k's advection and the species after
        step k's interactions (particle_advecter.py:233-235 with its quirk Q2, interaction_simulator.py:108-110).
        ``stride`` keeps every stride-th step.  Records travel in pairs of pinned buffers under the steps that follow
        them (lm_record_next_step).  ``packed=True`` sends the positions as lossless int16 ulp differences to the
        previous kept step instead (record.DeltaRecordPacker: about half the bytes over PCIe, the same file bit for bit).
        Returns (path, per-step species counts of the kept steps as an (nt, 3) array)."""
        from . import io as lmio
        from datetime import timedelta
        assert isinstance(dt, timedelta) and abs(dt.total_seconds() - self.dt) < 1e-9, "dt differs from the simulation's"
        n_steps = (end_time - start_time) // dt
        asm = lmio.RecordAssembler(self.n, n_steps, start_time, dt, stride, output_dir=output_dir, filename=filename)
        if packed:
            self._run_packed(asm, n_steps)
            self.check_faults()
            path = asm.write(output_dir, filename)
            return path, asm.counts
        rec = [tuple(torch.empty(self.n, dtype=t).pin_memory() for t in (torch.float32, torch.float32, torch.int8))
               for _ in range(2)]
        in_flight = []                                   # (step, buffer set) whose copies have been issued

        def drain():
            self.engine.host_copies_sync()
            for step, k in in_flight:
                asm.put(step, rec[k][0].numpy(), rec[k][1].numpy(), rec[k][2].numpy())
            in_flight.clear()

        for step in range(n_steps):
            if asm.wants(step):
                if len(in_flight) == 2:
                    drain()
                k = len(in_flight)
                self.step(record=rec[k])
                in_flight.append((step, k))
            else:
                self.step()
        drain()
        self.check_faults()                              # before anything is written: a truncated step must not reach the file
        path = asm.write(output_dir, filename)
        return path, asm.counts

    def _run_packed(self, asm, n_steps):
        """run_to_file's loop with the delta-packed position record: state in id order on the device (lm_state_get),
        lm_record_delta_pack, int16 deltas + escapes to pinned memory; species as plain int8.  One record in flight."""
        from .record import DeltaRecordPacker
        dev = self.engine.device
        packer = DeltaRecordPacker(self.n, device=dev)
        lon_d = torch.empty(self.n, dtype=torch.float32, device=dev)
        lat_d = torch.empty(self.n, dtype=torch.float32, device=dev)
        sp_pin = [torch.empty(self.n, dtype=torch.int8).pin_memory() for _ in range(2)]
        pending = []                                     # (step, species buffer) in push order

        def drain_one():
            step, k = pending.pop(0)
            lon, lat = packer.pop()
            self.engine.host_copies_sync()
            asm.put(step, lon, lat, sp_pin[k].numpy())

        kept = 0
        for step in range(n_steps):
            self.step()
            if asm.wants(step):
                if len(pending) == 2:
                    drain_one()
                k = kept % 2
                self.engine.state_get(lon_d, lat_d, None)
                packer.push(lon_d, lat_d)
                self.engine.state_get_host(None, None, sp_pin[k])
                pending.append((step, k))
                kept += 1
        while pending:
            drain_one()
        self.record_bytes_d2h = packer.bytes_d2h + asm.filled * self.n

    def stats(self):
        """Counters of the most recent step (synchronises)."""
        return self.engine.sync_stats()

    def download(self):
        """(lon, lat, species) in particle-id order as NumPy arrays."""
        dev = self.engine.device
        lon = torch.empty(self.n, dtype=torch.float32, device=dev)
        lat = torch.empty(self.n, dtype=torch.float32, device=dev)
        sp = torch.empty(self.n, dtype=torch.int8, device=dev)
        self.engine.state_get(lon, lat, sp)
        return lon.cpu().numpy(), lat.cpu().numpy(), sp.cpu().numpy()

    def record_to_host(self, lon_pin, lat_pin, sp_pin):
        """Asynchronous per-step record into pinned host tensors (call engine.host_copies_sync() before reading)."""
        self.engine.state_get_host(lon_pin, lat_pin, sp_pin)
