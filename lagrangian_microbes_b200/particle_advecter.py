"""ParticleAdvecter: the reference's advection driver on the B200 RK4 kernel.

Mirrors /root/reference/particle_advecter.py: same constructor arguments and attributes
(:70-121), same ``time_step(start_time, end_time, dt)`` (:123) and
``create_netcdf_file(start_time, end_time, dt)`` (:262), same chunk-pickle names and contents
(:201-214, :246-249), same initial-condition helpers (:26-66).  What changes is the hot loop
(:220-244): ``pset.execute(parcels.AdvectionRK4, ...)`` becomes one ``lm_advect_rk4`` launch per
step, the two O(N) Python loops (position read-back :233-235, diffusion kick :240-242) become a
device-to-host copy and an ``lm_diffuse`` launch.

Reference behaviours kept on purpose:
  Q1  every ``time_step`` call starts the particle clock at the first snapshot of
      ``start_time.year``'s dataset (Parcels builds a fresh ParticleSet whose time is
      grid.time[0]); ``start_time`` only labels the output.
  Q2  stored row n holds the state AFTER step n+1 and is labelled ``start_time + (n+1)*dt`` in the
      pickle (:226-235) but ``start_time + n*dt`` in particle_data.nc (:266).
``N_procs`` keeps its output meaning (number of contiguous particle tiles = pickles per chunk).  In one process all
tiles are advected together on the current CUDA device; under ``torchrun`` (an initialised ``torch.distributed`` group)
the tiles are dealt out over the ranks in contiguous blocks -- the reference's joblib workers (particle_advecter.py:143-148)
become one GPU each, with no communication on the data path, exactly as there.
"""
import logging
import os
from datetime import datetime, timedelta
from glob import glob

import numpy as np
from numpy import float32, linspace, repeat, tile

from . import io as lmio
from ._lib import StageTimes
from .utils import most_symmetric_integer_factorization, pretty_time
from .velocity_fields import oscar_dataset

logger = logging.getLogger(__name__)


class OutOfBoundsError(RuntimeError):
    """A particle left the velocity grid (Parcels raises this from pset.execute)."""


class TimeExtrapolationError(RuntimeError):
    """A sample time lies outside the dataset's time axis (Parcels raises this from pset.execute)."""


def uniform_particle_locations(N_particles, lat_min, lat_max, lon_min, lon_max):
    """Regular lattice initial condition (particle_advecter.py:26-35)."""
    N_particles_lat, N_particles_lon = most_symmetric_integer_factorization(N_particles)
    logger.info("Generating ({:d}, {:d}) particles along each (lat, lon).".format(N_particles_lat, N_particles_lon))
    particle_lons = repeat(linspace(lon_min, lon_max, N_particles_lon), N_particles_lat)
    particle_lats = tile(linspace(lat_min, lat_max, N_particles_lat), N_particles_lon)
    return particle_lons, particle_lats


def distribute_particles_across_tiles(particle_lons, particle_lats, tiles):
    """Contiguous equal index slabs, one per tile (particle_advecter.py:38-66)."""
    assert particle_lons.size == particle_lats.size
    N_particles = particle_lons.size
    assert (N_particles / tiles).is_integer()
    per_tile = N_particles // tiles
    lons = [particle_lons[i * per_tile:(i + 1) * per_tile] for i in range(tiles)]
    lats = [particle_lats[i * per_tile:(i + 1) * per_tile] for i in range(tiles)]
    return lons, lats


def tiles_of_rank(N_tiles, rank, world):
    """The contiguous block of tiles a rank advects (balanced: sizes differ by at most one; a rank may get none)."""
    return list(range(N_tiles * rank // world, N_tiles * (rank + 1) // world))


def _process_group():
    """(rank, world, dist) of the initialised torch.distributed group, or (0, 1, None)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


class HostFieldSet:
    """What Parcels' Grid/Field construction produces from the dataset (particle_advecter.py:160-184):
    first depth level, time as seconds since the first snapshot, float32 lon/lat/data, latitude
    flipped to ascending, NaN (land) -> 0."""

    def __init__(self, dataset):
        nominal_depth = dataset["depth"].values[0]
        sub = dataset.sel(depth=nominal_depth)
        times = sub["time"].values
        self.t0 = times[0]                       # calendar time of the first snapshot (numpy datetime64)
        self.time = np.array([(times[i] - times[0]) // np.timedelta64(1, "s") for i in range(times.size)],
                             dtype=np.float64)
        lon = np.asarray(sub["longitude"].values, dtype=np.float32)
        lat = np.asarray(sub["latitude"].values, dtype=np.float32)
        u = np.asarray(sub["u"].values, dtype=np.float32)
        v = np.asarray(sub["v"].values, dtype=np.float32)
        if lat[-1] < lat[0]:
            lat, u, v = lat[::-1], u[:, ::-1, :], v[:, ::-1, :]
        self.lon = np.ascontiguousarray(lon)
        self.lat = np.ascontiguousarray(lat)
        self.u = np.ascontiguousarray(np.nan_to_num(u, nan=0.0, posinf=np.inf, neginf=-np.inf))
        self.v = np.ascontiguousarray(np.nan_to_num(v, nan=0.0, posinf=np.inf, neginf=-np.inf))

    def to_device(self, device):
        import torch
        return tuple(torch.from_numpy(a).to(device) for a in (self.u, self.v, self.lon, self.lat))

    @classmethod
    def from_years(cls, years):
        """The datasets of consecutive years as ONE field: snapshots concatenated along time, seconds counted from the
        first snapshot of the first year.  (The reference opens one year per ``time_step`` call and restarts the particle
        clock at that file's first snapshot -- quirk Q1; this is what ``calendar_time=True`` uses instead.)"""
        parts = [cls(oscar_dataset(int(y))) for y in years]
        out = parts[0]
        for nxt in parts[1:]:
            assert np.array_equal(nxt.lon, out.lon) and np.array_equal(nxt.lat, out.lat), "the years' grids differ"
            shift = float((nxt.t0 - out.t0) // np.timedelta64(1, "s"))
            assert shift > out.time[-1], "the years' time axes overlap"
            out.time = np.concatenate((out.time, nxt.time + shift))
            out.u = np.ascontiguousarray(np.concatenate((out.u, nxt.u), axis=0))
            out.v = np.ascontiguousarray(np.concatenate((out.v, nxt.v), axis=0))
        return out

    def seconds_since_first_snapshot(self, when):
        """Particle-clock value of the calendar time ``when`` (datetime)."""
        return float((np.datetime64(when, "us") - self.t0.astype("datetime64[us]")) / np.timedelta64(1, "s"))


class StageClock:
    """Host mirror of the particle clock and Parcels' cached time index (parcels.h::search_time_index).

    All particles share the same time, so the per-sample decisions are made once per step here and
    passed to the kernel as ``lm_stage_times``.  The cached index only advances when the sample
    time strictly exceeds the next snapshot, and a sample exactly on a snapshot interpolates with
    fraction 1.0 from the previous bracket -- reproduced as is.
    """

    def __init__(self, time_axis, t0=None):
        self.time_axis = np.asarray(time_axis, dtype=np.float64)
        self.t = float(self.time_axis[0] if t0 is None else t0)
        self.ti = 0

    def _sample(self, t):
        ax = self.time_axis
        T = ax.size
        if t < ax[0] or t > ax[T - 1]:
            raise TimeExtrapolationError("sample time %r outside the dataset's time axis [%r, %r]" % (t, ax[0], ax[T - 1]))
        ti = self.ti
        while ti < T - 1 and t > ax[ti + 1]:
            ti += 1
        while ti > 0 and t < ax[ti]:
            ti -= 1
        self.ti = ti
        if ti < T - 1 and t > ax[ti]:
            return ti, 1, float(np.float32((np.float64(t) - ax[ti]) / (ax[ti + 1] - ax[ti])))
        return ti, 0, 0.0

    def next_step(self, dt_seconds):
        """Stage decisions for one RK4 step of float32 length dt, then advance the clock."""
        dt32 = float(np.float32(dt_seconds))
        st = StageTimes()
        for k, ts in enumerate((self.t, self.t + .5 * dt32, self.t + .5 * dt32, self.t + dt32)):
            st.ti[k], st.interp[k], st.frac[k] = self._sample(ts)
        self.t = self.t + dt32
        return st


class ParticleAdvecter:
    def __init__(
        self,
        particle_lons,
        particle_lats,
        N_procs=-1,
        velocity_field="OSCAR",
        output_dir=".",
        output_chunk_iters=100,
        Kh=0,
        seed=0,
        calendar_time=False,
    ):
        assert velocity_field == "OSCAR", "OSCAR is the only supported velocity field right now."
        assert 1 <= N_procs or N_procs == -1, "Number of processors N_procs must be a positive integer " \
                                              "or -1 (use all processors)."
        # The reference clamps to joblib.cpu_count() worker processes (particle_advecter.py:86-94); here a worker is a
        # GPU: N_procs fixes the number of tiles, dealt out over the ranks of the process group (one process: all
        # tiles on one GPU).  -1 ("all processors") -> one tile per rank.
        self.rank, self.world, self._dist = _process_group()
        N_procs = N_procs if N_procs >= 1 else self.world

        particle_lons = np.asarray(particle_lons)
        particle_lats = np.asarray(particle_lats)
        assert particle_lons.size == particle_lats.size
        N_particles = particle_lons.size
        assert (N_particles / N_procs).is_integer()

        output_dir = os.path.abspath(output_dir)
        if not os.path.exists(output_dir):
            logger.info("Creating directory: {:s}".format(output_dir))
            os.makedirs(output_dir)
        assert output_chunk_iters >= 1

        self.iteration = 0
        self.particle_lons, self.particle_lats = distribute_particles_across_tiles(particle_lons, particle_lats, N_procs)
        self.velocity_field = velocity_field
        self.N_particles = N_particles
        self.N_procs = N_procs
        self.particles_per_tile = N_particles // N_procs
        self.my_tiles = tiles_of_rank(N_procs, self.rank, self.world)
        self.output_dir = output_dir
        self.output_chunk_iters = output_chunk_iters
        self.Kh = Kh / 1e10  # [m^2/s] -> [deg^2/s] assuming 1 deg = 100 km (particle_advecter.py:121)
        self.seed = seed
        # False (default): the reference's behaviour, quirk Q1 -- every time_step call samples the velocity of
        # start_time.year's file from its FIRST snapshot on, whatever start_time is (particle_advecter.py:160,186-187).
        # True: the particle clock is the calendar -- a call starts at start_time inside its year's file, and a call that
        # runs past the end of a year continues in the next year's file ("year roll-over, fixing Q1").
        self.calendar_time = bool(calendar_time)
        self._engine = None
        self._field_year = None

    # -- device plumbing ---------------------------------------------------------------------------
    def _ensure_engine(self, year):
        import torch
        from .engine import Engine
        if self._engine is None:
            self._engine = Engine(max_particles=max(1, len(self.my_tiles) * self.particles_per_tile), max_cells=1 << 16, max_pairs=0)
        if self._field_year != year:
            fs = HostFieldSet.from_years(year) if isinstance(year, tuple) else HostFieldSet(oscar_dataset(year))
            self._fieldset = fs
            self._engine.set_field(*fs.to_device(self._engine.device))
            self._field_year = year
        return self._engine

    def time_step(self, start_time, end_time, dt):
        import torch
        if self.iteration != 0:
            logger.info("Restoring particle locations from disk...")
            pkl_files = sorted(glob(os.path.join(self.output_dir, "particle_locations_*.pickle")))
            import joblib
            for pkl_filepath in pkl_files[-self.N_procs:]:
                _, _, tile_id = lmio.parse_chunk_name(pkl_filepath)
                if tile_id not in self.my_tiles:
                    continue
                chunk = joblib.load(pkl_filepath)
                self.particle_lons[tile_id] = chunk["lon"][-1, :]
                self.particle_lats[tile_id] = chunk["lat"][-1, :]

        logger.info("Starting time stepping: {:} -> {:} (dt={:}) on {:d} tile(s)."
                    .format(start_time, end_time, dt, self.N_procs))
        if self.calendar_time:
            # the last sample of the last step is taken AT end_time: the next year's file is needed if that lies beyond
            # this year's last snapshot (the files hold 72 five-day snapshots: the last ~10 days of a year have none)
            years = list(range(start_time.year, end_time.year + 1))
            last = HostFieldSet(oscar_dataset(years[-1]))
            if last.seconds_since_first_snapshot(end_time) > last.time[-1]:
                years.append(years[-1] + 1)
            eng = self._ensure_engine(tuple(years))
        else:
            eng = self._ensure_engine(start_time.year)
        dev = eng.device
        per_tile = self.particles_per_tile
        mine = self.my_tiles
        N = len(mine) * per_tile                                 # this rank's particles: its tiles, in tile order
        dt_s = dt.total_seconds()

        # Parcels casts particle lon/lat to float32 (JITParticle); one fresh particle set per call (Q1)
        cat = lambda tiles: np.concatenate([tiles[t] for t in mine]) if mine else np.zeros(0)
        lon = torch.from_numpy(cat(self.particle_lons).astype(np.float32)).to(dev)
        lat = torch.from_numpy(cat(self.particle_lats).astype(np.float32)).to(dev)
        # global particle index of each local one: the diffusion kicks are keyed by it, whoever holds the particle
        ids = None
        if self.world > 1 and N:
            ids = torch.from_numpy(np.concatenate([np.arange(t * per_tile, (t + 1) * per_tile, dtype=np.int32) for t in mine])).to(dev)
        clock = StageClock(self._fieldset.time,
                           t0=self._fieldset.seconds_since_first_snapshot(start_time) if self.calendar_time else None)
        amp = float(np.sqrt(6 * np.fabs(np.float32(dt_s)) * self.Kh))

        t = start_time
        iteration = self.iteration
        while t < end_time:
            iters_remaining = (end_time - t) // dt
            iters_to_do = min(self.output_chunk_iters, iters_remaining)
            start_iter, end_iter = iteration, iteration + iters_to_do
            logger.info("Advecting particles (iteration {:05d} -> {:05d}): {:} -> {:}..."
                        .format(start_iter, end_iter, t, t + iters_to_do * dt))

            out_lon = torch.empty((iters_to_do, N), dtype=torch.float32).pin_memory()
            out_lat = torch.empty((iters_to_do, N), dtype=torch.float32).pin_memory()
            times = iters_to_do * [None]
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            eng.reset_stats()
            for n in range(iters_to_do):
                st = clock.next_step(dt_s)
                if N:
                    eng.advect_rk4(lon, lat, st, dt_s)                            # :222-223
                t = t + dt
                iteration += 1
                times[n] = t
                out_lon[n].copy_(lon, non_blocking=True)                          # :233-235
                out_lat[n].copy_(lat, non_blocking=True)
                if self.Kh > 0 and N:
                    eng.diffuse(lon, lat, amp, self.seed, iteration - 1, ids=ids)  # :240-242
            ev1.record()
            n_oob = eng.sync_stats().n_out_of_bounds                              # synchronises the stream
            n_oob = self._sum_over_ranks(n_oob)                                   # every rank raises, or none does
            if n_oob:
                raise OutOfBoundsError("%d particle-step(s) left the velocity grid in iterations %d..%d"
                                       % (n_oob, start_iter, end_iter))
            logger.info("Advecting + storing particles: {:s}.".format(pretty_time(ev0.elapsed_time(ev1) * 1e-3)))

            lon_np, lat_np = out_lon.numpy(), out_lat.numpy()
            for k, tile_id in enumerate(mine):
                sl = slice(k * per_tile, (k + 1) * per_tile)
                path = os.path.join(self.output_dir, lmio.chunk_pickle_name(start_iter, end_iter, tile_id))
                logger.info("Dumping intermediate output: {:s}".format(path))
                lmio.dump_chunk(path, times, np.ascontiguousarray(lat_np[:, sl]), np.ascontiguousarray(lon_np[:, sl]))

        # keep the final positions for callers that chain time_step without going through disk
        # (tiles of other ranks keep their previous values here; the next call restores from the pickles, as the reference does)
        final_lon, final_lat = lon.cpu().numpy(), lat.cpu().numpy()
        for k, tile_id in enumerate(mine):
            self.particle_lons[tile_id] = final_lon[k * per_tile:(k + 1) * per_tile]
            self.particle_lats[tile_id] = final_lat[k * per_tile:(k + 1) * per_tile]
        iters = (end_time - start_time) // dt
        self.iteration += iters
        self._barrier()                                          # every tile's pickles are on disk when any rank returns

    def _barrier(self):
        if self._dist is not None and self.world > 1:
            self._dist.barrier()

    def _sum_over_ranks(self, value):
        if self._dist is None or self.world == 1:
            return int(value)
        import torch
        on_gpu = self._dist.get_backend() == "nccl"
        v = torch.tensor([int(value)], dtype=torch.int64, device=self._engine.device if on_gpu else "cpu")
        self._dist.all_reduce(v)
        return int(v.item())

    def create_netcdf_file(self, start_time, end_time, dt):
        import joblib
        self._barrier()
        if self.rank != 0:                                       # one writer: the merge is rank 0's, as it is the parent's
            self._barrier()                                      # in the reference (particle_advecter.py:262-307)
            return
        iters = (end_time - start_time) // dt
        times = [start_time + n * dt for n in range(iters)]
        # (the reference merges into two dense (N, iters) host arrays, :269-270: 30 GB for its 490,000 x 7,670 run; the
        #  writer keeps them in memory while the file fits NetCDF-3 and goes through memory-mapped files beyond)
        nc_filepath = os.path.join(self.output_dir, "particle_data.nc")
        out = lmio.ParticleFileWriter(nc_filepath, {"longitude": float32, "latitude": float32}, self.N_particles, times)
        pkl_files = sorted(glob(os.path.join(self.output_dir, "particle_locations_*.pickle")))
        for pkl_filepath in pkl_files:
            logger.info("Collecting particle locations from {:s}...".format(pkl_filepath))
            t1, t2, tile_id = lmio.parse_chunk_name(pkl_filepath)
            chunk = joblib.load(pkl_filepath)
            i1, i2 = tile_id * self.particles_per_tile, (tile_id + 1) * self.particles_per_tile
            out.put_block(slice(i1, i2), t1, t2, longitude=np.transpose(chunk["lon"]), latitude=np.transpose(chunk["lat"]))
        logger.info("Writing particle locations to {:s}...".format(nc_filepath))
        out.close()
        for pkl_filepath in pkl_files:
            os.remove(pkl_filepath)
        self._barrier()
