"""Velocity input path, SURVEY.md 8(f) rank 2: land (NaN in the product -> 0, as Parcels does), the out-of-bounds policy
(Parcels raises OutOfBoundsError since the reference passes no recovery kernel: /root/reference/particle_advecter.py:222-223;
here the particle stays, is counted, and the host raises after the chunk), and the year roll-over behind
``ParticleAdvecter(calendar_time=True)`` (default: the reference's quirk Q1, particle_advecter.py:160,186-187)."""
import glob
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import rk4 as ork4

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _dataset_with_land(year, land_box=(30.0, 32.0, 206.0, 209.0)):
    """OSCAR-grid synthetic year with a NaN (land) block, as the product stores land."""
    from lagrangian_microbes_b200 import velocity_fields as vf
    vf.configure_synthetic(kind="random_fourier", seed=5, n_modes=8, rms_speed=0.3)
    vf.register_dataset_provider(None)                                   # the plain synthetic year underneath
    ds = vf.oscar_dataset(year)
    lat, lon = ds["latitude"].values, ds["longitude"].values
    u, v = ds["u"].values.copy(), ds["v"].values.copy()
    m = ((lat >= land_box[0]) & (lat <= land_box[1]))[:, None] & ((lon >= land_box[2]) & (lon <= land_box[3]))[None, :]
    u[:, :, m] = np.nan
    v[:, :, m] = np.nan
    return vf.SyntheticDataset({"time": ds["time"].values, "depth": ds["depth"].values, "latitude": lat, "longitude": lon,
                                "u": u, "v": v})


def test_land_is_still_water_and_leaving_the_grid_raises(tmp_path):
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import velocity_fields as vf
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet, OutOfBoundsError
    cache = {2017: _dataset_with_land(2017)}
    vf.register_dataset_provider(lambda year: cache[year])
    try:
        fs = HostFieldSet(vf.oscar_dataset(2017))
        assert (fs.u == 0).any() and not np.isnan(fs.u).any()
        ofs = ork4.FieldSet(fs.lon, fs.lat, fs.time, fs.u, fs.v)
        rng = np.random.default_rng(0)
        n = 4000
        lons = 204.0 + 7.0 * rng.random(n)
        lats = 28.0 + 6.0 * rng.random(n)
        inland = (lons > 206.5) & (lons < 208.5) & (lats > 30.5) & (lats < 31.5)        # deep inside the land block
        assert inland.sum() > 50
        adv = lm.ParticleAdvecter(lons, lats, N_procs=1, output_dir=str(tmp_path), output_chunk_iters=6)
        t0 = datetime(2017, 1, 1)
        adv.time_step(t0, t0 + timedelta(hours=12), timedelta(hours=1))
        lon_end, lat_end = np.concatenate(adv.particle_lons), np.concatenate(adv.particle_lats)
        l32, a32 = lons.astype(np.float32), lats.astype(np.float32)
        assert np.array_equal(lon_end[inland], l32[inland]) and np.array_equal(lat_end[inland], a32[inland])   # zero velocity on land
        t, ti = 0.0, 0
        for _ in range(12):
            l32, a32, ti, oob = ork4.rk4_step_f32(ofs, l32, a32, t, 3600.0, ti)
            assert oob == 0
            t += 3600.0
        assert np.array_equal(lon_end, l32) and np.array_equal(lat_end, a32)             # incl. the coast: partial land cells
        # a particle beyond the grid's northern edge (80N): Parcels raises OutOfBoundsError, so does the drop-in
        bad = lm.ParticleAdvecter(np.array([210.0, 211.0]), np.array([30.0, 80.5]), N_procs=1, output_dir=str(tmp_path / "bad"))
        with pytest.raises(OutOfBoundsError):
            bad.time_step(t0, t0 + timedelta(hours=2), timedelta(hours=1))
    finally:
        vf.register_dataset_provider(None)
        vf.configure_synthetic(kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2)


def test_calendar_time_runs_across_the_year_boundary_and_the_default_keeps_quirk_q1(tmp_path):
    import joblib
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import velocity_fields as vf
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet
    vf.configure_synthetic(kind="random_fourier", seed=3, n_modes=6, rms_speed=0.3)
    try:
        rng = np.random.default_rng(1)
        n = 3000
        lons, lats = 205.0 + 5.0 * rng.random(n), 28.0 + 5.0 * rng.random(n)
        start, dt = datetime(2017, 12, 28), timedelta(hours=1)
        end = start + 8 * 24 * dt                                                      # into 5 January 2018
        fs = HostFieldSet.from_years((2017, 2018))
        ofs = ork4.FieldSet(fs.lon, fs.lat, fs.time, fs.u, fs.v)
        # the calendar clock: 28 December of the concatenated field, over the gap between the two files
        cal = lm.ParticleAdvecter(lons, lats, N_procs=1, output_dir=str(tmp_path / "cal"), output_chunk_iters=500, calendar_time=True)
        cal.time_step(start, end, dt)
        l32, a32 = lons.astype(np.float32), lats.astype(np.float32)
        t, ti = fs.seconds_since_first_snapshot(start), 0
        assert t == 361 * 86400.0
        for _ in range(8 * 24):
            l32, a32, ti, oob = ork4.rk4_step_f32(ofs, l32, a32, t, 3600.0, ti)
            t += 3600.0
        assert ti == 72                                                                # the run ended inside the 2018 file
        assert np.array_equal(np.concatenate(cal.particle_lons), l32) and np.array_equal(np.concatenate(cal.particle_lats), a32)
        chunk = joblib.load(sorted(glob.glob(os.path.join(str(tmp_path / "cal"), "particle_locations_*.pickle")))[0])
        assert chunk["lon"].shape == (192, n) and np.array_equal(chunk["lon"][-1], l32)
        # default = the reference: the same call samples the 2017 file from its FIRST snapshot (January's velocities)
        ref = lm.ParticleAdvecter(lons, lats, N_procs=1, output_dir=str(tmp_path / "q1"), output_chunk_iters=500)
        ref.time_step(start, end, dt)
        fs17 = HostFieldSet(vf.oscar_dataset(2017))
        o17 = ork4.FieldSet(fs17.lon, fs17.lat, fs17.time, fs17.u, fs17.v)
        l32, a32 = lons.astype(np.float32), lats.astype(np.float32)
        t, ti = 0.0, 0
        for _ in range(8 * 24):
            l32, a32, ti, _ = ork4.rk4_step_f32(o17, l32, a32, t, 3600.0, ti)
            t += 3600.0
        assert np.array_equal(np.concatenate(ref.particle_lons), l32)
        assert not np.array_equal(np.concatenate(ref.particle_lons), np.concatenate(cal.particle_lons))
    finally:
        vf.configure_synthetic(kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2)
