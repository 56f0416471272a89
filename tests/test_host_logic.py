"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, and the host
mirror of the reference API (helpers, clocks, grids, file formats) behaves like the reference's."""
import ctypes
import os
import re
from datetime import datetime, timedelta

import numpy as np
import pytest

import lagrangian_microbes_b200 as lm
from lagrangian_microbes_b200 import _lib, io as lmio, utils, velocity_fields
from lagrangian_microbes_b200.particle_advecter import HostFieldSet, StageClock
from oracle import rk4 as ork4

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    so = _lib.build()
    assert os.path.exists(so)
    header = open(os.path.join(ROOT, "include", "lm_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*(lm_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(so)
    for name in sorted(declared):
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
    assert declared == set(_lib.EXPORTS)
    L2 = _lib.lib()                      # argtypes declared for all of them
    assert L2.lm_version() == 100
    assert L2.lm_error_string(_lib.LM_ENOSPC) == b"capacity exceeded"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from lagrangian_microbes_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(max_particles=10)
    with pytest.raises(NotImplementedError):
        lm.rock_paper_scissors_interaction({}, {}, 0, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lagrangian_microbes_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
            assert "scipy.spatial" not in src, fn
            assert "cuda_emu" not in src and "libemu" not in src, fn       # the CPU emulator of the kernels is test-only
    for bench_like in ("bench.py", "__graft_entry__.py"):
        assert "cuda_emu" not in open(os.path.join(ROOT, bench_like)).read(), bench_like


def test_utils_and_initial_condition():
    assert utils.most_symmetric_integer_factorization(490000) == (700, 700)
    assert utils.most_symmetric_integer_factorization(10000) == (100, 100)
    assert utils.most_symmetric_integer_factorization(12) == (3, 4)
    assert utils.factor(28) == [1, 2, 4, 7, 14, 28]
    assert utils.pretty_time(0.5) == "500 ms" and utils.pretty_time(120) == "2 mins"
    assert utils.pretty_filesize(2048) == "2.0 KiB"
    assert utils.closest_hour(np.datetime64("2017-01-01T10:29:59")) == datetime(2017, 1, 1, 10)
    assert utils.closest_hour(np.datetime64("2017-01-01T10:30:00")) == datetime(2017, 1, 1, 11)
    lons, lats = lm.uniform_particle_locations(N_particles=12, lat_min=25, lat_max=35, lon_min=205, lon_max=215)
    # (N_lat, N_lon) = (3, 4): lons = repeat(linspace(.., 4), 3), lats = tile(linspace(.., 3), 4)
    assert np.allclose(lons, np.repeat(np.linspace(205, 215, 4), 3)) and np.allclose(lats, np.tile(np.linspace(25, 35, 3), 4))
    tl, ta = lm.distribute_particles_across_tiles(lons, lats, 4)
    assert len(tl) == 4 and all(t.size == 3 for t in tl) and np.array_equal(np.concatenate(tl), lons)
    with pytest.raises(AssertionError):
        lm.distribute_particles_across_tiles(lons, lats, 5)


def test_interactions_factory_contract():
    np.random.seed(0)
    fn, params, props = lm.rock_paper_scissors(N_microbes=1000, pRS=0.5, pPR=0.6, pSP=0.7)
    assert props["species"].dtype == np.int8 and props["species"].shape == (1000,)
    assert set(np.unique(props["species"])) == {1, 2, 3}
    assert params == {"pRS": 0.5, "pPR": 0.6, "pSP": 0.7}
    assert fn is lm.rock_paper_scissors_interaction
    with pytest.raises(NotImplementedError):
        lm.InteractionSimulator(pair_interaction=(lambda *a: None, {}, {}), interaction_radius=0.01, output_dir="/tmp/lm_x")
    # interaction_norm is SciPy's p: 1, 2 and inf exist on the device, anything else needs pow()
    for p_bad in (3, 1.5, 0.5, "2"):
        with pytest.raises(NotImplementedError):
            lm.InteractionSimulator(pair_interaction=(fn, params, props), interaction_radius=0.01, interaction_norm=p_bad,
                                    output_dir="/tmp/lm_x")
    from lagrangian_microbes_b200 import _lib
    from lagrangian_microbes_b200.engine import norm_code
    assert [norm_code(p) for p in (1, 2, 2.0, np.inf, float("inf"))] == \
        [_lib.LM_NORM_1, _lib.LM_NORM_2, _lib.LM_NORM_2, _lib.LM_NORM_INF, _lib.LM_NORM_INF]
    for p_ok in (1, 2, np.inf):
        sim = lm.InteractionSimulator(pair_interaction=(fn, params, props), interaction_radius=0.01, interaction_norm=p_ok,
                                      output_dir="/tmp/lm_x")
        assert sim.interaction_norm == p_ok


def test_stage_clock_matches_oracle_time_search():
    ax = np.array([0.0, 400000.0, 864000.0, 1512000.0])
    clock = StageClock(ax)
    fs = type("F", (), {"time": ax})
    t, ti = 0.0, 0
    for _ in range(400):
        want, ti = ork4.stage_times(fs, t, 3600.0, ti)
        got = clock.next_step(3600.0)
        assert [(got.ti[k], bool(got.interp[k]), np.float32(got.frac[k])) for k in range(4)] == \
               [(w[0], bool(w[1]), np.float32(w[2])) for w in want]
        t += 3600.0
    with pytest.raises(lm.TimeExtrapolationError):
        for _ in range(100):
            clock.next_step(3600.0)


def test_make_grid_policy():
    from lagrangian_microbes_b200.engine import make_grid
    g = make_grid(205.0, 215.0, 25.0, 35.0, 0.01, 490000, 1 << 24, margin=0.5)
    h = 1.0 / g.inv_h
    assert h > 0.01 and g.x0 == 204.0 and g.y0 == 24.0
    assert g.ncx * g.ncy <= 2 * 490000 and g.ncx * h >= 216.0 - g.x0 - 0.5
    k = round(h / 0.01)
    assert abs(h - k * 0.01 * (1 + 2.0 ** -20)) < 1e-12
    g2 = make_grid(180.0, 240.0, 0.0, 60.0, 0.01, 100_000_000, 1 << 27, margin=0.0)
    assert round((1.0 / g2.inv_h) / 0.01) == 1          # dense enough for h = r


def test_synthetic_dataset_contract_and_fieldset_setup():
    velocity_fields.configure_synthetic(n_modes=4)
    try:
        ds = velocity_fields.oscar_dataset(2017)
        depth = ds["depth"].values[0]
        sub = ds.sel(depth=depth)
        assert sub["u"].values.shape == (72, 481, 1201) and sub["u"].values.dtype == np.float32
        lat = sub["latitude"].values
        assert lat[0] == 80.0 and lat[-1] == -80.0            # descending, like the product
        assert sub["longitude"].values[0] == 20.0 and abs(sub["longitude"].values[-1] - 420.0) < 1e-4
        t = sub["time"].values
        assert (t[1] - t[0]) // np.timedelta64(1, "s") == 432000
        fs = HostFieldSet(ds)
        assert fs.lat[0] < fs.lat[-1] and fs.time[1] == 432000.0 and fs.u.flags.c_contiguous
        assert np.array_equal(fs.u[:, 0, :], sub["u"].values[:, -1, :])   # flipped with the axis
        rms = np.sqrt(np.mean(fs.u[0].astype(np.float64) ** 2 + fs.v[0].astype(np.float64) ** 2))
        assert abs(rms - 0.2) < 1e-3
    finally:
        velocity_fields.configure_synthetic(n_modes=64)


def test_file_formats_round_trip(tmp_path):
    import joblib
    times = [datetime(2017, 1, 1) + timedelta(hours=k + 1) for k in range(3)]
    lat = np.arange(12, dtype=np.float32).reshape(3, 4)
    lon = lat + 100
    name = lmio.chunk_pickle_name(0, 3, 1)
    assert name == "particle_locations_00000_00003_tile01.pickle"
    lmio.dump_chunk(str(tmp_path / name), times, lat, lon)
    chunk = joblib.load(str(tmp_path / name))                      # what the reference's restore path does
    assert set(chunk) == {"time", "lat", "lon"} and np.array_equal(chunk["lon"], lon) and chunk["time"] == times
    assert lmio.parse_chunk_name(name) == (0, 3, 1)
    variables = {"longitude": lon.T.copy(), "latitude": lat.T.copy(), "species": np.ones((4, 3), dtype=np.int8)}
    path = str(tmp_path / "microbe_data.nc")
    lmio.write_particle_file(path, variables, times)
    back = lmio.read_particle_file(path)
    assert back.times == times
    for k, v in variables.items():
        assert np.array_equal(back[k], v) and back[k].dtype == v.dtype
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as nc:
        assert nc.variables["longitude"].dimensions == ("particle number", "time")
        assert list(nc.variables["particle number"][:]) == [1, 2, 3, 4]


# ---- the product's own files: oscar_vel<year>.nc read the way xarray.open_dataset decodes it --------------------
def _write_oscar_like(path, fill=None, packed=False):
    """A small file in the OSCAR product's layout (velocity_fields.py:21-32 opens such a file with xarray)."""
    from scipy.io import netcdf_file
    rng = np.random.default_rng(4)
    T, Y, X = 4, 7, 9
    days = np.array([8852, 8857, 8862, 8868], dtype=np.int32)               # irregular last step
    lat = 40.0 - np.arange(Y) / 3.0                                         # descending, like the product
    lon = 200.0 + np.arange(X) / 3.0
    u = rng.normal(0, 0.3, (T, 2, Y, X))
    v = rng.normal(0, 0.3, (T, 2, Y, X))
    land = np.zeros((Y, X), dtype=bool)
    land[2:4, 5:8] = True
    with netcdf_file(path, "w") as f:
        for name, n in (("time", T), ("depth", 2), ("latitude", Y), ("longitude", X)):
            f.createDimension(name, n)
        tv = f.createVariable("time", "i4", ("time",))
        tv[:] = days
        tv.units = "day since 1992-10-05 00:00:00"
        for name, vals, dt in (("depth", [15.0, 30.0], "f4"), ("latitude", lat, "f8"), ("longitude", lon, "f8")):
            var = f.createVariable(name, dt, (name,))
            var[:] = np.asarray(vals)
        for name, data in (("u", u), ("v", v)):
            if packed:                                                      # int16 with scale/offset and a fill value
                var = f.createVariable(name, "i2", ("time", "depth", "latitude", "longitude"))
                q = np.round((data - 0.25) / 1e-3).astype(np.int16)
                q[:, :, land] = -32767
                var[:] = q
                var.scale_factor = np.float64(1e-3)
                var.add_offset = np.float64(0.25)
                var._FillValue = np.int16(-32767)
            else:
                var = f.createVariable(name, "f8", ("time", "depth", "latitude", "longitude"))
                d = data.copy()
                d[:, :, land] = np.nan if fill is None else fill
                var[:] = d
                var.missing_value = np.float64(np.nan if fill is None else fill)
    return days, lat, lon, u, v, land


@pytest.mark.parametrize("kind", ["nan", "fill", "packed"])
def test_oscar_file_in_the_working_directory_is_read_like_xarray_would(tmp_path, monkeypatch, kind):
    from lagrangian_microbes_b200 import velocity_fields as vf
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet
    path = str(tmp_path / vf.oscar_dataset_filename(2017))
    days, lat, lon, u, v, land = _write_oscar_like(path, fill={"nan": None, "fill": -9999.0, "packed": None}[kind],
                                                   packed=kind == "packed")
    monkeypatch.chdir(tmp_path)
    vf.register_dataset_provider(None)
    ds = vf.oscar_dataset(2017)
    assert isinstance(ds, vf.NetcdfDataset) and vf.oscar_dataset(2017) is ds          # cached
    assert ds["depth"].values[0] == np.float32(15.0)
    t = ds["time"].values
    assert t.dtype == np.dtype("datetime64[ns]") and t[0] == np.datetime64("2016-12-30T00:00:00")
    assert ds["u"].values.shape == (4, 2, 7, 9)
    sub = ds.sel(depth=ds["depth"].values[0])
    assert sub["u"].values.shape == (4, 7, 9) and list(sub["depth"].values) == [15.0]
    got = sub["u"].values
    assert np.isnan(got[:, land]).all() and not np.isnan(got[:, ~land]).any()
    want = u[:, 0] if kind != "packed" else np.round((u[:, 0] - 0.25) / 1e-3) * 1e-3 + 0.25
    assert np.allclose(got[:, ~land], want[:, ~land], rtol=0, atol=1e-12)
    deeper = ds.sel(depth=29.0)["v"].values                                            # nearest level
    assert np.allclose(deeper[:, ~land], (v[:, 1] if kind != "packed" else np.round((v[:, 1] - 0.25) / 1e-3) * 1e-3 + 0.25)[:, ~land],
                       rtol=0, atol=1e-12)
    # what Parcels' field construction makes of it (particle_advecter.py:160-184)
    fs = HostFieldSet(ds)
    assert list(fs.time) == [0.0, 5 * 86400.0, 10 * 86400.0, 16 * 86400.0]
    assert fs.lat[0] < fs.lat[-1] and fs.lat.dtype == np.float32 and fs.u.dtype == np.float32
    assert np.array_equal(fs.lat, lat[::-1].astype(np.float32)) and np.array_equal(fs.lon, lon.astype(np.float32))
    assert (fs.u[:, land[::-1]] == 0).all() and np.array_equal(fs.u[:, ~land[::-1]], want[:, ::-1][:, ~land[::-1]].astype(np.float32))


def test_synthetic_dataset_round_trips_through_the_product_layout(tmp_path, monkeypatch):
    """save_dataset (the reference's dataset.to_netcdf, velocity_fields.py:30) + the file reader give back the
    dataset bit for bit; without a file, and with nothing to download from, the synthetic field is served."""
    from lagrangian_microbes_b200 import velocity_fields as vf
    monkeypatch.chdir(tmp_path)
    vf.register_dataset_provider(None)
    lon, lat = vf.oscar_grid()
    lon, lat = lon[540:560], lat[130:150]
    times_s = vf.OSCAR_DT_SECONDS * np.arange(3, dtype=np.int64)
    u, v = vf.synthetic_uv(lon, lat, times_s, n_modes=5, land=False)
    u[:, 3:5, 6:9] = np.nan
    time = (np.datetime64("2017-01-01T00:00:00", "s") + times_s.astype("timedelta64[s]")).astype("datetime64[ns]")
    ds = vf.SyntheticDataset({"time": time, "depth": np.array([15.0], dtype=np.float32), "latitude": lat.astype(np.float64),
                              "longitude": lon.astype(np.float64), "u": u[:, None], "v": v[:, None]})
    assert vf.oscar_dataset_path(2017) is None
    vf.save_dataset(ds, vf.oscar_dataset_filename(2017))
    back = vf.oscar_dataset(2017)
    assert isinstance(back, vf.NetcdfDataset)
    for name in ("time", "depth", "latitude", "longitude", "u", "v"):
        a, b = ds[name].values, back[name].values
        assert a.dtype == b.dtype and a.shape == b.shape, name
        assert np.array_equal(a, b, equal_nan=name in ("u", "v")), name
    # $LM_OSCAR_DIR is searched after the working directory
    other = tmp_path / "elsewhere"
    other.mkdir()
    monkeypatch.chdir(other)
    assert vf.oscar_dataset_path(2017) is None
    monkeypatch.setenv("LM_OSCAR_DIR", str(tmp_path))
    assert vf.oscar_dataset_path(2017) == str(tmp_path / "oscar_vel2017.nc")
    with pytest.raises(ValueError):
        vf.save_dataset(ds, "x.nc", time_units="day since 1992-10-05 07:00:00")      # not a whole number of days
    # a NetCDF-4 (HDF5) file is refused with a message that says what to do
    with open("oscar_vel2018.nc", "wb") as fh:
        fh.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(OSError, match="NetCDF-4"):
        vf.oscar_dataset(2018)


def test_cf_time_units():
    from lagrangian_microbes_b200.velocity_fields import decode_cf_time
    assert decode_cf_time(np.array([0, 1], dtype=np.int32), "days since 2000-01-01")[1] == np.datetime64("2000-01-02")
    assert decode_cf_time(np.array([1.5]), "hours since 2000-01-01T06:30:00")[0] == np.datetime64("2000-01-01T08:00:00")
    assert decode_cf_time(np.array([90], dtype=np.int64), "seconds since 1970-1-1 0:0:0")[0] == np.datetime64("1970-01-01T00:01:30")
    with pytest.raises(ValueError):
        decode_cf_time(np.array([1]), "fortnights since 2000-01-01")


def test_record_assembler_keeps_every_stride_th_step_in_the_reference_layout(tmp_path):
    from datetime import datetime, timedelta
    from lagrangian_microbes_b200 import io as lmio
    n, steps, stride = 50, 11, 3
    t0, dt = datetime(2018, 1, 1), timedelta(hours=1)
    asm = lmio.RecordAssembler(n, steps, t0, dt, stride)
    assert asm.kept_steps == [0, 3, 6, 9] and [asm.wants(k) for k in (0, 1, 3, 11)] == [True, False, True, False]
    rng = np.random.default_rng(0)
    cols = {}
    for k in asm.kept_steps:
        cols[k] = (rng.random(n).astype(np.float32), rng.random(n).astype(np.float32), rng.integers(1, 4, n).astype(np.int8))
        asm.put(k, *cols[k])
    path = asm.write(str(tmp_path / "out"))
    assert os.path.basename(path) == "microbe_data.nc"
    back = lmio.read_particle_file(path)                         # dims ("particle number", "time"), interaction_simulator.py:68-77
    assert back.times == [t0 + k * dt for k in asm.kept_steps]
    for j, k in enumerate(asm.kept_steps):
        assert np.array_equal(back["longitude"][:, j], cols[k][0]) and np.array_equal(back["species"][:, j], cols[k][2])
    short = lmio.RecordAssembler(n, steps, t0, dt, stride)
    short.put(0, *cols[0])
    with pytest.raises(AssertionError):
        short.write(str(tmp_path / "out2"))


def test_years_concatenate_into_one_calendar_field_and_the_clock_starts_at_start_time():
    """calendar_time=True (the fix of quirk Q1 behind an option; default = the reference's behaviour): consecutive years'
    files become one field with a continuous time axis, the particle clock of a time_step call is the calendar time
    inside it, and the stage decisions across the year boundary are the oracle's."""
    from datetime import datetime
    from lagrangian_microbes_b200 import velocity_fields
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet, StageClock
    from oracle import rk4 as ork4
    velocity_fields.configure_synthetic(kind="random_fourier", seed=3, n_modes=4, rms_speed=0.2)
    try:
        a = HostFieldSet(velocity_fields.oscar_dataset(2017))
        t_last_2017 = a.time[-1]
        fs = HostFieldSet.from_years((2017, 2018))
        assert fs.u.shape[0] == 144 and fs.time.shape == (144,) and np.all(np.diff(fs.time) > 0)
        assert fs.time[71] == t_last_2017 == 71 * 432000.0 and fs.time[72] == 365 * 86400.0        # 2018-01-01 since 2017-01-01
        assert np.array_equal(fs.u[:72], a.u)
        b = HostFieldSet(velocity_fields.oscar_dataset(2018))
        assert np.array_equal(fs.u[72:], b.u) and not np.array_equal(a.u[:5], b.u[:5])           # one flow, continued: 2018 differs from 2017
        start = datetime(2017, 12, 20, 6, 0, 0)
        t0 = fs.seconds_since_first_snapshot(start)
        assert t0 == (353 * 24 + 6) * 3600.0
        clock = StageClock(fs.time, t0=t0)
        ofs = ork4.FieldSet(fs.lon, fs.lat, fs.time, fs.u, fs.v)
        t, ti = t0, 0
        crossed = False
        for _ in range(24 * 14):                                       # two weeks of hourly steps: over the gap between the years' files
            st = clock.next_step(3600.0)
            want, ti = ork4.stage_times(ofs, t, 3600.0, ti)
            assert [(st.ti[k], st.interp[k], st.frac[k]) for k in range(4)] == [(w[0], int(w[1]), np.float32(w[2])) for w in want]
            crossed = crossed or st.ti[0] == 71
            t += 3600.0
        assert crossed and clock.ti == 72
    finally:
        velocity_fields.configure_synthetic(kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2)


def test_tiles_are_dealt_out_over_ranks_in_contiguous_balanced_blocks():
    """N_procs tiles (particle_advecter.py:143-148: one joblib worker each) -> one rank (GPU) each, in order."""
    from lagrangian_microbes_b200.particle_advecter import tiles_of_rank
    for tiles in (1, 2, 5, 6, 8, 40):
        for world in (1, 2, 3, 4, 8):
            blocks = [tiles_of_rank(tiles, r, world) for r in range(world)]
            assert sum(blocks, []) == list(range(tiles))                          # every tile once, in order
            sizes = [len(b) for b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert tiles_of_rank(6, 1, 4) == [1, 2] and tiles_of_rank(2, 3, 4) == [1] and tiles_of_rank(2, 0, 4) == []


def test_netcdf4_files_of_the_reference_are_recognised(tmp_path):
    """xarray.to_netcdf writes NetCDF-4 (HDF5) when netCDF4 is installed (particle_advecter.py:300-305): such a file is
    detected by its signature and read through netCDF4 / h5py when present, else refused with a message that says what
    it is -- not passed to SciPy's NetCDF-3 reader."""
    from lagrangian_microbes_b200 import io as lmio
    p = tmp_path / "particle_data.nc"
    p.write_bytes(lmio.HDF5_MAGIC + b"\0" * 64)
    try:
        import netCDF4  # noqa: F401
        have = True
    except ImportError:
        try:
            import h5py  # noqa: F401
            have = True
        except ImportError:
            have = False
    if not have:
        with pytest.raises(OSError, match="NetCDF-4"):
            lmio.read_particle_file(str(p))
    assert lmio._parse_time_units("seconds since 2017-01-01 00:00:00") == (1.0, datetime(2017, 1, 1))
    assert lmio._parse_time_units("hours since 2017-01-01T06:00:00") == (3600.0, datetime(2017, 1, 1, 6))
    assert lmio._parse_time_units("days since 1992-10-05") == (86400.0, datetime(1992, 10, 5))


@pytest.mark.parametrize("mode", ["in memory", "mapped", "mapped, one column per block"])
def test_particle_file_writer_fills_the_same_file_piece_by_piece(tmp_path, mode):
    """io.ParticleFileWriter: time columns (InteractionSimulator) or particle x time blocks (create_netcdf_file) instead
    of the reference's dense (N, Nt) host arrays; beyond NetCDF-3's variable limit the memory-mapped directory layout."""
    N, Nt = 37, 11
    times = [datetime(2017, 1, 1) + k * timedelta(hours=1) for k in range(Nt)]
    rng = np.random.default_rng(0)
    lon, lat = rng.random((N, Nt)).astype(np.float32), rng.random((N, Nt)).astype(np.float32)
    sp = rng.integers(1, 4, (N, Nt)).astype(np.int8)
    kw = {"in memory": {}, "mapped": dict(var_limit=100, block_bytes=4 * N * 9),
          "mapped, one column per block": dict(var_limit=100, block_bytes=1)}[mode]
    path = str(tmp_path / "microbe_data.nc")
    w = lmio.ParticleFileWriter(path, {"longitude": np.float32, "latitude": np.float32, "species": np.int8}, N, times, **kw)
    assert w.large == (mode != "in memory")
    for i in range(Nt):
        if i != 5:                                               # a column never written stays zero (the reference's zeros())
            w.put(i, longitude=lon[:, i], latitude=lat[:, i], species=sp[:, i])
    written = w.close()
    assert written == (path if mode == "in memory" else path + ".npz.d") and os.path.exists(written)
    f = lmio.read_particle_file(path)
    want_lon, want_sp = lon.copy(), sp.copy()
    want_lon[:, 5], want_sp[:, 5] = 0, 0
    assert np.array_equal(np.asarray(f["longitude"]), want_lon) and np.array_equal(np.asarray(f["species"]), want_sp)
    assert f.times == times and np.asarray(f["species"]).dtype == np.int8
    # blocks of particles x time ranges, as the chunk pickles arrive
    path2 = str(tmp_path / "particle_data.nc")
    w = lmio.ParticleFileWriter(path2, {"longitude": np.float32, "latitude": np.float32}, N, times, **kw)
    for rows in (slice(0, 20), slice(20, N)):
        for t1, t2 in ((0, 6), (6, Nt)):
            w.put_block(rows, t1, t2, longitude=lon[rows, t1:t2], latitude=lat[rows, t1:t2])
    w.close()
    f = lmio.read_particle_file(path2)
    assert np.array_equal(np.asarray(f["longitude"]), lon) and np.array_equal(np.asarray(f["latitude"]), lat)


def test_record_assembler_opened_on_its_destination_maps_large_records(tmp_path, monkeypatch):
    """FusedSimulation.run_to_file hands the assembler its destination up front: a record beyond NetCDF-3's variable limit
    is then filled through memory-mapped files, column block by column block; species counts are kept per column."""
    n, steps = 40, 9
    t0, dt = datetime(2018, 1, 1), timedelta(hours=1)
    rng = np.random.default_rng(1)
    cols = [(rng.random(n).astype(np.float32), rng.random(n).astype(np.float32), rng.integers(0, 5, n).astype(np.int8)) for _ in range(steps)]
    outs = []
    for tag, limit in (("nc3", None), ("mapped", 4 * n * 2)):
        if limit is not None:
            monkeypatch.setattr(lmio, "_NC3_VAR_LIMIT", limit)
        asm = lmio.RecordAssembler(n, steps, t0, dt, 2, output_dir=str(tmp_path / tag))
        for k in asm.kept_steps:
            asm.put(k, *cols[k])
        with pytest.raises(AssertionError):
            asm.write(str(tmp_path / "elsewhere"))                # opened on another destination
        path = asm.write(str(tmp_path / tag))
        assert path.endswith("microbe_data.nc" if limit is None else "microbe_data.nc.npz.d")
        back = lmio.read_particle_file(os.path.join(str(tmp_path / tag), "microbe_data.nc"))
        outs.append({k: np.array(back[k]) for k in ("longitude", "latitude", "species")})
        assert [list(c) for c in asm.counts] == [[int((cols[k][2] == s).sum()) for s in (1, 2, 3)] for k in asm.kept_steps]
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]) and outs[0][k].shape == (n, 5)
        assert np.array_equal(outs[0][k][:, 1], cols[2][("longitude", "latitude", "species").index(k)])


def test_files_without_columns_or_particles_round_trip(tmp_path):
    """Edge cases of the file layer: a time_step call with start_time == end_time (no columns) and an empty particle set."""
    for N, Nt in ((5, 0), (0, 3), (0, 0)):
        times = [datetime(2017, 1, 1) + k * timedelta(hours=1) for k in range(Nt)]
        path = str(tmp_path / ("f%d_%d.nc" % (N, Nt)))
        w = lmio.ParticleFileWriter(path, {"longitude": np.float32, "latitude": np.float32, "species": np.int8}, N, times)
        w.close()
        f = lmio.read_particle_file(path)
        assert np.asarray(f["longitude"]).shape == (N, Nt) and np.asarray(f["species"]).dtype == np.int8 and f.times == times
