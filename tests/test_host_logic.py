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
