"""RK4 oracle: analytic known-answer tests (the only anchors available -- parcels is absent, parity unpinned),
agreement of the NumPy and C float32-faithful restatements, and the float32-vs-float64 tolerance."""
import numpy as np
import pytest

from conftest import golden
from oracle import rk4 as ork4

DEG = 1852.0 * 60.0


def make_fs(ufun, vfun, T=4, Y=25, X=31, descending=False, t_axis=None):
    lon = (200.0 + np.arange(X) / 3.0).astype(np.float32)
    lat = (20.0 + np.arange(Y) / 3.0).astype(np.float32)
    time = np.arange(T) * 432000.0 if t_axis is None else np.asarray(t_axis, dtype=np.float64)
    tt, yy, xx = np.meshgrid(time, lat.astype(np.float64), lon.astype(np.float64), indexing="ij")
    u = ufun(tt, yy, xx).astype(np.float32)
    v = vfun(tt, yy, xx).astype(np.float32)
    if descending:
        return ork4.FieldSet(lon, lat[::-1], time, u[:, ::-1, :], v[:, ::-1, :])
    return ork4.FieldSet(lon, lat, time, u, v)


def particles(n=200, seed=0):
    rng = np.random.default_rng(seed)
    return (202.0 + 6.0 * rng.random(n)).astype(np.float32), (22.0 + 4.0 * rng.random(n)).astype(np.float32)


def test_zero_field_is_a_fixed_point():
    fs = make_fs(lambda t, y, x: 0 * x, lambda t, y, x: 0 * x)
    lon, lat = particles()
    for step in (ork4.rk4_step_f32, ork4.rk4_step_f64):
        a, b, _, oob = step(fs, lon, lat, 0.0, 3600.0)
        assert oob == 0 and np.array_equal(a, lon) and np.array_equal(b, lat)


def test_uniform_zonal_and_meridional_flow():
    U0, V0 = 0.25, -0.125                                     # float32-exact
    fs = make_fs(lambda t, y, x: 0 * x + U0, lambda t, y, x: 0 * x + V0)
    lon, lat = particles()
    a, b, _, _ = ork4.rk4_step_f64(fs, lon, lat, 0.0, 3600.0)
    lat64 = lat.astype(np.float64)
    assert np.allclose(b - lat64, V0 * 3600.0 / DEG, rtol=1e-12, atol=0)
    # dlon/dt = U0 / (DEG cos(lat(t))), lat(t) linear in t: closed form is an integral of sec; RK4 is 4th order
    lat_mid = lat64 + 0.5 * V0 * 3600.0 / DEG
    approx = U0 * 3600.0 / (DEG * np.cos(np.deg2rad(lat_mid)))
    assert np.allclose(a - lon.astype(np.float64), approx, rtol=1e-8)
    # pure zonal flow: exactly U0 dt / (DEG cos lat)
    fs2 = make_fs(lambda t, y, x: 0 * x + U0, lambda t, y, x: 0 * x)
    a2, b2, _, _ = ork4.rk4_step_f64(fs2, lon, lat, 0.0, 3600.0)
    assert np.array_equal(b2, lat64)
    assert np.allclose(a2 - lon, U0 * 3600.0 / (DEG * np.cos(lat64 * np.pi / 180)), rtol=1e-13)


def test_field_linear_in_space_and_time_matches_closed_form():
    # v = c0 + c1*(lat - 20): bilinear interpolation is exact; dlat/dt = v/DEG => exponential in t
    c0, c1 = 0.5, 0.125
    fs = make_fs(lambda t, y, x: 0 * x, lambda t, y, x: c0 + c1 * (y - 20.0))
    lon, lat = particles()
    dt = 3600.0
    _, b, _, _ = ork4.rk4_step_f64(fs, lon, lat, 0.0, dt)
    k = c1 / DEG
    y0 = lat.astype(np.float64) - 20.0
    exact = (y0 + c0 / c1) * np.exp(k * dt) - c0 / c1 + 20.0
    assert np.allclose(b, exact, rtol=1e-13)
    # v linear in time, uniform in space: RK4 (Simpson) integrates it exactly
    fs = make_fs(lambda t, y, x: 0 * x, lambda t, y, x: 0.25 + t / 432000.0 * 0.5, T=3)
    t0 = 100 * 3600.0
    _, b, _, _ = ork4.rk4_step_f64(fs, lon, lat, t0, dt, ti=0)
    v_mid = 0.25 + (t0 + dt / 2) / 432000.0 * 0.5
    assert np.allclose(b - lat.astype(np.float64), v_mid * dt / DEG, rtol=1e-6)   # frac is float32


def test_nodes_grid_lines_and_flip():
    rng = np.random.default_rng(1)
    coef = rng.normal(size=6)
    ufun = lambda t, y, x: coef[0] * np.sin(x) + coef[1] * np.cos(y) + coef[2]
    vfun = lambda t, y, x: coef[3] * np.sin(y) + coef[4] * np.cos(x) + coef[5]
    fs = make_fs(ufun, vfun)
    fs_flip = make_fs(ufun, vfun, descending=True)
    assert np.array_equal(fs.u, fs_flip.u) and np.array_equal(fs.lat, fs_flip.lat)
    # sample exactly at nodes: bilinear returns the node value (weights are exactly 0 / 1)
    xi, yi = np.array([0, 3, 10, 29]), np.array([0, 5, 11, 23])
    x, y = fs.lon[xi], fs.lat[yi]
    u, v, oob = ork4._sample(fs, x, y, 0, False, np.float32(0), True)
    assert not oob.any()
    want_u = (fs.u[0, yi, xi].astype(np.float64) / (DEG * np.cos(y.astype(np.float64) * np.pi / 180))).astype(np.float32)
    assert np.array_equal(u, want_u)
    # top/right edges are inside the domain (x == lon[-1] -> last cell, xsi = 1)
    u, v, oob = ork4._sample(fs, fs.lon[-1:], fs.lat[-1:], 0, False, np.float32(0), True)
    assert not oob.any()


def test_nan_land_is_zeroed_and_out_of_bounds_is_counted():
    def ufun(t, y, x):
        out = 0 * x + 0.5
        out[:, 5:9, 5:9] = np.nan
        return out
    fs = make_fs(ufun, lambda t, y, x: 0 * x)
    assert not np.isnan(fs.u).any() and fs.u[0, 6, 6] == 0.0
    lon = np.array([199.0, 205.0, 211.0, np.nan], dtype=np.float32)       # west of the grid, inside, east of it, NaN
    lat = np.array([25.0, 25.0, 25.0, 25.0], dtype=np.float32)
    a, b, _, oob = ork4.rk4_step_f32(fs, lon, lat, 0.0, 3600.0)
    assert oob == 3 and a[0] == lon[0] and a[2] == lon[2] and a[1] > lon[1]


def test_time_index_hysteresis_and_extrapolation():
    ax = np.array([0.0, 432000.0, 864000.0])
    # exactly on a snapshot coming from below: cached index stays, interpolate with fraction 1
    assert ork4.search_time_index(ax, 432000.0, 0) == (0, True, np.float32(1.0))
    # just past it: index advances
    ti, interp, frac = ork4.search_time_index(ax, 432000.0 + 1800.0, 0)
    assert (ti, interp) == (1, True) and frac == np.float32(1800.0 / 432000.0)
    # exactly on the first snapshot: hold
    assert ork4.search_time_index(ax, 0.0, 0) == (0, False, np.float32(0))
    # last snapshot exactly, reached from the last bracket
    assert ork4.search_time_index(ax, 864000.0, 1) == (1, True, np.float32(1.0))
    with pytest.raises(ork4.TimeExtrapolationError):
        ork4.search_time_index(ax, 864000.0 + 1.0, 1)
    fs = make_fs(lambda t, y, x: 0 * x, lambda t, y, x: 0 * x, T=3)
    st, ti = ork4.stage_times(fs, 428400.0, 3600.0, 0)        # step ending exactly on snapshot 1
    assert [s[0] for s in st] == [0, 0, 0, 0] and st[3][2] == np.float32(1.0)
    st, ti = ork4.stage_times(fs, 432000.0, 3600.0, ti)       # next step: stage 1 still in bracket 0
    assert [s[0] for s in st] == [0, 1, 1, 1]


def test_c_and_numpy_float32_restatements_agree_bitwise_and_match_golden():
    g = golden("rk4_small.npz")
    fs = ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"])
    ln, an = g["lon0"].copy(), g["lat0"].copy()
    lc, ac = g["lon0"].copy(), g["lat0"].copy()
    t, ti, tic = 0.0, 0, 0
    for _ in range(int(g["steps"])):
        ln, an, ti, oob = ork4.rk4_step_f32(fs, ln, an, t, 3600.0, ti)
        tic, oobc = ork4.rk4_step_c(fs, lc, ac, t, 3600.0, tic)
        assert oob == oobc and oob >= 5
        t += 3600.0
    assert np.array_equal(ln, lc) and np.array_equal(an, ac)
    assert np.array_equal(ln, g["lon_f32"]) and np.array_equal(an, g["lat_f32"])


def test_float32_path_is_within_1e6_relative_of_float64_per_step():
    """north_star tolerance: positions within 1e-6 relative of the fp64 RK4 on the same field."""
    g = golden("rk4_small.npz")
    fs = ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"])
    lon, lat = g["lon0"], g["lat0"]
    for t, ti in ((0.0, 0), (396000.0, 0), (432000.0 * 2 + 1800.0, 2)):
        a, b, _, _ = ork4.rk4_step_f32(fs, lon, lat, t, 3600.0, ti)
        c, d, _, _ = ork4.rk4_step_f64(fs, lon, lat, t, 3600.0, ti)
        assert np.max(np.abs(a - c) / np.abs(c)) < 1e-6 and np.max(np.abs(b - d) / np.abs(d)) < 1e-6
