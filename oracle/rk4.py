"""RK4 advection oracle (TEST INFRASTRUCTURE ONLY) -- PARITY UNPINNED.

The reference advects with one call per step of

    pset.execute(parcels.AdvectionRK4, runtime=dt, dt=dt, ...)      particle_advecter.py:222-223

on a ``RectilinearZGrid(lon, lat, depth, time, mesh="spherical")`` (:177) with
``Field(..., interp_method="linear")`` U and V (:182-183) and ``JITParticle`` (:186-187).
The arithmetic therefore lives in **parcels 2.0.0beta2** (environment.yml:80), which is not
under /root/reference, not installed and not installable offline, and the reference has no
test or golden vector touching it.  This file RESTATES the published algorithm of that
release (parcels/kernels/advection.py::AdvectionRK4, parcels/include/parcels.h:
search_indices_rectilinear, spatial_interpolation_bilinear,
temporal_interpolation_structured_grid, search_time_index; parcels/tools/converters.py:
Geographic, GeographicPolar; the JIT code generator's float locals) -- it is anchored by
analytic known-answer tests only (tests/test_oracle_rk4.py), hence "parity unpinned".

Restated semantics (the JIT-compiled C path that ``JITParticle`` selects):

Setup (Field/Grid construction)
  * data cast to float32, NaN -> 0; if lat is descending, lat and the data's y axis are flipped;
    lon/lat grid arrays float32; grid time float64 seconds since the first snapshot
    (particle_advecter.py:172-173); particle lon/lat float32, particle time float64 starting at
    grid.time[0] = 0; particle dt float32.

Sample S(t, y, x) for F in (U, V)  [x, y float32]
  * time index ``ti`` cached on the particle: ``while ti < T-1 and t > time[ti+1]: ti += 1``;
    ``while ti > 0 and t < time[ti]: ti -= 1``; t outside [time[0], time[-1]] raises
    TimeExtrapolationError.
  * horizontal: xi with lon[xi] <= x <= lon[xi+1] (local linear search; the +-360 wrap
    branches of the spherical search are unreachable for grid 20..420 and 20 <= x < 245),
    ``xsi = (x - lon[xi]) / (lon[xi+1] - lon[xi])`` evaluated in **float32**, stored as
    double; likewise yi / eta; x or y outside the grid raises OutOfBoundsError.
  * bilinear, double arithmetic, result rounded to float32:
        f = (1-xsi)*(1-eta)*d[yi][xi] + xsi*(1-eta)*d[yi][xi+1] + xsi*eta*d[yi+1][xi+1] + (1-xsi)*eta*d[yi+1][xi]
  * if ``ti < T-1 and t > time[ti]``: ``f = f0 + (f1 - f0) * (float)((t - t0) / (t1 - t0))`` in
    float32, with f0, f1 the bilinear values at ti and ti+1; else f = f0.
  * unit conversion (mesh="spherical"), double arithmetic, rounded to float32:
        u *= 1.0 / (1852. * 60. * cos(y * M_PI / 180))        (y = the SAMPLE point's latitude)
        v *= 1.0 / (1852. * 60.)

AdvectionRK4 (locals are C floats; expressions promote to double through the .5 / 6. literals)
        (u1, v1) = S(t, lat, lon);          lon1 = lon + u1*.5*dt;  lat1 = lat + v1*.5*dt
        (u2, v2) = S(t + .5*dt, lat1, lon1); lon2 = lon + u2*.5*dt;  lat2 = lat + v2*.5*dt
        (u3, v3) = S(t + .5*dt, lat2, lon2); lon3 = lon + u3*dt;     lat3 = lat + v3*dt
        (u4, v4) = S(t + dt, lat3, lon3)
        lon += (u1 + 2*u2 + 2*u3 + u4) / 6. * dt;   lat += (v1 + 2*v2 + 2*v3 + v4) / 6. * dt
    (stages 1-2 and the final update are evaluated in double because of the .5 / 6. literals and
    rounded to float32 on assignment; the stage-3 line has no double literal, so with the
    float32 ``dt`` of that release it is pure float32; the four-term sums are float32 adds.)
    A failed sample leaves the particle unchanged and raises.

Two implementations are provided:
  ``rk4_step_f32``  the float32-faithful restatement above (NumPy float32/float64 array ops
                    are IEEE single/double, so this reproduces the C semantics exactly, up to
                    libm's cos).
  ``rk4_step_f64``  the same algorithm carried entirely in float64 (positions, weights, sums)
                    on the same float32 grid data: the "exact" answer the north-star tolerance
                    (1e-6 relative on positions) is measured against.
``oracle/rk4_c.c`` is the float32-faithful restatement again in C (the timed CPU baseline).
"""
import ctypes

import numpy as np

F32 = np.float32
F64 = np.float64


class FieldSet:
    """Grid + U, V as Parcels holds them after construction (see module docstring: Setup)."""

    def __init__(self, lon, lat, time, u, v):
        lon = np.ascontiguousarray(lon, dtype=F32)
        lat = np.asarray(lat, dtype=F32)
        u = np.asarray(u, dtype=F32)
        v = np.asarray(v, dtype=F32)
        assert u.ndim == 3 and u.shape == v.shape == (len(time), lat.size, lon.size)
        if lat[-1] < lat[0]:                       # descending latitude -> flip to ascending
            lat = lat[::-1]
            u = u[:, ::-1, :]
            v = v[:, ::-1, :]
        u = np.where(np.isnan(u), F32(0), u)       # land -> 0
        v = np.where(np.isnan(v), F32(0), v)
        self.lon = lon
        self.lat = np.ascontiguousarray(lat)
        self.time = np.ascontiguousarray(time, dtype=F64)
        self.u = np.ascontiguousarray(u, dtype=F32)
        self.v = np.ascontiguousarray(v, dtype=F32)
        assert np.all(np.diff(self.lon) > 0) and np.all(np.diff(self.lat) > 0) and np.all(np.diff(self.time) > 0)


class TimeExtrapolationError(RuntimeError):
    pass


class OutOfBoundsError(RuntimeError):
    pass


def search_time_index(time_axis, t, ti):
    """parcels.h::search_time_index + the interpolate/hold decision.  Returns (ti, interp, frac32)."""
    T = time_axis.size
    if t < time_axis[0] or t > time_axis[T - 1]:
        raise TimeExtrapolationError("t=%r outside [%r, %r]" % (t, time_axis[0], time_axis[T - 1]))
    while ti < T - 1 and t > time_axis[ti + 1]:
        ti += 1
    while ti > 0 and t < time_axis[ti]:
        ti -= 1
    if ti < T - 1 and t > time_axis[ti]:
        t0, t1 = F64(time_axis[ti]), F64(time_axis[ti + 1])
        return ti, True, F32((F64(t) - t0) / (t1 - t0))
    return ti, False, F32(0)


def _search_axis(vals, x):
    """Index i with vals[i] <= x <= vals[i+1]; out-of-range mask returned separately."""
    n = vals.size
    oob = (x < vals[0]) | (x > vals[n - 1]) | ~np.isfinite(x)
    i = np.searchsorted(vals, x, side="right") - 1
    i = np.clip(i, 0, n - 2)
    return i, oob


def _sample(fs, x, y, ti, interp, frac32, faithful):
    """Sample (U, V) at float32 (faithful) or float64 points; returns converted (u, v), oob mask."""
    if faithful:
        xi, oobx = _search_axis(fs.lon, x)
        yi, ooby = _search_axis(fs.lat, y)
        xsi = ((x - fs.lon[xi]) / (fs.lon[xi + 1] - fs.lon[xi])).astype(F64)      # float32 ops
        eta = ((y - fs.lat[yi]) / (fs.lat[yi + 1] - fs.lat[yi])).astype(F64)
    else:
        lon64, lat64 = fs.lon.astype(F64), fs.lat.astype(F64)
        xi, oobx = _search_axis(lon64, x)
        yi, ooby = _search_axis(lat64, y)
        xsi = (x - lon64[xi]) / (lon64[xi + 1] - lon64[xi])
        eta = (y - lat64[yi]) / (lat64[yi + 1] - lat64[yi])
    oob = oobx | ooby

    def bilinear(d):
        d00 = d[yi, xi].astype(F64)
        d01 = d[yi, xi + 1].astype(F64)
        d11 = d[yi + 1, xi + 1].astype(F64)
        d10 = d[yi + 1, xi].astype(F64)
        val = (1 - xsi) * (1 - eta) * d00 + xsi * (1 - eta) * d01 + xsi * eta * d11 + (1 - xsi) * eta * d10
        return val.astype(F32) if faithful else val

    out = []
    for data in (fs.u, fs.v):
        f0 = bilinear(data[ti])
        if interp:
            f1 = bilinear(data[ti + 1])
            if faithful:
                f = f0 + (f1 - f0) * F32(frac32)                                   # float32 ops
            else:
                f = f0 + (f1 - f0) * F64(frac32)
        else:
            f = f0
        out.append(f)
    u, v = out
    y64 = y.astype(F64)
    cu = 1.0 / (1852. * 60. * np.cos(y64 * np.pi / 180))
    cv = 1.0 / (1852. * 60.)
    u = u.astype(F64) * cu
    v = v.astype(F64) * cv
    if faithful:
        u, v = u.astype(F32), v.astype(F32)
    return u, v, oob


def stage_times(fs, t, dt, ti):
    """The four sample times of one RK4 step and their cached-index decisions.

    Returns ([(ti, interp, frac32)] * 4, ti_after).  Stages 2 and 3 share t + dt/2.
    """
    dt32 = F64(F32(dt))
    out = []
    for ts in (F64(t), F64(t) + .5 * dt32, F64(t) + .5 * dt32, F64(t) + dt32):
        ti, interp, frac = search_time_index(fs.time, ts, ti)
        out.append((ti, interp, frac))
    return out, ti


def rk4_step_f32(fs, lon, lat, t, dt, ti=0):
    """One float32-faithful AdvectionRK4 step.  lon/lat float32 arrays (not modified).

    Returns (lon_new, lat_new, ti_after, n_out_of_bounds).  Out-of-bounds particles are left
    unchanged (Parcels would raise; the count is what the device reports).
    """
    lon = np.asarray(lon, dtype=F32)
    lat = np.asarray(lat, dtype=F32)
    dt32 = F32(dt)
    dtd = F64(dt32)
    st, ti_after = stage_times(fs, t, dt, ti)
    lon64, lat64 = lon.astype(F64), lat.astype(F64)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        u1, v1, o1 = _sample(fs, lon, lat, *st[0], True)
        lon1 = (lon64 + u1.astype(F64) * .5 * dtd).astype(F32)
        lat1 = (lat64 + v1.astype(F64) * .5 * dtd).astype(F32)
        u2, v2, o2 = _sample(fs, lon1, lat1, *st[1], True)
        lon2 = (lon64 + u2.astype(F64) * .5 * dtd).astype(F32)
        lat2 = (lat64 + v2.astype(F64) * .5 * dtd).astype(F32)
        u3, v3, o3 = _sample(fs, lon2, lat2, *st[2], True)
        lon3 = lon + u3 * dt32                  # no double literal in this line: float32 ops
        lat3 = lat + v3 * dt32
        u4, v4, o4 = _sample(fs, lon3, lat3, *st[3], True)
        su = u1 + F32(2) * u2 + F32(2) * u3 + u4                                   # float32 adds
        sv = v1 + F32(2) * v2 + F32(2) * v3 + v4
        lon_new = (lon64 + su.astype(F64) / 6. * dtd).astype(F32)
        lat_new = (lat64 + sv.astype(F64) / 6. * dtd).astype(F32)
    oob = o1 | o2 | o3 | o4
    lon_new = np.where(oob, lon, lon_new)
    lat_new = np.where(oob, lat, lat_new)
    return lon_new, lat_new, ti_after, int(oob.sum())


def rk4_step_f64(fs, lon, lat, t, dt, ti=0):
    """Same step with float64 positions/weights/sums on the same float32 grid data."""
    lon = np.asarray(lon, dtype=F64)
    lat = np.asarray(lat, dtype=F64)
    dtd = F64(F32(dt))
    st, ti_after = stage_times(fs, t, dt, ti)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        u1, v1, o1 = _sample(fs, lon, lat, *st[0], False)
        u2, v2, o2 = _sample(fs, lon + u1 * .5 * dtd, lat + v1 * .5 * dtd, *st[1], False)
        u3, v3, o3 = _sample(fs, lon + u2 * .5 * dtd, lat + v2 * .5 * dtd, *st[2], False)
        u4, v4, o4 = _sample(fs, lon + u3 * dtd, lat + v3 * dtd, *st[3], False)
        lon_new = lon + (u1 + 2 * u2 + 2 * u3 + u4) / 6. * dtd
        lat_new = lat + (v1 + 2 * v2 + 2 * v3 + v4) / 6. * dtd
    oob = o1 | o2 | o3 | o4
    return np.where(oob, lon, lon_new), np.where(oob, lat, lat_new), ti_after, int(oob.sum())


# ---------------------------------------------------------------------------------------------
# C restatement (oracle/rk4_c.c), OpenMP over particles -- the timed CPU baseline for advection.
# ---------------------------------------------------------------------------------------------
def rk4_step_c(fs, lon, lat, t, dt, ti=0, threads=0):
    """float32-faithful step through oracle/rk4_c.c.  Updates lon/lat (float32, contiguous) IN PLACE.

    Returns (ti_after, n_out_of_bounds).
    """
    from .rps import load_c
    lib = load_c()
    fn = lib.rk4_step_f32
    fn.restype = ctypes.c_int64
    assert lon.dtype == F32 and lat.dtype == F32 and lon.flags.c_contiguous and lat.flags.c_contiguous
    st, ti_after = stage_times(fs, t, dt, ti)
    tis = (ctypes.c_int * 4)(*[s[0] for s in st])
    interp = (ctypes.c_int * 4)(*[int(s[1]) for s in st])
    frac = (ctypes.c_float * 4)(*[float(s[2]) for s in st])
    T, Y, X = fs.u.shape
    n_oob = fn(ctypes.c_void_p(lon.ctypes.data), ctypes.c_void_p(lat.ctypes.data), ctypes.c_int64(lon.size),
               ctypes.c_void_p(fs.u.ctypes.data), ctypes.c_void_p(fs.v.ctypes.data),
               ctypes.c_void_p(fs.lon.ctypes.data), ctypes.c_void_p(fs.lat.ctypes.data),
               ctypes.c_int(T), ctypes.c_int(Y), ctypes.c_int(X),
               tis, interp, frac, ctypes.c_float(dt), ctypes.c_int(threads))
    return ti_after, int(n_oob)
