"""Velocity-field inputs: the reference's ``oscar_dataset(year)`` contract.

Mirrors /root/reference/velocity_fields.py:21-32.  The reference opens ``oscar_vel<year>.nc`` from the
working directory with xarray, downloading the OSCAR third-degree surface-current product over OPeNDAP
first when the file is missing.  Neither the network nor xarray exists here, so
``oscar_dataset`` returns a small dataset object that supports exactly the accesses the advecter makes
(/root/reference/particle_advecter.py:160-180):

    ds["depth"].values[0];  ds.sel(depth=d);  sub["time"|"latitude"|"longitude"|"u"|"v"].values

* ``oscar_vel<year>.nc`` present in the working directory (or in ``$LM_OSCAR_DIR``): read with
  ``scipy.io.netcdf_file`` (NetCDF-3 classic / 64-bit offset, the format of the product's files) and decoded
  the way ``xarray.open_dataset`` decodes it -- CF time units to datetime64[ns], ``_FillValue`` /
  ``missing_value`` to NaN, ``scale_factor`` / ``add_offset`` applied (``NetcdfDataset``);
* otherwise (there is nothing to download from): an analytic (steady) or random-Fourier (time-varying) eddy
  field on the OSCAR grid: longitude float32(20 + k/3), k=0..1200; latitude float32(80 - j/3), j=0..480
  (DESCENDING, like the product); 72 snapshots 432000 s apart; depth [15.0]; ``u``/``v`` float32
  (time, depth, latitude, longitude) in m/s.  ``save_dataset`` writes any such dataset in the product's layout
  (the reference's ``dataset.to_netcdf(dataset_filepath)``, velocity_fields.py:30).

Any other source can be plugged in with ``register_dataset_provider``.
"""
import os
import re

import numpy as np

OSCAR_NX, OSCAR_NY, OSCAR_NT = 1201, 481, 72
OSCAR_DT_SECONDS = 432000            # 5 days


def oscar_dataset_filename(year):
    """velocity_fields.py:17-18."""
    return "oscar_vel" + str(year) + ".nc"


class _Var:
    def __init__(self, values):
        self.values = values


class SyntheticDataset:
    """Dict-like dataset with ``.sel(depth=...)`` -- the subset of xarray.Dataset the advecter uses."""

    def __init__(self, variables):
        self._vars = dict(variables)

    def __getitem__(self, name):
        return _Var(self._vars[name])

    def sel(self, depth):
        depths = self._vars["depth"]
        idx = int(np.argmin(np.abs(depths - depth)))
        sub = dict(self._vars)
        sub["u"] = self._vars["u"][:, idx]
        sub["v"] = self._vars["v"][:, idx]
        sub["depth"] = depths[idx:idx + 1]
        return SyntheticDataset(sub)


def oscar_grid():
    lon = (20.0 + np.arange(OSCAR_NX) / 3.0).astype(np.float32)
    lat = (80.0 - np.arange(OSCAR_NY) / 3.0).astype(np.float32)       # descending, like OSCAR
    return lon, lat


def _fourier_modes(n_modes, seed, steady):
    rng = np.random.default_rng(seed)
    wavelength = 10.0 ** rng.uniform(0.0, 1.0, n_modes)               # 1..10 degrees
    kmag = 2.0 * np.pi / wavelength
    angle = rng.uniform(0.0, 2.0 * np.pi, n_modes)
    k = kmag * np.cos(angle)
    l = kmag * np.sin(angle)
    amp = kmag ** (-5.0 / 3.0) * rng.normal(1.0, 0.25, n_modes)
    period_days = rng.uniform(10.0, 60.0, n_modes)
    omega = np.zeros(n_modes) if steady else 2.0 * np.pi / (period_days * 86400.0) * rng.choice([-1.0, 1.0], n_modes)
    phase = rng.uniform(0.0, 2.0 * np.pi, n_modes)
    return k, l, amp, omega, phase


def synthetic_uv(lon, lat, times_s, kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2,
                 divergence_fraction=0.15, land=False):
    """Sample the synthetic eddy field on (times, lat, lon) -> two float32 arrays (T, Y, X).

    psi = sum_m a_m cos(k_m x + l_m y - w_m t + p_m); rotational part u = -dpsi/dy, v = dpsi/dx,
    plus ``divergence_fraction`` of the same modes as a potential (convergent) part so that
    particles cluster into filaments as they do in the real surface currents.
    ``kind="steady"`` sets every w_m = 0 (BASELINE config 1); evaluated in float64, cast to float32.
    """
    assert kind in ("steady", "random_fourier")
    x = np.asarray(lon, dtype=np.float64)
    y = np.asarray(lat, dtype=np.float64)
    t = np.asarray(times_s, dtype=np.float64)
    k, l, amp, omega, phase = _fourier_modes(n_modes, seed, kind == "steady")
    eps = float(divergence_fraction)
    cu = amp * (l - eps * k)                      # u = sum cu_m sin(theta_m)
    cv = amp * (-k - eps * l)                     # v = sum cv_m sin(theta_m)
    sx, cx = np.sin(np.outer(k, x)), np.cos(np.outer(k, x))            # (M, X)
    u = np.empty((t.size, y.size, x.size), dtype=np.float32)
    v = np.empty_like(u)
    scale = None
    for n in range(t.size):
        beta = np.outer(y, l) - omega * t[n] + phase                  # (Y, M)
        cb, sb = np.cos(beta), np.sin(beta)
        # sin(kx + beta) = sin(kx) cos(beta) + cos(kx) sin(beta)
        un = (cb * cu) @ sx + (sb * cu) @ cx
        vn = (cb * cv) @ sx + (sb * cv) @ cx
        if scale is None:
            scale = rms_speed / np.sqrt(np.mean(un * un + vn * vn))
        u[n] = (scale * un).astype(np.float32)
        v[n] = (scale * vn).astype(np.float32)
    if land:
        # a few rectangular "islands" of NaN (the product is NaN over land; Parcels zeroes them)
        rng = np.random.default_rng(seed + 1)
        for _ in range(6):
            j0 = int(rng.integers(0, y.size - 12))
            i0 = int(rng.integers(0, x.size - 12))
            u[:, j0:j0 + 9, i0:i0 + 9] = np.nan
            v[:, j0:j0 + 9, i0:i0 + 9] = np.nan
    return u, v


_CONFIG = dict(kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2, divergence_fraction=0.15, land=False)
_PROVIDER = None
_CACHE = {}


def configure_synthetic(**kwargs):
    """Change the synthetic field served by ``oscar_dataset`` (kind, seed, n_modes, rms_speed, ...)."""
    for key in kwargs:
        assert key in _CONFIG, key
    _CONFIG.update(kwargs)
    _CACHE.clear()


def register_dataset_provider(fn):
    """Install ``fn(year) -> dataset`` (e.g. a real OSCAR reader) in place of the synthetic field."""
    global _PROVIDER
    _PROVIDER = fn
    _CACHE.clear()


# ------------------------------------------------------------------------------------------------------
# the product's own files
# ------------------------------------------------------------------------------------------------------
_TIME_UNIT_NS = {"day": 86400 * 10**9, "hour": 3600 * 10**9, "minute": 60 * 10**9, "second": 10**9,
                 "millisecond": 10**6, "microsecond": 10**3}


def decode_cf_time(values, units):
    """``<unit>(s) since <date>[ <time>]`` -> datetime64[ns], as xarray's decode_times does for the standard
    calendar (OSCAR: int32 ``day since 1992-10-05 00:00:00``)."""
    m = re.match(r"\s*(\w+?)s?\s+since\s+(\d{1,4}-\d{1,2}-\d{1,2})(?:[T\s]+(\d{1,2}:\d{1,2}(?::\d{1,2}(?:\.\d+)?)?))?",
                 units.strip(), re.IGNORECASE)
    if m is None or m.group(1).lower() not in _TIME_UNIT_NS:
        raise ValueError("cannot decode time units %r" % units)
    y, mo, d = (int(x) for x in m.group(2).split("-"))
    base = np.datetime64("%04d-%02d-%02d" % (y, mo, d), "ns")
    if m.group(3):
        hms = [float(x) for x in m.group(3).split(":")] + [0.0, 0.0]
        base = base + np.timedelta64(int(round((hms[0] * 3600 + hms[1] * 60 + hms[2]) * 1e9)), "ns")
    unit_ns = _TIME_UNIT_NS[m.group(1).lower()]
    values = np.asarray(values)
    if np.issubdtype(values.dtype, np.integer):
        offs = values.astype(np.int64) * unit_ns
    else:
        offs = np.round(values.astype(np.float64) * unit_ns).astype(np.int64)
    return base + offs.astype("timedelta64[ns]")


def _attr(var, name):
    a = getattr(var, name, None)
    if a is None:
        return None
    a = np.asarray(a)
    return a.reshape(-1)[0] if a.size else None


def decode_cf_variable(var, data):
    """Mask-and-scale of one variable as xarray.open_dataset applies it: fill / missing values become NaN,
    then ``data * scale_factor + add_offset``; float32 stays float32 unless scaling promotes it."""
    data = np.asarray(data)
    fills = [f for f in (_attr(var, "_FillValue"), _attr(var, "missing_value")) if f is not None]
    scale, offset = _attr(var, "scale_factor"), _attr(var, "add_offset")
    if not fills and scale is None and offset is None:
        return np.array(data)
    if np.issubdtype(data.dtype, np.floating) and scale is None and offset is None:
        out = np.array(data)
    else:
        out = data.astype(np.float32 if data.dtype.itemsize <= 2 and scale is None and offset is None else np.float64)
    for f in fills:
        if not (isinstance(f, (float, np.floating)) and np.isnan(f)):
            out[data == f] = np.nan
    if scale is not None:
        out = out * np.float64(scale)
    if offset is not None:
        out = out + np.float64(offset)
    return out


class NetcdfDataset:
    """A NetCDF-3 file with the dataset accessors the advecter uses.  Variables are read (and decoded) on first
    access; ``sel(depth=...)`` picks the nearest depth level of every variable that has a depth dimension."""

    def __init__(self, path, _depth_index=None, _file=None):
        from scipy.io import netcdf_file
        self.path = path
        if _file is None:
            with open(path, "rb") as fh:
                magic = fh.read(4)
            if magic[:3] != b"CDF":
                kind = "NetCDF-4 / HDF5" if magic[1:4] == b"HDF" else "not a NetCDF-3"
                raise OSError("%s is a %s file; this reader handles NetCDF-3 (classic, 64-bit offset). Convert it "
                              "with `nccopy -k classic`, or plug another reader in with register_dataset_provider()."
                              % (path, kind))
            _file = netcdf_file(path, "r", mmap=False)
        self._file = _file
        self._depth_index = _depth_index
        self._cache = {}

    def _depth_dim(self):
        return "depth" if "depth" in self._file.dimensions else None

    def __getitem__(self, name):
        if name not in self._cache:
            var = self._file.variables[name]
            data = np.asarray(var[:])
            data = data.astype(data.dtype.newbyteorder("="))         # the file is big-endian; xarray hands out native arrays
            dd = self._depth_dim()
            if self._depth_index is not None and dd is not None and dd in var.dimensions:
                axis = var.dimensions.index(dd)
                if name == dd:
                    data = data[self._depth_index:self._depth_index + 1]
                else:
                    data = np.take(data, self._depth_index, axis=axis)
            units = getattr(var, "units", b"")
            units = units.decode() if isinstance(units, bytes) else str(units)
            if " since " in units:
                values = decode_cf_time(data, units)
            else:
                values = decode_cf_variable(var, data)
            self._cache[name] = _Var(values)
        return self._cache[name]

    def sel(self, depth):
        depths = np.asarray(self._file.variables["depth"][:], dtype=np.float64)
        idx = int(np.argmin(np.abs(depths - float(depth))))
        return NetcdfDataset(self.path, _depth_index=idx, _file=self._file)

    def close(self):
        self._file.close()


def save_dataset(dataset, path, time_units="day since 1992-10-05 00:00:00"):
    """Write a dataset (synthetic or otherwise) as NetCDF-3 in the layout of the OSCAR product: dimensions
    (time, depth, latitude, longitude), integer ``time`` in ``time_units``, NaN as the missing value --
    the counterpart of ``dataset.to_netcdf(dataset_filepath)`` (velocity_fields.py:30)."""
    from scipy.io import netcdf_file
    t = np.asarray(dataset["time"].values).astype("datetime64[ns]")
    ref = decode_cf_time(np.zeros(1, dtype=np.int64), time_units)[0]
    unit_ns = (decode_cf_time(np.ones(1, dtype=np.int64), time_units)[0] - ref) // np.timedelta64(1, "ns")
    tv = (t - ref) // np.timedelta64(1, "ns")
    if np.any(tv % unit_ns):
        raise ValueError("the time axis is not a whole number of %r" % time_units)
    u, v = np.asarray(dataset["u"].values), np.asarray(dataset["v"].values)
    with netcdf_file(path, "w", version=2) as f:
        for name, n in (("time", t.size), ("depth", u.shape[1]), ("latitude", u.shape[2]), ("longitude", u.shape[3])):
            f.createDimension(name, n)
        tvar = f.createVariable("time", "i4", ("time",))
        tvar[:] = (tv // unit_ns).astype(np.int32)
        tvar.units = time_units
        for name, dtype in (("depth", "f4"), ("latitude", "f8"), ("longitude", "f8")):
            var = f.createVariable(name, dtype, (name,))
            var[:] = np.asarray(dataset[name].values)
        for name, data in (("u", u), ("v", v)):
            var = f.createVariable(name, "f4" if data.dtype == np.float32 else "f8", ("time", "depth", "latitude", "longitude"))
            var[:] = data
            var.units = "meter/sec"
            var.missing_value = np.array(np.nan, dtype=data.dtype if data.dtype == np.float32 else np.float64)
    return path


def oscar_dataset_path(year):
    """Where ``oscar_dataset`` looks for the product's file: the working directory (as the reference does,
    velocity_fields.py:22-23), then ``$LM_OSCAR_DIR``.  None if it is in neither."""
    name = oscar_dataset_filename(year)
    for d in (os.getcwd(), os.environ.get("LM_OSCAR_DIR")):
        if d and os.path.isfile(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


def oscar_dataset(year):
    """velocity_fields.py:21-32 -- same name, same return contract: the product's file when it is there, the
    synthetic field otherwise (the reference would download it; there is no network here)."""
    if _PROVIDER is not None:
        return _PROVIDER(year)
    path = oscar_dataset_path(year)
    if path is not None:
        key = ("file", path, os.path.getmtime(path))
        if key not in _CACHE:
            _CACHE.clear()
            _CACHE[key] = NetcdfDataset(path)
        return _CACHE[key]
    key = (year,) + tuple(sorted(_CONFIG.items()))
    if key not in _CACHE:
        lon, lat = oscar_grid()
        times_s = OSCAR_DT_SECONDS * np.arange(OSCAR_NT, dtype=np.int64)
        first = np.datetime64("%04d-01-01T00:00:00" % year, "s")
        # the synthetic flow is a function of ABSOLUTE time (seconds since 2017-01-01, so 2017 is what it always was): the
        # files of consecutive years continue one flow, which is what a run across a year boundary needs
        since_epoch = int((first - np.datetime64("2017-01-01T00:00:00", "s")) // np.timedelta64(1, "s"))
        u, v = synthetic_uv(lon, lat, times_s + since_epoch, **_CONFIG)
        time = first + times_s.astype("timedelta64[s]")
        _CACHE.clear()                                  # hold one year at a time (332 MB each)
        _CACHE[key] = SyntheticDataset({
            "time": time.astype("datetime64[ns]"),
            "depth": np.array([15.0], dtype=np.float32),
            "latitude": lat.astype(np.float64),
            "longitude": lon.astype(np.float64),
            "u": u[:, None],                            # (time, depth, latitude, longitude)
            "v": v[:, None],
        })
    return _CACHE[key]
