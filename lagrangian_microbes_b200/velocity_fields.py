"""Velocity-field inputs: the reference's ``oscar_dataset(year)`` contract, synthetic data.

Mirrors /root/reference/velocity_fields.py:21-32.  The reference downloads the OSCAR
third-degree surface-current product over OPeNDAP and opens it with xarray; neither the
network nor xarray exists here, so ``oscar_dataset`` returns a small
dataset object that supports exactly the accesses the advecter makes
(/root/reference/particle_advecter.py:160-180):

    ds["depth"].values[0];  ds.sel(depth=d);  sub["time"|"latitude"|"longitude"|"u"|"v"].values

filled with an analytic (steady) or random-Fourier (time-varying) eddy field on the OSCAR
grid: longitude float32(20 + k/3), k=0..1200; latitude float32(80 - j/3), j=0..480
(DESCENDING, like the product); 72 snapshots 432000 s apart; depth [15.0];
``u``/``v`` float32 (time, depth, latitude, longitude) in m/s.

A user with the real NetCDF files can register any object with the same accessors via
``register_dataset_provider``.
"""
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


def oscar_dataset(year):
    """velocity_fields.py:21-32 -- same name, same return contract, synthetic contents."""
    if _PROVIDER is not None:
        return _PROVIDER(year)
    key = (year,) + tuple(sorted(_CONFIG.items()))
    if key not in _CACHE:
        lon, lat = oscar_grid()
        times_s = OSCAR_DT_SECONDS * np.arange(OSCAR_NT, dtype=np.int64)
        u, v = synthetic_uv(lon, lat, times_s, **_CONFIG)
        time = np.datetime64("%04d-01-01T00:00:00" % year, "s") + times_s.astype("timedelta64[s]")
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
