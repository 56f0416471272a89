"""File formats either side of the hot path: the reference's chunk pickles
and its two NetCDF files, written without xarray/netCDF4 (absent here).

* chunk pickles  ``particle_locations_{start:05d}_{end:05d}_tile{id:02d}.pickle``:
  ``joblib.dump({"time": [...], "lat": f32(iters, n), "lon": f32(iters, n)}, compress=("zlib", 3))``
  -- byte-compatible with /root/reference/particle_advecter.py:201-214,246-249.
* ``particle_data.nc`` (particle_advecter.py:262-303) and ``microbe_data.nc``
  (interaction_simulator.py:68-77,119-122): dims ("particle number", "time"), float32
  ``longitude`` / ``latitude`` (+ int8 ``species``), coordinate ``particle number`` = 1..N.
  Written as NetCDF-3 (64-bit offset) through ``scipy.io.netcdf_file``, which xarray opens; the
  time coordinate is CF-encoded ("seconds since <start>").  NetCDF-3 caps a variable at 4 GiB; past
  that the same arrays go to ``<name>.npz.d/<var>.npy`` and readers here accept either.
"""
import os
import pickle
from datetime import datetime, timedelta

import numpy as np

_NC3_VAR_LIMIT = (1 << 32) - 4


def chunk_pickle_name(start_iter, end_iter, tile_id):
    return "particle_locations_" + str(start_iter).zfill(5) + "_" + str(end_iter).zfill(5) + \
           "_tile" + str(tile_id).zfill(2) + ".pickle"


def dump_chunk(filepath, times, lat, lon):
    import joblib
    with open(filepath, "wb") as f:
        joblib.dump({"time": list(times), "lat": lat, "lon": lon}, f, compress=("zlib", 3),
                    protocol=pickle.HIGHEST_PROTOCOL)


def parse_chunk_name(filename):
    """-> (start_iter, end_iter, tile_id), as particle_advecter.py:287-290 does."""
    parts = os.path.basename(filename).split("_")
    return int(parts[2]), int(parts[3]), int(parts[4][4:6])


class ParticleFile:
    """In-memory view of particle_data.nc / microbe_data.nc: dict of (N, Nt) arrays + times."""

    def __init__(self, variables, times):
        self.variables = variables
        self.times = list(times)

    def __getitem__(self, name):
        return self.variables[name]


def write_particle_file(filepath, variables, times):
    """variables: {"longitude": f32 (N, Nt), "latitude": ..., ["species": int8 (N, Nt)]}."""
    first = next(iter(variables.values()))
    N, Nt = first.shape
    t0 = times[0] if len(times) else datetime(1970, 1, 1)
    secs = np.array([(t - t0).total_seconds() for t in times], dtype=np.float64)
    # (a file without time columns -- a time_step call with start_time == end_time -- has no NetCDF-3 spelling with time as
    #  the second dimension: a zero-length dimension is the record dimension there; it takes the directory layout)
    if Nt > 0 and all(v.nbytes <= _NC3_VAR_LIMIT for v in variables.values()):
        from scipy.io import netcdf_file
        with netcdf_file(filepath, "w", version=2) as nc:
            nc.createDimension("particle number", N)
            nc.createDimension("time", Nt)
            pn = nc.createVariable("particle number", np.dtype("int32"), ("particle number",))
            pn[:] = np.arange(1, N + 1, dtype=np.int32)
            tv = nc.createVariable("time", np.dtype("float64"), ("time",))
            tv[:] = secs
            tv.units = "seconds since " + t0.strftime("%Y-%m-%d %H:%M:%S")
            tv.calendar = "proleptic_gregorian"
            for name, arr in variables.items():
                var = nc.createVariable(name, arr.dtype, ("particle number", "time"))
                var[:] = arr
        return filepath
    d = filepath + ".npz.d"
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, "time_seconds.npy"), secs)
    with open(os.path.join(d, "time_origin.txt"), "w") as f:
        f.write(t0.strftime("%Y-%m-%d %H:%M:%S"))
    for name, arr in variables.items():
        np.save(os.path.join(d, name + ".npy"), arr)
    return d


HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _read_netcdf4(filepath):
    """particle_data.nc / microbe_data.nc as the REFERENCE writes them when netCDF4 is installed next to xarray
    (``xarray.Dataset.to_netcdf``: NetCDF-4 = HDF5; particle_advecter.py:300-305, interaction_simulator.py:119-121).
    SciPy reads NetCDF-3 only, so these go through whichever HDF5 reader the environment has."""
    try:
        import netCDF4
    except ImportError:
        netCDF4 = None
    if netCDF4 is not None:
        with netCDF4.Dataset(filepath, "r") as nc:
            nc.set_auto_mask(False)
            names = [k for k in nc.variables if k not in ("particle number", "time")]
            variables = {k: np.ascontiguousarray(nc.variables[k][:]) for k in names}
            return variables, np.array(nc.variables["time"][:], dtype=np.float64), str(nc.variables["time"].units)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(filepath, "r") as f:
            names = [k for k in f.keys() if k not in ("particle number", "time")]
            variables = {k: np.ascontiguousarray(f[k][...]) for k in names}
            units = f["time"].attrs["units"]
            return variables, np.array(f["time"][...], dtype=np.float64), units.decode() if isinstance(units, bytes) else str(units)
    raise OSError("%s is a NetCDF-4 (HDF5) file -- what xarray writes when netCDF4 is installed -- and neither netCDF4 nor "
                  "h5py is importable here; convert it with `nccopy -k classic` (or xarray's to_netcdf(format="
                  "'NETCDF3_64BIT')), or install one of the two readers" % filepath)


def _parse_time_units(units):
    """'<unit> since <origin>' of a CF time axis -> (seconds per unit, origin)."""
    unit, _, origin = units.partition(" since ")
    scale = {"seconds": 1.0, "second": 1.0, "minutes": 60.0, "hours": 3600.0, "hour": 3600.0, "days": 86400.0, "day": 86400.0}[unit.strip()]
    origin = origin.strip().replace("T", " ")
    for fmt in ("%Y-%m-%d %H:%M:%S.%f", "%Y-%m-%d %H:%M:%S", "%Y-%m-%d %H:%M", "%Y-%m-%d"):
        try:
            return scale, datetime.strptime(origin, fmt)
        except ValueError:
            pass
    raise ValueError("time units not understood: %r" % units)


class ParticleFileWriter:
    """``particle_data.nc`` / ``microbe_data.nc`` written piece by piece.

    The reference allocates the whole (N, Nt) arrays on the host and hands them to xarray (particle_advecter.py:269-270,
    interaction_simulator.py:80-82): 33.8 GB for its own 490,000-microbe, 7,670-step run.  This writer keeps the same
    file for runs that fit NetCDF-3 (arrays in memory, ``write_particle_file`` at ``close()``) and, for larger ones, fills
    the ``<file>.npz.d/`` layout that ``write_particle_file`` falls back to -- memory-mapped ``.npy`` files, written in
    blocks of whole time columns, so the host holds ``block_bytes`` instead of the run.  Cells never written stay 0, as in
    the reference's ``zeros`` arrays."""

    def __init__(self, filepath, variables, N, times, var_limit=None, block_bytes=256 << 20):
        self.filepath, self.N, self.times = filepath, int(N), list(times)
        self.Nt = len(self.times)
        self.dtypes = {k: np.dtype(v) for k, v in variables.items()}
        limit = _NC3_VAR_LIMIT if var_limit is None else var_limit
        self.large = any(self.N * self.Nt * dt.itemsize > limit for dt in self.dtypes.values())
        self._n_buf, self._t0 = 0, 0
        if not self.large:
            self.arrays = {k: np.zeros((self.N, self.Nt), dtype=dt) for k, dt in self.dtypes.items()}
            return
        self.dir = filepath + ".npz.d"
        os.makedirs(self.dir, exist_ok=True)
        self.arrays = {k: np.lib.format.open_memmap(os.path.join(self.dir, k + ".npy"), mode="w+", dtype=dt, shape=(self.N, self.Nt))
                       for k, dt in self.dtypes.items()}
        per_column = self.N * sum(dt.itemsize for dt in self.dtypes.values())
        self.block = int(max(1, min(self.Nt, block_bytes // max(per_column, 1))))
        self._buf = {k: np.zeros((self.N, self.block), dtype=dt) for k, dt in self.dtypes.items()}

    def put(self, i, **columns):
        """Time column ``i`` of the named variables (arrays of N values)."""
        if not self.large:
            for k, v in columns.items():
                self.arrays[k][:, i] = v
            return
        if self._n_buf and (i != self._t0 + self._n_buf or self._n_buf == self.block):
            self._flush()
        if self._n_buf == 0:
            self._t0 = i
            for b in self._buf.values():
                b[:] = 0
        for k, v in columns.items():
            self._buf[k][:, self._n_buf] = v
        self._n_buf += 1

    def put_block(self, rows, t1, t2, **blocks):
        """A sub-block: particles ``rows`` (a slice) x time columns [t1, t2), arrays of shape (rows, t2 - t1)."""
        if self.large:
            self._flush()
        for k, v in blocks.items():
            self.arrays[k][rows, t1:t2] = v

    def _flush(self):
        if self.large and self._n_buf:
            # only the variables that were given columns carry data; the others hold zeros, which is what the file has
            for k, b in self._buf.items():
                self.arrays[k][:, self._t0:self._t0 + self._n_buf] = b[:, :self._n_buf]
            self._n_buf = 0

    def close(self):
        """-> the path written (the ``.nc`` file, or the ``.npz.d`` directory of a large run)."""
        if not self.large:
            return write_particle_file(self.filepath, self.arrays, self.times)
        self._flush()
        t0 = self.times[0] if self.times else datetime(1970, 1, 1)
        np.save(os.path.join(self.dir, "time_seconds.npy"), np.array([(t - t0).total_seconds() for t in self.times], dtype=np.float64))
        with open(os.path.join(self.dir, "time_origin.txt"), "w") as f:
            f.write(t0.strftime("%Y-%m-%d %H:%M:%S"))
        for a in self.arrays.values():
            a.flush()
        self.arrays = self._buf = None
        return self.dir


def read_particle_file(filepath):
    if os.path.isfile(filepath):
        with open(filepath, "rb") as f:
            magic = f.read(8)
        if magic == HDF5_MAGIC:
            variables, secs, units = _read_netcdf4(filepath)
            scale, t0 = _parse_time_units(units)
            return ParticleFile(variables, [t0 + timedelta(seconds=float(s) * scale) for s in secs])
        from scipy.io import netcdf_file
        with netcdf_file(filepath, "r", mmap=False) as nc:
            names = [k for k in nc.variables if k not in ("particle number", "time")]
            # NetCDF-3 stores big-endian; hand back native-endian arrays
            variables = {k: np.ascontiguousarray(nc.variables[k][:]).astype(nc.variables[k][:].dtype.newbyteorder("="))
                         for k in names}
            secs = np.array(nc.variables["time"][:], dtype=np.float64)
            units = nc.variables["time"].units
            units = units.decode() if isinstance(units, bytes) else units
        t0 = datetime.strptime(units.replace("seconds since ", ""), "%Y-%m-%d %H:%M:%S")
    else:
        d = filepath + ".npz.d"
        if not os.path.isdir(d):
            raise FileNotFoundError(filepath)
        secs = np.load(os.path.join(d, "time_seconds.npy"))
        with open(os.path.join(d, "time_origin.txt")) as f:
            t0 = datetime.strptime(f.read().strip(), "%Y-%m-%d %H:%M:%S")
        variables = {fn[:-4]: np.load(os.path.join(d, fn), mmap_mode="r") for fn in sorted(os.listdir(d))
                     if fn.endswith(".npy") and fn != "time_seconds.npy"}
    times = [t0 + timedelta(seconds=float(s)) for s in secs]
    return ParticleFile(variables, times)


def write_png(filepath, rgb):
    """8-bit RGB PNG from a uint8 (height, width, 3) array (zlib + CRCs by hand: matplotlib / PIL are not required)."""
    import struct
    import zlib
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    assert rgb.ndim == 3 and rgb.shape[2] == 3
    h, w, _ = rgb.shape
    raw = np.empty((h, 1 + 3 * w), dtype=np.uint8)
    raw[:, 0] = 0                                       # filter type 0 on every scanline
    raw[:, 1:] = rgb.reshape(h, 3 * w)

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(filepath, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))
    return filepath


class RecordAssembler:
    """The (N, Nt) record of a run, filled one step at a time, written in the layout of the reference's
    ``microbe_data.nc`` (interaction_simulator.py:62-77, :108-110, :119-122).  ``stride`` keeps every stride-th step
    (at 9 B per microbe-step the dense record of 490,000 microbes x 7,670 steps is 33.8 GB, which
    the reference holds in RAM: docs/disk_usage.txt, interaction_simulator.py:62-66).  With ``output_dir`` given up
    front, a record beyond NetCDF-3's variable limit goes straight to memory-mapped files in blocks of time columns
    (ParticleFileWriter) instead of dense host arrays.  ``counts`` = microbes of species 1 / 2 / 3 per kept step."""

    def __init__(self, n_particles, n_steps, start_time, dt, stride=1, output_dir=None, filename="microbe_data.nc"):
        assert stride >= 1 and n_steps >= 0
        self.stride = int(stride)
        self.n_steps = int(n_steps)
        self.kept_steps = list(range(0, self.n_steps, self.stride))
        nt = len(self.kept_steps)
        self.times = [start_time + k * dt for k in self.kept_steps]
        self.counts = np.zeros((nt, 3), dtype=np.int64)
        self.filled = 0
        self._n = int(n_particles)
        self._spec = {"longitude": np.float32, "latitude": np.float32, "species": np.int8}
        self._target = None
        self._writer = None
        if output_dir is not None:
            self._open(output_dir, filename)
        else:                                   # destination not known yet: in memory, whatever the size
            self._writer = ParticleFileWriter(None, self._spec, self._n, self.times, var_limit=float("inf"))

    def _open(self, output_dir, filename):
        os.makedirs(output_dir, exist_ok=True)
        self._target = os.path.join(output_dir, filename)
        self._writer = ParticleFileWriter(self._target, self._spec, self._n, self.times)

    def wants(self, step):
        return 0 <= step < self.n_steps and step % self.stride == 0

    def put(self, step, lon, lat, species):
        assert self.wants(step)
        col = step // self.stride
        self._writer.put(col, longitude=lon, latitude=lat, species=species)
        sp = np.asarray(species)
        self.counts[col] = [int((sp == s).sum()) for s in (1, 2, 3)]
        self.filled += 1

    def write(self, output_dir, filename="microbe_data.nc"):
        assert self.filled == len(self.kept_steps), "%d of %d columns filled" % (self.filled, len(self.kept_steps))
        os.makedirs(output_dir, exist_ok=True)
        target = os.path.join(output_dir, filename)
        if self._target is None:
            self._writer.filepath = target
        else:
            assert os.path.abspath(target) == os.path.abspath(self._target), "the assembler was opened on %s" % self._target
        return self._writer.close()


# ---- the delta-packed position record (csrc/record.cu, lm_record_delta_pack) ------------------------------------------
DELTA_ESCAPE = -32768


def _mono_key(x):
    """float32 array -> int64 keys monotone in the value (negative: ~bits, else bits | 0x80000000): a bijection on all
    2^32 bit patterns, the same map as csrc/record.cu::mono_key."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.int64)


def _from_mono_key(k):
    k = k.astype(np.uint32)
    return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32).view(np.float32)


def unpack_delta_record(prev_lon, prev_lat, dlon, dlat, escapes):
    """Host decoder of ``lm_record_delta_pack``: the float32 (lon, lat) of a step, BIT-EXACT, from the previous step's
    record, the int16 ulp differences and the escape list (uint32 (m, 2): slot = 2 i + (0 lon | 1 lat), raw bits)."""
    out = []
    for prev, d in ((prev_lon, dlon), (prev_lat, dlat)):
        d = np.asarray(d, dtype=np.int16)
        out.append(_from_mono_key(_mono_key(prev) + np.where(d == DELTA_ESCAPE, 0, d.astype(np.int64))))
    esc = np.asarray(escapes, dtype=np.uint32).reshape(-1, 2)
    if esc.shape[0]:
        idx, coord = (esc[:, 0] >> 1).astype(np.int64), esc[:, 0] & 1
        for c in (0, 1):
            sel = coord == c
            out[c].view(np.uint32)[idx[sel]] = esc[sel, 1]
    n_marked = int((np.asarray(dlon) == DELTA_ESCAPE).sum() + (np.asarray(dlat) == DELTA_ESCAPE).sum())
    if n_marked != esc.shape[0]:
        raise ValueError("delta record: %d escape markers but %d escape entries (list overflowed?)" % (n_marked, esc.shape[0]))
    return out[0], out[1]
