"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's snapshot analyses.

Only tests/ may import this module; the product (lagrangian_microbes_b200/) never does.

Pair-distance histogram: /root/reference/sandbox/pairwise_distance_histogram_distributed.jl
    haversine_distance32                      :33-44
    pairwise_distance_histogram_1point        :54-64   bin = round(Int8, 10f0 * log10(max(1, d))); sub_hist[bin] += 1
    (plot_pairwise_histogram calls it once per species on that species' microbes, :127-134, with 70 bins, :180)
PARITY UNPINNED: the reference file is Julia, there is no Julia here, and the reference holds no output of it.  The
restatement below follows the source line by line in float32 NumPy; Julia's sinpi / cospi are replaced by a float64
evaluation rounded to float32 (Julia's are within 1 ulp of that).  Because libm differences of 1 ulp move pairs
that sit within float32 rounding of a bin edge, parity of a histogram is stated as BOUNDS: `pdh_bounds` evaluates every
pair in float64 and returns, per bin, how many pairs are in it beyond doubt and how many could be.

Frame: /root/reference/microbe_plotter.py:132-146 -- one plt.scatter over all microbes, colours by species, markers
drawn in array order.  `raster_reference` restates the raster lm_rasterize documents (include/lm_b200.h) with exact
integer results.
"""
import numpy as np

R32 = np.float32(6371.228e3)      # pairwise_distance_histogram_distributed.jl:103


def _sinpi32(x):
    """sinpi for float32 x with |x| <= 2, evaluated in float64 (exact zeros at integers, like Julia's)."""
    x = np.asarray(x, dtype=np.float64)
    r = x - 2.0 * np.round(x / 2.0)                    # [-1, 1]
    r = np.where(r > 0.5, 1.0 - r, np.where(r < -0.5, -1.0 - r, r))     # sin(pi r) = sin(pi (1 - r))
    return np.sin(np.pi * r).astype(np.float32)


def _cospi32(x):
    x = np.asarray(x, dtype=np.float64)
    return _sinpi32_f64(0.5 - np.abs(x)).astype(np.float32)


def _sinpi32_f64(x):
    r = x - 2.0 * np.round(x / 2.0)
    r = np.where(r > 0.5, 1.0 - r, np.where(r < -0.5, -1.0 - r, r))
    return np.sin(np.pi * r)


def haversine_distance32(lat1, lon1, lat2, lon2, radius=R32):
    """pairwise_distance_histogram_distributed.jl:33-44, float32 throughout."""
    f = np.float32
    lat1, lon1, lat2, lon2 = (np.asarray(v, dtype=f) for v in (lat1, lon1, lat2, lon2))
    c1 = _cospi32(lat1 / f(180.0))
    c2 = _cospi32(lat2 / f(180.0))
    dlat = lat2 - lat1
    dlon = lon2 - lon1
    d1 = _sinpi32(dlat / f(360.0))
    d2 = _sinpi32(dlon / f(360.0))
    t = d2 * d2 * c1 * c2
    a = d1 * d1 + t
    c = f(2.0) * np.arcsin(np.minimum(f(1.0), np.sqrt(a)))
    return f(radius) * c


def pair_distance_hist_reference(lats, lons, bins=70, radius=R32):
    """:54-64 summed over i (what `sum(subhists)` does, :136-138).  Returns int64[bins + 2]: [b] for b = 0..bins
    (b = 0 is what the reference's 1-based `sub_hist[bin]` cannot hold), [bins + 1] = pairs beyond the last bin."""
    lats = np.asarray(lats, dtype=np.float32)
    lons = np.asarray(lons, dtype=np.float32)
    n = lats.size
    hist = np.zeros(bins + 2, dtype=np.int64)
    for i in range(n - 1):
        d = haversine_distance32(lats[i], lons[i], lats[i + 1:], lons[i + 1:], radius)
        b = np.rint(np.float32(10.0) * np.log10(np.maximum(np.float32(1.0), d))).astype(np.int64)   # round: ties to even
        hist += np.bincount(np.minimum(b, bins + 1), minlength=bins + 2)
    return hist


def pdh_bounds(lats, lons, bins=70, radius=float(R32), tol=5e-5):
    """(lower, upper) int64[bins + 2]: float64 haversine of the float32 coordinates; a pair whose 10 log10(d) is within
    `tol` of a bin edge k + 1/2 counts for the lower bound of neither neighbour and for the upper bound of both.
    tol = 5e-5: a float32 value near 70 carries 4e-6 of rounding by itself, the float32 haversine argument about ten
    ulps (1.3e-6 in 10 log10 d)."""
    lats = np.asarray(lats, dtype=np.float32).astype(np.float64)
    lons = np.asarray(lons, dtype=np.float32).astype(np.float64)
    n = lats.size
    lower = np.zeros(bins + 2, dtype=np.int64)
    upper = np.zeros(bins + 2, dtype=np.int64)
    cl = np.cos(np.deg2rad(lats))
    for i in range(n - 1):
        s1 = _sinpi32_f64((lats[i + 1:] - lats[i]) / 360.0)
        s2 = _sinpi32_f64((lons[i + 1:] - lons[i]) / 360.0)
        a = s1 * s1 + cl[i] * cl[i + 1:] * s2 * s2
        d = 2.0 * radius * np.arcsin(np.minimum(1.0, np.sqrt(a)))
        y = 10.0 * np.log10(np.maximum(1.0, d))
        b = np.floor(y + 0.5).astype(np.int64)
        frac = y + 0.5 - b                                   # in [0, 1): distance above the bin's lower edge
        near_lo = (frac < tol) & (d > 1.0)                   # could belong to b - 1   (y = 0 exactly: max(1, d) clamps)
        near_hi = frac > 1.0 - tol                           # could belong to b + 1
        sure = ~(near_lo | near_hi)
        bc = np.minimum(b, bins + 1)
        lower += np.bincount(bc[sure], minlength=bins + 2)
        upper += np.bincount(bc, minlength=bins + 2)
        upper += np.bincount(np.minimum(np.maximum(b[near_lo] - 1, 0), bins + 1), minlength=bins + 2)
        upper += np.bincount(np.minimum(b[near_hi] + 1, bins + 1), minlength=bins + 2)
    return lower, upper


def raster_reference(lon, lat, species, lon_min, lon_max, lat_min, lat_max, width, height):
    """(counts uint32[3][height][width], top int32[height][width]) exactly as include/lm_b200.h::lm_rasterize states."""
    lon = np.asarray(lon, dtype=np.float32).astype(np.float64)
    lat = np.asarray(lat, dtype=np.float32).astype(np.float64)
    n = lon.size
    species = np.ones(n, dtype=np.int8) if species is None else np.asarray(species, dtype=np.int8)
    sx = float(width) / (lon_max - lon_min)
    sy = float(height) / (lat_max - lat_min)
    fx = np.floor((lon - lon_min) * sx)
    fy = np.floor((lat - lat_min) * sy)
    inside = (fx >= 0) & (fx < width) & (fy >= 0) & (fy < height)
    idx = np.nonzero(inside)[0]
    col = fx[idx].astype(np.int64)
    row = height - 1 - fy[idx].astype(np.int64)
    pix = row * width + col
    counts = np.zeros((3, height * width), dtype=np.uint32)
    for s in (1, 2, 3):
        sel = species[idx] == s
        counts[s - 1] = np.bincount(pix[sel], minlength=height * width).astype(np.uint32)
    top = np.full(height * width, -1, dtype=np.int32)
    np.maximum.at(top, pix, idx.astype(np.int32))
    return counts.reshape(3, height, width), top.reshape(height, width)


def compose_reference(counts, top, species, palette, mode):
    """uint8[height][width][3]; palette uint8[4][3] (background, rock, paper, scissors); mode 0 last drawn, 1 plurality."""
    palette = np.asarray(palette, dtype=np.uint8)
    h, w = top.shape
    if mode == 0:
        s = np.zeros((h, w), dtype=np.int64)
        has = top >= 0
        sp = np.ones(int(top.max()) + 1 if has.any() else 1, dtype=np.int8) if species is None else np.asarray(species)
        v = sp[top[has]].astype(np.int64)
        v[(v < 1) | (v > 3)] = 0
        s[has] = v
    else:
        c1, c2, c3 = (counts[k].astype(np.int64) for k in range(3))
        s = np.where((c1 >= c2) & (c1 >= c3), 1, np.where(c2 >= c3, 2, 3))
        s[(c1 + c2 + c3) == 0] = 0
    return palette[s]
