"""Radius pair-search oracle (TEST INFRASTRUCTURE ONLY).

Reference call site (/root/reference/interaction_simulator.py:88-98):

    microbe_locations = stack((lon[:, i], lat[:, i]), axis=-1)      # float32 (N, 2)
    kdt = cKDTree(np.array(microbe_locations))
    microbe_pairs = kdt.query_pairs(r=interaction_radius, p=interaction_norm)

``query_pairs_reference`` is that call verbatim (SciPy is importable here and on the
GPU box, so this is the reference's own third-party implementation, not a port).
``query_pairs_bruteforce`` restates the predicate SciPy applies for p=2:

    x, y  = float32 positions widened to float64
    s     = fl64(dx*dx); s = fl64(s + fl64(dy*dy));   pair  <=>  s <= fl64(r*r)

(inclusive, i < j, coincident points are pairs) on a uniform cell grid; it is
checked against cKDTree in tests/test_oracle_pairs.py.

``interaction_norm`` (interaction_simulator.py:27) is handed to SciPy as ``p``.  SciPy 1.2.1 .. 1.18
(ckdtree/src/rectangle.h, query_pairs.cxx) keeps distances as d**p: the bound is ``r*r`` for p=2,
``pow(r, 1) = r`` for p=1, ``r`` for p=inf, and the point-to-point value is accumulated from 0 as
``+= fabs(d)`` (p=1) / ``fmax(., fabs(d))`` (p=inf).  Restated here for p in {1, 2, inf}; other p use pow()
and are outside the device's contract.
"""
import numpy as np


def stack_locations(lon32, lat32):
    """interaction_simulator.py:88-89 -- float32 (N, 2), x = lon, y = lat."""
    lon32 = np.asarray(lon32, dtype=np.float32)
    lat32 = np.asarray(lat32, dtype=np.float32)
    return np.stack((lon32, lat32), axis=-1)


def query_pairs_reference(lon32, lat32, r, p=2):
    """Exactly the reference's library calls; returns the Python ``set`` it iterates over."""
    from scipy.spatial import cKDTree
    kdt = cKDTree(np.array(stack_locations(lon32, lat32)))
    return kdt.query_pairs(r=r, p=p)


def query_pairs_reference_array(lon32, lat32, r, p=2):
    """Same tree/query, ndarray output, sorted lexicographically -- for large cases."""
    from scipy.spatial import cKDTree
    kdt = cKDTree(np.array(stack_locations(lon32, lat32)))
    a = kdt.query_pairs(r=r, p=p, output_type="ndarray")
    return sort_pairs(a)


def sort_pairs(pairs):
    """(P,2) integer array -> int64, each row (min,max), rows sorted lexicographically."""
    a = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if a.shape[0] == 0:
        return a
    lo = np.minimum(a[:, 0], a[:, 1])
    hi = np.maximum(a[:, 0], a[:, 1])
    order = np.lexsort((hi, lo))
    return np.stack((lo[order], hi[order]), axis=-1)


def pairs_from_set(pair_set):
    return sort_pairs(np.array(sorted(pair_set), dtype=np.int64).reshape(-1, 2))


def within_radius(x_a, y_a, x_b, y_b, r, p=2):
    """The exact fp64 predicate (see module docstring).  Inputs float32-valued arrays; p in {1, 2, inf}."""
    dx = x_a.astype(np.float64) - x_b.astype(np.float64)
    dy = y_a.astype(np.float64) - y_b.astype(np.float64)
    if p == 2:
        s = dx * dx
        s = s + dy * dy
        return s <= np.float64(r) * np.float64(r)
    if p == 1:
        return np.abs(dx) + np.abs(dy) <= np.float64(r)
    if p == np.inf:
        return np.maximum(np.abs(dx), np.abs(dy)) <= np.float64(r)
    raise ValueError("p must be 1, 2 or inf")


def cell_index(v32, origin, inv_h, ncell):
    """Cell coordinate used by the device binning kernel (csrc/bin.cu: cell_coord):

        c = floor((double(v) - origin) * inv_h), clamped to [0, ncell-1]

    Same IEEE operations in the same order on both sides, so the oracle reproduces the
    device's cell assignment bit for bit.
    """
    q = np.floor((np.asarray(v32, dtype=np.float32).astype(np.float64) - np.float64(origin)) * np.float64(inv_h))
    q = np.clip(q, 0, ncell - 1)
    return q.astype(np.int64)


def query_pairs_bruteforce(lon32, lat32, r, p=2):
    """All pairs within r (Minkowski p in {1, 2, inf}) via a cell grid + the exact predicate; sorted (P,2) int64.
    Any p >= 1 bounds |dx| and |dy| by r, so cells of edge >= r and the half stencil serve every norm."""
    lon32 = np.asarray(lon32, dtype=np.float32)
    lat32 = np.asarray(lat32, dtype=np.float32)
    n = lon32.size
    if n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    h = float(r) * (1.0 + 1e-6) if r > 0 else 1.0
    x0, y0 = float(lon32.min()), float(lat32.min())
    cx = np.floor((lon32.astype(np.float64) - x0) / h).astype(np.int64)
    cy = np.floor((lat32.astype(np.float64) - y0) / h).astype(np.int64)
    ncx = int(cx.max()) + 1
    key = cy * ncx + cx
    order = np.argsort(key, kind="stable")
    skey = key[order]
    out = []
    # half stencil: same cell, E, NW, N, NE
    for dxc, dyc in ((0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)):
        ncx_ok = (cx + dxc >= 0) & (cx + dxc < ncx)
        nkey = (cy + dyc) * ncx + (cx + dxc)
        lo = np.searchsorted(skey, nkey, side="left")
        hi = np.searchsorted(skey, nkey, side="right")
        cnt = np.where(ncx_ok, hi - lo, 0)
        tot = int(cnt.sum())
        if tot == 0:
            continue
        a = np.repeat(np.arange(n), cnt)
        start = np.repeat(lo, cnt)
        off = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        b = order[start + off]
        if dxc == 0 and dyc == 0:
            keep = a < b
            a, b = a[keep], b[keep]
        ok = within_radius(lon32[a], lat32[a], lon32[b], lat32[b], r, p)
        out.append(np.stack((a[ok], b[ok]), axis=-1))
    if not out:
        return np.zeros((0, 2), dtype=np.int64)
    return sort_pairs(np.concatenate(out, axis=0))
