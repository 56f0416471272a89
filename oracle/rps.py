"""Rock-paper-scissors oracle (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/interactions.py:13-40 (``rock_paper_scissors_interaction``)
and the sequential in-place pair loop of /root/reference/interaction_simulator.py:104-105:

    for pair in microbe_pairs:
        pair_interaction(parameters, microbe_properties, pair[0], pair[1])

The reference draws ``np.random.rand()`` only when the two species differ
(interactions.py:17-20).  Parity is defined with an injected per-pair stream
``u[k]``: pair k consumes ``u[k]`` iff its species differ at the
moment it is processed -- which is what patching ``np.random.rand`` with
``lambda: u[k]`` does to the unmodified function (tests/golden/make_golden.py).

Pinned by tests/golden/rps_*.npz, produced by the UNMODIFIED reference function.
"""
import ctypes
import os
import subprocess

import numpy as np

ROCK, PAPER, SCISSORS = 1, 2, 3   # interactions.py:5

_HERE = os.path.dirname(os.path.abspath(__file__))


def rps_pair(species, p1, p2, r, pRS, pPR, pSP):
    """One call of interactions.py:13-40 with the random draw ``r`` supplied by the caller.

    Returns True iff the draw was consumed (species differed).
    """
    if species[p1] != species[p2]:                       # :17
        s1, s2 = species[p1], species[p2]                # :18
        winner = None                                    # :22
        if s1 == ROCK and s2 == SCISSORS:                # :24-35
            winner = p1 if r < pRS else p2
        elif s1 == ROCK and s2 == PAPER:
            winner = p2 if r < pPR else p1
        elif s1 == PAPER and s2 == ROCK:
            winner = p1 if r < pPR else p2
        elif s1 == PAPER and s2 == SCISSORS:
            winner = p2 if r < pSP else p1
        elif s1 == SCISSORS and s2 == ROCK:
            winner = p2 if r < pRS else p1
        elif s1 == SCISSORS and s2 == PAPER:
            winner = p1 if r < pSP else p2
        if winner == p1:                                 # :37-40
            species[p2] = species[p1]
        elif winner == p2:
            species[p1] = species[p2]
        return True
    return False


def rps_sequential_py(species, pairs, u, pRS, pPR, pSP):
    """Pure-Python sequential loop (small cases).  Mutates and returns ``species`` (int8)."""
    species = np.asarray(species)
    assert species.dtype == np.int8
    pairs = np.asarray(pairs).reshape(-1, 2)
    u = np.asarray(u, dtype=np.float64)
    draws = 0
    for k in range(pairs.shape[0]):
        draws += rps_pair(species, int(pairs[k, 0]), int(pairs[k, 1]), float(u[k]), pRS, pPR, pSP)
    return species, draws


# ---------------------------------------------------------------------------------------------
# C restatement (same rule, same order) for cases too large for a Python loop.
# ---------------------------------------------------------------------------------------------
_lib = None


def build_c(force=False):
    """Compile oracle/rps_seq.c + oracle/rk4_c.c into oracle/_build/liboracle.so (gcc -O2 -fopenmp)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    srcs = [os.path.join(_HERE, "rps_seq.c"), os.path.join(_HERE, "rk4_c.c")]
    if (not force) and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(out_dir, exist_ok=True)
    # -ffp-contract=off: the restated C must not fuse multiply-adds (Parcels' JIT'd C on
    # x86-64 without -march flags has no FMA); keeps the arithmetic IEEE-reproducible.
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", so] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return so


def load_c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.rps_sequential.restype = ctypes.c_int64
        _lib.rps_sequential.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_double, ctypes.c_double, ctypes.c_double]
    return _lib


def rps_sequential_c(species, pairs, u, pRS, pPR, pSP):
    """C sequential loop.  Mutates and returns ``species`` (int8 C-contiguous) and the draw count."""
    lib = load_c()
    species = np.ascontiguousarray(species)
    assert species.dtype == np.int8
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int64).reshape(-1, 2))
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.shape[0] == pairs.shape[0]
    draws = lib.rps_sequential(species.ctypes.data, pairs.ctypes.data, u.ctypes.data, pairs.shape[0],
                               float(pRS), float(pPR), float(pSP))
    return species, int(draws)


# ---------------------------------------------------------------------------------------------
# Canonical pair order of the fused device path ("cell-phase order").
# ---------------------------------------------------------------------------------------------
def cell_phase_order(pairs, lon32, lat32, grid):
    """Return ``pairs`` (rows i<j, original ids) re-ordered into the device's canonical order.

    grid = dict(x0, y0, inv_h, ncx, ncy) as reported by ``lm_get_grid``.  Every pair joins two
    particles whose cells differ by at most one in each direction.  Order key:

        phase   0: same cell                                   unit = that cell
                1+(cx&1): east neighbour  (cx,cy)-(cx+1,cy)    unit = west cell
                3*(cy&1)+3: north-west    (cx,cy)-(cx-1,cy+1)  unit = south cell
                3*(cy&1)+4: north         (cx,cy)-(cx  ,cy+1)
                3*(cy&1)+5: north-east    (cx,cy)-(cx+1,cy+1)
        then unit (anchor cell key cy*ncx+cx), then id of the particle in the anchor cell,
        then id of the other particle (same cell: smaller id, larger id).

    Units inside one phase touch disjoint particles, so their relative order is immaterial;
    the device runs them concurrently.
    """
    from .pairs import cell_index
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    i, j = pairs[:, 0], pairs[:, 1]
    cx = cell_index(lon32, grid["x0"], grid["inv_h"], grid["ncx"])
    cy = cell_index(lat32, grid["y0"], grid["inv_h"], grid["ncy"])
    cxi, cyi, cxj, cyj = cx[i], cy[i], cx[j], cy[j]
    assert np.all(np.abs(cxi - cxj) <= 1) and np.all(np.abs(cyi - cyj) <= 1), "pair spans non-adjacent cells"
    same = (cxi == cxj) & (cyi == cyj)
    # anchor = i unless j's cell is "before" i's: lower row, or same row and smaller cx
    j_anchor = (cyj < cyi) | ((cyj == cyi) & (cxj < cxi))
    a = np.where(j_anchor, j, i)
    b = np.where(j_anchor, i, j)
    cxa, cya, cxb, cyb = cx[a], cy[a], cx[b], cy[b]
    d = cxb - cxa
    phase = np.where(same, 0,
                     np.where(cya == cyb, 1 + (cxa & 1), 3 * (cya & 1) + 4 + d))
    unit = cya * np.int64(grid["ncx"]) + cxa
    order = np.lexsort((b, a, unit, phase))
    return pairs[order], phase[order]


# ---------------------------------------------------------------------------------------------
# Canonical pair order of the fused tile kernel ("tile-round order", DESIGN.md §4.3, csrc/interact.cu).
# ---------------------------------------------------------------------------------------------
TILE_W, TILE_H = 32, 16      # cells per tile of the device kernel: part of the definition of the order


def cell_ranks(lon32, lat32, grid):
    """Cell coordinates, rank of every particle among the particles of its cell (by id), cell occupancy."""
    from .pairs import cell_index
    n = np.asarray(lon32).shape[0]
    cx = cell_index(lon32, grid["x0"], grid["inv_h"], grid["ncx"]).astype(np.int64)
    cy = cell_index(lat32, grid["y0"], grid["inv_h"], grid["ncy"]).astype(np.int64)
    key = cy * np.int64(grid["ncx"]) + cx
    order = np.lexsort((np.arange(n), key))                    # storage order of the device: (cell, id)
    sk = key[order]
    pos = np.arange(n, dtype=np.int64)
    first = np.r_[True, sk[1:] != sk[:-1]] if n else np.zeros(0, dtype=bool)
    run_start = np.maximum.accumulate(np.where(first, pos, 0)) if n else pos
    rank = np.empty(n, dtype=np.int64)
    rank[order] = pos - run_start
    occ = np.bincount(key, minlength=int(grid["ncx"]) * int(grid["ncy"]))[key] if n else pos
    return cx, cy, rank, occ.astype(np.int64)


def unit_is_light(same, ma, mb, tile=True, heavy_min=1024):
    """Part of the definition of the orders.  Fused tile kernel (csrc/interact.cu::unit_is_light): a unit of two cells is
    LIGHT when m_a * m_b <= 256 and m_b <= 64, a unit of one cell when m <= 23; otherwise it is HEAVY.  Hybrid path
    (csrc/pairs.cu, ``heavy_dirs``): LIGHT when its candidate pairs -- m_a * m_b, one cell: m (m - 1) / 2 -- do not exceed
    ``heavy_min`` = 1,024 (LM_OPT_HEAVY_MIN)."""
    if tile:
        return np.where(same, ma <= 23, (ma * mb <= 256) & (mb <= 64))
    return np.where(same, ma * (ma - 1) // 2 <= heavy_min, ma * mb <= heavy_min)


def tile_round_order(pairs, lon32, lat32, grid, tile=(TILE_W, TILE_H)):
    """Return ``pairs`` (rows i<j, original ids) re-ordered into the canonical order of the fused tile kernel.

    Cells are grouped into tiles of ``tile`` = (32, 16) cells, origin at cell (0, 0) of the grid.  A *unit* is one
    cell (the pairs inside it) or two adjacent cells (half stencil: E, NW, N, NE of the anchor cell).  Order key
    (phase, unit, then inside the unit):

      phase  0..8   units inside one tile: 0 same cell | 1 + (cx & 1) east | 3 + (cy & 1) north-west |
                    5 + (cy & 1) north | 7 + (cy & 1) north-east
             9..14  units across a tile boundary: 9 east | 10, 11 north-west, north-east across a vertical
                    boundary only | 12, 13, 14 north-west, north, north-east across a horizontal boundary
      unit   anchor cell (the western / southern one)
      inside a LIGHT unit (``unit_is_light``): (rank in the anchor cell, rank in the other cell) lexicographic, ranks by
             particle id (one cell: smaller rank, larger rank)
      inside a HEAVY unit: ROUNDS OF MATCHINGS -- inside a round no microbe occurs twice, so a round is order-free
             and the device resolves it in parallel:
             two cells, m_a and m_b microbes, M = max(m_a, m_b): round k in [0, M) pairs rank i of the anchor cell
                    with rank (i + k) mod M of the other cell; slot = i
             one cell, m microbes, M = m rounded up to even (rank M - 1 is a phantom when m is odd): the circle
                    method of round-robin tournaments -- round k in [0, M - 1) pairs rank M - 1 with rank k
                    (slot 0) and rank (k + j) mod (M - 1) with rank (k - j) mod (M - 1) for j in [1, M / 2) (slot j)

    Units of one phase touch disjoint microbes; their relative order is immaterial.
    """
    return _round_order(pairs, lon32, lat32, grid, tile)


def cell_round_order(pairs, lon32, lat32, grid, heavy_min=1024):
    """The canonical order of the HYBRID device path (LM_OPT_INTERACT_MODE = 2, the default): the nine phases of
    ``cell_phase_order`` (0 same cell | 1 + (cx & 1) east | 3 (cy & 1) + 3, + 4, + 5 north-west, north, north-east), units
    by anchor cell, and inside a unit the rule of ``tile_round_order``: LIGHT units (``unit_is_light`` with tile=False) in (rank in the
    anchor cell, rank in the other cell) lexicographic order -- which is ``cell_phase_order``'s (id_a, id_b) -- and HEAVY
    units in rounds of matchings.  The round-1 pipeline resolves the light units, a device-wide queue of heavy units is
    resolved round by round by whole warps / CTAs (csrc/interact.cu::interact_heavy_kernel)."""
    return _round_order(pairs, lon32, lat32, grid, None, heavy_min)


def _round_order(pairs, lon32, lat32, grid, tile, heavy_min=1024):
    tw, th = tile if tile is not None else (1 << 30, 1 << 30)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    cx, cy, rank, occ = cell_ranks(lon32, lat32, grid)
    i, j = pairs[:, 0], pairs[:, 1]
    cxi, cyi, cxj, cyj = cx[i], cy[i], cx[j], cy[j]
    assert np.all(np.abs(cxi - cxj) <= 1) and np.all(np.abs(cyi - cyj) <= 1), "pair spans non-adjacent cells"
    same = (cxi == cxj) & (cyi == cyj)
    j_anchor = (cyj < cyi) | ((cyj == cyi) & (cxj < cxi))
    a = np.where(j_anchor, j, i)
    b = np.where(j_anchor, i, j)
    cxa, cya, cxb, cyb = cx[a], cy[a], cx[b], cy[b]
    d = cxb - cxa
    if tile is not None:
        inner = np.where(same, 0, np.where(cya == cyb, 1 + (cxa & 1), 5 + 2 * d + (cya & 1)))
    else:
        inner = np.where(same, 0, np.where(cya == cyb, 1 + (cxa & 1), 3 * (cya & 1) + 4 + d))
    cross_v = (cxa // tw) != (cxb // tw)
    cross_h = (cya // th) != (cyb // th)
    outer = np.where(cya == cyb, 9, np.where(cross_h, 13 + d, np.where(d < 0, 10, 11)))
    phase = np.where(cross_v | cross_h, outer, inner)
    unit = cya * np.int64(grid["ncx"]) + cxa
    ra, rb, ma, mb = rank[a], rank[b], occ[a], occ[b]
    light = unit_is_light(same, ma, mb, tile is not None, heavy_min)
    # heavy units: rounds and slots
    big = np.maximum(ma, mb)
    rnd_x = np.mod(rb - ra, big)
    slot_x = ra
    p, q = np.minimum(ra, rb), np.maximum(ra, rb)
    m_even = ma + (ma & 1)
    n1 = np.maximum(m_even - 1, 1)
    fixed = q == m_even - 1                                    # the player that stays put (only real when m is even)
    rnd_s = np.where(fixed, p, np.mod((p + q) * (m_even // 2), n1))
    jp = np.mod(p - rnd_s, n1)
    jq = np.mod(q - rnd_s, n1)
    slot_s = np.where(fixed, 0, np.where((jp >= 1) & (jp < m_even // 2), jp, jq))
    k1 = np.where(light, np.where(same, p, ra), np.where(same, rnd_s, rnd_x))
    k2 = np.where(light, np.where(same, q, rb), np.where(same, slot_s, slot_x))
    order = np.lexsort((k2, k1, unit, phase))
    return pairs[order], phase[order]


def canonical_order(pairs, lon32, lat32, grid, mode=2):
    """The device's canonical pair order: LM_OPT_INTERACT_MODE 2 (default, hybrid) = cell-round order, 1 (fused tile
    kernel) = tile-round order, 0 (round-1 pipeline) = cell-phase order."""
    if mode == 2:
        return cell_round_order(pairs, lon32, lat32, grid)
    if mode == 1:
        return tile_round_order(pairs, lon32, lat32, grid)
    return cell_phase_order(pairs, lon32, lat32, grid)


# ---------------------------------------------------------------------------------------------
# The reference's pair function at the reference's COST (bench.py --impl reference times it).
# ---------------------------------------------------------------------------------------------
def reference_pair_interaction(parameters, microbe_properties, p1, p2):
    """interactions.py:13-40 restated with the reference's own signature and work per call: the species array and the
    three probabilities are looked up in the dicts on every call and the draw is ``np.random.rand()`` (NumPy's global
    MT19937), taken only when the species differ -- so a loop over it costs what the reference's loop
    (interaction_simulator.py:104-105) costs, which ``rps_pair`` above (draw injected, probabilities as arguments)
    would understate.  /root/reference cannot travel to the GPU box, hence a restatement; tests/test_oracle_rps.py
    checks it call by call against the unmodified function in this container."""
    species = microbe_properties["species"]
    pRS, pPR, pSP = parameters["pRS"], parameters["pPR"], parameters["pSP"]
    if species[p1] != species[p2]:
        s1, s2 = species[p1], species[p2]
        r = np.random.rand()
        winner = None
        if s1 == ROCK and s2 == SCISSORS:
            winner = p1 if r < pRS else p2
        elif s1 == ROCK and s2 == PAPER:
            winner = p2 if r < pPR else p1
        elif s1 == PAPER and s2 == ROCK:
            winner = p1 if r < pPR else p2
        elif s1 == PAPER and s2 == SCISSORS:
            winner = p2 if r < pSP else p1
        elif s1 == SCISSORS and s2 == ROCK:
            winner = p2 if r < pRS else p1
        elif s1 == SCISSORS and s2 == PAPER:
            winner = p1 if r < pSP else p2
        if winner == p1:
            species[p2] = species[p1]
        elif winner == p2:
            species[p1] = species[p2]
