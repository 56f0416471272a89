"""NumPy emulation of the pair-search -> resolver hand-off (csrc/pairs.cu: hits / rec / rec2) and line-by-line
transliterations of the resolvers that consume it.  Test infrastructure: the CUDA kernels cannot run without a GPU, their
index arithmetic can -- this is how the tiled resolver's logic was checked before its first hardware run.

Layout restated from csrc/pairs.cu (header comment + find_pairs_kernel):
    storage order     particles sorted by (cell key = cy * ncx + cx, id); cell_start = exclusive scan of cell occupancy
    segment           the particles of one cell inside one 32-particle chunk of the storage order
    entry             b_rel (24 bits, relative to the partner cell's start) | a_rel << 24 (anchor index inside the segment)
                      | decision bits << 29 (bit k: u < p_k for pRS, pPR, pSP)
    streams           the entries of one (segment, direction) are contiguous, sorted by (anchor, partner);
                      directions: 0 same cell, 1 E, 2 NW, 3 N, 4 NE
    rec[d][cell]      (first entry, count) of the cell's FIRST segment;  rec2[d][chunk] of a segment continuing a cell at
                      the start of a chunk.  Records of empty cells / unused chunks are stale (garbage here).
"""
import numpy as np

from oracle.pairs import cell_index

B_REL_MASK = (1 << 24) - 1
STALE = (123456789, 77)


def rps_apply(s1, s2, dec):
    """csrc/pairs.cu::rps_apply -- species both microbes end up with (s1 != s2, both in 1..3)."""
    d = (s1 - s2) % 3
    w, l = (s1, s2) if d == 1 else (s2, s1)
    return w if (dec >> (w - 1)) & 1 else l


def build(lon, lat, pairs, u, p, grid):
    """Storage order + hand-off for the pair set `pairs` (original ids, i < j) with per-pair draws u (aligned with pairs)."""
    ncx, ncy = grid["ncx"], grid["ncy"]
    n = lon.size
    cx = cell_index(lon, grid["x0"], grid["inv_h"], ncx)
    cy = cell_index(lat, grid["y0"], grid["inv_h"], ncy)
    key = cy.astype(np.int64) * ncx + cx
    ids = np.lexsort((np.arange(n), key)).astype(np.int64)         # storage slot -> original id
    slot = np.empty(n, dtype=np.int64)
    slot[ids] = np.arange(n)
    cell_start = np.zeros(ncx * ncy + 1, dtype=np.int64)
    np.cumsum(np.bincount(key, minlength=ncx * ncy), out=cell_start[1:])
    skey = key[ids]                                                  # cell of each storage slot
    # anchor / partner in storage slots, direction
    i, j = slot[pairs[:, 0]], slot[pairs[:, 1]]
    ci, cj = skey[i], skey[j]
    j_anchor = (cj < ci) | ((cj == ci) & (j < i))                    # lower row, or same row and lower column; same cell: lower slot
    a = np.where(j_anchor, j, i)
    b = np.where(j_anchor, i, j)
    ca, cb = skey[a], skey[b]
    dx, dy = (cb % ncx) - (ca % ncx), (cb // ncx) - (ca // ncx)
    d = np.where(dy == 0, np.where(dx == 0, 0, 1), 3 + dx)
    assert np.all((dy == 0) | (dy == 1)) and np.all(np.abs(dx) <= 1) and np.all((dy == 1) | (dx >= 0))
    dec = (u < p[0]).astype(np.int64) | ((u < p[1]).astype(np.int64) << 1) | ((u < p[2]).astype(np.int64) << 2)
    seg_first = np.maximum(cell_start[ca], (a >> 5) << 5)
    entry = (b - cell_start[cb]) | ((a - seg_first) << 24) | (dec << 29)
    assert np.all(b - cell_start[cb] <= B_REL_MASK) and np.all(a - seg_first < 32)
    # streams: arbitrary placement in hits[] (here: by chunk block of 8 warps, direction-major like the kernel, then segment)
    order = np.lexsort((b, a, seg_first, d, a >> 8))
    hits = entry[order].astype(np.int64)
    a_o, d_o, sf_o, ca_o = a[order], d[order], seg_first[order], ca[order]
    rec = np.empty((5, ncx * ncy, 2), dtype=np.int64); rec[:] = STALE
    rec2 = np.empty((5, n // 32 + 2, 2), dtype=np.int64); rec2[:] = STALE
    # every segment of a non-empty cell gets a record in every direction (count 0 when it has no pairs)
    for c in np.nonzero(np.diff(cell_start))[0]:
        s0 = cell_start[c]
        while s0 < cell_start[c + 1]:
            for dd in range(5):
                (rec[dd, c] if s0 == cell_start[c] else rec2[dd, s0 >> 5])[:] = (0, 0)
            s0 = (s0 | 31) + 1
    if hits.size:
        brk = np.nonzero((np.diff(sf_o) != 0) | (np.diff(d_o) != 0) | (np.diff(a_o >> 8) != 0))[0] + 1
        starts = np.concatenate(([0], brk)); ends = np.concatenate((brk, [hits.size]))
        for s, e in zip(starts, ends):
            tgt = rec[d_o[s], ca_o[s]] if sf_o[s] == cell_start[ca_o[s]] else rec2[d_o[s], sf_o[s] >> 5]
            tgt[:] = (s, e - s)
    return dict(ids=ids, cell_start=cell_start, hits=hits, rec=rec, rec2=rec2, ncx=ncx, ncy=ncy)


def phase_geom(ph):
    if ph == 0:
        return 0, 0, 0, 0                       # mode SAME
    if ph <= 2:
        return 1, ph - 1, 0, 1                  # mode EAST, parity
    return 2, (ph - 3) // 3, (ph - 3) % 3 - 1, 3 + (ph - 3) % 3 - 1


def _walk_unit(H, d_idx, cell, ob, spA, offA, spB, offB):
    """One lane walking one unit: resolve_tiled_kernel's inner loop (and resolve_phase_kernel's PB == 1 walk)."""
    cs0, cs1 = H["cell_start"][cell], H["cell_start"][cell + 1]
    first, count = H["rec"][d_idx, cell]
    cur_a, sa, sa0, a0 = -1, 0, 0, cs0
    while True:
        for k in range(first, first + count):
            en = int(H["hits"][k])
            a, b = a0 + ((en >> 24) & 31), ob + (en & B_REL_MASK)
            if a != cur_a:
                if cur_a >= 0 and sa != sa0:
                    spA[offA + cur_a] = sa
                cur_a = a
                sa = sa0 = int(spA[offA + a])
            sb = int(spB[offB + b])
            if sa != sb and 1 <= sa <= 3 and 1 <= sb <= 3:
                sa = rps_apply(sa, sb, en >> 29)
                spB[offB + b] = sa
        a0 = (a0 | 31) + 1
        if a0 >= cs1:
            break
        first, count = H["rec2"][d_idx, a0 >> 5]
    if cur_a >= 0 and sa != sa0:
        spA[offA + cur_a] = sa


def resolve_phases(H, sp, first, last, rows_owned=None, rows_local=None):
    """launch_resolve_phases: every unit of every phase on the live species (storage order), in place."""
    ncx = H["ncx"]
    rows_local = H["ncy"] if rows_local is None else rows_local
    rows_owned = rows_local if rows_owned is None else rows_owned
    for ph in range(first, last + 1):
        mode, parity, dr, d_idx = phase_geom(ph)
        for cy in range(rows_local):
            for cx in range(ncx):
                oy, ox = cy, cx
                if mode == 0:
                    on = cy < rows_owned
                elif mode == 1:
                    ox = cx + 1
                    on = cy < rows_owned and (cx & 1) == parity and ox < ncx
                else:
                    oy, ox = cy + 1, cx + dr
                    on = (cy & 1) == parity and oy < rows_local and 0 <= ox < ncx
                cell = cy * ncx + cx
                if not on or H["cell_start"][cell + 1] <= H["cell_start"][cell]:
                    continue
                _walk_unit(H, d_idx, cell, H["cell_start"][oy * ncx + ox], sp, 0, sp, 0)
    return sp


def resolve_tiled(H, sp, first, last, tile_x=64, tile_y=16, hx=6, hy=2, rows_owned=None, rows_local=None):
    """resolve_tiled_kernel, CTA by CTA: private species of the loaded region (indexed particle + delta[row]), every
    unit inside the loaded region phase by phase, interior written back.  Returns the new species (storage order)."""
    ncx, cell_start = H["ncx"], H["cell_start"]
    rows_local = H["ncy"] if rows_local is None else rows_local
    rows_owned = rows_local if rows_owned is None else rows_owned
    sp_in, sp_out = sp.copy(), sp.copy()
    tiles_x = (ncx + tile_x - 1) // tile_x
    tiles_y = (rows_local + tile_y - 1) // tile_y
    for blk in range(tiles_x * tiles_y):
        tx, ty = blk % tiles_x, blk // tiles_x
        ix0, ix1 = tx * tile_x, min(ncx, tx * tile_x + tile_x)
        iy0, iy1 = ty * tile_y, min(rows_local, ty * tile_y + tile_y)
        lx0, lx1 = max(0, ix0 - hx), min(ncx, ix1 + hx)
        ly0, ly1 = max(0, iy0 - hy), min(rows_local, iy1 + hy)
        lw, lh = lx1 - lx0, ly1 - ly0
        p0 = [int(cell_start[(ly0 + t) * ncx + lx0]) for t in range(lh)]
        cnt = [int(cell_start[(ly0 + t) * ncx + lx1]) - p0[t] for t in range(lh)]
        delta, off = [], 0
        for t in range(lh):
            delta.append(off - p0[t]); off += cnt[t]
        tsp = np.full(off, -99, dtype=np.int64)
        for t in range(lh):
            tsp[delta[t] + p0[t]: delta[t] + p0[t] + cnt[t]] = sp_in[p0[t]: p0[t] + cnt[t]]
        for ph in range(first, last + 1):
            mode, parity, dr, d_idx = phase_geom(ph)
            for uu in range(lw * lh):
                ry = uu // lw
                cy, cx = ly0 + ry, lx0 + (uu - ry * lw)
                oy, ox = cy, cx
                if mode == 0:
                    on = cy < rows_owned
                elif mode == 1:
                    ox = cx + 1
                    on = cy < rows_owned and (cx & 1) == parity and ox < lx1
                else:
                    oy, ox = cy + 1, cx + dr
                    on = (cy & 1) == parity and oy < ly1 and lx0 <= ox < lx1
                if not on:
                    continue
                cell = cy * ncx + cx
                if cell_start[cell + 1] <= cell_start[cell]:
                    continue
                _walk_unit(H, d_idx, cell, int(cell_start[oy * ncx + ox]), tsp, delta[ry], tsp, delta[oy - ly0])
        for y in range(iy0, iy1):
            q0, q1 = int(cell_start[y * ncx + ix0]), int(cell_start[y * ncx + ix1])
            sp_out[q0:q1] = tsp[delta[y - ly0] + q0: delta[y - ly0] + q1]
    return sp_out
