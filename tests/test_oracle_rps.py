"""RPS oracle: restatements (Python + C) against goldens made by the UNMODIFIED reference function."""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs
from oracle import philox
from oracle import rps as orps

CASES = ["rps_uniform", "rps_clustered", "rps_oddspecies", "rps_knots"]


def _grid(g):
    return dict(x0=float(g["grid"][0]), y0=float(g["grid"][1]), inv_h=float(g["grid"][2]),
                ncx=int(g["grid_n"][0]), ncy=int(g["grid_n"][1]))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("impl", ["py", "c"])
def test_restatement_matches_reference_in_reference_order(name, impl):
    g = golden(name + ".npz")
    fn = orps.rps_sequential_py if impl == "py" else orps.rps_sequential_c
    sp, draws = fn(g["species0"].copy(), g["pairs_ref_order"], g["u_ref"], float(g["pRS"]), float(g["pPR"]), float(g["pSP"]))
    assert np.array_equal(sp, g["species_ref"])
    assert 0 < draws <= g["pairs_ref_order"].shape[0]


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_in_cell_phase_order(name):
    g = golden(name + ".npz")
    sp, _ = orps.rps_sequential_c(g["species0"].copy(), g["pairs_cell_order"], philox.pair_uniforms(
        g["pairs_cell_order"][:, 0], g["pairs_cell_order"][:, 1], int(g["step"]), int(g["seed"])),
        float(g["pRS"]), float(g["pPR"]), float(g["pSP"]))
    assert np.array_equal(sp, g["species_cell"])


@pytest.mark.parametrize("name", CASES)
def test_cell_phase_order_is_reproducible_and_conflict_free(name):
    g = golden(name + ".npz")
    lon, lat, grid = g["lon"], g["lat"], _grid(g)
    pairs = opairs.sort_pairs(g["pairs_ref_order"])
    order, phase = orps.cell_phase_order(pairs, lon, lat, grid)
    assert np.array_equal(order.astype(np.int32), g["pairs_cell_order"])
    assert np.all(np.diff(phase) >= 0) and phase.min() >= 0 and phase.max() <= 8
    # inside a phase, a particle belongs to exactly one unit (anchor cell)
    cx = opairs.cell_index(lon, grid["x0"], grid["inv_h"], grid["ncx"])
    cy = opairs.cell_index(lat, grid["y0"], grid["inv_h"], grid["ncy"])
    key = cy * grid["ncx"] + cx
    for ph in range(9):
        sel = order[phase == ph]
        if sel.shape[0] == 0:
            continue
        ka, kb = key[sel[:, 0]], key[sel[:, 1]]
        anchor = np.minimum(ka, kb) if ph != 3 and ph != 6 else None
        if ph == 0:
            assert np.array_equal(ka, kb)
            continue
        # unit id = unordered cell pair; every cell may appear in at most one unit of the phase
        units = {(int(min(a, b)), int(max(a, b))) for a, b in zip(ka, kb)}
        cells = [c for u in units for c in u]
        assert len(cells) == len(set(cells))


def test_order_matters_and_stream_is_per_pair():
    g = golden("rps_uniform.npz")
    # the two orders give different species fields: the sequential semantics are order-sensitive
    assert not np.array_equal(g["species_ref"], g["species_cell"])
    # draws are consumed only when species differ: equal-species pairs leave everything untouched
    sp = np.ones(10, dtype=np.int8)
    out, draws = orps.rps_sequential_py(sp.copy(), np.array([[0, 1], [2, 3]]), np.array([0.1, 0.9]), 0.5, 0.5, 0.5)
    assert draws == 0 and np.array_equal(out, sp)


def test_rule_table():
    R, P, S = 1, 2, 3
    # (s1, s2, r, expected pair of species) with pRS=.2, pPR=.5, pSP=.8
    table = [(R, S, 0.1, (R, R)), (R, S, 0.3, (S, S)), (S, R, 0.1, (R, R)), (S, R, 0.3, (S, S)),
             (R, P, 0.4, (P, P)), (R, P, 0.6, (R, R)), (P, R, 0.4, (P, P)), (P, R, 0.6, (R, R)),
             (P, S, 0.7, (S, S)), (P, S, 0.9, (P, P)), (S, P, 0.7, (S, S)), (S, P, 0.9, (P, P)),
             (R, S, 0.2, (S, S))]      # strict '<': r == pRS is the backward outcome
    for s1, s2, r, want in table:
        sp = np.array([s1, s2], dtype=np.int8)
        orps.rps_pair(sp, 0, 1, r, 0.2, 0.5, 0.8)
        assert tuple(sp) == want
        sp = np.array([s1, s2], dtype=np.int8)
        orps.rps_sequential_c(sp, np.array([[0, 1]]), np.array([r]), 0.2, 0.5, 0.8)
        assert tuple(sp) == want


# ---- the tile-round order of the fused tile kernel (LM_OPT_INTERACT_MODE = 1) ------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_in_cell_round_order(name):
    """species_round: the UNMODIFIED reference function fed the cell-round order of the hybrid device path."""
    g = golden(name + ".npz")
    lon, lat, grid = g["lon"], g["lat"], _grid(g)
    order, _ = orps.cell_round_order(opairs.sort_pairs(g["pairs_ref_order"]), lon, lat, grid)
    assert np.array_equal(order.astype(np.int32), g["pairs_round_order"])
    sp, _ = orps.rps_sequential_c(g["species0"].copy(), order, philox.pair_uniforms(order[:, 0], order[:, 1], int(g["step"]), int(g["seed"])),
                                  float(g["pRS"]), float(g["pPR"]), float(g["pSP"]))
    assert np.array_equal(sp, g["species_round"])
    if name == "rps_knots":
        assert not np.array_equal(g["species_round"], g["species_cell"])       # the heavy units' rounds do change the outcome


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_in_tile_round_order(name):
    """species_tile was produced by the UNMODIFIED reference function fed the tile-round order (make_golden.py)."""
    g = golden(name + ".npz")
    order = g["pairs_tile_order"]
    sp, _ = orps.rps_sequential_c(g["species0"].copy(), order, philox.pair_uniforms(order[:, 0], order[:, 1], int(g["step"]), int(g["seed"])),
                                  float(g["pRS"]), float(g["pPR"]), float(g["pSP"]))
    assert np.array_equal(sp, g["species_tile"])


@pytest.mark.parametrize("name", CASES)
def test_tile_round_order_is_reproducible_and_its_rounds_are_matchings(name):
    g = golden(name + ".npz")
    lon, lat, grid = g["lon"], g["lat"], _grid(g)
    pairs = opairs.sort_pairs(g["pairs_ref_order"])
    order, phase = orps.tile_round_order(pairs, lon, lat, grid)
    assert np.array_equal(order.astype(np.int32), g["pairs_tile_order"])
    assert np.array_equal(opairs.sort_pairs(order), pairs)                     # a permutation of the pair set
    assert np.all(np.diff(phase) >= 0) and phase.min() >= 0 and phase.max() <= 14
    # inside one phase no microbe may occur in two units
    cx, cy, rank, occ = orps.cell_ranks(lon, lat, grid)
    key = cy * grid["ncx"] + cx
    for ph in np.unique(phase):
        sel = order[phase == ph]
        ka, kb = key[sel[:, 0]], key[sel[:, 1]]
        unit_of_cell = {}
        for a, b in zip(ka.tolist(), kb.tolist()):
            u = (min(a, b), max(a, b))
            for c in u:
                assert unit_of_cell.setdefault(c, u) == u, "phase %d: cell %d in two units" % (ph, c)
    # LIGHT units: (rank in the anchor cell, rank in the other cell) lexicographic.  HEAVY units: rounds of matchings --
    # walk the order: a new unit or a microbe already seen in the current round starts a new round; a unit then needs at
    # most max(m_a, m_b) rounds (two cells) or m - 1 + (m odd) rounds (one cell)
    rounds, cur_unit, seen, prev = {}, None, set(), None
    n_heavy_pairs = 0
    for k in range(order.shape[0]):
        i, j = int(order[k, 0]), int(order[k, 1])
        ki, kj = int(key[i]), int(key[j])
        same = ki == kj
        # anchor = the microbe of the western / southern cell (same cell: the smaller rank)
        a, b = (i, j) if ((cy[i], cx[i]) < (cy[j], cx[j]) or (same and rank[i] < rank[j])) else (j, i)
        unit = (int(phase[k]), min(ki, kj), max(ki, kj))
        light = bool(orps.unit_is_light(np.bool_(same), occ[a], occ[b]))
        if light:
            if unit == cur_unit:
                assert (rank[a], rank[b]) > prev, "light unit %r not in lexicographic rank order" % (unit,)
            prev = (rank[a], rank[b])
        else:
            n_heavy_pairs += 1
            if unit != cur_unit or i in seen or j in seen:
                rounds[unit] = rounds.get(unit, 0) + 1
                seen = set()
            seen.add(i)
            seen.add(j)
        if unit != cur_unit:
            cur_unit = unit
            if light:
                seen = set()
    for (ph, ca, cb), nr in rounds.items():
        ma, mb = int((key == ca).sum()), int((key == cb).sum())
        bound = max(ma, mb) if ca != cb else ma - 1 + (ma & 1)
        assert nr <= bound, "unit %r needs %d rounds, bound %d" % ((ph, ca, cb), nr, bound)
    if name in ("rps_clustered", "rps_knots"):
        assert n_heavy_pairs > 1000 and len(rounds) > 5          # these cases do exercise the heavy order


def test_reference_cost_pair_function_equals_the_unmodified_reference_call_by_call():
    """oracle.rps.reference_pair_interaction (what bench.py --impl reference loops over on the GPU box, where
    /root/reference does not exist) against /root/reference/interactions.py::rock_paper_scissors_interaction: same
    NumPy global stream, same species after every call."""
    import os
    import sys
    if not os.path.exists("/root/reference/interactions.py"):
        pytest.skip("the reference tree is only present in the authoring container")
    sys.path.insert(0, "/root/reference")
    import interactions as ref
    g = golden("rps_oddspecies.npz")
    params = {"pRS": 0.3, "pPR": 0.7, "pSP": 0.5}
    a, b = {"species": g["species0"].copy()}, {"species": g["species0"].copy()}
    pairs = [tuple(map(int, p)) for p in g["pairs_ref_order"][:1500]]
    np.random.seed(123)
    for p1, p2 in pairs:
        ref.rock_paper_scissors_interaction(params, a, p1, p2)
    state_ref = np.random.get_state()[2]
    np.random.seed(123)
    for p1, p2 in pairs:
        orps.reference_pair_interaction(params, b, p1, p2)
    assert np.array_equal(a["species"], b["species"]) and np.random.get_state()[2] == state_ref      # same draws consumed
    assert int((a["species"] != g["species0"]).sum()) > 50
