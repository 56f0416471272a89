"""Pair-search oracle: the restated predicate against the reference's own library call and goldens."""
import numpy as np
import pytest

from conftest import golden
from oracle import pairs as opairs

CASES = ["exact345", "dups_collinear", "lattice", "uniform", "tiny_lat", "empty", "single"]


@pytest.mark.parametrize("name", CASES)
def test_bruteforce_matches_golden_ckdtree(name):
    g = golden("pairs_cases.npz")
    lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
    want = g[name + "_pairs"].astype(np.int64)
    got = opairs.query_pairs_bruteforce(lon, lat, r)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["exact345", "uniform", "tiny_lat"])
def test_live_ckdtree_matches_golden(name):
    """SciPy on this box still returns what it returned when the fixtures were made."""
    g = golden("pairs_cases.npz")
    lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
    live = opairs.pairs_from_set(opairs.query_pairs_reference(lon, lat, r))
    assert np.array_equal(live, g[name + "_pairs"].astype(np.int64))
    assert np.array_equal(opairs.query_pairs_reference_array(lon, lat, r), live)


def test_inclusive_radius_and_coincident_points():
    g = golden("pairs_cases.npz")
    pr = {tuple(p) for p in g["exact345_pairs"]}
    assert (0, 1) in pr and (1, 2) in pr          # 3-4-5 triangles: distance exactly r
    assert (0, 2) not in pr                       # 2r apart
    lon = np.array([210.0, 210.0, 210.0], dtype=np.float32)
    lat = np.array([30.0, 30.0, 30.0], dtype=np.float32)
    assert opairs.query_pairs_bruteforce(lon, lat, 0.0).shape[0] == 3     # coincident points pair up even at r=0


def test_random_clouds_against_ckdtree():
    rng = np.random.default_rng(7)
    for n, r in ((50, 0.3), (500, 0.05), (3000, 0.02), (3000, 0.005)):
        lon = (200 + rng.random(n)).astype(np.float32)
        lat = (20 + rng.random(n)).astype(np.float32)
        want = opairs.pairs_from_set(opairs.query_pairs_reference(lon, lat, r))
        assert np.array_equal(opairs.query_pairs_bruteforce(lon, lat, r), want)


def test_cell_index_clamps():
    v = np.array([199.0, 200.0, 200.0151, 200.02, 250.0], dtype=np.float32)
    c = opairs.cell_index(v, 200.0, 100.0, 10)
    assert list(c) == [0, 0, 1, 2, 9]          # below the origin -> 0, beyond the last cell -> n-1


# ---- the other norms SciPy evaluates without pow(): interaction_norm = 1, inf (interaction_simulator.py:27,98) ----
NORM_CASES = ["exact", "uniform", "blob", "tiny_lat", "dups_collinear"]
NORMS = [("p1", 1), ("pinf", np.inf)]


@pytest.mark.parametrize("tag,p", NORMS)
@pytest.mark.parametrize("name", NORM_CASES)
def test_bruteforce_other_norms_match_golden_ckdtree(name, tag, p):
    g = golden("pairs_norms.npz")
    lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
    want = g["%s_pairs_%s" % (name, tag)].astype(np.int64)
    assert np.array_equal(opairs.query_pairs_bruteforce(lon, lat, r, p=p), want)


@pytest.mark.parametrize("tag,p", NORMS)
def test_live_ckdtree_other_norms_match_golden(tag, p):
    g = golden("pairs_norms.npz")
    for name in ("exact", "uniform", "tiny_lat"):
        lon, lat, r = g[name + "_lon"], g[name + "_lat"], float(g[name + "_r"])
        live = opairs.query_pairs_reference_array(lon, lat, r, p=p)
        assert np.array_equal(live, g["%s_pairs_%s" % (name, tag)].astype(np.int64))


def test_other_norms_are_inclusive_and_nest():
    """Points at distance exactly r count in every norm; ball(p=1) within ball(p=2) within ball(p=inf)."""
    g = golden("pairs_norms.npz")
    p1 = {tuple(q) for q in g["exact_pairs_p1"]}
    pinf = {tuple(q) for q in g["exact_pairs_pinf"]}
    assert (0, 1) in p1 and (1, 2) in p1 and (0, 3) in p1       # |dx| + |dy| == r exactly
    assert (0, 4) in pinf and (0, 4) not in p1                   # (0.375, 0.375): on the max-norm ball only
    assert (3, 6) in pinf and (0, 6) not in pinf                 # max(|dx|, |dy|) == r  /  dx = 2r
    lon, lat, r = g["uniform_lon"], g["uniform_lat"], float(g["uniform_r"])
    s1 = {tuple(q) for q in g["uniform_pairs_p1"]}
    s2 = {tuple(q) for q in opairs.query_pairs_bruteforce(lon, lat, r)}
    si = {tuple(q) for q in g["uniform_pairs_pinf"]}
    assert s1 < s2 < si


def test_random_clouds_other_norms_against_ckdtree():
    rng = np.random.default_rng(17)
    for n, r in ((50, 0.3), (500, 0.05), (3000, 0.02)):
        lon = (200 + rng.random(n)).astype(np.float32)
        lat = (rng.random(n) - 0.5).astype(np.float32)          # latitudes of both signs
        for p in (1, np.inf):
            want = opairs.query_pairs_reference_array(lon, lat, r, p=p)
            assert np.array_equal(opairs.query_pairs_bruteforce(lon, lat, r, p=p), want)
