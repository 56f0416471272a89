"""BASELINE config 1 -- the reference's own CPU-runnable case -- end to end on the GPU against the CPU oracle loop:
490,000 microbes on the 700 x 700 lattice over 25-35N, 205-215E (README.md:7 of the reference;
sandbox/pairwise_distance_histogram_distributed.jl:150), species from the reference's factory after
np.random.seed(0) (interactions.py:45), pRS = pPR = pSP = 0.55, r = 0.01 degrees, steady synthetic velocity on the
OSCAR 1/3-degree grid, dt = 1 h, 24 steps.  Every step: positions vs the float32-faithful RK4 restatement (bitwise)
and the float64 one (1e-6 relative), the pair set vs cKDTree.query_pairs on the same positions, species vs the
sequential reference rule in the canonical order."""
import numpy as np
import pytest

from oracle import pairs as opairs
from oracle import philox
from oracle import rk4 as ork4
from oracle import rps as orps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_config1_24_steps_against_the_cpu_oracle():
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import velocity_fields
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet
    from lagrangian_microbes_b200.simulation import FusedSimulation

    n, r, p, seed = 490_000, 0.01, (0.55, 0.55, 0.55), 0
    velocity_fields.configure_synthetic(kind="steady", seed=0, n_modes=64, rms_speed=0.2)
    try:
        hfs = HostFieldSet(velocity_fields.oscar_dataset(2017))
        fs = ork4.FieldSet(hfs.lon, hfs.lat, hfs.time, hfs.u, hfs.v)
        lons, lats = lm.uniform_particle_locations(n, 25, 35, 205, 215)
        assert lons.size == n and abs((lons.max() - lons.min()) / 699 - 0.014306) < 1e-5      # the 700 x 700 lattice
        np.random.seed(0)
        _, _, props = lm.rock_paper_scissors(n, *p)
        sp0 = props["species"].copy()
        assert sp0.dtype == np.int8
        sim = FusedSimulation(lons, lats, sp0, r, *p, hfs, dt_seconds=3600.0, seed=seed, emit_pairs=True,
                              pair_capacity=8 * n, regrid_every=8, grid_margin=0.5)
        lon_prev, lat_prev = lons.astype(np.float32), lats.astype(np.float32)
        sp_ref = sp0.copy()
        t, ti, total_pairs, bit_mismatch = 0.0, 0, 0, 0
        def checked_step(step):
            nonlocal lon_prev, lat_prev, sp_ref, t, ti, total_pairs, bit_mismatch
            grid = sim.grid.as_dict()
            st = sim.step(check=True)
            gl, ga, gs = sim.download()
            a32, b32 = lon_prev.copy(), lat_prev.copy()          # rk4_step_c advances its arrays in place
            ti_new, oob = ork4.rk4_step_c(fs, a32, b32, t, 3600.0, ti)
            a64, b64, _, _ = ork4.rk4_step_f64(fs, lon_prev, lat_prev, t, 3600.0, ti)
            assert oob == 0 and st.n_out_of_bounds == 0
            bit_mismatch += int((gl != a32).sum() + (ga != b32).sum())
            assert np.max(np.abs(gl - a64) / np.abs(a64)) < 1e-6 and np.max(np.abs(ga - b64) / np.abs(b64)) < 1e-6
            want_pairs = opairs.query_pairs_reference_array(gl, ga, r)
            assert st.n_pairs == want_pairs.shape[0], "step %d" % step
            assert np.array_equal(opairs.sort_pairs(sim.pairs[:st.n_pairs].cpu().numpy()), want_pairs)
            order, _ = orps.canonical_order(want_pairs, gl, ga, grid)
            u = philox.pair_uniforms(order[:, 0], order[:, 1], step, seed)
            sp_ref, _ = orps.rps_sequential_c(sp_ref, order, u, *p)
            assert np.array_equal(gs, sp_ref), "species differ at step %d" % step
            assert list(st.species_count[1:]) == [int((sp_ref == k).sum()) for k in (1, 2, 3)]
            lon_prev, lat_prev, t, ti = gl, ga, t + 3600.0, ti_new
            total_pairs += st.n_pairs
            return st.n_pairs

        # the configuration as specified: 24 hourly steps.  The lattice spacing (0.0143 degrees) exceeds r and a day of
        # this flow does not close it, so the interaction phase finds nothing -- which must also be reproduced.
        for step in range(24):
            checked_step(step)
        pairs_24 = total_pairs
        # let the flow strain the lattice for four more days (GPU only), re-anchor the oracle on the GPU state and
        # check six more steps of the same population, now with pairs
        for step in range(24, 120):
            sim.step()
        lon_prev, lat_prev, sp_ref = sim.download()
        t = 120 * 3600.0
        ti = sim.clock.ti
        for step in range(120, 126):
            checked_step(step)
        print("config 1: %d pairs in the first 24 steps, %d in steps 120-125, %d positions differ bitwise from the float32 "
              "oracle" % (pairs_24, total_pairs - pairs_24, bit_mismatch))
        assert total_pairs - pairs_24 > 0
        assert bit_mismatch <= 1e-4 * 2 * n * 30
    finally:
        velocity_fields.configure_synthetic(kind="random_fourier", seed=0, n_modes=64, rms_speed=0.2)
