"""GPU tests of the reference-facing driver API (ParticleAdvecter / InteractionSimulator, the call
sequence of rock_paper_scissors_example.py:22-36) and size-independent properties at large N."""
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import pairs as opairs
from oracle import philox
from oracle import rk4 as ork4
from oracle import rps as orps

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reference_call_sequence_end_to_end(tmp_path):
    """uniform_particle_locations -> ParticleAdvecter.time_step (x2) -> create_netcdf_file ->
    rock_paper_scissors -> InteractionSimulator.time_step, checked against the oracle pipeline."""
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import io as lmio, velocity_fields
    from lagrangian_microbes_b200.engine import make_grid
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet

    velocity_fields.configure_synthetic(n_modes=8, rms_speed=0.4, seed=3)
    try:
        N, r = 10000, 0.02
        start, mid, end, dt = datetime(2017, 1, 1), datetime(2017, 1, 1, 7), datetime(2017, 1, 1, 12), timedelta(hours=1)
        lons, lats = lm.uniform_particle_locations(N_particles=N, lat_min=30, lat_max=32.2, lon_min=208, lon_max=210.2)
        out = str(tmp_path / "run")
        pa = lm.ParticleAdvecter(lons, lats, N_procs=4, velocity_field="OSCAR", output_dir=out, output_chunk_iters=5, Kh=0)
        assert (pa.N_particles, pa.N_procs, pa.particles_per_tile, pa.iteration) == (N, 4, 2500, 0)
        pa.time_step(start, mid, dt)
        assert pa.iteration == 7
        names = sorted(os.listdir(out))
        assert names[0] == "particle_locations_00000_00005_tile00.pickle" and len(names) == 8
        pa.time_step(mid, end, dt)          # restores from the newest pickles; clock restarts (reference quirk Q1)
        assert pa.iteration == 12
        pa.create_netcdf_file(start, end, dt)
        assert sorted(os.listdir(out)) == ["particle_data.nc"]

        # oracle advection: two calls, each starting the particle clock at t = 0
        fs_h = HostFieldSet(velocity_fields.oscar_dataset(2017))
        fs = ork4.FieldSet(fs_h.lon, fs_h.lat, fs_h.time, fs_h.u, fs_h.v)
        pdata = lmio.read_particle_file(os.path.join(out, "particle_data.nc"))
        assert pdata["longitude"].shape == (N, 12) and pdata["longitude"].dtype == np.float32
        lon, lat = lons.astype(np.float32), lats.astype(np.float32)
        col = 0
        for nsteps in (7, 5):
            t, ti = 0.0, 0
            for _ in range(nsteps):
                lon, lat, ti, oob = ork4.rk4_step_f32(fs, lon, lat, t, 3600.0, ti)
                t += 3600.0
                assert oob == 0
                glon, glat = pdata["longitude"][:, col], pdata["latitude"][:, col]
                assert np.max(np.abs(glon - lon) / lon) < 1e-6 and np.max(np.abs(glat - lat) / lat) < 1e-6
                lon, lat = glon.copy(), glat.copy()       # continue from the GPU's stored positions
                col += 1

        np.random.seed(0)
        rps = lm.rock_paper_scissors(N_microbes=N, pRS=0.55, pPR=0.55, pSP=0.55)
        sp0 = rps[2]["species"].copy()
        isim = lm.InteractionSimulator(pair_interaction=rps, interaction_radius=r, advection_dir=out, output_dir=out, seed=9)
        isim.time_step(start, end, dt)
        assert isim.iteration == 12
        mdata = lmio.read_particle_file(os.path.join(out, "microbe_data.nc"))
        assert mdata["species"].dtype == np.int8 and mdata["species"].shape == (N, 12)
        assert np.array_equal(mdata["longitude"], pdata["longitude"])
        sp = sp0.copy()
        for i in range(12):
            lo, la = pdata["longitude"][:, i], pdata["latitude"][:, i]
            want_pairs = opairs.query_pairs_reference_array(lo, la, r)
            assert isim.pairs_found[i] == want_pairs.shape[0]
            grid = make_grid(float(lo.min()), float(lo.max()), float(la.min()), float(la.max()), r, N,
                             isim._engine.max_cells, margin=0.0).as_dict()
            order, _ = orps.canonical_order(want_pairs, lo, la, grid)
            u = philox.pair_uniforms(order[:, 0], order[:, 1], i, 9)
            sp, _ = orps.rps_sequential_c(sp, order, u, 0.55, 0.55, 0.55)
            assert np.array_equal(mdata["species"][:, i], sp)
        assert np.array_equal(rps[2]["species"], sp)      # mutated in place like the reference's dict
        assert sum(isim.pairs_found) > 0
    finally:
        velocity_fields.configure_synthetic(n_modes=64, rms_speed=0.2, seed=0)


def test_diffusion_run_is_reproducible_and_statistically_right(tmp_path):
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import io as lmio, velocity_fields
    velocity_fields.configure_synthetic(n_modes=4, rms_speed=0.0001)
    try:
        N = 40000
        lons, lats = lm.uniform_particle_locations(N_particles=N, lat_min=30, lat_max=31, lon_min=208, lon_max=209)
        outs = []
        for k in range(2):
            out = str(tmp_path / ("run%d" % k))
            pa = lm.ParticleAdvecter(lons, lats, N_procs=1, output_dir=out, output_chunk_iters=10, Kh=100, seed=1)
            pa.time_step(datetime(2017, 1, 1), datetime(2017, 1, 1, 10), timedelta(hours=1))
            pa.create_netcdf_file(datetime(2017, 1, 1), datetime(2017, 1, 1, 10), timedelta(hours=1))
            outs.append(lmio.read_particle_file(os.path.join(out, "particle_data.nc")))
        assert np.array_equal(outs[0]["latitude"], outs[1]["latitude"])
        # 9 kicks precede the 10th stored row; each has variance amp^2/3 (uniform on [-amp, amp])
        amp = np.sqrt(6 * 3600.0 * 100 / 1e10)
        d = outs[0]["latitude"][:, 9].astype(np.float64) - lats
        assert abs(d.std() - amp * np.sqrt(9 / 3.0)) < 0.03 * amp * np.sqrt(3.0)
    finally:
        velocity_fields.configure_synthetic(n_modes=64, rms_speed=0.2)


def test_large_n_properties():
    """BASELINE-size inputs (5M microbes at the config-3 density): properties that need no oracle run."""
    from lagrangian_microbes_b200.engine import Engine, make_grid
    n, r = 5_000_000, 0.01
    side = np.sqrt(n / 100000.0)            # 1e5 / deg^2 = 10M in 10x10 deg  -> rho = pi r^2 n_areal / 2 = 15.7
    rng = np.random.default_rng(0)
    lon = (205 + side * rng.random(n)).astype(np.float32)
    lat = (25 + side * rng.random(n)).astype(np.float32)
    sp0 = rng.integers(1, 4, n).astype(np.int8)
    eng = Engine(max_particles=n, max_cells=1 << 24, max_pairs=int(17.0 * n))
    try:
        eng.set_grid(make_grid(205, 205 + side, 25, 25 + side, r, n, eng.max_cells, margin=0.0))
        cap = int(17.0 * n)
        out = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
        dl, da = dev(lon), dev(lat)
        npairs = eng.find_pairs(dl, da, r, out)
        rho = npairs / n
        edge = 1 - 2 * (4 / (3 * np.pi)) * r / side            # boundary correction of the expected pair count
        assert abs(rho - 15.708 * edge) < 0.05
        pr = out[:npairs].long()
        i, j = pr[:, 0], pr[:, 1]
        assert bool((i < j).all()) and int(i.min()) >= 0 and int(j.max()) < n
        dx = dl.double()[i] - dl.double()[j]
        dy = da.double()[i] - da.double()[j]
        assert bool((dx * dx + dy * dy <= r * r).all())          # every emitted pair satisfies the predicate
        keys = i * n + j
        assert int(torch.unique(keys).numel()) == npairs          # no duplicates
        # a sub-box checked exhaustively against cKDTree
        box = (lon < 205.4) & (lat < 25.4)
        idx = np.nonzero(box)[0]
        want = opairs.query_pairs_reference_array(lon[idx], lat[idx], r)
        inbox = torch.from_numpy(box).cuda()
        sel = inbox[i] & inbox[j]
        got = pr[sel].cpu().numpy()
        remap = -np.ones(n, dtype=np.int64)
        remap[idx] = np.arange(idx.size)
        assert np.array_equal(opairs.sort_pairs(remap[got]), want)
        # idempotence + the fused path finds the same number of pairs and conserves the population
        del pr, keys, dx, dy, sel
        npairs2 = eng.find_pairs(dl, da, r, out)
        assert npairs2 == npairs
        species = dev(sp0.copy())
        eng.interact_rps(dl, da, species, r, 0.55, 0.55, 0.55, 1, 1)
        st = eng.sync_stats()
        assert st.n_pairs == npairs
        sp1 = species.cpu().numpy()
        assert sp1.min() >= 1 and sp1.max() <= 3 and (sp1 != sp0).any()
        # same inputs, same stream -> bit-identical species (atomics decide no outcome)
        species_b = dev(sp0.copy())
        eng.interact_rps(dl, da, species_b, r, 0.55, 0.55, 0.55, 1, 1)
        assert bool((species_b == species).all())
    finally:
        eng.close()


def test_runs_beyond_netcdf3_go_through_memory_mapped_files(tmp_path, monkeypatch):
    """The reference's own headline run (490,000 microbes x 7,670 steps) does not fit NetCDF-3 variables, and its dense
    (N, Nt) host arrays (particle_advecter.py:269-270, interaction_simulator.py:80-82) would be 33.8 GB: beyond the
    limit the drop-in classes write ``<file>.npz.d/`` (memory-mapped .npy files, filled in blocks) and read it back
    lazily.  Here the limit is lowered so that a small run takes that path: same numbers as the in-memory path."""
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import io as lmio, velocity_fields
    velocity_fields.configure_synthetic(n_modes=8, rms_speed=0.4, seed=3)
    try:
        N, r = 3000, 0.03
        start, end, dt = datetime(2017, 1, 1), datetime(2017, 1, 1, 7), timedelta(hours=1)
        lons, lats = lm.uniform_particle_locations(N_particles=N, lat_min=30, lat_max=31.2, lon_min=208, lon_max=209.2)
        results = []
        for tag, limit in (("nc3", None), ("mapped", 4 * N * 3)):          # "mapped": variables above three columns
            if limit is not None:
                monkeypatch.setattr(lmio, "_NC3_VAR_LIMIT", limit)
            out = str(tmp_path / tag)
            pa = lm.ParticleAdvecter(lons, lats, N_procs=3, output_dir=out, output_chunk_iters=4, Kh=0)
            pa.time_step(start, end, dt)
            pa.create_netcdf_file(start, end, dt)
            np.random.seed(0)
            rps = lm.rock_paper_scissors(N_microbes=N, pRS=0.55, pPR=0.55, pSP=0.55)
            isim = lm.InteractionSimulator(pair_interaction=rps, interaction_radius=r, advection_dir=out, output_dir=out, seed=9)
            isim.time_step(start, end, dt)
            names = sorted(os.listdir(out))
            assert names == (["microbe_data.nc", "particle_data.nc"] if limit is None else ["microbe_data.nc.npz.d", "particle_data.nc.npz.d"])
            m = lmio.read_particle_file(os.path.join(out, "microbe_data.nc"))
            results.append({k: np.array(m[k]) for k in ("longitude", "latitude", "species")})
            assert m.times == [start + k * dt for k in range(7)] and sum(isim.pairs_found) > 0
        for k in ("longitude", "latitude", "species"):
            assert results[0][k].shape == (N, 7) and np.array_equal(results[0][k], results[1][k]), k
    finally:
        velocity_fields.configure_synthetic(n_modes=64, rms_speed=0.2, seed=0)
