"""FusedSimulation.run_to_file: the reference's end product (microbe_data.nc, interaction_simulator.py:62-77, :108-122)
straight from the fused loop, against a twin simulation stepped by hand."""
from datetime import datetime, timedelta

import numpy as np
import pytest

# (File name: sorts after every verified GPU test.)  Composed of verified pieces
# (step(record=...), host_copies_sync); first run on hardware: the round-1 driver run, green.
pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("stride", [1, 3])
def test_run_to_file_equals_stepping_by_hand(tmp_path, stride):
    from lagrangian_microbes_b200 import io as lmio
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from test_gpu_strips import P, R, particles, small_fs
    fs = small_fs()
    lon, lat, sp = particles(30000, 9)
    mk = lambda: FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, seed=4, emit_pairs=False, regrid_every=4)  # noqa: E731
    a, b = mk(), mk()
    t0, dt, steps = datetime(2018, 1, 1), timedelta(hours=1), 10
    path, counts = a.run_to_file(str(tmp_path), t0, t0 + steps * dt, dt, stride=stride)
    data = lmio.read_particle_file(path)
    kept = list(range(0, steps, stride))
    assert data.times == [t0 + k * dt for k in kept] and counts.shape == (len(kept), 3)
    col = 0
    for k in range(steps):
        b.step()
        if k % stride == 0:
            wl, wa, ws = b.download()
            assert np.array_equal(data["longitude"][:, col], wl) and np.array_equal(data["latitude"][:, col], wa), "step %d" % k
            assert np.array_equal(data["species"][:, col], ws), "species, step %d" % k
            assert list(counts[col]) == [int((ws == s).sum()) for s in (1, 2, 3)]
            col += 1
    assert col == len(kept)
