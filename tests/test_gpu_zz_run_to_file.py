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


@pytest.mark.parametrize("passes", [2, 5])
def test_record_scatter_in_id_windows_gives_the_same_file(tmp_path, passes):
    """LM_OPT_SCATTER_PASSES: the scatter to particle order taken in id windows (what a 12.5 M-microbe handle does by
    default so that a window's sectors complete in L2) writes the same record as the single pass."""
    from lagrangian_microbes_b200 import _lib, io as lmio
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from test_gpu_strips import P, R, particles, small_fs
    fs = small_fs()
    lon, lat, sp = particles(30011, 10)                           # not a multiple of anything
    files = []
    for k, n_pass in enumerate((1, passes)):
        sim = FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, seed=4, emit_pairs=False, regrid_every=4)
        sim.engine.set_option(_lib.LM_OPT_SCATTER_PASSES, n_pass)
        t0, dt = datetime(2018, 1, 1), timedelta(hours=1)
        path, _ = sim.run_to_file(str(tmp_path / ("p%d" % k)), t0, t0 + 5 * dt, dt, stride=2)
        files.append(lmio.read_particle_file(path))
        wl, wa, ws = sim.download()                               # lm_state_get takes the same windows
        assert np.array_equal(files[-1]["longitude"][:, -1], wl) and np.array_equal(files[-1]["species"][:, -1], ws)
        sim.engine.close()
    for name in ("longitude", "latitude", "species"):
        assert np.array_equal(files[0][name], files[1][name]), name
    with pytest.raises(_lib.LmError):
        FusedSimulation(lon[:100], lat[:100], sp[:100], R, *P, fs).engine.set_option(_lib.LM_OPT_SCATTER_PASSES, 65)
