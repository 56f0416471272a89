"""``ParticleAdvecter(N_procs=k)`` under a torch.distributed group: the reference's joblib workers
(particle_advecter.py:86-94, 143-148 -- contiguous particle tiles, one worker each, no communication) become one rank
(= one GPU) each.  Bar: the merged ``particle_data.nc`` of a W-rank run equals the single-process run bit for bit,
diffusion kicks included (they are keyed by the global particle index, lm_diffuse_ids).

The ranks of this test share GPU 0 when the box has fewer GPUs than ranks (gloo group); with enough GPUs they take one
each (nccl group)."""
import os
import socket
from datetime import datetime, timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N, TILES = 12000, 6
START, MID, END, DT = datetime(2017, 1, 1), datetime(2017, 1, 1, 5), datetime(2017, 1, 1, 9), timedelta(hours=1)


def run_advecter(out, kh):
    import lagrangian_microbes_b200 as lm
    from lagrangian_microbes_b200 import velocity_fields
    velocity_fields.configure_synthetic(n_modes=8, rms_speed=0.4, seed=3)
    try:
        lons, lats = lm.uniform_particle_locations(N_particles=N, lat_min=30, lat_max=32.2, lon_min=208, lon_max=210.2)
        pa = lm.ParticleAdvecter(lons[:N], lats[:N], N_procs=TILES, velocity_field="OSCAR", output_dir=out, output_chunk_iters=3, Kh=kh)
        pa.time_step(START, MID, DT)
        pa.time_step(MID, END, DT)               # restores this rank's tiles from the newest pickles
        pa.create_netcdf_file(START, END, DT)
        return pa
    finally:
        velocity_fields.configure_synthetic()


def _worker(rank, world, port, out, kh):
    import torch.distributed as dist
    one_each = torch.cuda.device_count() >= world
    torch.cuda.set_device(rank if one_each else 0)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl" if one_each else "gloo", rank=rank, world_size=world)
    try:
        pa = run_advecter(out, kh)
        from lagrangian_microbes_b200.particle_advecter import tiles_of_rank
        assert pa.my_tiles == tiles_of_rank(TILES, rank, world) and (pa.rank, pa.world) == (rank, world)
    except BaseException:                        # do not wait for the other ranks in a barrier they will never reach
        import traceback
        traceback.print_exc()
        os._exit(1)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,kh", [(2, 0.0), (3, 25.0), (4, 25.0)])
def test_tiles_over_ranks_equal_one_process(tmp_path, world, kh):
    import torch.multiprocessing as mp
    from lagrangian_microbes_b200 import io as lmio
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_w = str(tmp_path / "ranks")
    os.makedirs(out_w)
    mp.spawn(_worker, args=(world, port, out_w, kh), nprocs=world, join=True)
    assert sorted(os.listdir(out_w)) == ["particle_data.nc"]           # rank 0 merged every rank's pickles and removed them
    out_1 = str(tmp_path / "one")
    run_advecter(out_1, kh)
    a = lmio.read_particle_file(os.path.join(out_w, "particle_data.nc"))
    b = lmio.read_particle_file(os.path.join(out_1, "particle_data.nc"))
    assert a["longitude"].shape == (N, 9)
    assert np.array_equal(a["longitude"], b["longitude"]) and np.array_equal(a["latitude"], b["latitude"])
    # the tiles moved (and, with Kh > 0, were kicked)
    assert np.abs(a["longitude"][:, -1] - a["longitude"][:, 0]).max() > 1e-4
