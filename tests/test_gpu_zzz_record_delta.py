"""The delta-packed position record on the GPU (csrc/record.cu, lm_record_delta_pack; record.DeltaRecordPacker): the
decoded stream of a stepping simulation must equal the float32 positions the reference would store
(particle_advecter.py:233-235, interaction_simulator.py:108-110) BIT FOR BIT -- against the plain download of the same
step, and against oracle/record.py on the raw deltas."""
import ctypes

import numpy as np
import pytest

from oracle import record as orec

# (File name: sorts after every verified GPU test.)  Also executed on the CPU emulator
# (tests/test_record_delta.py); first run on hardware: the round-1 driver run, green.
pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("n,offset", [(0, 0), (5, 0), (4096, 0), (100003, 0), (100003, 1)])
def test_delta_pack_against_the_oracle(n, offset):
    from lagrangian_microbes_b200 import _lib, io as lmio
    from test_record_delta import _records
    L = _lib.lib()
    prev_lon, prev_lat, lon, lat = _records(n, seed=n + offset)
    dev = lambda a: torch.from_numpy(np.concatenate([np.zeros(offset, a.dtype), a])).cuda()[offset:]   # offset 1: not 16-byte aligned
    tp, ta, tl, tt = dev(prev_lon), dev(prev_lat), dev(lon), dev(lat)
    dl = torch.full((n + offset,), 77, dtype=torch.int16, device="cuda")[offset:]
    da = torch.full((n + offset,), 77, dtype=torch.int16, device="cuda")[offset:]
    cap = 64
    esc = torch.zeros((cap, 2), dtype=torch.int32, device="cuda")
    cnt = torch.full((1,), 99, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.lm_record_delta_pack(p(tp), p(ta), p(tl), p(tt), n, p(dl), p(da), p(esc), cap, p(cnt),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "lm_record_delta_pack")
    torch.cuda.synchronize()
    m = int(cnt.item())
    sample = slice(0, min(n, 3000))                              # the oracle is a Python loop
    w_dl, w_da, _ = orec.pack_reference(prev_lon[sample], prev_lat[sample], lon[sample], lat[sample])
    assert np.array_equal(dl.cpu().numpy()[sample], w_dl) and np.array_equal(da.cpu().numpy()[sample], w_da)
    esc_h = esc.cpu().numpy().view(np.uint32)[:m]
    got_lon, got_lat = lmio.unpack_delta_record(prev_lon, prev_lat, dl.cpu().numpy(), da.cpu().numpy(), esc_h)
    assert np.array_equal(got_lon.view(np.uint32), lon.view(np.uint32))
    assert np.array_equal(got_lat.view(np.uint32), lat.view(np.uint32))


def test_packed_record_stream_of_a_stepping_simulation_is_bit_exact():
    from lagrangian_microbes_b200.record import DeltaRecordPacker
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from test_gpu_strips import P, R, particles, small_fs
    fs = small_fs()
    lon, lat, sp = particles(30000, 9)
    sim = FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, seed=4, emit_pairs=False, regrid_every=4)
    n = len(lon)
    packer = DeltaRecordPacker(n, escape_capacity=16)            # small list: the overflow -> key frame path is reachable
    lon_d = torch.empty(n, dtype=torch.float32, device="cuda"); lat_d = torch.empty_like(lon_d)
    want = []
    for k in range(8):
        sim.step()
        sim.engine.state_get(lon_d, lat_d, None)
        if k == 5:
            lon_d[:40] += 2.0                                    # 40 far jumps > 16 escapes: this step is resent as a key frame
        packer.push(lon_d, lat_d)
        want.append((lon_d.cpu().numpy().copy(), lat_d.cpu().numpy().copy()))
        if k >= 1:                                               # one record in flight, as a pipelined caller would
            got = packer.pop()
            assert np.array_equal(got[0].view(np.uint32), want[k - 1][0].view(np.uint32)), "lon, step %d" % (k - 1)
            assert np.array_equal(got[1].view(np.uint32), want[k - 1][1].view(np.uint32)), "lat, step %d" % (k - 1)
    got = packer.pop()
    assert np.array_equal(got[0].view(np.uint32), want[-1][0].view(np.uint32)) and np.array_equal(got[1].view(np.uint32), want[-1][1].view(np.uint32))
    assert packer.bytes_d2h < 8 * n * 8                          # fewer bytes than the plain record of eight steps


@pytest.mark.parametrize("stride", [1, 3])
def test_run_to_file_packed_writes_the_same_file(tmp_path, stride):
    """FusedSimulation.run_to_file(packed=True) against the plain record of a twin simulation: every column of
    microbe_data.nc identical bit for bit, fewer bytes over the link."""
    from datetime import datetime, timedelta
    from lagrangian_microbes_b200 import io as lmio
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from test_gpu_strips import P, R, particles, small_fs
    fs = small_fs()
    lon, lat, sp = particles(30000, 9)
    mk = lambda: FusedSimulation(lon, lat, sp, R, *P, fs, dt_seconds=3600.0, seed=4, emit_pairs=False, regrid_every=4)  # noqa: E731
    a, b = mk(), mk()
    t0, dt, steps = datetime(2018, 1, 1), timedelta(hours=1), 9
    pa, ca = a.run_to_file(str(tmp_path / "plain"), t0, t0 + steps * dt, dt, stride=stride)
    pb, cb = b.run_to_file(str(tmp_path / "packed"), t0, t0 + steps * dt, dt, stride=stride, packed=True)
    da, db = lmio.read_particle_file(pa), lmio.read_particle_file(pb)
    for k in ("longitude", "latitude"):
        assert np.array_equal(np.array(da[k]).view(np.uint32), np.array(db[k]).view(np.uint32)), k
    assert np.array_equal(np.array(da["species"]), np.array(db["species"])) and np.array_equal(ca, cb)
    assert b.record_bytes_d2h < 9 * len(lon) * len(range(0, steps, stride))
