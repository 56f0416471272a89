"""Developer tool, NOT part of the product (tests/test_host_layer_emulated.py runs it in a subprocess): runs the torch-side host layer (Engine, FusedSimulation,
run_to_file and its record pipeline) against the CPU emulator of the kernels (tests/cuda_emu) when no GPU is at hand.
torch.cuda is monkeypatched to the CPU inside this process only, and the script must be run with ``python -O`` so that the
wrappers' ``is_cuda`` asserts are stripped:

    python -O tools/host_layer_on_emulator.py

Checks FusedSimulation.run_to_file (stride 2, regridding on) against a twin simulation stepped by hand -- the logic of
tests/test_gpu_zz_run_to_file.py -- and prints the per-step species census."""
import ctypes, os, sys, contextlib, tempfile, numpy as np, torch
from datetime import datetime, timedelta
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'cuda_emu')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
assert not __debug__, 'run with python -O (the wrappers\' is_cuda asserts must be stripped for the emulator)'
import emu_build
from lagrangian_microbes_b200 import _lib
_lib._lib = _lib.declare(ctypes.CDLL(emu_build.build()))
torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
class _S: cuda_stream = 0
torch.cuda.current_stream = lambda *a, **k: _S()
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.cuda.synchronize = lambda *a, **k: None
_real_device = torch.device
class _Dev:
    index = 0
    type = "cpu"
torch.Tensor.pin_memory = lambda self, *a, **k: self
_real_to = torch.Tensor.to
def _to(self, *a, **k):
    a = tuple(x for x in a if not isinstance(x, _Dev)); k.pop("device", None) if isinstance(k.get("device"), _Dev) else None
    return _real_to(self, *a, **k) if (a or k) else self
torch.Tensor.to = _to
import lagrangian_microbes_b200.engine as eng
_dev = _Dev()
def fake_device(*a, **k): return _dev
eng.torch.device = fake_device
_real_empty, _real_zeros = torch.empty, torch.zeros
def strip(fn):
    def f(*a, **k):
        if isinstance(k.get("device"), _Dev): k.pop("device")
        return fn(*a, **k)
    return f
torch.empty, torch.zeros = strip(_real_empty), strip(_real_zeros)
from lagrangian_microbes_b200.simulation import FusedSimulation
from lagrangian_microbes_b200 import io as lmio
from conftest import golden
from oracle import rk4 as ork4
g = golden("rk4_small.npz")
fs0 = ork4.FieldSet(g["grid_lon"], g["grid_lat"], g["grid_time"], g["u"], g["v"])
class HostFS:
    u, v, lon, lat, time = fs0.u, fs0.v, fs0.lon, fs0.lat, fs0.time
    @staticmethod
    def to_device(device): return tuple(torch.from_numpy(a) for a in (fs0.u, fs0.v, fs0.lon, fs0.lat))
rng = np.random.default_rng(9)
n = n_sim = 800
lon = (201.5 + 0.2*rng.random(n)).astype(np.float32); lat = (32.5 + 0.15*rng.random(n)).astype(np.float32)
sp = rng.integers(1,4,n).astype(np.int8)
mk = lambda: FusedSimulation(lon, lat, sp, 0.01, 0.55, 0.6, 0.9, HostFS, dt_seconds=3600.0, seed=4, emit_pairs=False, regrid_every=2, grid_margin=0.1, max_cells=1<<14, pair_capacity=40*n)
a, b = mk(), mk()
t0, dt, steps, stride = datetime(2018,1,1), timedelta(hours=1), 5, 2
with tempfile.TemporaryDirectory() as d:
    path, counts = a.run_to_file(d, t0, t0 + steps*dt, dt, stride=stride)
    data = lmio.read_particle_file(path)
    col = 0
    for k in range(steps):
        b.step()
        if k % stride == 0:
            wl, wa, ws = b.download()
            assert np.array_equal(data["longitude"][:, col], wl) and np.array_equal(data["latitude"][:, col], wa), k
            assert np.array_equal(data["species"][:, col], ws), k
            assert list(counts[col]) == [int((ws == s).sum()) for s in (1,2,3)]
            col += 1
    print("run_to_file ok:", data["species"].shape, data.times, counts.tolist())
    plain = {k: np.array(data[k]) for k in ("longitude", "latitude", "species")}


# ---- record.DeltaRecordPacker (lm_record_delta_pack + io.unpack_delta_record): a pipelined stream with overflow steps ----
class _Ev:
    def record(self, *a): pass
    def synchronize(self): pass
torch.cuda.Event = _Ev
_S.synchronize = lambda self: None
class _FakeCudaTensor(torch.Tensor):
    is_cuda = property(lambda self: True)
from lagrangian_microbes_b200.record import DeltaRecordPacker
n = 5003
packer = DeltaRecordPacker(n, escape_capacity=16, device=_dev)
cur_lon = (205 + 10*rng.random(n)).astype(np.float32); cur_lat = (25 + 10*rng.random(n)).astype(np.float32)
want = []
for k in range(7):
    cur_lon = (cur_lon + 0.02*(rng.random(n) - 0.5)).astype(np.float32); cur_lat = (cur_lat + 0.02*(rng.random(n) - 0.5)).astype(np.float32)
    if k == 3: cur_lon[:5] += 3.0            # 5 escapes: fits the list
    if k == 5: cur_lon[:40] += 2.0           # 40 escapes: overflows, resent as a key frame
    packer.push(torch.from_numpy(cur_lon.copy()).as_subclass(_FakeCudaTensor), torch.from_numpy(cur_lat.copy()).as_subclass(_FakeCudaTensor))
    want.append((cur_lon.copy(), cur_lat.copy()))
    if k >= 1:
        g = packer.pop()
        assert np.array_equal(g[0].view(np.uint32), want[k-1][0].view(np.uint32)) and np.array_equal(g[1].view(np.uint32), want[k-1][1].view(np.uint32)), k
g = packer.pop()
assert np.array_equal(g[0].view(np.uint32), want[-1][0].view(np.uint32)) and np.array_equal(g[1].view(np.uint32), want[-1][1].view(np.uint32))
print("delta record stream ok: %d B over the link for %d plain" % (packer.bytes_d2h, 8*n*7))

# ---- run_to_file(packed=True): the same file as the plain record, bit for bit --------------------------------------------
c = mk()
with tempfile.TemporaryDirectory() as d:
    path, counts2 = c.run_to_file(d, t0, t0 + steps*dt, dt, stride=stride, packed=True)
    data2 = lmio.read_particle_file(path)
    for k in ("longitude", "latitude", "species"):
        a2 = np.array(data2[k])
        assert np.array_equal(a2.view(np.uint32) if a2.dtype == np.float32 else a2, plain[k].view(np.uint32) if a2.dtype == np.float32 else plain[k]), k
    assert np.array_equal(counts2, counts)
    print("run_to_file(packed=True) ok: %d B of record over the link, plain %d" % (c.record_bytes_d2h, 9 * n_sim * data2["species"].shape[1]))

# ---- DeltaRecordPacker.pop(decode=False): the packed stream as stored, decoded later (decode on read) -----------------
packer = DeltaRecordPacker(n, escape_capacity=64, device=_dev)
stored, truth = [], []
for k in range(4):
    cur_lon = (cur_lon + 0.02*(rng.random(n) - 0.5)).astype(np.float32); cur_lat = (cur_lat + 0.02*(rng.random(n) - 0.5)).astype(np.float32)
    packer.push(torch.from_numpy(cur_lon.copy()).as_subclass(_FakeCudaTensor), torch.from_numpy(cur_lat.copy()).as_subclass(_FakeCudaTensor))
    truth.append((cur_lon.copy(), cur_lat.copy()))
    stored.append(tuple(np.array(x) if isinstance(x, np.ndarray) else x for x in packer.pop(decode=False)))
assert [r[0] for r in stored] == ["key", "delta", "delta", "delta"]
state = (stored[0][1], stored[0][2])
for k in range(1, 4):
    state = lmio.unpack_delta_record(state[0], state[1], *stored[k][1:])
    assert np.array_equal(state[0].view(np.uint32), truth[k][0].view(np.uint32)) and np.array_equal(state[1].view(np.uint32), truth[k][1].view(np.uint32))
print("packed stream decoded on read ok")
