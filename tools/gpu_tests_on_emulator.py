"""Developer tool, NOT part of the product or the test suite: runs the bodies of GPU tests (tests/test_gpu_zz_*.py -- the
ones written when no GPU was at hand) against the CPU emulator of the kernels (tests/cuda_emu), through the product's own
torch-side wrappers.  Inside this process only, torch.cuda is monkeypatched to the CPU and tensors are wrapped in a
subclass that reports is_cuda = True, so that Engine's argument checks hold.  Slow (minutes: the tests keep their GPU
sizes); it found one bug in a test helper before any of this had seen hardware.

    python tools/gpu_tests_on_emulator.py [analysis] [tiled-golden] [tiled-live] [tiled-knots] [tiled-fused] [record] [record-sim] [driver]
"""
import contextlib
import ctypes
import os
import pathlib
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))

import emu_build  # noqa: E402
from lagrangian_microbes_b200 import _lib  # noqa: E402

_lib._lib = _lib.declare(ctypes.CDLL(emu_build.build()))          # what _lib.lib() hands to the wrappers


class FakeCuda(torch.Tensor):
    is_cuda = property(lambda self: True)


class _Stream:
    cuda_stream = 0


class _Dev:
    index = 0
    type = "cpu"


def _strip(fn):
    def f(*a, **k):
        if isinstance(k.get("device"), (_Dev, str)) or k.get("device") is not None:
            k.pop("device")
        return fn(*a, **k)
    return f


_real_to = torch.Tensor.to


def _to(self, *a, **k):
    a = tuple(x for x in a if not isinstance(x, _Dev))
    k.pop("device", None)
    return (_real_to(self, *a, **k) if (a or k) else self).as_subclass(FakeCuda)


torch.Tensor.cuda = lambda self, *a, **k: self.as_subclass(FakeCuda)
torch.Tensor.cpu = lambda self: self.as_subclass(torch.Tensor)
torch.Tensor.to = _to
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
torch.cuda.current_stream = lambda *a, **k: _Stream()
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.device = lambda *a, **k: _Dev()
torch.empty, torch.zeros, torch.full = _strip(torch.empty), _strip(torch.zeros), _strip(torch.full)

from lagrangian_microbes_b200.engine import Engine  # noqa: E402


def factory(**kw):
    return Engine(**kw)


def timed(label, fn, *a):
    t = time.time()
    fn(*a)
    print("%-60s ok  %6.1f s" % (label, time.time() - t), flush=True)


what = sys.argv[1:] or ["analysis", "tiled-golden"]
if "analysis" in what:
    import test_gpu_zz_analysis as A
    for kind in ("patch", "clustered", "global"):
        for n in (2, 257, 1025, 3000):
            timed("pair-distance histogram bounds %s n=%d" % (kind, n), A.test_pair_distance_histogram_inside_the_oracle_bounds, kind, n)
    timed("pair-distance histogram edge cases", A.test_pair_distance_histogram_edge_cases)
    for mode in (0, 1):
        for with_species in (True, False):
            timed("rasteriser mode %d species %s" % (mode, with_species), A.test_rasteriser_is_bit_exact, mode, with_species)
    with tempfile.TemporaryDirectory() as d:
        timed("MicrobePlotter frames", A.test_microbe_plotter_writes_the_reference_s_frames, pathlib.Path(d))
if any(w.startswith("tiled") for w in what):
    import test_gpu_zz_resolver_tiled as T
    if "tiled-golden" in what:
        for name in ("rps_uniform", "rps_clustered", "rps_oddspecies"):
            for smem in (32768, 1024):
                timed("tiled golden %s smem %d" % (name, smem), T.test_golden_species, factory, name, smem)
        timed("tiled option validation", T.test_mode_option_validation, factory)
    if "tiled-live" in what:
        for kind in ("knots", "uniform", "crowded"):
            timed("tiled live species %s (7 settings)" % kind, T.test_live_species, factory, kind)
    if "tiled-knots" in what:
        timed("tiled knots on the whole CTA (4 settings)", T.test_knots_on_the_whole_cta, factory)
    if "tiled-fused" in what:
        timed("tiled fused loop vs nine phases (12 steps)", T.test_fused_loop_equals_the_phase_launches)
if "record" in what or "record-sim" in what:
    class _Ev:
        def __init__(self, *a, **k): pass
        def record(self, *a): pass
        def synchronize(self): pass
    torch.cuda.Event = _Ev
    _Stream.synchronize = lambda self: None
    import test_gpu_zzz_record_delta as Rd
    if "record" in what:
        for n, off in ((0, 0), (5, 0), (4096, 0), (100003, 0), (100003, 1)):
            timed("delta pack vs oracle n=%d offset=%d" % (n, off), Rd.test_delta_pack_against_the_oracle, n, off)
    if "record-sim" in what:
        timed("packed record stream of a stepping simulation", Rd.test_packed_record_stream_of_a_stepping_simulation_is_bit_exact)
        with tempfile.TemporaryDirectory() as d:
            timed("run_to_file packed = plain (stride 3)", Rd.test_run_to_file_packed_writes_the_same_file, pathlib.Path(d), 3)
if "driver" in what:
    # the reference's call sequence through the drop-in classes (uniform_particle_locations -> ParticleAdvecter.time_step x 2
    # -> create_netcdf_file -> rock_paper_scissors -> InteractionSimulator.time_step) against the oracle pipeline
    class _Ev2:
        def __init__(self, *a, **k): pass
        def record(self, *a): pass
        def synchronize(self): pass
        def elapsed_time(self, other): return 0.0
    torch.cuda.Event = _Ev2
    import test_gpu_driver as D
    with tempfile.TemporaryDirectory() as d:
        if "driver-mapped-only" not in what:
            timed("reference call sequence end to end (N = 10,000)", D.test_reference_call_sequence_end_to_end, pathlib.Path(d))

    class _MonkeyPatch:                                        # pytest's fixture, as far as the test uses it
        def setattr(self, obj, name, value):
            setattr(obj, name, value)
    with tempfile.TemporaryDirectory() as d:
        from lagrangian_microbes_b200 import io as _lmio
        saved = _lmio._NC3_VAR_LIMIT
        try:
            timed("runs beyond NetCDF-3 through memory-mapped files", D.test_runs_beyond_netcdf3_go_through_memory_mapped_files, pathlib.Path(d), _MonkeyPatch())
        finally:
            _lmio._NC3_VAR_LIMIT = saved
print("all requested GPU test bodies passed on the emulator")
