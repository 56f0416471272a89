"""Thin Python wrapper over the C ABI (include/lm_b200.h) on PyTorch CUDA tensors.

PyTorch is used here only for device memory, streams and (elsewhere) torch.distributed; every
computation is a hand-written CUDA kernel in liblm_b200.so.  There is no CPU path: constructing
an ``Engine`` without a CUDA device raises.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import Grid, RpsParams, StageTimes, Stats, Strip, StripBuffers, check


def make_grid(lon_min, lon_max, lat_min, lat_max, radius, n_particles, max_cells, margin=0.5, cells_per_particle=2.0):
    """Host policy for the binning grid.

    Cell edge h = k * r * (1 + 2^-20) with the smallest integer k >= 1 whose grid over the
    margin-padded bounding box has at most ``min(max_cells, max(cells_per_particle * n, 2^16))``
    cells.  The origin is integral so that ``double(x) - x0`` is exact.  Particles that leave the
    box are clamped into edge cells by the device (correct, only slower).
    """
    r = float(radius)
    base = (r if r > 0 else 1e-3) * (1.0 + 2.0 ** -20)
    x0 = math.floor(lon_min - margin)
    y0 = math.floor(lat_min - margin)
    sx = (lon_max + margin) - x0
    sy = (lat_max + margin) - y0
    budget = int(min(max_cells, max(cells_per_particle * n_particles, 1 << 16)))
    k = max(1, int(math.ceil(math.sqrt(max(sx * sy, 1e-30) / (base * base) / budget))))
    while True:
        h = base * k
        ncx = int(sx / h) + 2
        ncy = int(sy / h) + 2
        if ncx * ncy <= budget:
            break
        k += 1
    return Grid(float(x0), float(y0), 1.0 / h, ncx, ncy)


def norm_code(p):
    """LM_NORM_* for the ``p`` of ``cKDTree.query_pairs(r, p)`` (interaction_simulator.py:27,98).  SciPy evaluates
    p = 1, 2 and infinity with plain adds / multiplies / max, which the device reproduces bit for bit; any other p
    goes through pow() and is not offered."""
    if isinstance(p, (int, float, np.integer, np.floating)) and not isinstance(p, bool):
        if p == 2:
            return _lib.LM_NORM_2
        if p == 1:
            return _lib.LM_NORM_1
        if p == math.inf:
            return _lib.LM_NORM_INF
    raise NotImplementedError("interaction_norm=%r: the device offers the Minkowski norms p = 1, 2 and inf" % (p,))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32(t, n=None):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous CUDA float32"
    if n is not None:
        assert t.numel() == n
    return t


class Engine:
    """One lm_handle bound to one CUDA device."""

    # LM_OPT_INTERACT_MODE / LM_OPT_ADVECT_MODE a new Engine starts with (None: the library's defaults -- fused tile
    # kernel, bit-faithful RK4).  Class attributes so that a test module or a benchmark can pin the whole stack
    # (FusedSimulation, StripSet, the drop-in classes) to one path.
    DEFAULT_INTERACT_MODE = None
    DEFAULT_ADVECT_MODE = None

    def __init__(self, max_particles, max_cells=None, max_pairs=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("lagrangian_microbes_b200 needs a CUDA device: the hot path has no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_particles = int(max_particles)
        self.max_cells = int(max_cells if max_cells is not None else max(4 * max_particles, 1 << 20))
        self.max_pairs = int(max_pairs)
        h = ctypes.c_void_p()
        check(self.L.lm_create(ctypes.byref(h), self.device.index, self.max_particles, self.max_cells, self.max_pairs),
              "lm_create")
        self.h = h
        self._field = None            # keeps the borrowed field tensors alive
        self._n_pairs_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.interact_mode = 2                       # the library's default (hybrid)
        # class attribute, else the environment (LM_INTERACT_MODE / LM_ADVECT_MODE: A/B runs of the tools), else the library's
        im = self.DEFAULT_INTERACT_MODE if self.DEFAULT_INTERACT_MODE is not None else os.environ.get("LM_INTERACT_MODE")
        am = self.DEFAULT_ADVECT_MODE if self.DEFAULT_ADVECT_MODE is not None else os.environ.get("LM_ADVECT_MODE")
        if im is not None:
            self.set_option(_lib.LM_OPT_INTERACT_MODE, int(im))
        if am is not None:
            self.set_option(_lib.LM_OPT_ADVECT_MODE, int(am))

    def close(self):
        if getattr(self, "h", None):
            self.L.lm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- inputs --------------------------------------------------------------------------------
    def set_field(self, u, v, lon, lat):
        """u, v float32 (T, Y, X) NaN-free; lon (X,), lat (Y,) ascending -- CUDA tensors (borrowed)."""
        u, v, lon, lat = _f32(u), _f32(v), _f32(lon), _f32(lat)
        T, Y, X = u.shape
        assert v.shape == u.shape and lon.numel() == X and lat.numel() == Y
        check(self.L.lm_set_field(self.h, _ptr(u), _ptr(v), _ptr(lon), _ptr(lat), T, Y, X), "lm_set_field")
        self._field = (u, v, lon, lat)

    def update_field_data(self, u, v):
        """Swap in another window of time levels (same axes); u, v CUDA float32 (T, Y, X), borrowed."""
        u, v = _f32(u), _f32(v)
        assert self._field is not None and u.shape == v.shape and u.shape[1:] == self._field[0].shape[1:]
        check(self.L.lm_update_field_data(self.h, _ptr(u), _ptr(v), u.shape[0]), "lm_update_field_data")
        self._field = (u, v, self._field[2], self._field[3])

    def set_grid(self, grid):
        check(self.L.lm_set_grid(self.h, ctypes.byref(grid)), "lm_set_grid")

    def get_grid(self):
        g = Grid()
        check(self.L.lm_get_grid(self.h, ctypes.byref(g)), "lm_get_grid")
        return g

    # ---- stateless operators ---------------------------------------------------------------------
    def advect_rk4(self, lon, lat, stage_times, dt):
        n = lon.numel()
        check(self.L.lm_advect_rk4(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), n, ctypes.byref(stage_times),
                                   float(dt), self._stream()), "lm_advect_rk4")

    def diffuse(self, lon, lat, amp_deg, seed, step, ids=None):
        """``ids``: CUDA int32 [n] global particle ids keying the kicks (default: the array index)."""
        n = lon.numel()
        if ids is None:
            check(self.L.lm_diffuse(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), n, float(amp_deg), int(seed), int(step),
                                    self._stream()), "lm_diffuse")
            return
        assert ids.is_cuda and ids.dtype == torch.int32 and ids.is_contiguous() and ids.numel() == n
        check(self.L.lm_diffuse_ids(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), _ptr(ids), n, float(amp_deg), int(seed),
                                    int(step), self._stream()), "lm_diffuse_ids")

    def find_pairs(self, lon, lat, r, pairs_out):
        """pairs_out: CUDA int32 (cap, 2).  Returns the number of pairs found (host int, synchronises)."""
        n = lon.numel()
        assert pairs_out.dtype == torch.int32 and pairs_out.is_contiguous() and pairs_out.shape[1] == 2
        check(self.L.lm_find_pairs(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), n, float(r), _ptr(pairs_out),
                                   pairs_out.shape[0], _ptr(self._n_pairs_dev), self._stream()), "lm_find_pairs")
        st = self.sync_stats()
        return st.n_pairs

    def interact_rps(self, lon, lat, species, r, pRS, pPR, pSP, seed, step, pairs_out=None):
        n = lon.numel()
        assert species.dtype == torch.int8 and species.is_cuda and species.is_contiguous() and species.numel() == n
        prm = RpsParams(float(pRS), float(pPR), float(pSP), int(seed), int(step))
        cap = pairs_out.shape[0] if pairs_out is not None else 0
        check(self.L.lm_interact_rps(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), _ptr(species), n, float(r),
                                     ctypes.byref(prm), _ptr(pairs_out), cap, _ptr(self._n_pairs_dev),
                                     self._stream()), "lm_interact_rps")

    def pair_uniforms(self, pairs, seed, step):
        assert pairs.dtype == torch.int32 and pairs.is_cuda and pairs.is_contiguous()
        P = pairs.shape[0]
        u = torch.empty(P, dtype=torch.float64, device=pairs.device)
        check(self.L.lm_pair_uniforms(_ptr(pairs), P, int(seed), int(step), _ptr(u), self._stream()), "lm_pair_uniforms")
        return u

    def resolve_rps(self, pairs, u, species, pRS, pPR, pSP):
        """Explicit-order resolver; returns the number of rounds."""
        assert pairs.dtype == torch.int32 and pairs.is_contiguous() and u.dtype == torch.float64 and u.is_contiguous()
        assert species.dtype == torch.int8 and species.is_contiguous()
        rounds = ctypes.c_int32(0)
        check(self.L.lm_resolve_rps(self.h, _ptr(pairs), _ptr(u), pairs.shape[0], _ptr(species), species.numel(),
                                    float(pRS), float(pPR), float(pSP), ctypes.byref(rounds), self._stream()),
              "lm_resolve_rps")
        return rounds.value

    # ---- resident pipeline -------------------------------------------------------------------------
    def state_set(self, lon, lat, species=None, ids=None):
        n = lon.numel()
        if species is not None:
            assert species.dtype == torch.int8 and species.is_cuda and species.numel() == n
        if ids is not None:
            assert ids.dtype == torch.int32 and ids.is_cuda and ids.numel() == n
        check(self.L.lm_state_set(self.h, _ptr(_f32(lon)), _ptr(_f32(lat, n)), _ptr(species), _ptr(ids), n,
                                  self._stream()), "lm_state_set")

    def state_size(self):
        return int(self.L.lm_state_size(self.h))

    def step(self, flags, stage_times=None, dt=0.0, diffuse_amp_deg=0.0, r=0.0, rps=None, pairs_out=None):
        cap = pairs_out.shape[0] if pairs_out is not None else 0
        check(self.L.lm_step(self.h, int(flags), ctypes.byref(stage_times) if stage_times is not None else None,
                             float(dt), float(diffuse_amp_deg), float(r),
                             ctypes.byref(rps) if rps is not None else None, _ptr(pairs_out), cap, self._stream()),
              "lm_step")

    # ---- latitude strips (multi-GPU): staged step, the caller exchanges the buffers between the stages ----
    def strip_alloc(self, send_cap, ghost_cap, row_cap):
        check(self.L.lm_strip_alloc(self.h, int(send_cap), int(ghost_cap), int(row_cap)), "lm_strip_alloc")

    def set_strip(self, row0, rows_owned, has_south, has_north):
        st = Strip(int(row0), int(rows_owned), int(bool(has_south)), int(bool(has_north)))
        check(self.L.lm_set_strip(self.h, ctypes.byref(st)), "lm_set_strip")

    def peer_export(self):
        """lm_strip_peer_export as ``bytes`` (to hand to a neighbour, possibly through torch.distributed)."""
        e = _lib.PeerExport()
        check(self.L.lm_strip_peer_export(self.h, ctypes.byref(e)), "lm_strip_peer_export")
        return bytes(e)

    def peer_connect(self, side, export_bytes, use_ipc):
        """Connect to the neighbour on ``side`` (0 south, 1 north) whose ``peer_export()`` is ``export_bytes``."""
        e = _lib.PeerExport.from_buffer_copy(export_bytes)
        check(self.L.lm_strip_peer_connect(self.h, int(side), ctypes.byref(e), 1 if use_ipc else 0), "lm_strip_peer_connect")

    def step_push(self, kind):
        check(self.L.lm_step_push(self.h, int(kind), self._stream()), "lm_step_push")

    def strip_buffers(self):
        """Exchange buffers as uint8 CUDA tensors (no copy): dict name -> tensor / [south, north] pair."""
        b = StripBuffers()
        check(self.L.lm_strip_buffers_get(self.h, ctypes.byref(b)), "lm_strip_buffers_get")

        def view(ptr, nbytes):
            return torch.as_tensor(_CudaArray(int(ptr), int(nbytes), 1, "|u1"), device=self.device)
        return {"mig_send": [view(b.mig_send[k], b.mig_bytes) for k in range(2)],
                "mig_recv": [view(b.mig_recv[k], b.mig_bytes) for k in range(2)],
                "ghost_send": view(b.ghost_send, b.ghost_bytes), "ghost_recv": view(b.ghost_recv, b.ghost_bytes),
                "gsp_send": view(b.gsp_send, b.species_bytes), "gsp_recv": view(b.gsp_recv, b.species_bytes),
                "gret_send": view(b.gret_send, b.species_bytes), "gret_recv": view(b.gret_recv, b.species_bytes)}

    def step_move(self, flags, stage_times=None, dt=0.0, diffuse_amp_deg=0.0, rps=None):
        check(self.L.lm_step_move(self.h, int(flags), ctypes.byref(stage_times) if stage_times is not None else None,
                                  float(dt), float(diffuse_amp_deg), ctypes.byref(rps) if rps is not None else None,
                                  self._stream()), "lm_step_move")

    def step_bin(self):
        check(self.L.lm_step_bin(self.h, self._stream()), "lm_step_bin")

    def step_interact_begin(self, r=0.0, pairs_out=None):
        cap = pairs_out.shape[0] if pairs_out is not None else 0
        check(self.L.lm_step_interact_begin(self.h, float(r), _ptr(pairs_out), cap, self._stream()),
              "lm_step_interact_begin")

    def step_interact_end(self):
        check(self.L.lm_step_interact_end(self.h, self._stream()), "lm_step_interact_end")

    def step_finish(self):
        check(self.L.lm_step_finish(self.h, self._stream()), "lm_step_finish")

    def state_get(self, lon_out=None, lat_out=None, species_out=None):
        check(self.L.lm_state_get(self.h, _ptr(lon_out), _ptr(lat_out), _ptr(species_out), self._stream()), "lm_state_get")

    def state_get_host(self, lon_host=None, lat_host=None, species_host=None):
        """Outputs are pinned CPU tensors (or None); copies complete after ``host_copies_sync``."""
        for t in (lon_host, lat_host, species_host):
            assert t is None or (not t.is_cuda and t.is_pinned() and t.is_contiguous())
        check(self.L.lm_state_get_host(self.h, _ptr(lon_host), _ptr(lat_host), _ptr(species_host), self._stream()),
              "lm_state_get_host")

    def record_next_step(self, lon_host=None, lat_host=None, species_host=None):
        """Arm the next ``step`` to write its per-step record into pinned CPU tensors, overlapped with the step."""
        for t in (lon_host, lat_host, species_host):
            assert t is None or (not t.is_cuda and t.is_pinned() and t.is_contiguous())
        check(self.L.lm_record_next_step(self.h, _ptr(lon_host), _ptr(lat_host), _ptr(species_host)), "lm_record_next_step")

    def record_next_step_ids(self, ids_host, lon_host=None, lat_host=None, species_host=None):
        """Arm the next step to write its record in STORAGE order with the particle ids beside it (pinned CPU tensors with
        room for max_particles): the form that works for a latitude strip.  ``record_count()`` particles were written."""
        for t in (ids_host, lon_host, lat_host, species_host):
            assert t is None or (not t.is_cuda and t.is_pinned() and t.is_contiguous() and t.numel() >= self.max_particles)
        assert ids_host is not None and ids_host.dtype == torch.int32
        check(self.L.lm_record_next_step_ids(self.h, _ptr(ids_host), _ptr(lon_host), _ptr(lat_host), _ptr(species_host)),
              "lm_record_next_step_ids")

    def record_count(self):
        return int(self.L.lm_record_count(self.h))

    def host_copies_sync(self):
        check(self.L.lm_host_copies_sync(self.h), "lm_host_copies_sync")

    def state_view(self, rows=None):
        """Raw resident arrays in storage (cell, id) order as torch tensors (no copy): lon, lat, species, ids,
        cell_start (``rows`` = number of local cell rows of a strip; default: the whole grid)."""
        p = [ctypes.c_void_p() for _ in range(5)]
        check(self.L.lm_state_view(self.h, *[ctypes.byref(x) for x in p]), "lm_state_view")
        n = self.state_size()
        g = self.get_grid()
        rows = g.ncy if rows is None else int(rows)

        def view(ptr, count, dtype, itemsize):
            if count == 0:
                return torch.empty(0, dtype=dtype, device=self.device)
            arr = _CudaArray(ptr.value, count, itemsize, np.dtype({torch.float32: "f4", torch.int8: "i1", torch.int32: "i4"}[dtype]).str)
            return torch.as_tensor(arr, device=self.device)
        return (view(p[0], n, torch.float32, 4), view(p[1], n, torch.float32, 4), view(p[2], n, torch.int8, 1),
                view(p[3], n, torch.int32, 4), view(p[4], g.ncx * rows + 1, torch.int32, 4))

    # ---- status ----------------------------------------------------------------------------------
    def sync_stats(self, raise_on_overflow=True, allow_misrouted=False):
        st = Stats()
        rc = self.L.lm_sync_stats(self.h, ctypes.byref(st), self._stream())
        if rc == _lib.LM_ENOSPC and not raise_on_overflow:
            return st
        if rc == _lib.LM_ESTATE and allow_misrouted and st.n_misrouted > 0:
            return st
        check(rc, "lm_sync_stats")
        return st

    def reset_stats(self):
        check(self.L.lm_reset_stats(self.h, self._stream()), "lm_reset_stats")

    def phase_times(self):
        """[advect, bin, pair search, RPS resolution, stats] device milliseconds of the last step run with
        LM_STEP_TIMING."""
        ms = (ctypes.c_float * 5)()
        check(self.L.lm_phase_times(self.h, ms), "lm_phase_times")
        return list(ms)

    def set_option(self, option, value):
        check(self.L.lm_set_option(self.h, int(option), int(value)), "lm_set_option")
        if int(option) == _lib.LM_OPT_INTERACT_MODE:
            self.interact_mode = int(value)

    def set_norm(self, p):
        """The Minkowski norm of the radius query: p = 1, 2 (default) or math.inf (``query_pairs(r, p)``)."""
        self.set_option(_lib.LM_OPT_NORM, norm_code(p))

    def join(self):
        """Order the current stream after the RPS phases still running on the handle's internal stream."""
        check(self.L.lm_join(self.h, self._stream()), "lm_join")

    def launch_count(self):
        return int(self.L.lm_launch_count(self.h))


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr, count, itemsize, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}
