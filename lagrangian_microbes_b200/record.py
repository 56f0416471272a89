"""The per-step position record over PCIe at 4 B instead of 8 B per microbe-step ("on-GPU
quantise/delta-pack"): ``lm_record_delta_pack`` (csrc/record.cu) sends step k as int16 ulp differences to step k-1,
lossless, and ``lm_record_delta_unpack_host`` (host threads; NumPy twin: ``io.unpack_delta_record``) restores the float32 arrays the reference stores
(/root/reference/particle_advecter.py:233-235, interaction_simulator.py:108-110) bit for bit.

    packer = DeltaRecordPacker(n)                    # device + pinned buffers, two sets
    for k in range(steps):
        sim.step(); sim.engine.state_get(lon_d, lat_d, sp_d)
        packer.push(lon_d, lat_d)                    # async: pack + D2H on the current stream
        ...                                          # (the next step may be enqueued here)
        lon, lat = packer.pop()                      # float32 numpy, bit-exact

The first record and any step whose escape list overflows travel as plain float32 (a key frame).  Escapes are not
only far jumps: within about a degree of the equator (or of longitude 0) float32 ulps are so fine that one step's
displacement exceeds 32767 of them, so those microbes always go through the list -- 1-2 % of BASELINE config 4's microbes
(lat 0-60); the default list holds n / 16 entries (0.5 B per microbe of D2H)."""
import ctypes

import numpy as np
import torch

from . import _lib


def unpack_delta_record_native(prev_lon, prev_lat, dlon, dlat, escapes, n_threads=None, out=None):
    """``io.unpack_delta_record`` through the library's multi-threaded host decoder (``lm_record_delta_unpack_host``):
    float32 numpy arrays (lon, lat), bit-exact.  ``out`` = (lon, lat) arrays to fill (may be the ``prev`` arrays)."""
    import os
    L = _lib.lib()
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    prev_lon, prev_lat = c(prev_lon, np.float32), c(prev_lat, np.float32)
    dlon, dlat = c(dlon, np.int16), c(dlat, np.int16)
    esc = c(escapes, np.uint32).reshape(-1, 2)
    n = prev_lon.size
    assert prev_lat.size == n and dlon.size == n and dlat.size == n
    lon, lat = out if out is not None else (np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32))
    assert lon.dtype == lat.dtype == np.float32 and lon.size == lat.size == n and lon.flags.c_contiguous and lat.flags.c_contiguous
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    threads = int(n_threads) if n_threads else max(1, min(16, (os.cpu_count() or 1)))
    rc = L.lm_record_delta_unpack_host(p(prev_lon), p(prev_lat), p(dlon), p(dlat), p(esc), esc.shape[0], n, p(lon), p(lat), threads)
    if rc == _lib.LM_EINVAL:
        raise ValueError("delta record: escape markers and escape entries disagree (list overflowed?)")
    _lib.check(rc, "lm_record_delta_unpack_host")
    return lon, lat


class DeltaRecordPacker:
    def __init__(self, n, escape_capacity=None, device=None):
        self.n = int(n)
        self.cap = int(escape_capacity if escape_capacity is not None else max(1024, self.n // 16))
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.L = _lib.lib()
        self.prev = None                                       # device (lon, lat) of the last pushed record
        self.host_prev = None                                  # the same, decoded on the host
        mk = lambda dt, *shape: torch.empty(*shape, dtype=dt, device=dev)
        pin = lambda dt, *shape: torch.empty(*shape, dtype=dt).pin_memory()
        self.keep = [(mk(torch.float32, self.n), mk(torch.float32, self.n)) for _ in range(2)]
        self.d_dev = [(mk(torch.int16, self.n), mk(torch.int16, self.n)) for _ in range(2)]
        self.esc_dev = [mk(torch.int32, self.cap, 2) for _ in range(2)]
        self.cnt_dev = [mk(torch.int32, 1) for _ in range(2)]
        self.d_host = [(pin(torch.int16, self.n), pin(torch.int16, self.n)) for _ in range(2)]
        self.esc_host = [pin(torch.int32, self.cap, 2) for _ in range(2)]
        self.cnt_host = [pin(torch.int32, 1) for _ in range(2)]
        self.key_host = [(pin(torch.float32, self.n), pin(torch.float32, self.n)) for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.pending = []                                      # (slot, kind) in push order
        self.pushed = 0
        self.bytes_d2h = 0

    def push(self, lon_dev, lat_dev):
        """Enqueue the record of one step (device float32 [n], particle-id order) on the current stream."""
        assert len(self.pending) < 2, "pop() before pushing a third record"
        assert lon_dev.is_cuda and lat_dev.is_cuda and lon_dev.dtype == lat_dev.dtype == torch.float32
        assert lon_dev.numel() == lat_dev.numel() == self.n and lon_dev.is_contiguous() and lat_dev.is_contiguous()
        slot = self.pushed % 2
        keep_lon, keep_lat = self.keep[slot]
        keep_lon.copy_(lon_dev); keep_lat.copy_(lat_dev)        # the caller's arrays may be overwritten by the next step
        if self.prev is None:
            self._key_frame(slot)
        else:
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            dl, da = self.d_dev[slot]
            _lib.check(self.L.lm_record_delta_pack(p(self.prev[0]), p(self.prev[1]), p(keep_lon), p(keep_lat), self.n, p(dl), p(da),
                                                   p(self.esc_dev[slot]), self.cap, p(self.cnt_dev[slot]), stream),
                       "lm_record_delta_pack")
            self.d_host[slot][0].copy_(dl, non_blocking=True)
            self.d_host[slot][1].copy_(da, non_blocking=True)
            self.esc_host[slot].copy_(self.esc_dev[slot], non_blocking=True)
            self.cnt_host[slot].copy_(self.cnt_dev[slot], non_blocking=True)
            self.pending.append((slot, "delta"))
            self.bytes_d2h += 4 * self.n + 8 * self.cap + 4
        self.prev = (keep_lon, keep_lat)
        self.events[slot].record()
        self.pushed += 1

    def _key_frame(self, slot):
        self.key_host[slot][0].copy_(self.keep[slot][0], non_blocking=True)
        self.key_host[slot][1].copy_(self.keep[slot][1], non_blocking=True)
        self.pending.append((slot, "key"))
        self.bytes_d2h += 8 * self.n

    def pop(self, decode=True):
        """The oldest pushed record as float32 numpy arrays (lon, lat), bit-exact.  ``decode=False`` hands out the record
        as it crossed the link instead -- ("key", lon, lat) or ("delta", dlon, dlat, escapes), views of the pinned buffers,
        valid until the second push from now -- for callers that store the packed stream and decode on read."""
        slot, kind = self.pending.pop(0)
        self.events[slot].synchronize()
        if kind == "delta" and int(self.cnt_host[slot][0]) > self.cap:
            self._key_frame(slot)                              # escape list overflowed: resend this step as a key frame
            self.pending.pop()
            torch.cuda.current_stream().synchronize()
            kind = "key"
        if kind == "delta":
            m = int(self.cnt_host[slot][0])
            esc = self.esc_host[slot].numpy()[:m].view(np.uint32)
            dlon, dlat = self.d_host[slot][0].numpy(), self.d_host[slot][1].numpy()
            if not decode:
                self.host_prev = None
                return "delta", dlon, dlat, esc
            assert self.host_prev is not None, "decode=True after decode=False: the previous record was not decoded"
            lon, lat = unpack_delta_record_native(self.host_prev[0], self.host_prev[1], dlon, dlat, esc)
        else:
            if not decode:
                self.host_prev = None
                return "key", self.key_host[slot][0].numpy(), self.key_host[slot][1].numpy()
            lon, lat = self.key_host[slot][0].numpy().copy(), self.key_host[slot][1].numpy().copy()
        self.host_prev = (lon, lat)
        return lon, lat
