"""The delta-packed position record (csrc/record.cu, lm_record_delta_pack) on the CPU: the kernel
EXECUTED on the emulator through the C ABI against oracle/record.py, and the round trip through the product's host
decoder (io.unpack_delta_record) -- bit-exact float32, NaNs / infinities / signed zeros / sign changes included, aligned
and unaligned arrays (vector and scalar kernel paths), ragged tails, empty input, escape-list overflow."""
import ctypes
import os
import sys

import numpy as np
import pytest

from oracle import record as orec

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))


@pytest.fixture(scope="module")
def abi():
    import emu_build
    from lagrangian_microbes_b200 import _lib
    return _lib.declare(ctypes.CDLL(emu_build.build()))


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _aligned(n, dtype, offset_items=0):
    """array of n items whose first item sits offset_items past a 64-byte boundary"""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + 128 + offset_items * item, dtype=np.uint8)
    start = (-raw.ctypes.data) % 64 + offset_items * item
    return raw[start:start + n * item].view(dtype)


def _records(n, seed):
    rng = np.random.default_rng(seed)
    prev_lon = (205 + 10 * rng.random(n)).astype(np.float32)
    prev_lat = (25 + 10 * rng.random(n)).astype(np.float32)
    lon = (prev_lon + 0.03 * (rng.random(n) - 0.5)).astype(np.float32)       # ~ +-1000 ulps at 210 degrees
    lat = (prev_lat + 0.03 * (rng.random(n) - 0.5)).astype(np.float32)
    if n >= 16:
        lon[1] = prev_lon[1] + 3.0                                           # far jump: escape
        lat[2] = -prev_lat[2]                                                # sign change: escape
        lon[3], prev_lon[3] = np.float32(1e-41), np.float32(-1e-41)          # subnormals either side of zero: small delta
        lat[4], prev_lat[4] = np.float32(-0.0), np.float32(0.0)              # signed zeros: delta -1
        lon[5] = np.nan
        lat[6], prev_lat[6] = np.inf, np.float32(3.0e38)
        prev_lon[7] = np.nan; lon[7] = 210.0
        lon[n - 1] = prev_lon[n - 1]                                         # delta 0 at the ragged tail
        k32767 = prev_lat[8].view(np.uint32) + np.uint32(32767)
        lat[8] = k32767.view(np.float32)                                     # the largest delta that still fits
        lat[9] = (prev_lat[9].view(np.uint32) + np.uint32(32768)).view(np.float32)   # the smallest that does not
    return prev_lon, prev_lat, lon, lat


def _run(abi, prev_lon, prev_lat, lon, lat, cap, offset=0):
    from lagrangian_microbes_b200 import _lib
    n = len(lon)
    bufs = []
    for a in (prev_lon, prev_lat, lon, lat):
        b = _aligned(n, np.float32, offset); b[:] = a; bufs.append(b)
    dlon, dlat = _aligned(n, np.int16, offset), _aligned(n, np.int16, offset)
    dlon[:] = 77; dlat[:] = 77
    esc = np.full((max(cap, 1), 2), 0xABCDABCD, dtype=np.uint32)
    cnt = np.full(1, 99, dtype=np.uint32)
    rc = abi.lm_record_delta_pack(*[_ptr(b) for b in bufs], n, _ptr(dlon), _ptr(dlat), _ptr(esc), cap, _ptr(cnt), None)
    assert rc == _lib.LM_OK
    return dlon.copy(), dlat.copy(), esc, int(cnt[0])


@pytest.mark.parametrize("n,offset", [(0, 0), (1, 0), (3, 0), (4, 0), (1023, 0), (1024, 0), (1030, 0), (1030, 1), (2049, 3)])
def test_delta_pack_executed_against_the_oracle_and_round_trip(abi, n, offset):
    from lagrangian_microbes_b200 import io as lmio
    prev_lon, prev_lat, lon, lat = _records(n, seed=n + offset)
    dlon, dlat, esc, m = _run(abi, prev_lon, prev_lat, lon, lat, cap=64, offset=offset)
    want_dlon, want_dlat, want_esc = orec.pack_reference(prev_lon, prev_lat, lon, lat)
    assert np.array_equal(dlon, want_dlon) and np.array_equal(dlat, want_dlat)
    assert m == len(want_esc) and sorted(map(tuple, esc[:m].tolist())) == want_esc
    if n >= 16:
        assert m >= 6 and dlat[8] == 32767 and dlat[9] == orec.ESCAPE and dlat[4] == -1
    got_lon, got_lat = lmio.unpack_delta_record(prev_lon, prev_lat, dlon, dlat, esc[:m])
    assert np.array_equal(got_lon.view(np.uint32), lon.view(np.uint32))       # bit for bit, NaN payloads included
    assert np.array_equal(got_lat.view(np.uint32), lat.view(np.uint32))
    # the library's own host decoder (lm_record_delta_unpack_host), separate outputs and in place
    out_lon, out_lat = np.full(n, 7, dtype=np.float32), np.full(n, 7, dtype=np.float32)
    e = np.ascontiguousarray(esc[:m])
    assert abi.lm_record_delta_unpack_host(_ptr(prev_lon), _ptr(prev_lat), _ptr(dlon), _ptr(dlat), _ptr(e), m, n,
                                           _ptr(out_lon), _ptr(out_lat), 3) == 0
    assert np.array_equal(out_lon.view(np.uint32), lon.view(np.uint32)) and np.array_equal(out_lat.view(np.uint32), lat.view(np.uint32))
    in_lon, in_lat = prev_lon.copy(), prev_lat.copy()
    assert abi.lm_record_delta_unpack_host(_ptr(in_lon), _ptr(in_lat), _ptr(dlon), _ptr(dlat), _ptr(e), m, n,
                                           _ptr(in_lon), _ptr(in_lat), 1) == 0
    assert np.array_equal(in_lon.view(np.uint32), lon.view(np.uint32)) and np.array_equal(in_lat.view(np.uint32), lat.view(np.uint32))


def test_escape_overflow_is_counted_not_written_past_the_list(abi):
    from lagrangian_microbes_b200 import io as lmio
    n = 500
    prev_lon, prev_lat, lon, lat = _records(n, seed=5)
    lon[100:140] += 5.0                                                       # 40 more escapes
    dlon, dlat, esc, m = _run(abi, prev_lon, prev_lat, lon, lat, cap=8)
    assert m > 8 and m == len(orec.pack_reference(prev_lon, prev_lat, lon, lat)[2])
    assert (esc[:8, 0] != 0xABCDABCD).all() and esc.shape[0] == 8
    with pytest.raises(ValueError):
        lmio.unpack_delta_record(prev_lon, prev_lat, dlon, dlat, esc[:8])
    from lagrangian_microbes_b200 import _lib
    out, out2 = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    e = np.ascontiguousarray(esc[:8])
    assert abi.lm_record_delta_unpack_host(_ptr(prev_lon), _ptr(prev_lat), _ptr(dlon), _ptr(dlat), _ptr(e), 8, n, _ptr(out), _ptr(out2),
                                           2) == _lib.LM_EINVAL
    bad = np.array([[2 * n + 1, 0]], dtype=np.uint32)                         # an entry outside the arrays
    assert abi.lm_record_delta_unpack_host(_ptr(prev_lon), _ptr(prev_lat), _ptr(dlon), _ptr(dlat), _ptr(bad), 1, n, _ptr(out), _ptr(out2),
                                           2) == _lib.LM_EINVAL


def test_host_decoder_of_the_shipped_library_many_threads():
    """The product's own .so on the CPU (a host function: no device call): 300,000 microbes over 4 threads, packed by the
    NumPy twin of the kernel's arithmetic (io._mono_key), decoded natively and by io.unpack_delta_record."""
    from lagrangian_microbes_b200 import io as lmio
    from lagrangian_microbes_b200.record import unpack_delta_record_native
    n = 300_000
    prev_lon, prev_lat, lon, lat = _records(n, seed=77)
    esc = []
    ds = []
    for c, (prev, cur) in enumerate(((prev_lon, lon), (prev_lat, lat))):
        d = lmio._mono_key(cur) - lmio._mono_key(prev)
        far = np.abs(d) > 32767
        ds.append(np.where(far, lmio.DELTA_ESCAPE, d).astype(np.int16))
        esc += [(2 * int(i) + c, int(cur.view(np.uint32)[i])) for i in np.flatnonzero(far)]
    esc = np.array(esc, dtype=np.uint32).reshape(-1, 2)
    a = unpack_delta_record_native(prev_lon, prev_lat, ds[0], ds[1], esc, n_threads=4)
    b = lmio.unpack_delta_record(prev_lon, prev_lat, ds[0], ds[1], esc)
    for got in (a, b):
        assert np.array_equal(got[0].view(np.uint32), lon.view(np.uint32)) and np.array_equal(got[1].view(np.uint32), lat.view(np.uint32))
    with pytest.raises(ValueError):
        unpack_delta_record_native(prev_lon, prev_lat, ds[0], ds[1], esc[:-1])


def test_delta_pack_rejects_bad_arguments(abi):
    from lagrangian_microbes_b200 import _lib
    a = np.zeros(8, dtype=np.float32); d = np.zeros(8, dtype=np.int16); e = np.zeros((4, 2), dtype=np.uint32); c = np.zeros(1, dtype=np.uint32)
    P = _ptr
    assert abi.lm_record_delta_pack(P(a), P(a), P(a), P(a), -1, P(d), P(d), P(e), 4, P(c), None) == _lib.LM_EINVAL
    assert abi.lm_record_delta_pack(P(a), P(a), P(a), P(a), 8, P(d), P(d), P(e), 4, None, None) == _lib.LM_EINVAL
    assert abi.lm_record_delta_pack(P(a), P(a), P(a), P(a), 8, P(d), P(d), None, 4, P(c), None) == _lib.LM_EINVAL
    assert abi.lm_record_delta_pack(None, P(a), P(a), P(a), 8, P(d), P(d), P(e), 4, P(c), None) == _lib.LM_EINVAL
    assert abi.lm_record_delta_pack(P(a), P(a), P(a), P(a), 1 << 31, P(d), P(d), P(e), 4, P(c), None) == _lib.LM_EINVAL


def test_round_trip_holds_for_arbitrary_bit_patterns(abi):
    """Property: for ANY pair of float32 bit patterns (NaN payloads, subnormals, both signs) pack -> decode returns the
    current record bit for bit, through the kernel (emulated) and both decoders; deltas agree with the oracle."""
    from hypothesis import given, settings, strategies as st
    from lagrangian_microbes_b200 import io as lmio

    bits = st.integers(min_value=0, max_value=2 ** 32 - 1)
    near = st.integers(min_value=-40000, max_value=40000)        # straddles the int16 limit

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(bits, bits, near, bits), min_size=1, max_size=40))
    def check(rows):
        prev_lon = np.array([r[0] for r in rows], dtype=np.uint32).view(np.float32)
        lon = np.array([r[1] for r in rows], dtype=np.uint32).view(np.float32)
        prev_lat = np.array([r[3] for r in rows], dtype=np.uint32).view(np.float32)
        # latitude: a neighbour of prev_lat in key space, clipped to the key range
        k = np.clip(lmio._mono_key(prev_lat) + np.array([r[2] for r in rows], dtype=np.int64), 0, 2 ** 32 - 1)
        lat = lmio._from_mono_key(k)
        n = len(rows)
        dlon, dlat, esc, m = _run(abi, prev_lon, prev_lat, lon, lat, cap=2 * n)
        w_dlon, w_dlat, w_esc = orec.pack_reference(prev_lon, prev_lat, lon, lat)
        assert np.array_equal(dlon, w_dlon) and np.array_equal(dlat, w_dlat) and sorted(map(tuple, esc[:m].tolist())) == w_esc
        a = lmio.unpack_delta_record(prev_lon, prev_lat, dlon, dlat, esc[:m])
        out = (np.empty(n, np.float32), np.empty(n, np.float32))
        e = np.ascontiguousarray(esc[:m])
        assert abi.lm_record_delta_unpack_host(_ptr(prev_lon), _ptr(prev_lat), _ptr(dlon), _ptr(dlat), _ptr(e), m, n,
                                               _ptr(out[0]), _ptr(out[1]), 1) == 0
        for got in (a, out):
            assert np.array_equal(got[0].view(np.uint32), lon.view(np.uint32)) and np.array_equal(got[1].view(np.uint32), lat.view(np.uint32))

    check()
