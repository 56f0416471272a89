"""TEST INFRASTRUCTURE ONLY (never imported by the product): CPU restatement of the delta-packed position record of
csrc/record.cu / lm_record_delta_pack.  The format is this framework's own (the reference stores plain float32 columns:
/root/reference/particle_advecter.py:233-235, interaction_simulator.py:108-110), so there is no reference vector to pin
it against (PARITY UNPINNED by construction); its contract is the round trip -- decoding must return the reference's float32 arrays bit for bit -- and
this restatement states the encoding independently of both the kernel and the product's decoder (struct-based, one
value at a time)."""
import struct

import numpy as np

ESCAPE = -32768


def key_of(x):
    """monotone integer key of one float32 value (all bit patterns)"""
    (b,) = struct.unpack("<I", struct.pack("<f", x)) if not isinstance(x, np.float32) else (int(np.float32(x).view(np.uint32)),)
    return (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)


def pack_reference(prev_lon, prev_lat, lon, lat):
    """-> dlon int16[n], dlat int16[n], escapes as a sorted list of (slot, raw bits)"""
    n = len(lon)
    dlon, dlat = np.zeros(n, dtype=np.int16), np.zeros(n, dtype=np.int16)
    esc = []
    for c, (prev, cur, out) in enumerate(((prev_lon, lon, dlon), (prev_lat, lat, dlat))):
        for i in range(n):
            d = key_of(np.float32(cur[i])) - key_of(np.float32(prev[i]))
            if -32767 <= d <= 32767:
                out[i] = d
            else:
                out[i] = ESCAPE
                esc.append((2 * i + c, int(np.float32(cur[i]).view(np.uint32))))
    return dlon, dlat, sorted(esc)
