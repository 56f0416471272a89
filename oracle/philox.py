"""Philox counter-based RNGs (4x32-10 and 2x32-10), NumPy restatement (TEST INFRASTRUCTURE ONLY).

The reference draws its per-pair random number from NumPy's global MT19937
stream (``np.random.rand()``, /root/reference/interactions.py:20), one draw per
pair whose species differ, in CPython-set iteration order.  That stream cannot be
reproduced in parallel, so parity is defined with an injected *per-pair* stream
(SURVEY.md §8c): ``u(i, j, step, seed)`` below.  The CUDA resolver evaluates the
same function on the device (csrc/philox.cuh); tests compare them bit for bit.

Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
Philox-4x32 and Philox-2x32 with 10 rounds; known-answer vectors from Random123's
kat_vectors are checked in tests/test_philox.py.

Per-pair stream (shared with the device code, csrc/philox.cuh):
    key(seed, step) = word 0 of Philox4x32-10(counter = (step & 0xffffffff, step >> 32, 'RPS1', 0),
                                              key = (seed & 0xffffffff, seed >> 32))
    (x0, x1)        = Philox2x32-10(counter = (i, j), key = key(seed, step))     i < j the pair's particle ids
    u               = ((x0 >> 5) * 2**26 + (x1 >> 6)) / 2**53                    (53-bit, in [0, 1))
Philox2x32 yields exactly the 64 bits one draw needs at half the multiplications of Philox4x32.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays.

    Returns four uint32 arrays (x0, x1, x2, x3).
    """
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0            # 32x32 -> 64 bit products, exact in uint64
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def u53(x0, x1):
    """Two uint32 words -> float64 in [0, 1) with 53 random bits."""
    a = (np.asarray(x0, dtype=np.uint64) >> np.uint64(5)).astype(np.float64)
    b = (np.asarray(x1, dtype=np.uint64) >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


_M2 = np.uint64(0xD256D193)


def philox2x32_10(c0, c1, k):
    """Vectorised Philox2x32-10.  c0, c1 broadcastable uint32-valued arrays, k a 32-bit key.  Returns (x0, x1)."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c0, c1 = np.broadcast_arrays(c0, c1)
    k = int(k) & 0xFFFFFFFF
    for _ in range(10):
        p = _M2 * c0
        hi, lo = p >> _S32, p & _MASK
        c0, c1 = hi ^ np.uint64(k) ^ c1, lo
        k = (k + _W0) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32)


PAIR_STREAM_TAG = 0x52505331          # 'RPS1'


def pair_stream_key(step, seed):
    """The 32-bit Philox2x32 key of one step's per-pair stream."""
    step, seed = int(step), int(seed)
    x0, _, _, _ = philox4x32_10(step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, PAIR_STREAM_TAG, 0,
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return int(x0)


def pair_uniforms(i, j, step, seed):
    """Per-pair uniform u(i, j, step, seed) in [0,1), float64.  Requires i < j elementwise."""
    i = np.asarray(i, dtype=np.uint64)
    j = np.asarray(j, dtype=np.uint64)
    x0, x1 = philox2x32_10(i, j, pair_stream_key(step, seed))
    return u53(x0, x1)


def particle_uniforms(pid, step, seed, stream):
    """Per-particle uniforms for the diffusion kick: returns (u_a, u_b) in [0,1).

    counter = (pid, stream, step_lo, step_hi); key = seed.  ``stream`` separates this
    use from the per-pair stream (pairs always have c1 = j > i >= 0, diffusion uses
    c1 = 0xD1FF0000 | stream).
    """
    pid = np.asarray(pid, dtype=np.uint64)
    step = int(step)
    seed = int(seed)
    x0, x1, x2, x3 = philox4x32_10(pid, 0xD1FF0000 | (int(stream) & 0xFFFF), step & 0xFFFFFFFF,
                                   (step >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return u53(x0, x1), u53(x2, x3)
