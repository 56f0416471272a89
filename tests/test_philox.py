"""Philox4x32-10 restatement against Random123's published known-answer vectors."""
import numpy as np

from oracle import philox


def test_kat_vectors():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for c, k, want in kat:
        got = philox.philox4x32_10(*[np.array([v]) for v in c], *k)
        assert tuple(int(x[0]) for x in got) == want


def test_uniform_range_and_determinism():
    i = np.arange(0, 100000, dtype=np.int64)
    u = philox.pair_uniforms(i, i + 1, step=3, seed=9)
    assert u.dtype == np.float64 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(u, philox.pair_uniforms(i, i + 1, step=3, seed=9))
    assert not np.array_equal(u, philox.pair_uniforms(i, i + 1, step=4, seed=9))
    # 64-bit step / seed use both key / counter words
    assert not np.array_equal(philox.pair_uniforms(i, i + 1, 1, 1), philox.pair_uniforms(i, i + 1, 1 + 2**32, 1))
    assert not np.array_equal(philox.pair_uniforms(i, i + 1, 1, 1), philox.pair_uniforms(i, i + 1, 1, 1 + 2**32))


def test_philox2x32_kat_vectors():
    """Random123 kat_vectors, philox2x32 with 10 rounds."""
    kat = [((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
           ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
           ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]
    for c, k, want in kat:
        got = philox.philox2x32_10(np.array([c[0]]), np.array([c[1]]), k)
        assert tuple(int(x[0]) for x in got) == want


def test_pair_stream_is_philox2x32_keyed_per_step():
    i, j = np.array([5, 7, 2**31 + 3]), np.array([9, 8, 2**32 - 1])
    key = philox.pair_stream_key(12, 2**40 + 5)
    assert 0 <= key < 2**32 and key != philox.pair_stream_key(13, 2**40 + 5) != philox.pair_stream_key(12, 2**40 + 6)
    x0, x1 = philox.philox2x32_10(i, j, key)
    assert np.array_equal(philox.pair_uniforms(i, j, 12, 2**40 + 5), philox.u53(x0, x1))
