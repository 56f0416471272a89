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
