"""Philox4x32-10 known-answer vectors (Random123) and the 53-bit double construction."""
import numpy

from oracle import anymdp_oracle as ao

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_kats():
    for ctr, key, want in KATS:
        got = ao.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in got) == want


def test_philox_vectorised():
    ctrs = numpy.arange(7, dtype=numpy.uint64)
    got = ao.philox4x32_10((ctrs, 3, 9, 0), (5, 6))
    for i in range(7):
        one = ao.philox4x32_10((i, 3, 9, 0), (5, 6))
        assert [int(g[i]) for g in got] == [int(x) for x in one]


def test_u53():
    assert float(ao.u53(0x6627e8d5, 0xe169c58d)) == 0.39904647231489565
    assert float(ao.u53(0xffffffff, 0xffffffff)) == 1.0 - 2.0 ** -53
    assert float(ao.u53(0, 0)) == 0.0


def test_engine_draws_shape_and_ranges():
    U = ao.engine_draws(1234, numpy.arange(1000), 7, da=2, do=3, na=5)
    assert U.shape == (1000, 24)
    for k in range(3):
        for kind in (ao.SLOT_TRANS, ao.SLOT_RESET, ao.SLOT_OBS, ao.SLOT_RESET_OBS):
            col = U[:, kind + 8 * k]
            assert (col >= 0).all() and (col < 1).all()
        a = U[:, ao.SLOT_ACTION + 8 * k]
        assert set(numpy.unique(a)) <= set(range(5))
        z = U[:, ao.SLOT_Z + 8 * k]
        assert abs(z.mean()) < 0.15 and 0.85 < z.std() < 1.15
