import numpy as np

from oracle import philox as ph


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, o in kat:
        r = ph.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))
        assert [int(x) for x in r] == list(o)


def test_normals_moments_and_sharding_invariance():
    z = ph.normals(42, 0, np.arange(20000), 3, 100)
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3 and abs((z ** 4).mean() - 3) < 3e-2
    a = ph.xi_tensor(7, 1, 0, 16, 6, 5)
    b = np.concatenate([ph.xi_tensor(7, 1, 0, 8, 6, 5), ph.xi_tensor(7, 1, 8, 8, 6, 5)])
    assert np.array_equal(a, b)
    assert not np.array_equal(a, ph.xi_tensor(7, 2, 0, 16, 6, 5))
