"""MT19937 jump-ahead groundwork (tools/mt_jump.py) against numpy's own legacy generator.  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import mt_jump as mj  # noqa: E402


def numpy_key_after_blocks(seed, blocks):
    rs = np.random.RandomState(seed)
    rs.randint(0, 2 ** 32, size=624 * blocks, dtype=np.uint32)     # one raw word per draw: regenerates `blocks` times
    return rs.get_state()[1]


def test_step_words_is_numpys_recurrence():
    key0 = np.random.RandomState(7).get_state()[1]
    assert np.array_equal(mj.step_words(key0, 624 * 3), numpy_key_after_blocks(7, 3))


def test_jump_polynomial_reproduces_sequential_generation():
    phi = mj.char_poly()
    assert phi.bit_length() - 1 == 19937
    for seed, blocks in ((1, 5), (4000000000, 40)):
        J = 624 * blocks
        g = mj.pow_x_mod(J, phi)
        key0 = np.random.RandomState(seed).get_state()[1]
        got = mj.jump_state(key0, g)
        ref = numpy_key_after_blocks(seed, blocks)
        assert np.array_equal(got[1:], ref[1:]) and (int(got[0]) ^ int(ref[0])) & 0x80000000 == 0
