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


def test_device_table_and_parallel_combination():
    """The generated header (csrc/mt_jump_tab.h) holds Q_k = x^(2^k L0) mod phi; the device kernel's formulation
    h[j] = XOR_{i in Q} x[i + j] reproduces numpy's state after u * 512 blocks for u = 1, 2, 3 (binary composition)."""
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectra_without_windows_b200", "csrc",
                       "mt_jump_tab.h")
    polys = mj.read_header(hdr)
    assert len(polys) == mj.K_TAB and all(p.bit_length() <= 19937 for p in polys)
    phi = mj.char_poly()
    assert polys[0] == mj.pow_x_mod(mj.L0, phi) and polys[1] == mj.pow_x_mod(2 * mj.L0, phi)
    key0 = np.random.RandomState(2024).get_state()[1]
    same = lambda got, ref: np.array_equal(got[1:], ref[1:]) and (int(got[0]) ^ int(ref[0])) & 0x80000000 == 0
    w1 = mj.combine(key0, polys[0])
    assert same(w1, numpy_key_after_blocks(2024, mj.SUB_BLOCKS))
    w2 = mj.combine(key0, polys[1])
    assert same(w2, numpy_key_after_blocks(2024, 2 * mj.SUB_BLOCKS))
    w3 = mj.combine(w2, polys[0])                       # u = 3 = 0b11: the rounds compose
    assert same(w3, numpy_key_after_blocks(2024, 3 * mj.SUB_BLOCKS))
