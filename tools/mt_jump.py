"""MT19937 jump-ahead (Haramoto et al. 2008) -- groundwork for splitting ONE numpy-legacy Gaussian field over many
CTAs ("Next", item 3).  GF(2) polynomials are Python integers (bit i = coefficient of x^i).

    phi = char_poly()                    characteristic polynomial of the one-word state transition (degree 19937)
    g   = pow_x_mod(J, phi)              x^J mod phi: the jump by J words is g(F) applied to the state (Horner)
    st  = jump_state(key, g)             624-word window after J one-word steps of `key` (numpy RandomState key array)

Checked against numpy's own generator in tests/test_mt_jump_cpu.py.  Nothing on the product path imports this yet."""
import numpy as np

N, M = 624, 397
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF
DEG = 19937


def generate(window, count):
    """The 624 words of `window` followed by the next `count` words of the MT19937 recurrence (untempered)."""
    out = [int(v) for v in window] + [0] * count
    for n in range(count):
        y = (out[n] & UPPER) | (out[n + 1] & LOWER)
        out[n + N] = out[n + M] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    return out


def step_words(window, count):
    """Advance a 624-word window by `count` one-word steps of the MT19937 recurrence; returns the new window."""
    return np.asarray(generate(window, count)[count:], dtype=np.uint32)


def char_poly():
    """Berlekamp-Massey over GF(2) on 2*19937 + 64 terms of one output bit of the recurrence -> minimal polynomial of
    the sequence = characteristic polynomial of the one-word transition (the generator has full period 2^19937 - 1)."""
    key = np.random.RandomState(12345).get_state()[1]
    words = generate(key, 2 * DEG + 64)[N:]            # generated words only (word 0 of a seeded key is not all state)
    bits = [v & 1 for v in words]
    C, B, L, m = 1, 1, 0, 1                            # connection polynomials, bit i = c_i (c_0 = 1)
    R = 0                                              # bit i = s_{n-i}
    for n, bit in enumerate(bits):
        R = (R << 1) | bit
        d = bin(C & R).count("1") & 1                  # discrepancy sum_{i=0..L} c_i s_{n-i}
        if d:
            T = C
            C ^= B << m
            if 2 * L <= n:
                L, B, m = n + 1 - L, T, 1
            else:
                m += 1
        else:
            m += 1
    assert L == DEG, L
    return _reciprocal(C, DEG)                         # sum_i c_i x^(L - i)


def _reciprocal(C, deg):
    return int(bin(C)[2:].zfill(deg + 1)[::-1], 2)


def mulmod(a, b, phi):
    """carry-less a*b mod phi"""
    r = 0
    while b:
        low = b & -b
        r ^= a << (low.bit_length() - 1)
        b ^= low
    return _mod(r, phi)


def _mod(r, phi):
    dp = phi.bit_length() - 1
    while r.bit_length() - 1 >= dp:
        r ^= phi << (r.bit_length() - 1 - dp)
    return r


def pow_x_mod(J, phi):
    """x^J mod phi by square-and-multiply (squaring = spreading the bits apart)."""
    result, base = 1, 2          # 1, x
    while J:
        if J & 1:
            result = mulmod(result, base, phi)
        J >>= 1
        if J:
            base = _mod(_square(base), phi)
    return result


def _square(a):
    s = bin(a)[2:]
    return int("0".join(s), 2)


def jump_state(key, g):
    """Horner evaluation of g(F) on the window `key`: h <- F h (+ key if the coefficient is set), from the top coefficient
    down.  F is the one-word step.  The low 31 bits of word 0 of the result are not part of the state (don't care)."""
    key = np.asarray(key, dtype=np.uint32)
    h = np.zeros(N, dtype=np.uint32)
    started = False
    for i in range(g.bit_length() - 1, -1, -1):
        if started:
            h = step_words(h, 1)
        if (g >> i) & 1:
            h = h ^ key
            started = True
    return h
