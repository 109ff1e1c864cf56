"""numpy model of the index algebra of the one-exchange z kernels for nz = 256 (csrc/fft_sm100.cu: pass_z8f_kernel forward,
pass_z8i_kernel inverse): a quarter-warp owns a row, lane b < 8 holds x_c[b + 8 r] (r < 16) in registers, 128 = 16 x 8 with the
16-point stage over r, the twiddle W_128^(b q), ONE exchange (store at q * 9 + b, load at q' * 9 + b') and the 8-point stage over
the old lane index; the table-driven Hermitian pre- and r2c post-processing of pass_z16p_fused_kernel with H = 128.  Also the
forward half of the nz = 512 / 1024 kernels (pass_z16f / pass_z32f: the post tables, the radix-2 join).  CPU only: it pins the
derivations, the kernels themselves are checked on the GPU (tests/test_gpu_fft_sm100.py)."""
import numpy as np
import pytest

H, NZ = 128, 256


def dft(u, n, sign):
    """natural-order n-point DFT along the last axis; sign = -1: forward (e^{-i...}), +1: inverse (unnormalised)."""
    j = np.arange(n)
    return u @ np.exp(sign * 2j * np.pi * np.outer(j, j) / float(n))


def tables(nz, h, scale):
    k = np.arange(h // 2 + 1)
    w = np.exp(-2j * np.pi * k / nz)                       # twn: (c, -s)
    c, s = w.real, -w.imag
    half = 0.5 * scale
    post_a = half * ((1.0 - s) - 1j * c)                   # half * (1 + w.y, -w.x)
    post_b = half * ((1.0 + s) + 1j * c)                   # half * (1 - w.y,  w.x)
    pre_a = (1.0 - s) + 1j * c                             # (1 + w.y, w.x)
    pre_b = (1.0 + s) + 1j * c                             # (1 - w.y, w.x)
    return post_a, post_b, pre_a, pre_b


@pytest.mark.parametrize("kz_out", [1, 30, 34, 64])
def test_z8f_forward_index_algebra(kz_out):
    rng = np.random.default_rng(kz_out)
    x = rng.normal(size=NZ)
    scale = 0.83
    post_a, post_b, _, _ = tables(NZ, H, scale)
    b, r = np.meshgrid(np.arange(8), np.arange(16), indexing="ij")
    m = b + 8 * r
    u = x[2 * m] + 1j * x[2 * m + 1]                       # registers u[b, r]
    a = dft(u, 16, -1)                                     # stage 0 over r -> A[b, q]
    q = np.arange(16)[None, :]
    a = a * np.exp(-2j * np.pi * (np.arange(8)[:, None] * q) / H)       # W_128^(b q)
    buf = np.zeros(16 * 9, dtype=complex)
    buf[(q * 9 + np.arange(8)[:, None])] = a               # exchange: store at q * 9 + b
    V = np.zeros(H, dtype=complex)
    for c in range(8):                                     # lane c reads q = c and q = c + 8 for all eight b
        for j in range(2):
            qq = c + 8 * j
            v = dft(np.array([buf[qq * 9 + bb] for bb in range(8)]), 8, -1)   # stage 1 over b -> s
            V[qq + 16 * np.arange(8)] = v                  # natural-order copy: k = q + 16 s
    T = np.array([V[k] * post_a[k] + np.conj(V[0 if k == 0 else H - k]) * post_b[k] for k in range(kz_out)])
    ref = np.fft.rfft(x)[:kz_out] * scale
    assert np.allclose(T, ref, rtol=0, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("kz_in", [1, 12, 34, 64])
def test_z8i_inverse_index_algebra(kz_in):
    rng = np.random.default_rng(100 + kz_in)
    Z = np.zeros(H + 1, dtype=complex)
    Z[:kz_in] = rng.normal(size=kz_in) + 1j * rng.normal(size=kz_in)
    _, _, pre_a, pre_b = tables(NZ, H, 1.0)
    x_c = np.zeros((8, 16), dtype=complex)
    buf = np.zeros(16 * 9, dtype=complex)
    for c in range(8):                                     # lane c gathers Y[q + 16 s], q = c + 8 j
        for j in range(2):
            qq = c + 8 * j
            y = np.zeros(8, dtype=complex)
            for s in range(8):
                k = qq + 16 * s
                if k < kz_in:
                    v = Z[k].real if k == 0 else Z[k]
                    y[s] = v * pre_a[k]
                elif H - k < kz_in:
                    y[s] = np.conj(Z[H - k]) * pre_b[H - k]
            g = dft(y, 8, +1)                              # 8-point inverse over s -> index b
            g = g * np.exp(2j * np.pi * (np.arange(8) * qq) / H)       # conj W_128^(b q)
            buf[qq * 9 + np.arange(8)] = g                 # exchange: store at q * 9 + b'
    for b in range(8):
        u = np.array([buf[q * 9 + b] for q in range(16)])
        x_c[b] = dft(u, 16, +1)                            # 16-point inverse over q -> r: x_c[b + 8 r]
    x = np.empty(NZ)
    b, r = np.meshgrid(np.arange(8), np.arange(16), indexing="ij")
    x[2 * (b + 8 * r)] = x_c.real
    x[2 * (b + 8 * r) + 1] = x_c.imag
    Zfull = Z.copy()
    Zfull[0] = Zfull[0].real
    ref = np.fft.irfft(Zfull, n=NZ) * NZ
    assert np.allclose(x, ref, rtol=0, atol=1e-10 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("nz", [512, 1024])
def test_forward_post_tables_and_radix2_join(nz):
    """pass_z16f (nz = 512): T[k] = V[k] postA[k] + conj(V[H-k]) postB[k] from the half-length transform V of x_c;
    pass_z32f (nz = 1024): the same V from the two 256-point transforms of the even / odd samples of x_c, joined by
    V[j] = E[j] + W_512^j O[j], V[j + 256] = E[j] - W_512^j O[j], with the 1/2 folded into the input."""
    rng = np.random.default_rng(nz)
    x = rng.normal(size=nz)
    h = nz // 2
    scale = 1.7
    xc = x[0::2] + 1j * x[1::2]
    ref = np.fft.rfft(x)[:h // 2] * scale
    if nz == 512:
        V = np.fft.fft(xc)
        post_a, post_b, _, _ = tables(nz, h, scale)
        T = np.array([V[k] * post_a[k] + np.conj(V[0 if k == 0 else h - k]) * post_b[k] for k in range(h // 2)])
    else:
        half = 0.5 * scale
        E, O = np.fft.fft(half * xc[0::2]), np.fft.fft(half * xc[1::2])
        j = np.arange(256)
        t = O * np.exp(-2j * np.pi * j / 512.0)
        V = np.concatenate([E + t, E - t])
        k = np.arange(h // 2)
        v2 = np.conj(V[(h - k) % h])
        ev, dv = V[k] + v2, V[k] - v2
        T = ev + (-1j * dv) * np.exp(-2j * np.pi * k / nz)
    assert np.allclose(T, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
