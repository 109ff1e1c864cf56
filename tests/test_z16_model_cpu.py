"""numpy model of the index algebra of pass_z16_fused_kernel (csrc/fft_sm100.cu): lane (b < 16) / register (r < 16)
layout, the single exchange per transform (store at 16 b + r, load at b + 16 r), the pruned Hermitian pre-processing
and the r2c post-processing that pairs V[k] with V[H-k] from lane 16 - b, register 15 - r.  CPU only: it pins the
derivation, the kernel itself is checked on the GPU (tests/test_gpu_fft_sm100.py)."""
import numpy as np
import pytest

H, NZ = 256, 512


def dft16(u, sign):
    """natural-order 16-point DFT along the last axis; sign = -1: e^{-i...} (forward), +1: inverse (unnormalised)."""
    j = np.arange(16)
    return u @ np.exp(sign * 2j * np.pi * np.outer(j, j) / 16.0)


def two_stage(regs, sign):
    """regs[b, r] = element b + 16 r  ->  out[b, r] = transform at index b + 16 r."""
    t = dft16(regs, sign)                              # stage 0, no twiddles
    buf = np.empty(H, dtype=complex)
    b, r = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    buf[16 * b + r] = t                                # exchange: store at 16 b + r ...
    u = buf[b + 16 * r]                                # ... load at b + 16 r
    u = u * np.exp(sign * 2j * np.pi * (b * r) / H)    # inter-stage twiddles W_256^(b r)
    return dft16(u, sign)


@pytest.mark.parametrize("kz_in,kz_out", [(124, 124), (30, 124), (1, 64), (128, 128)])
def test_z16_index_algebra(kz_in, kz_out):
    rng = np.random.default_rng(kz_in)
    Z = np.zeros(H + 1, dtype=complex)
    Z[:kz_in] = rng.normal(size=kz_in) + 1j * rng.normal(size=kz_in)
    w = rng.normal(size=NZ)
    scale = 0.37
    b, r = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    k = b + 16 * r
    # Hermitian pre-processing of the staged row (only k or H-k below kz_in contribute)
    zk = np.where(k < kz_in, Z[np.minimum(k, H)], 0.0)
    zh = np.where(H - k < kz_in, Z[H - k], 0.0)
    zk = np.where(k == 0, zk.real, zk)
    zh = np.where(k == 0, zh.real, zh)
    a, d = zk + np.conj(zh), zk - np.conj(zh)
    y = a + 1j * d * np.exp(2j * np.pi * k / NZ)
    xc = two_stage(y, +1)                              # x[2m] + i x[2m+1] at m = b + 16 r
    x = np.empty(NZ)
    x[2 * k] = xc.real
    x[2 * k + 1] = xc.imag
    Zfull = Z.copy()
    Zfull[0] = Zfull[0].real
    assert np.allclose(x, np.fft.irfft(Zfull, n=NZ) * NZ, rtol=0, atol=1e-10 * np.abs(x).max())
    # weights (1/2 of the post-processing and the caller's scale folded in), forward transform from the same registers
    half = 0.5 * scale
    v = two_stage(xc.real * w[2 * k] * half + 1j * xc.imag * w[2 * k + 1] * half, -1)   # V[k], k = b + 16 r
    T = np.zeros(kz_out, dtype=complex)
    for bb in range(16):
        for rr in range(8):                            # only k < 128 is produced
            kk = bb + 16 * rr
            if kk >= kz_out:
                continue
            v1 = v[bb, rr]
            v2 = v[0, (16 - rr) & 15] if bb == 0 else v[16 - bb, 15 - rr]      # V[H - k]: own lane for b = 0
            ev, dv = v1 + np.conj(v2), v1 - np.conj(v2)
            T[kk] = ev + (-1j * dv) * np.exp(-2j * np.pi * kk / NZ)
    ref = np.fft.rfft(x * w)[:kz_out] * scale
    assert np.allclose(T, ref, rtol=0, atol=1e-10 * np.abs(ref).max())
