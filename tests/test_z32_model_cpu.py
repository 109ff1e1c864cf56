"""numpy model of the index algebra of pass_z32_fused_kernel (csrc/fft_sm100.cu; nz = 1024, one row per warp): the two
half-warps (p = lane >> 4) run the even / odd 256-point transforms of a 512-point half-length FFT in the two-exchange
layout of the nz = 512 kernel; a radix-2 step joins them (decimation in frequency on the way in, in time on the way
out, through one xor-16 shuffle), and the r2c post-processing pairs V[k] with V[512-k].  CPU only."""
import numpy as np
import pytest

from test_z16_model_cpu import two_stage

H, NZ = 512, 1024


@pytest.mark.parametrize("kz_in,kz_out", [(122, 122), (17, 122), (256, 200), (1, 128)])
def test_z32_index_algebra(kz_in, kz_out):
    rng = np.random.default_rng(100 + kz_in)
    Z = np.zeros(H + 1, dtype=complex)
    Z[:kz_in] = rng.normal(size=kz_in) + 1j * rng.normal(size=kz_in)
    w = rng.normal(size=NZ)
    scale = 1.7
    b, r = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    n = b + 16 * r                                     # 0..255, register r of lane b in BOTH half-warps

    def Y(k):                                          # Hermitian pre-processing, half-length index k < 512
        zk = np.where(k < kz_in, Z[np.minimum(k, H)], 0.0)
        zh = np.where(H - k < kz_in, Z[H - k], 0.0)
        zk = np.where(k == 0, zk.real, zk)
        zh = np.where(k == 0, zh.real, zh)
        return (zk + np.conj(zh)) + 1j * (zk - np.conj(zh)) * np.exp(2j * np.pi * k / NZ)

    y0, y1 = Y(n), Y(n + 256)
    regs = [y0 + y1, (y0 - y1) * np.exp(2j * np.pi * n / 512.0)]      # p = 0: A[n];  p = 1: B[n] tw[n]
    xc = [two_stage(regs[p], +1) for p in (0, 1)]                     # x_c[2q + p] at q = b + 16 r
    x = np.empty(NZ)
    for p in (0, 1):
        x[4 * n + 2 * p] = xc[p].real
        x[4 * n + 2 * p + 1] = xc[p].imag
    Zfull = Z.copy()
    Zfull[0] = Zfull[0].real
    assert np.allclose(x, np.fft.irfft(Zfull, n=NZ) * NZ, rtol=0, atol=1e-10 * np.abs(x).max())
    half = 0.5 * scale
    EO = [two_stage(xc[p].real * w[4 * n + 2 * p] * half + 1j * xc[p].imag * w[4 * n + 2 * p + 1] * half, -1) for p in (0, 1)]
    tw = np.exp(-2j * np.pi * n / 512.0)
    V = [EO[0] + tw * EO[1], EO[0] - tw * EO[1]]       # after the xor-16 exchange: p = 0 holds V[j], p = 1 holds V[j + 256]
    T = np.zeros(kz_out, dtype=complex)
    for bb in range(16):
        for rr in range(16):
            kk = bb + 16 * rr
            if kk >= kz_out:
                continue
            v1 = V[0][bb, rr]
            if kk == 0:
                v2 = V[0][0, 0]                        # V[512] = V[0]
            elif bb == 0:
                v2 = V[1][0, 16 - rr]                  # V[512 - k] = V[(256 - k) + 256], 256 - k = 16 (16 - r)
            else:
                v2 = V[1][16 - bb, 15 - rr]
            ev, dv = v1 + np.conj(v2), v1 - np.conj(v2)
            T[kk] = ev + (-1j * dv) * np.exp(-2j * np.pi * kk / NZ)
    ref = np.fft.rfft(x * w)[:kz_out] * scale
    assert np.allclose(T, ref, rtol=0, atol=1e-10 * np.abs(ref).max())
