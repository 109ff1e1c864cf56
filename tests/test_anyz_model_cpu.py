"""numpy model of the any-length z pass (csrc/fft_any_passes.cuh: gpass_z_kernel): two real rows A, B of ANY length n (odd lengths
included) ride one complex transform, z = x_A + i x_B.  Inverse: Z[k] = X_A[k] + i X_B[k] over the Hermitian extension of the two
pruned half-spectra, the imaginary parts of the self-conjugate modes dropped; forward: X_A[k] = (Z[k] + conj Z[n-k]) / 2,
X_B[k] = -i (Z[k] - conj Z[n-k]) / 2.  CPU only."""
import numpy as np
import pytest


@pytest.mark.parametrize("n,kz_in,kz_out", [(15, 8, 8), (20, 11, 6), (175, 40, 34), (274, 124, 124), (182, 1, 92)])
def test_packed_pair_z_pass(n, kz_in, kz_out):
    rng = np.random.default_rng(n)
    nzh = n // 2 + 1
    QA = np.zeros(nzh, dtype=complex)
    QB = np.zeros(nzh, dtype=complex)
    QA[:kz_in] = rng.normal(size=kz_in) + 1j * rng.normal(size=kz_in)
    QB[:kz_in] = rng.normal(size=kz_in) + 1j * rng.normal(size=kz_in)
    Z = np.zeros(n, dtype=complex)
    for k in range(nzh):
        xa, xb = QA[k], QB[k]
        self_conj = k == 0 or 2 * k == n
        if self_conj:
            xa, xb = xa.real, xb.real
        Z[k] = complex(xa.real - xb.imag, xa.imag + xb.real)            # X_A + i X_B
        if not self_conj:
            Z[n - k] = complex(xa.real + xb.imag, xb.real - xa.imag)    # conj X_A + i conj X_B
    z = np.fft.ifft(Z) * n                                              # unnormalised inverse
    refA, refB = np.fft.irfft(QA, n=n) * n, np.fft.irfft(QB, n=n) * n   # irfft drops the same imaginary parts
    assert np.allclose(z.real, refA, atol=1e-10 * np.abs(refA).max()) and np.allclose(z.imag, refB, atol=1e-10 * np.abs(refB).max())
    w = rng.normal(size=(2, n))
    zf = np.fft.fft(z.real * w[0] + 1j * z.imag * w[1])
    k = np.arange(kz_out)
    a, b = zf[k], zf[(n - k) % n]
    XA = 0.5 * (a + np.conj(b))
    XB = 0.5 * ((a.imag + b.imag) - 1j * (a.real - b.real))
    assert np.allclose(XA, np.fft.rfft(refA * w[0])[:kz_out], atol=1e-9 * np.abs(XA).max())
    assert np.allclose(XB, np.fft.rfft(refB * w[1])[:kz_out], atol=1e-9 * np.abs(XB).max())
