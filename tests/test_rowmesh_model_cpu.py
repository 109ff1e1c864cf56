"""numpy model of the row mesh (csrc/rowmesh.cu): the band-limited product FT[w IFT[U]] restricted to the box |k_i| <= b_i, evaluated
on another grid M of the same box -- smaller than the mesh (truncated weight spectrum), equal, or LARGER (the n-periodic weight
spectrum turns the mesh's cyclic convolution, wrap-around included, into a linear one).  `band_copy` restates band_copy_kernel
(half spectra, Hermitian partner for negative reduced z indices).  CPU only: it pins the derivation; the device path is checked
against the reference goldens on the GPU (tests/test_gpu_fft_any.py::test_enlarged_row_mesh_matches_goldens)."""
import itertools

import numpy as np
import pytest


def signed(j, n):
    return j if 2 * j < n else j - n


def band_copy(spec_f, n, M, b):
    """dst(d) = src(d mod n) on the half spectrum of the M grid (band_copy_kernel)."""
    nx, ny, nz = n
    mx, my, mz = M
    mzh = mz // 2 + 1
    out = np.zeros((mx, my, mzh), dtype=complex)
    for jx, jy, dz in itertools.product(range(mx), range(my), range(mzh)):
        dx, dy = signed(jx, mx), signed(jy, my)
        keep = (2 * jx != mx) if mx < nx else (mx == nx or abs(dx) <= 2 * b[0])
        keep = keep and ((2 * jy != my) if my < ny else (my == ny or abs(dy) <= 2 * b[1]))
        keep = keep and ((dz != mzh - 1 or mz % 2 == 1) if mz < nz else (mz == nz or dz <= 2 * b[2]))
        if not keep:
            continue
        sx, sy, sz = int(np.fmod(dx, nx)), int(np.fmod(dy, ny)), dz % nz
        if 2 * sx > nx:
            sx -= nx
        if 2 * sx <= -nx:
            sx += nx
        if 2 * sy > ny:
            sy -= ny
        if 2 * sy <= -ny:
            sy += ny
        if 2 * sz > nz:
            sz -= nz
        cj = sz < 0
        if cj:
            sx, sy, sz = -sx, -sy, -sz
        v = spec_f[sx % nx, sy % ny, sz]
        out[jx, jy, dz] = np.conj(v) if cj else v
    return out


def box_modes(n, b):
    return [(kx, ky, kz) for kx in range(-b[0], b[0] + 1) for ky in range(-b[1], b[1] + 1) for kz in range(-b[2], b[2] + 1)]


@pytest.mark.parametrize("n,b,M", [
    ((10, 9, 8), (2, 2, 1), (16, 9, 16)),      # x and z enlarged, y on the mesh's own (odd) length; 2b wraps around n on x
    ((10, 9, 8), (2, 2, 1), (10, 16, 8)),      # only y enlarged
    ((20, 18, 16), (2, 2, 2), (12, 18, 16)),   # x smaller than the mesh
    ((7, 6, 12), (3, 1, 2), (16, 8, 12)),      # b = 3 of n = 7: the box is the whole axis, every difference wraps
])
def test_row_mesh_product_is_exact_in_the_box(n, b, M):
    rng = np.random.default_rng(sum(n) + sum(M))
    w = rng.normal(size=n)                                   # the weight map n W: NOT band-limited
    # a real field confined to the box
    U = np.zeros(n, dtype=complex)
    for k in box_modes(n, b):
        U[k[0] % n[0], k[1] % n[1], k[2] % n[2]] = rng.normal() + 1j * rng.normal()
    u = np.fft.ifftn(U).real * 2.0                           # real part: Hermitian-symmetrised input
    U = np.fft.fftn(u)
    T_ref = np.fft.fftn(w * u)                               # the reference's product on the mesh
    # row mesh: w~ = IFT_M[band-limit(FT_n[w])] / N_n (unnormalised c2r), rows normalised by the full mesh
    Nn, NM = float(np.prod(n)), float(np.prod(M))
    w_t = np.fft.irfftn(band_copy(np.fft.rfftn(w), n, M, b), s=M, axes=(0, 1, 2)) * NM / Nn
    UM = np.zeros(M, dtype=complex)
    for k in box_modes(n, b):
        UM[k[0] % M[0], k[1] % M[1], k[2] % M[2]] = U[k[0] % n[0], k[1] % n[1], k[2] % n[2]]
    uM = np.fft.ifftn(UM) * NM / Nn                          # the same band-limited function on the M grid
    assert np.abs(uM.imag).max() < 1e-12 * np.abs(uM.real).max()
    T_M = np.fft.fftn(w_t * uM.real) * (Nn / NM)
    scale = np.abs(T_ref).max()
    for k in box_modes(n, b):
        got = T_M[k[0] % M[0], k[1] % M[1], k[2] % M[2]]
        ref = T_ref[k[0] % n[0], k[1] % n[1], k[2] % n[2]]
        assert abs(got - ref) < 1e-12 * scale, (k, got, ref)
