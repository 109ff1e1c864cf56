"""Edge cases through the C ABI: k_min > 0 binning on the hand-written FFT passes, error returns
(no silent fallbacks), empty inputs, and workspace checks."""
import ctypes as C

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from spectra_without_windows_b200 import engine
    return engine


def test_kmin_positive_nonuniform_bins_on_own_fft(eng):
    """k_min > 0 (modes below the first shell belong to no bin), lmax = 2, non-cubic power-of-two mesh:
    fused chains (backend 1) vs the live oracle."""
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import synthetic as syn
    shape, box, bc = (64, 128, 128), (1280.0, 2304.0, 2560.0), (2100.0, -80.0, 55.0)
    kmin, kmax, dk, lmax = 0.015, 0.115, 0.025, 2
    n_gal = syn.survey_nbar(box, shape, bc, n0=3e-4)
    nbar_r = n_gal / 0.04
    mesh = eng.Mesh(box, shape, bc, kmin, kmax, dk, lmax)
    ctx = eng.Context(mesh, 0, ("ops", "pk_fisher", "pk_data"))
    assert ctx.fft_backend == 1
    ctx.set_fields_from_randoms_density(nbar_r, 0.04, 1.04)
    S = orc.Setup(box, shape, bc, nbar_r, 0.04, 1.04, kmin, kmax, dk, lmax)
    a = S.grf(9)
    F, qbar = ctx.pk_fisher(a)
    qb_o, F_o = orc.pk_fisher_realisation(S, a)
    assert relerr(F.cpu().numpy(), F_o) < 1e-9 and relerr(qbar.cpu().numpy(), qb_o) < 1e-9
    d = np.random.default_rng(4).normal(size=shape) * 1e-4
    assert relerr(ctx.pk_data_q(d).cpu().numpy(), orc.pk_data_q(S, d)) < 1e-9
    ctx.close()


def test_errors_are_loud(eng, world):
    from spectra_without_windows_b200 import _lib
    w = world("toyA")
    # k_max at/above Nyquist is rejected (the R2C formulation needs every binned mode below Nyquist)
    with pytest.raises(Exception, match="Nyquist"):
        eng.Context(eng.Mesh(w.box, w.grid, w.box_center, 0.0, 0.2, 0.02, 4), 0)
    with pytest.raises(Exception, match="l-max"):
        eng.Context(eng.Mesh(w.box, w.grid, w.box_center, 0.0, 0.12, 0.02, 6), 0)
    mesh = eng.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = eng.Context(mesh, 0, ("pk_fisher",))
    a = ctx.zeros(mesh.shape)
    with pytest.raises(Exception, match="fields not set"):
        ctx.pk_fisher(a)
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    lib = ctx.lib
    assert lib.sww_pk_fisher_realisation(ctx._h, 0, 0, 0) != 0 and b"null" in lib.sww_last_error()
    # a workspace that is too small is refused, never silently re-allocated
    small = ctx.empty((1024,))
    assert lib.sww_bind_workspace(ctx._h, small.data_ptr(), 8192) != 0 and b"too small" in lib.sww_last_error()
    ctx.close()


def test_zero_field_and_linearity(eng, world):
    """F is a quadratic form in the field and q-bar too: zero in -> zero out, scaling by s -> s^2."""
    w = world("toyD")
    mesh = eng.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = eng.Context(mesh, 0, ("pk_fisher",))
    nbar_A = ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    F0, q0 = ctx.pk_fisher(np.zeros(mesh.shape))
    assert float(F0.abs().max()) == 0.0 and float(q0.abs().max()) == 0.0
    a = np.random.default_rng(0).normal(size=mesh.shape) * np.sqrt(nbar_A)
    F1, q1 = ctx.pk_fisher(a)
    F2, q2 = ctx.pk_fisher(-3.0 * a)
    assert relerr(F2.cpu().numpy(), 9.0 * F1.cpu().numpy()) < 1e-12
    assert relerr(q2.cpu().numpy(), 9.0 * q1.cpu().numpy()) < 1e-12
    ctx.close()
