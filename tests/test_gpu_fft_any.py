"""The any-length hand-written passes (csrc/fft_any.cuh, csrc/fft_any_passes.cuh): meshes whose axes are not powers of
two -- every production mesh of the reference (paramfiles/boss_nC.param:23,33, the Nseries bispectrum mesh) and the toy
worlds of the goldens -- run on fft_backend = 1.  Generic transforms against numpy; every pipeline against the goldens of
the unmodified reference scripts (1e-9) and against the cuFFT backend (round-off)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from spectra_without_windows_b200 import engine
    return engine


def oracle_setup(w, which):
    from oracle import sww_oracle as orc
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    return orc, orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)


SMALL = [(20, 18, 15), (32, 24, 20), (15, 20, 18), (7, 9, 11), (64, 18, 128), (24, 64, 50), (30, 128, 64), (128, 45, 33),
         (97, 12, 121)]
PRODUCTION = [(258, 496, 274), (172, 330, 182), (167, 307, 175)]


@pytest.mark.parametrize("shape", SMALL + PRODUCTION)
def test_any_length_transforms(eng, shape):
    """r2c / c2r / c2c of mixed, odd, prime and production sizes against numpy (pocketfft)."""
    box = tuple(20.0 * s for s in shape)
    mesh = eng.Mesh(box, shape, (0.0, 0.0, 0.0), 0.0, 0.1, 0.05, 0)
    ctx = eng.Context(mesh, 0, ("ops",))
    assert ctx.fft_backend == 1
    rng = np.random.default_rng(1)
    x = rng.normal(size=shape)
    ref = np.fft.rfftn(x)
    X = ctx.fft_r2c(x).cpu().numpy()
    assert relerr(X, ref) < 1e-13
    back = ctx.fft_c2r(ref).cpu().numpy()
    assert relerr(back, x) < 1e-13
    own, _ = ctx.launch_counts()
    assert own >= 6                              # three passes per transform, all of them the library's own kernels
    xc = x + 1j * rng.normal(size=shape)
    refc = np.fft.fftn(xc)
    assert relerr(ctx.fft_c2c(xc).cpu().numpy(), refc) < 1e-13
    assert relerr(ctx.fft_c2c(refc, inverse=True).cpu().numpy(), xc) < 1e-13
    ctx.close()


def test_backend_switch_off(eng, world):
    """SWW_NO_ANY_LENGTH=1 restores the round-1 behaviour: meshes that are not powers of two take cuFFT."""
    w = world("toyA")
    mesh = eng.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    os.environ["SWW_NO_ANY_LENGTH"] = "1"
    try:
        ctx = eng.Context(mesh, 0, ("ops",))
        assert ctx.fft_backend == 0
        with pytest.raises(Exception, match="power-of-two"):
            ctx.set_fft_backend(1)
        ctx.close()
    finally:
        del os.environ["SWW_NO_ANY_LENGTH"]


@pytest.mark.parametrize("name", ["toyA", "toyC"])
def test_pk_goldens_on_any_length_passes(eng, world, golden, name):
    w, g = world(name), golden(name)
    kmin, kmax, dk, lmax = w.pk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = eng.Context(mesh, 0, ("ops", "pk_fisher", "pk_data"))
    assert ctx.fft_backend == 1
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    orc, S = oracle_setup(w, "pk")
    for r in (1, 2):
        a = S.grf(r)
        F, qbar = ctx.pk_fisher(a)
        F, qbar = F.cpu().numpy(), qbar.cpu().numpy()
        assert relerr(F, g["pk_fish_%d" % r]) < TOL
        assert relerr(qbar, g["pk_qbar_%d" % r]) < TOL
        F2, q2 = ctx.pk_fisher(a)
        assert np.array_equal(F, F2.cpu().numpy()) and np.array_equal(qbar, q2.cpu().numpy())   # deterministic
    own, lib = ctx.launch_counts()
    assert own > 0
    # in-pass binning instead of the shell-ordered store
    os.environ["SWW_NO_SHELL_STORE"] = "1"
    try:
        c2 = eng.Context(mesh, 0, ("pk_fisher",))
        c2.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
        Fb, qb = c2.pk_fisher(S.grf(1))
        assert relerr(Fb.cpu().numpy(), g["pk_fish_1"]) < TOL and relerr(qb.cpu().numpy(), g["pk_qbar_1"]) < TOL
        c2.close()
    finally:
        del os.environ["SWW_NO_SHELL_STORE"]
    # cuFFT backend: round-off agreement
    ctx.set_fft_backend(0)
    F0, q0 = ctx.pk_fisher(S.grf(1))
    ctx.set_fft_backend(1)
    F1, q1 = ctx.pk_fisher(S.grf(1))
    assert relerr(F1.cpu().numpy(), F0.cpu().numpy()) < 1e-12
    assert relerr(q1.cpu().numpy(), q0.cpu().numpy()) < 1e-12
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    q = ctx.pk_data_q(diff).cpu().numpy()
    assert relerr(q, g["pk_q"]) < TOL
    p = np.inner(np.linalg.inv(g["pk_fish_comb"]), q - g["pk_bias_comb"])
    assert relerr(p, g["pk_p_f64"]) < 1e-9
    x = np.random.default_rng(3).normal(size=mesh.shape)
    maps = S.applyC_alpha(x)
    for l_i in range(S.n_l):
        a_i = min(2, S.n_k - 1)
        got = ctx.apply_c_alpha_single(x, S.nbar, S.v_cell, l_i, a_i).cpu().numpy()
        assert relerr(got, np.real(maps[l_i * S.n_k + a_i])) < 1e-12
    for p_ in (0, 2):
        got = ctx.ylm_project(x, p_).cpu().numpy()
        for l_i in range(S.n_l):
            assert relerr(got[l_i], S.ylm_sum(x, l_i, p_)[:, :, :mesh.nzh]) < 1e-12
    ctx.close()


@pytest.mark.parametrize("name", ["toyA", "toyC"])
def test_bk_goldens_on_any_length_passes(eng, world, golden, name):
    w, g = world(name), golden(name)
    kmin, kmax, dk, lmax = w.bk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = eng.Context(mesh, 0, ("ops", "bk_fisher", "bk_data"))
    assert ctx.fft_backend == 1
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    orc, S = oracle_setup(w, "bk")
    aA, aB = S.grf(2), S.grf(3)
    D = ctx.bk_delta().cpu().numpy()
    assert relerr(D, orc.delta_abc(S, orc.triangle_bins(S.kmin, S.kmax, S.dk, S.n_k))) < 1e-11
    F, (gA, tgA, gB, tgB) = ctx.bk_fisher_pair(aA, aB)
    assert relerr(F.cpu().numpy(), g["bk_fish_1"]) < TOL
    assert relerr(gA.cpu().numpy()[:, :, :, :, 3], g["bk_g_plane"]) < TOL
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    gd = ctx.bk_gmaps(diff, data=True)
    q = ctx.bk_q3(gd)
    ctx.bk_q1_accumulate(gd, gA, tgA, 2, q)
    ctx.bk_q1_accumulate(gd, gB, tgB, 2, q)
    nt = ctx.n_tri
    q = q.cpu().numpy()
    qa = np.concatenate([q[:, l] / D[l * nt:(l + 1) * nt] for l in range(mesh.n_l)])
    p = np.matmul(np.linalg.inv(g["bk_fish_comb"]), qa)
    assert relerr(p, g["bk_p_f64"]) < 1e-9
    ctx.close()


@pytest.mark.parametrize("shape,kmax,dk", [((258, 496, 274), 0.41, 0.08), ((167, 307, 175), 0.11, 0.02),
                                          ((172, 330, 182), 0.16, 0.04)])
def test_production_meshes_agree_with_cufft_backend(eng, shape, kmax, dk):
    """The reference's production meshes at full size (box 1800 x 3450 x 1900, paramfiles/boss_nC.param): P_l Fisher matrix
    and q-bar of one realisation on a masked n(r), own passes against the cuFFT backend (itself pinned to the reference
    goldens at toy sizes), with no library FFT call on the own path."""
    from spectra_without_windows_b200 import synthetic as syn
    box = (1800.0, 3450.0, 1900.0)
    bc = (2.0 * box[0], 30.0, -20.0)
    mesh = eng.Mesh(box, shape, bc, 0.0, kmax, dk, 4)
    n_gal = syn.survey_nbar(box, shape, bc, n0=3e-4)
    nbar_r = n_gal / 0.05
    # the cuFFT context is created with the any-length passes switched off, so that its workspace is laid out for cuFFT
    os.environ["SWW_NO_ANY_LENGTH"] = "1"
    try:
        ref = eng.Context(mesh, 0, ("pk_fisher",))
    finally:
        del os.environ["SWW_NO_ANY_LENGTH"]
    assert ref.fft_backend == 0
    nbar_A = ref.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    a = np.random.default_rng(2).normal(size=shape) * np.sqrt(nbar_A)
    F0, q0 = ref.pk_fisher(a)
    F0, q0 = F0.cpu().numpy(), q0.cpu().numpy()
    ref.close()
    ctx = eng.Context(mesh, 0, ("pk_fisher",))
    assert ctx.fft_backend == 1
    ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    F1, q1 = ctx.pk_fisher(a)
    assert relerr(F1.cpu().numpy(), F0) < 1e-11
    assert relerr(q1.cpu().numpy(), q0) < 1e-11
    F2, _ = ctx.pk_fisher(a)
    assert np.array_equal(F1.cpu().numpy(), F2.cpu().numpy())
    ctx.close()


@pytest.mark.parametrize("name,dims", [("toyA", (64, 64, 128)), ("toyA", (32, 24, 128)), ("toyA", (64, 24, 20)), ("toyA", (32, 64, 20)),
                                       ("toyC", (64, 64, 128)), ("toyC", (20, 18, 128)), ("toyC", (20, 64, 15))])
def test_enlarged_row_mesh_matches_goldens(eng, world, golden, name, dims):
    """Row mesh LARGER than the mesh (csrc/rowmesh.cu): the Fisher rows of a mesh whose axis lengths carry large prime
    factors run on a power-of-two grid against the n-periodic spectrum of the weight map -- the mesh's cyclic convolution,
    wrap-around included, as a linear one.  Forced here on the toy worlds, whose k_max sits at 0.6-0.9 of Nyquist (so the
    wrap-around terms are exercised), per axis and on all axes, against the goldens of the unmodified reference script."""
    w, g = world(name), golden(name)
    kmin, kmax, dk, lmax = w.pk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    os.environ["SWW_ROW_MESH_DIMS"] = "%d,%d,%d" % dims
    try:
        ctx = eng.Context(mesh, 0, ("pk_fisher",))
    finally:
        del os.environ["SWW_ROW_MESH_DIMS"]
    assert ctx.fft_backend == 1 and ctx.row_mesh_shape() == dims
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    orc, S = oracle_setup(w, "pk")
    for r in (1, 2):
        F, qbar = ctx.pk_fisher(S.grf(r))
        assert relerr(F.cpu().numpy(), g["pk_fish_%d" % r]) < TOL
        assert relerr(qbar.cpu().numpy(), g["pk_qbar_%d" % r]) < TOL
    ctx.close()


@pytest.mark.parametrize("row_mesh", ["auto", "off"])
@pytest.mark.parametrize("shape,kmax,dk", [((258, 496, 274), 0.41, 0.08), ((167, 307, 175), 0.11, 0.02),
                                          ((172, 330, 182), 0.16, 0.04)])
def test_production_meshes_against_oracle_golden(eng, golden, shape, kmax, dk, row_mesh):
    """q-bar and F of one realisation at the FULL size of the reference's production meshes against the numpy restatement of the
    reference (oracle/make_golden_production.py -> tests/golden/production_meshes.npz; 1e-9 like every golden), with the rows on
    the row mesh the cost model picks and on the mesh itself (the any-length fused z pass: 274 = 2 * 137, 175 = 7 * 5 * 5)."""
    from oracle import make_golden_production as mg
    g = golden("production_meshes")
    tag = "%dx%dx%d" % shape
    bc, nbar_r = mg.inputs(shape)
    mesh = eng.Mesh(mg.BOX, shape, bc, 0.0, kmax, dk, 4)
    if row_mesh == "off":
        os.environ["SWW_ROW_MESH"] = "0"
    try:
        ctx = eng.Context(mesh, 0, ("pk_fisher",))
    finally:
        os.environ.pop("SWW_ROW_MESH", None)
    assert ctx.fft_backend == 1
    if row_mesh == "off":
        assert ctx.row_mesh_shape() is None
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, mg.ALPHA, mg.SHOT)
    a = np.random.default_rng(mg.SEED).normal(size=shape) * np.sqrt(nbar_A)
    F, qbar = ctx.pk_fisher(a)
    assert relerr(F.cpu().numpy(), g["F_" + tag]) < TOL
    assert relerr(qbar.cpu().numpy(), g["qbar_" + tag]) < TOL
    ctx.close()
