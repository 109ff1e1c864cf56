"""The hand-written sm_100a FFT passes (fft_backend=1): generic transforms against numpy, and
every pipeline re-run on the fused / pruned chains against the reference goldens (1e-9)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from spectra_without_windows_b200 import engine
    return engine


def ctx_for(eng, w, which, pipelines):
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = eng.Context(mesh, 0, pipelines)
    ctx.set_fft_backend(1)
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    return mesh, ctx


def oracle_setup(w, which):
    from oracle import sww_oracle as orc
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    return orc, orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)


def test_backend_rejects_unsupported_mesh(eng):
    shape = (700, 16, 16)                # neither a power of two nor within the any-length limit (640)
    mesh = eng.Mesh(tuple(10.0 * s for s in shape), shape, (0.0, 0.0, 0.0), 0.0, 0.1, 0.05, 0)
    ctx = eng.Context(mesh, 0, ("ops",))
    assert ctx.fft_backend == 0          # auto-selection fell back to cuFFT
    with pytest.raises(Exception, match="power-of-two"):
        ctx.set_fft_backend(1)
    ctx.close()


@pytest.mark.parametrize("shape", [(64, 64, 128), (128, 64, 256), (64, 256, 128), (64, 64, 512), (512, 64, 1024),
                                   (64, 1024, 128)])
def test_generic_transforms(eng, shape):
    box = tuple(20.0 * s for s in shape)
    mesh = eng.Mesh(box, shape, (0.0, 0.0, 0.0), 0.0, 0.1, 0.05, 0)
    ctx = eng.Context(mesh, 0, ("ops",))
    ctx.set_fft_backend(1)
    rng = np.random.default_rng(1)
    x = rng.normal(size=shape)
    X = ctx.fft_r2c(x).cpu().numpy()
    ref = np.fft.rfftn(x)
    assert relerr(X, ref) < 1e-13
    back = ctx.fft_c2r(ref).cpu().numpy()
    assert relerr(back, x) < 1e-13
    ctx.close()


def test_pk_pipelines_on_fused_chains(eng, world, golden):
    w, g = world("toyD"), golden("toyD")
    mesh, ctx = ctx_for(eng, w, "pk", ("ops", "pk_fisher", "pk_data"))
    orc, S = oracle_setup(w, "pk")
    for r in (1, 2):
        a = S.grf(r)
        F, qbar = ctx.pk_fisher(a)
        F, qbar = F.cpu().numpy(), qbar.cpu().numpy()
        assert relerr(F, g["pk_fish_%d" % r]) < TOL
        assert relerr(qbar, g["pk_qbar_%d" % r]) < TOL
        F2, q2 = ctx.pk_fisher(a)
        assert np.array_equal(F, F2.cpu().numpy()) and np.array_equal(qbar, q2.cpu().numpy())   # deterministic
    # the cuFFT backend agrees with the fused chains to round-off
    ctx.set_fft_backend(0)
    F0, q0 = ctx.pk_fisher(S.grf(1))
    ctx.set_fft_backend(1)
    F1, q1 = ctx.pk_fisher(S.grf(1))
    assert relerr(F1.cpu().numpy(), F0.cpu().numpy()) < 1e-12
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    q = ctx.pk_data_q(diff).cpu().numpy()
    assert relerr(q, g["pk_q"]) < TOL
    x = np.random.default_rng(3).normal(size=mesh.shape)
    maps = S.applyC_alpha(x)
    for l_i in range(S.n_l):
        got = ctx.apply_c_alpha_single(x, S.nbar, S.v_cell, l_i, 3).cpu().numpy()
        assert relerr(got, np.real(maps[l_i * S.n_k + 3])) < 1e-12
    ctx.close()


def test_bk_pipelines_on_fused_chains(eng, world, golden):
    w, g = world("toyD"), golden("toyD")
    mesh, ctx = ctx_for(eng, w, "bk", ("ops", "bk_fisher", "bk_data"))
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
    assert relerr(p, g["bk_p"][:, 3:].T.ravel()) < 2e-8
    assert relerr(p, g["bk_p_f64"]) < 1e-9
    ctx.close()


@pytest.mark.parametrize("shape", [(64, 128, 512), (128, 64, 1024)])
def test_fused_chains_agree_with_cufft_backend(eng, shape):
    """Every radix plan of the fused z pass (H = 256, 512) against the cuFFT backend (itself pinned to
    the oracle at small sizes): P_l Fisher matrix and q-bar of one realisation on a masked n(r)."""
    from spectra_without_windows_b200 import synthetic as syn
    box = tuple(18.0 * s for s in shape)
    bc = (2.0 * box[0], 30.0, -20.0)
    mesh = eng.Mesh(box, shape, bc, 0.0, 0.12, 0.02, 4)
    ctx = eng.Context(mesh, 0, ("pk_fisher",))
    assert ctx.fft_backend == 1
    ctx.set_fft_backend(0)
    n_gal = syn.survey_nbar(box, shape, bc, n0=3e-4)
    nbar_r = n_gal / 0.05
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    a = np.random.default_rng(2).normal(size=shape) * np.sqrt(nbar_A)
    F0, q0 = ctx.pk_fisher(a)
    F0, q0 = F0.cpu().numpy(), q0.cpu().numpy()
    ctx.set_fft_backend(1)
    F1, q1 = ctx.pk_fisher(a)
    assert relerr(F1.cpu().numpy(), F0) < 1e-11
    assert relerr(q1.cpu().numpy(), q0) < 1e-11
    ctx.close()


def test_fused_z_pass_nz512_two_exchange_kernel(eng):
    """nz = 512 takes the two-exchange fused z pass (pass_z16_fused_kernel): it must agree with the generic
    warp-per-row kernel (SWW_NO_Z16=1) and with the cuFFT backend to round-off, and be deterministic."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    grid, box, bc = (64, 64, 512), (1280.0, 1408.0, 5120.0), (1700.0, 60.0, -45.0)
    mesh = eng.Mesh(box, grid, bc, 0.0, 0.12, 0.02, 4)
    ctx = eng.Context(mesh, 0, ("ops", "pk_fisher"))
    ctx.set_fft_backend(1)
    nbar_r = syn.survey_nbar(box, grid, bc, n0=6e-3)
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    a = np.random.default_rng(7).normal(size=grid) * np.sqrt(nbar_A)
    F1, q1 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    F1b, _ = ctx.pk_fisher(a)
    assert np.array_equal(F1, F1b.cpu().numpy())
    own0, _ = ctx.launch_counts()
    os.environ["SWW_NO_Z16"] = "1"
    try:
        F2, q2 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    finally:
        del os.environ["SWW_NO_Z16"]
    ctx.set_fft_backend(0)
    F0, q0 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    assert np.all(np.isfinite(F1)) and np.max(np.abs(F1)) > 0
    assert relerr(F1, F2) < 1e-12 and relerr(q1, q2) < 1e-12
    assert relerr(F1, F0) < 1e-11 and relerr(q1, q0) < 1e-11
    ctx.close()


@pytest.mark.skipif(not __import__("os").environ.get("SWW_TEST_Z32"), reason="experimental nz = 1024 kernel: opt-in (SWW_TEST_Z32=1)")
def test_fused_z_pass_nz1024_experimental_kernel(eng):
    """pass_z32_fused_kernel (SWW_Z32=1) against the generic fused z pass and the cuFFT backend on a 64 x 64 x 1024 mesh."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    grid, box, bc = (64, 64, 1024), (1280.0, 1408.0, 10240.0), (1700.0, 60.0, -45.0)
    mesh = eng.Mesh(box, grid, bc, 0.0, 0.12, 0.02, 4)
    ctx = eng.Context(mesh, 0, ("ops", "pk_fisher"))
    ctx.set_fft_backend(1)
    nbar_r = syn.survey_nbar(box, grid, bc, n0=6e-3)
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    a = np.random.default_rng(7).normal(size=grid) * np.sqrt(nbar_A)
    F_ref, q_ref = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    os.environ["SWW_Z32"] = "1"
    try:
        F1, q1 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
        F1b, _ = ctx.pk_fisher(a)
    finally:
        del os.environ["SWW_Z32"]
    assert np.array_equal(F1, F1b.cpu().numpy())
    assert relerr(F1, F_ref) < 1e-12 and relerr(q1, q_ref) < 1e-12
    ctx.set_fft_backend(0)
    F0, q0 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    assert relerr(F1, F0) < 1e-11 and relerr(q1, q0) < 1e-11
    ctx.close()


def test_row_mesh_matches_full_mesh(eng):
    """Coarse row mesh (csrc/rowmesh.cu): on a 128 x 128 x 256 mesh whose k_max box is b = (14, 14, 28) the Fisher rows run
    on a 64 x 64 x 128 child grid against the band-limited weight map.  F and q-bar must agree with the full-mesh rows
    (SWW_ROW_MESH=0) and with the cuFFT backend to round-off -- the identity is exact, not an approximation."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    grid, box, bc = (128, 128, 256), (1000.0, 1000.0, 2000.0), (1400.0, 60.0, -45.0)
    mesh = eng.Mesh(box, grid, bc, 0.0, 0.09, 0.015, 4)
    nbar_r = syn.survey_nbar(box, grid, bc, n0=6e-3)
    a = None
    res = {}
    for tag, env in (("row", None), ("full", "0")):
        if env is not None:
            os.environ["SWW_ROW_MESH"] = env
        try:
            ctx = eng.Context(mesh, 0, ("ops", "pk_fisher"))
        finally:
            os.environ.pop("SWW_ROW_MESH", None)
        assert ctx.fft_backend == 1
        nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
        if a is None:
            a = np.random.default_rng(11).normal(size=grid) * np.sqrt(nbar_A)
        F, q = (t.cpu().numpy() for t in ctx.pk_fisher(a))
        F2, _ = ctx.pk_fisher(a)
        assert np.array_equal(F, F2.cpu().numpy())          # deterministic
        res[tag] = (F, q, ctx.row_mesh_shape())
        if tag == "full":
            ctx.set_fft_backend(0)
            res["cufft"] = tuple(t.cpu().numpy() for t in ctx.pk_fisher(a))
        ctx.close()
    assert res["row"][2] == (64, 64, 128) and res["full"][2] is None
    assert np.max(np.abs(res["full"][0])) > 0
    assert relerr(res["row"][0], res["full"][0]) < 1e-11 and relerr(res["row"][1], res["full"][1]) < 1e-11
    assert relerr(res["row"][0], res["cufft"][0]) < 1e-11 and relerr(res["row"][1], res["cufft"][1]) < 1e-11


def test_tile_major_planes_and_bulk_copy_f2(eng):
    """Opt-in tile-major Q / R planes (SWW_TILE_MAJOR=b) and the bulk-asynchronous-copy F2 kernel (SWW_F2_BULK=1,
    cp.async.bulk + mbarrier) must reproduce the default row-major chain bit for bit: same arithmetic, other data movement."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    grid, box, bc = (64, 128, 512), (1280.0, 2816.0, 5120.0), (1700.0, 60.0, -45.0)
    mesh = eng.Mesh(box, grid, bc, 0.0, 0.12, 0.02, 4)
    ctx = eng.Context(mesh, 0, ("ops", "pk_fisher"))
    nbar_A = ctx.set_fields_from_randoms_density(syn.survey_nbar(box, grid, bc, n0=6e-3), 0.05, 1.05)
    a = np.random.default_rng(3).normal(size=grid) * np.sqrt(nbar_A)
    F0, q0 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
    # ... and so must the chain without the sphere skipping of the column passes (skipped columns are exact zeros)
    for env in ({"SWW_TILE_MAJOR": "b"}, {"SWW_TILE_MAJOR": "r", "SWW_F2_BULK": "1", "SWW_NO_SPHERE_SKIP": "1"}, {"SWW_TILE_MAJOR": "q"},
                {"SWW_NO_SPHERE_SKIP": "1"}):
        os.environ.update(env)
        try:
            F1, q1 = (t.cpu().numpy() for t in ctx.pk_fisher(a))
        finally:
            for k in env:
                del os.environ[k]
        assert np.array_equal(F1, F0) and np.array_equal(q1, q0), env
    ctx.close()


@pytest.mark.parametrize("shape", [(64, 64, 128), (128, 64, 256), (64, 256, 512)])
def test_complex_transforms_on_own_passes(eng, shape):
    """ft / ift (src/opt_utilities.py:502-555) as complex 3-D transforms on the hand-written passes (complex z pass + the
    two column passes over the full box) against numpy.fft, and against the cuFFT Z2Z path."""
    box = tuple(20.0 * s for s in shape)
    mesh = eng.Mesh(box, shape, (0.0, 0.0, 0.0), 0.0, 0.1, 0.05, 0)
    ctx = eng.Context(mesh, 0, ("ops",))
    assert ctx.fft_backend == 1
    rng = np.random.default_rng(2)
    x = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    own0, fft0 = ctx.launch_counts()
    X = ctx.fft_c2c(x).cpu().numpy()
    own1, _ = ctx.launch_counts()
    assert own1 - own0 >= 3                                   # the passes ran (no cuFFT plan involved)
    assert relerr(X, np.fft.fftn(x)) < 1e-13
    assert relerr(ctx.fft_c2c(X, inverse=True).cpu().numpy(), x) < 1e-13
    ctx.set_fft_backend(0)
    assert relerr(ctx.fft_c2c(x).cpu().numpy(), X) < 1e-13
    ctx.close()


@pytest.mark.parametrize("shape", [(64, 64, 256), (64, 64, 512), (64, 128, 1024)])
def test_forward_only_16x16_z_kernels(eng, shape):
    """The forward transforms FT[f Y_lm(r^)] of a realisation's head on pass_z8f_kernel (nz = 256) / pass_z16f_kernel (512) / pass_z32f_kernel (1024):
    a mesh whose k_max box stays below nz / 4, P_l Fisher matrix and q-bar against the cuFFT backend and against the
    three-stage warp-per-row kernel they replace (SWW_NO_Z16F=1)."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    box = tuple(6.0 * s for s in shape)
    bc = (2.0 * box[0], 30.0, -20.0)
    mesh = eng.Mesh(box, shape, bc, 0.0, 0.12, 0.03, 4)
    ctx = eng.Context(mesh, 0, ("pk_fisher",))
    assert ctx.fft_backend == 1
    n_gal = syn.survey_nbar(box, shape, bc, n0=3e-4)
    nbar_r = n_gal / 0.05
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    a = np.random.default_rng(4).normal(size=shape) * np.sqrt(nbar_A)
    F1, q1 = ctx.pk_fisher(a)
    F1, q1 = F1.cpu().numpy(), q1.cpu().numpy()
    os.environ["SWW_NO_Z16F"] = "1"
    try:
        F2, q2 = ctx.pk_fisher(a)
    finally:
        del os.environ["SWW_NO_Z16F"]
    assert relerr(F1, F2.cpu().numpy()) < 1e-12 and relerr(q1, q2.cpu().numpy()) < 1e-12
    ctx.close()
    os.environ["SWW_ROW_MESH"] = "0"          # the cuFFT context lays its workspace out for the full mesh
    try:
        ref = eng.Context(mesh, 0, ("pk_fisher",))
    finally:
        del os.environ["SWW_ROW_MESH"]
    ref.set_fft_backend(0)
    ref.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    F0, q0 = ref.pk_fisher(a)
    assert relerr(F1, F0.cpu().numpy()) < 1e-11 and relerr(q1, q0.cpu().numpy()) < 1e-11
    ref.close()


def test_inverse_only_z8_kernel(eng):
    """The opt-in batched inverse z pass of the bispectrum phi maps at nz = 256 (SWW_Z8I=1, pass_z8i_kernel: one exchange,
    harmonics as row polynomials with one reciprocal per cell) against the default warp-per-row kernel with cached harmonics:
    the whole B_l Fisher matrix of one pair, l <= 4."""
    import os
    from spectra_without_windows_b200 import synthetic as syn
    shape = (64, 64, 256)
    box = tuple(6.0 * s for s in shape)
    bc = (2.0 * box[0], 30.0, -20.0)
    mesh = eng.Mesh(box, shape, bc, 0.0, 0.12, 0.03, 4)
    ctx = eng.Context(mesh, 0, ("bk_fisher",))
    assert ctx.fft_backend == 1
    n_gal = syn.survey_nbar(box, shape, bc, n0=3e-4)
    nbar_r = n_gal / 0.05
    nbar_A = ctx.set_fields_from_randoms_density(nbar_r, 0.05, 1.05)
    rng = np.random.default_rng(6)
    aA = rng.normal(size=shape) * np.sqrt(nbar_A)
    aB = rng.normal(size=shape) * np.sqrt(nbar_A)
    F1, maps = ctx.bk_fisher_pair(aA, aB)               # default: the warp-per-row kernel with cached harmonics
    F1 = F1.cpu().numpy().copy()
    os.environ["SWW_Z8I"] = "1"                         # opt-in kernel (measured slower, kept for the record: DESIGN.md)
    try:
        F2, maps = ctx.bk_fisher_pair(aA, aB, maps)
    finally:
        del os.environ["SWW_Z8I"]
    F2 = F2.cpu().numpy()
    assert np.isfinite(F1).all() and relerr(F1, F2) < 1e-11
    ctx.close()
