"""GPU parity tests of the P_l(k) path: CUDA (through the C ABI) vs the CPU oracle and the
golden vectors of the unmodified reference scripts.  Tolerance: 1e-9 norm-wise
(max|d|/max|ref|), the north-star tolerance against the FP64-promoted reference."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

WORLDS = ["toyA", "toyB", "toyC"]
TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from spectra_without_windows_b200 import engine
    return engine


def make_ctx(eng, w, which="pk"):
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = eng.get_context(mesh, 0, ("ops", "pk_fisher", "pk_data"))
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    return mesh, ctx


def oracle_setup(w, which="pk"):
    from oracle import sww_oracle as orc
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    return orc, orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)


@pytest.mark.parametrize("name", WORLDS)
def test_binmap_bit_exact(name, world, eng):
    w = world(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    ref = np.full(S.k_norm.shape, 255, dtype=np.uint8)
    for i in range(S.n_k):
        ref[S.filt(i, S.k_norm)] = i
    got = ctx.binmap().cpu().numpy()
    assert np.array_equal(got, ref[:, :, :mesh.nzh])


@pytest.mark.parametrize("name", WORLDS)
def test_operators(name, world, eng):
    w = world(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    rng = np.random.default_rng(3)
    x = rng.normal(size=mesh.shape)
    # ft / ift mirrors
    xc = x + 1j * rng.normal(size=mesh.shape)
    assert relerr(ctx.fft_c2c(xc).cpu().numpy(), orc.ft(xc)) < 1e-13
    assert relerr(ctx.fft_c2c(xc, inverse=True).cpu().numpy(), orc.ift(xc)) < 1e-13
    assert relerr(ctx.fft_r2c(x).cpu().numpy(), orc.ft(x)[:, :, :mesh.nzh]) < 1e-13
    assert relerr(ctx.fft_c2r(ctx.fft_r2c(x)).cpu().numpy(), x) < 1e-13
    # FKP operators
    got = ctx.apply_cinv_fkp(x, S.nbar_weight, S.v_cell, S.shot_fac).cpu().numpy()
    assert relerr(got, np.real(orc.applyCinv_fkp(x, S.nbar_weight, S.v_cell, S.shot_fac))) < 1e-14
    got = ctx.apply_n(x, S.nbar, S.v_cell, S.shot_fac).cpu().numpy()
    assert relerr(got, orc.applyN(x, S.nbar, S.v_cell, S.shot_fac)) < 1e-14
    # sum_m Y_lm(k) FT[x Y_lm(r)] with and without MAS^2
    for p in (0, 2):
        got = ctx.ylm_project(x, p).cpu().numpy()
        for l_i in range(S.n_l):
            ref = S.ylm_sum(x, l_i, p)[:, :, :mesh.nzh]
            assert relerr(got[l_i], ref) < 1e-12, (p, l_i)
    # one derivative map per multipole
    maps = S.applyC_alpha(x)
    for l_i in range(S.n_l):
        a = min(2, S.n_k - 1)
        got = ctx.apply_c_alpha_single(x, S.nbar, S.v_cell, l_i, a).cpu().numpy()
        assert relerr(got, np.real(maps[l_i * S.n_k + a])) < 1e-12


@pytest.mark.parametrize("name", WORLDS)
def test_pk_fisher_vs_reference_golden(name, world, golden, eng):
    w, g = world(name), golden(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    for r in (1, 2):
        a = S.grf(r)
        F, qbar = ctx.pk_fisher(a)
        F, qbar = F.cpu().numpy(), qbar.cpu().numpy()
        assert relerr(F, g["pk_fish_%d" % r]) < TOL
        assert relerr(qbar, g["pk_qbar_%d" % r]) < TOL
        assert np.array_equal(F, F.T)
        # and against the live oracle
        qb_o, F_o = orc.pk_fisher_realisation(S, a)
        assert relerr(F, F_o) < TOL and relerr(qbar, qb_o) < TOL


@pytest.mark.parametrize("name", WORLDS)
def test_pk_fisher_deterministic(name, world, eng):
    w = world(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    a = S.grf(5)
    F1, q1 = ctx.pk_fisher(a)
    F2, q2 = ctx.pk_fisher(a)
    assert np.array_equal(F1.cpu().numpy(), F2.cpu().numpy())
    assert np.array_equal(q1.cpu().numpy(), q2.cpu().numpy())


@pytest.mark.parametrize("name", WORLDS)
def test_paint_and_pk_data(name, world, golden, eng):
    w, g = world(name), golden(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    ref, _ = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    assert relerr(diff.cpu().numpy(), ref) < 1e-12
    assert relerr(diff.cpu().numpy()[:, :, 2], g["diff_plane"]) < 1e-12
    q = ctx.pk_data_q(diff).cpu().numpy()
    assert relerr(q, g["pk_q"]) < TOL
    p = np.inner(np.linalg.inv(g["pk_fish_comb"]), q - g["pk_bias_comb"])
    assert relerr(p, g["pk_p"][:, 1:].T.ravel()) < 2e-8     # '%.8e' text
    assert relerr(p, g["pk_p_f64"]) < 1e-9                  # the script's FP64 p_alpha (pk/compute_pk_data.py:250)


def test_paint_edge_cases(world, eng):
    w = world("toyC")
    mesh, ctx = make_ctx(eng, w)
    from oracle import sww_oracle as orc
    # empty catalogue, particles exactly on cell centres / half-way points / outside the box (periodic wrap)
    out = ctx.paint(np.zeros((0, 3)), np.zeros(0), w.box_center)
    assert float(out.abs().sum()) == 0.0
    H = np.asarray(w.box) / np.asarray(w.grid)
    pos = np.asarray(w.box_center) + np.array([[0, 0, 0], [0.5, 0.5, 0.5], [-0.5, 1.5, 2.5], [100.0, -37.25, 19.5],
                                               [-9.999, 8.0, -7.5]]) * H
    wts = np.array([1.0, 2.0, -1.5, 0.25, 3.0])
    got = ctx.paint(pos, wts, w.box_center).cpu().numpy()
    ref = orc.tsc_paint(pos, wts, w.box, w.grid, w.box_center)
    assert relerr(got, ref) < 1e-14
    assert abs(got.sum() - wts.sum()) < 1e-12


@pytest.mark.parametrize("scheme", ["TSC", "CIC"])
def test_tile_binned_painter_parity_and_determinism(scheme, eng):
    """The tile-binned painter (csrc/paint.cu: 8^3 tiles, fixed-point shared-memory accumulation, colour phases) against
    the numpy restatement for both windows, on a clustered catalogue with weights of both signs and particles outside the
    box (periodic wrap), on a mesh with odd tile counts on two axes; bit-identical from run to run and against the
    order-dependent first version (global FP64 atomics) to round-off."""
    import os
    from oracle import sww_oracle as orc
    grid, box, bc = (72, 40, 64), (900.0, 500.0, 800.0), (100.0, -20.0, 40.0)
    mesh = eng.Mesh(box, grid, bc, 0.0, 0.05, 0.025, 0)
    ctx = eng.Context(mesh, 0, ("ops",))
    rng = np.random.default_rng(17)
    n = 400000
    centres = rng.uniform(-0.6, 0.6, size=(40, 3)) * np.asarray(box)
    pos = np.asarray(bc) + centres[rng.integers(0, 40, n)] + rng.normal(scale=25.0, size=(n, 3))
    wts = rng.uniform(-0.5, 2.0, n)
    ref = (orc.tsc_paint if scheme == "TSC" else orc.cic_paint)(pos, wts, box, grid, bc)
    a = ctx.paint(pos, wts, bc, scheme=scheme, scale=0.37).cpu().numpy()
    b = ctx.paint(pos, wts, bc, scheme=scheme, scale=0.37).cpu().numpy()
    assert np.array_equal(a, b)                                   # bit-deterministic
    assert relerr(a, 0.37 * ref) < 1e-13
    assert abs(a.sum() - 0.37 * wts.sum()) < 1e-9 * np.abs(wts).sum()
    os.environ["SWW_PAINT_V1"] = "1"
    try:
        v1 = ctx.paint(pos, wts, bc, scheme=scheme, scale=0.37).cpu().numpy()
    finally:
        del os.environ["SWW_PAINT_V1"]
    assert relerr(a, v1) < 1e-13
    # accumulation into an existing mesh (grid_data paints the randoms on top of the galaxies)
    base = ctx.dev(rng.normal(size=grid))
    base0 = base.cpu().numpy().copy()
    out = ctx.paint(pos[:1000], wts[:1000], bc, scheme=scheme, out=base).cpu().numpy()
    sub = (orc.tsc_paint if scheme == "TSC" else orc.cic_paint)(pos[:1000], wts[:1000], box, grid, bc)
    assert relerr(out - base0, sub) < 1e-12
    ctx.close()
