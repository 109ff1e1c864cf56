"""GPU parity tests of the B_l(k1,k2,k3) path vs the CPU oracle and the reference goldens."""
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


def make_ctx(eng, w):
    kmin, kmax, dk, lmax = w.bk_bins
    mesh = eng.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = eng.get_context(mesh, 0, ("ops", "bk_fisher", "bk_data"))
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    return mesh, ctx


def oracle_setup(w):
    from oracle import sww_oracle as orc
    kmin, kmax, dk, lmax = w.bk_bins
    return orc, orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)


def test_gram_f64_tensor_core(eng, world):
    w = world("toyC")
    mesh, ctx = make_ctx(eng, w)
    rng = np.random.default_rng(0)
    for (m, n, K) in [(5, 3, 1000), (130, 70, 5400), (64, 64, 4097), (200, 9, 15360)]:
        A, B = rng.normal(size=(m, K)), rng.normal(size=(n, K))
        C = ctx.gram(A, B).cpu().numpy()
        assert relerr(C, A @ B.T) < 1e-13, (m, n, K)
    A, B = rng.normal(size=(17, 2048)), rng.normal(size=(11, 2048))
    C0 = rng.normal(size=(17, 11))
    C = ctx.gram(A, B, out=ctx.dev(C0), accumulate=True).cpu().numpy()
    assert relerr(C, C0 + A @ B.T) < 1e-13


@pytest.mark.parametrize("name", WORLDS)
def test_delta_abc(name, world, eng):
    w = world(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    D = ctx.bk_delta().cpu().numpy()
    ref = orc.delta_abc(S, orc.triangle_bins(S.kmin, S.kmax, S.dk, S.n_k))
    assert relerr(D, ref) < 1e-11


@pytest.mark.parametrize("name", WORLDS)
def test_bk_fisher_pair_and_data(name, world, golden, eng):
    w, g = world(name), golden(name)
    mesh, ctx = make_ctx(eng, w)
    orc, S = oracle_setup(w)
    aA, aB = S.grf(2), S.grf(3)
    F, (gA, tgA, gB, tgB) = ctx.bk_fisher_pair(aA, aB)
    F = F.cpu().numpy()
    assert F.shape == g["bk_fish_1"].shape
    assert relerr(F, g["bk_fish_1"]) < TOL
    assert np.array_equal(F, F.T)
    # the g-maps handed to the data step (saved by the reference as .npz)
    assert relerr(gA.cpu().numpy()[:, :, :, :, 3], g["bk_g_plane"]) < TOL
    assert relerr(tgA.cpu().numpy()[:, :, :, :, 3], g["bk_tg_plane"]) < TOL
    # data: paint, g-maps with the MAS factor, 3-field and 1-field terms, Delta, solve
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    gd = ctx.bk_gmaps(diff, data=True)
    q = ctx.bk_q3(gd)
    N_mc = 2
    ctx.bk_q1_accumulate(gd, gA, tgA, N_mc, q)
    ctx.bk_q1_accumulate(gd, gB, tgB, N_mc, q)
    D = ctx.bk_delta().cpu().numpy()
    nt = ctx.n_tri
    q = q.cpu().numpy()
    qa = np.concatenate([q[:, l] / D[l * nt:(l + 1) * nt] for l in range(mesh.n_l)])
    q_o = orc.bk_data_q(S, diff.cpu().numpy(),
                        mc_maps=[orc.compute_g_a(S, aA), orc.compute_g_a(S, aB)], N_mc=N_mc)
    assert relerr(qa, q_o) < TOL
    p = np.matmul(np.linalg.inv(g["bk_fish_comb"]), qa)
    assert relerr(p, g["bk_p"][:, 3:].T.ravel()) < 2e-8     # '%.8e' text
    assert relerr(p, g["bk_p_f64"]) < 1e-9                  # the script's FP64 p_alpha (bk/compute_bk_data.py:417)
