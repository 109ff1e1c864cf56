"""Pins the CPU oracle (oracle/sww_oracle.py) against the golden vectors produced by the
UNMODIFIED reference scripts (oracle/make_golden.py, FP64-promoted mode).  CPU only."""
import numpy as np
import pytest

from conftest import relerr
from oracle import sww_oracle as orc

WORLDS = ["toyA", "toyB", "toyC", "toyD"]
TOL = 1e-12      # two FP64 evaluations of the same algorithm


def setup_for(w, which):
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    return orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)


@pytest.mark.parametrize("name", WORLDS)
def test_pk_fisher_matches_reference_script(name, world, golden):
    w, g = world(name), golden(name)
    S = setup_for(w, "pk")
    for r in (1, 2):
        qbar, F = orc.pk_fisher_realisation(S, S.grf(r))
        assert relerr(F, g["pk_fish_%d" % r]) < TOL
        assert relerr(qbar, g["pk_qbar_%d" % r]) < TOL


@pytest.mark.parametrize("name", WORLDS)
def test_paint_and_pk_data(name, world, golden):
    w, g = world(name), golden(name)
    S = setup_for(w, "pk")
    diff, _ = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    assert relerr(diff[:, :, 2], g["diff_plane"]) < TOL
    assert relerr(S.MAS[:, :, 3], g["MAS_plane"]) < 1e-15
    q = orc.pk_data_q(S, diff)
    assert relerr(q, g["pk_q"]) < TOL
    p = orc.pk_solve(g["pk_fish_comb"], q, g["pk_bias_comb"])
    n_k = S.n_k
    p_txt = g["pk_p"][:, 1:].T.ravel()           # text rows are k | P0 P2 P4 -> alpha = l*n_k + a
    assert p.size == n_k * S.n_l
    assert relerr(p, p_txt) < 2e-8               # limited by the '%.8e' text format
    assert relerr(p, g["pk_p_f64"]) < 1e-10     # the script's own FP64 p_alpha (pk/compute_pk_data.py:250)


@pytest.mark.parametrize("name", WORLDS)
def test_applyC_alpha_single(name, world, golden):
    w, g = world(name), golden(name)
    S = setup_for(w, "pk")
    np.random.seed(7)
    a = np.random.normal(size=tuple(w.grid))
    maps = S.applyC_alpha(a)
    m = np.real(maps[1 * S.n_k + 2])
    assert relerr(m[:, :, 1], g["pk_Calpha_l2_a2_plane"]) < TOL


@pytest.mark.parametrize("name", WORLDS)
def test_bk_fisher_and_data(name, world, golden):
    w, g = world(name), golden(name)
    S = setup_for(w, "bk")
    F, (gA, tgA), (gB, tgB) = orc.bk_fisher_pair(S, S.grf(2), S.grf(3), return_g=True)
    assert F.shape == g["bk_fish_1"].shape
    assert relerr(F, g["bk_fish_1"]) < 1e-11
    gplane = np.asarray([[np.real(m[:, :, 3]) for m in row] for row in gA])
    assert relerr(gplane, g["bk_g_plane"]) < TOL
    tgplane = np.asarray([[np.real(m[:, :, 3]) for m in row] for row in tgA])
    assert relerr(tgplane, g["bk_tg_plane"]) < TOL
    diff, _ = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    q = orc.bk_data_q(S, diff, mc_maps=[(gA, tgA), (gB, tgB)], N_mc=2)
    p = np.matmul(np.linalg.inv(g["bk_fish_comb"]), q)
    nt = len(orc.triangle_bins(S.kmin, S.kmax, S.dk, S.n_k))
    p_txt = g["bk_p"][:, 3:].T.ravel()
    assert p.size == nt * S.n_l
    assert relerr(p, p_txt) < 2e-8
    assert relerr(p, g["bk_p_f64"]) < 1e-9      # the script's own FP64 p_alpha (bk/compute_bk_data.py:417)


@pytest.mark.parametrize("name", WORLDS)
def test_shipped_precision_gap_documented(name, golden):
    """The reference as shipped (complex64) differs from its FP64 promotion at the ~1e-7..1e-5
    level; this is why the 1e-9 parity target is stated against the promoted run."""
    g = golden(name)
    for k in ("pk_fish_1", "pk_qbar_1", "bk_fish_1"):
        e = relerr(g["c64_" + k], g[k])
        assert 1e-10 < e < 1e-4, (k, e)


def _variant_setup(w, which, variant):
    kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
    kw = {}
    if variant == "pix":
        kw["include_pix"] = True
    if variant == "rand":      # n(r) from the painted randoms (src/opt_utilities.py:262-267)
        _, rnd = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
        kw["nbar_override"] = rnd
    return orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, **kw)


@pytest.mark.parametrize("variant", ["pix", "rand"])
def test_param_booleans_match_reference_scripts(variant, world, golden):
    """include_pix = True and rand_nbar = True (the remaining .param booleans) against the unmodified
    reference scripts run with those settings."""
    w, g = world("toyA"), golden("toyA_" + variant)
    S = _variant_setup(w, "pk", variant)
    for r in (1, 2):
        qbar, F = orc.pk_fisher_realisation(S, S.grf(r))
        assert relerr(F, g["pk_fish_%d" % r]) < 1e-11
        assert relerr(qbar, g["pk_qbar_%d" % r]) < 1e-11
    diff, _ = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    q = orc.pk_data_q(S, diff)
    p = orc.pk_solve(g["pk_fish_comb"], q, g["pk_bias_comb"])
    assert relerr(p, g["pk_p"][:, 1:].T.ravel()) < 2e-8
    assert relerr(p, g["pk_p_f64"]) < 1e-10
    Sb = _variant_setup(w, "bk", variant)
    F, (gA, tgA), (gB, tgB) = orc.bk_fisher_pair(Sb, Sb.grf(2), Sb.grf(3), return_g=True)
    assert relerr(F, g["bk_fish_1"]) < 1e-10
    qb = orc.bk_data_q(Sb, diff, mc_maps=[(gA, tgA), (gB, tgB)], N_mc=2)
    pb = np.matmul(np.linalg.inv(g["bk_fish_comb"]), qb)
    assert relerr(pb, g["bk_p"][:, 3:].T.ravel()) < 2e-8
    assert relerr(pb, g["bk_p_f64"]) < 1e-9


# ---- maximum-likelihood weights (src/covariances_pk.py:10-95, 222-327): restatement vs reference functions
def _planes(m):
    nz = m.shape[2]
    return {"p3": m[:, :, 3], "pny": m[:, :, nz // 2], "sums": np.asarray([m.sum(), (m * m).sum(), np.abs(m).max()])}


def _check_planes(m, g, key, tol):
    for kk, v in _planes(np.asarray(m, dtype=np.complex128)).items():
        assert relerr(v, g[key + "_" + kk]) < tol, (key, kk)


@pytest.mark.parametrize("name", ["toyA", "toyC"])     # toyC: odd mesh sizes (no Nyquist planes), k_min > 0, lmax = 2
@pytest.mark.parametrize("pix", [False, True])
def test_ml_operators_match_reference_functions(pix, name, world, golden):
    w, g = world(name), golden(name + "_ml")
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, include_pix=pix)
    pk_map = orc.pk_map_from_table(S, g["ml_pk_table"])
    tag = "pix" if pix else "nopix"
    x = np.random.default_rng(5).normal(size=tuple(w.grid))
    _check_planes(orc.applyS(S, x, pk_map), g, "ml_S_" + tag, 1e-12)
    _check_planes(orc.applyC(S, x, pk_map), g, "ml_C_" + tag, 1e-12)
    diff, _ = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    info = {}
    Ci = orc.applyCinv(S, diff, pk_map, rel_tol=1e-6, max_it=30, info=info)
    assert "after %d iterations" % info["iterations"] in str(g["ml_cg_log_" + tag])
    _check_planes(Ci, g, "ml_Cinv_d_" + tag, 1e-10)
    _check_planes(orc.applyCinv(S, diff, pk_map, abs_tol=1e-3, max_it=4), g, "ml_Cinv_d_it4_" + tag, 1e-11)
    if not pix:
        assert relerr(orc.pk_data_q_ml(S, diff, pk_map), g["ml_pk_q"]) < 1e-10


def test_ml_fisher_and_spectrum(world, golden):
    w, g = world("toyA"), golden("toyA_ml")
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
    pk_map = orc.pk_map_from_table(S, g["ml_pk_table"])
    qbar, F = orc.pk_fisher_realisation_ml(S, S.grf(1), pk_map)
    assert relerr(F, g["ml_pk_fish_1"]) < 1e-10
    assert relerr(qbar, g["ml_pk_qbar_1"]) < 1e-10
    Sp = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, include_pix=True)
    qbar, F = orc.pk_fisher_realisation_ml(Sp, Sp.grf(1), pk_map)
    assert relerr(F, g["ml_pk_fish_pix_1"]) < 1e-10 and relerr(qbar, g["ml_pk_qbar_pix_1"]) < 1e-10
    # the unmodified pk/compute_pk_data.py run with weights = ML on those Fisher / bias files
    fish = 0.5 * (g["ml_pk_fish_1"] + g["ml_pk_fish_2"])
    bias = 0.5 * (g["ml_pk_qbar_1"] + g["ml_pk_qbar_2"])
    p = orc.pk_solve(fish, g["ml_pk_q"], bias)
    assert relerr(p, g["ml_pk_p"][:, 1:].T.ravel()) < 2e-8
    assert "_ML_" in str(g["ml_pk_txt_name"])
