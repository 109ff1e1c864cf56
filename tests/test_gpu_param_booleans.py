"""include_pix = True and rand_nbar = True (the remaining .param booleans) on the GPU path against the
golden files of the unmodified reference scripts run with those settings, at function and script level."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("variant", ["pix", "rand"])
def test_pipelines(variant, world, golden):
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import engine
    w, g = world("toyA"), golden("toyA_" + variant)
    for which in ("pk", "bk"):
        kmin, kmax, dk, lmax = w.pk_bins if which == "pk" else w.bk_bins
        mesh = engine.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
        ctx = engine.Context(mesh, 0, ("ops", "pk_fisher", "pk_data", "bk_fisher", "bk_data"))
        nbar_A = np.array(w.nbar_r, copy=True)
        nbar_A[nbar_A == 0] = min(nbar_A[nbar_A != 0]) * 0.01
        nbar_mask = w.nbar_r * w.alpha
        diff, rnd = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center, return_randoms=True)
        ctx.set_fields(rnd if variant == "rand" else nbar_mask, nbar_mask, nbar_A, w.shot_fac)
        ctx.set_include_pix(variant == "pix")
        S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
        if which == "pk":
            for r in (1, 2):
                F, qbar = ctx.pk_fisher(S.grf(r))
                assert relerr(F.cpu().numpy(), g["pk_fish_%d" % r]) < TOL
                assert relerr(qbar.cpu().numpy(), g["pk_qbar_%d" % r]) < TOL
            q = ctx.pk_data_q(diff).cpu().numpy()
            p = np.inner(np.linalg.inv(g["pk_fish_comb"]), q - g["pk_bias_comb"])
            assert relerr(p, g["pk_p"][:, 1:].T.ravel()) < 2e-8
            assert relerr(p, g["pk_p_f64"]) < 1e-9
        else:
            F, (gA, tgA, gB, tgB) = ctx.bk_fisher_pair(S.grf(2), S.grf(3))
            assert relerr(F.cpu().numpy(), g["bk_fish_1"]) < TOL
            assert relerr(gA.cpu().numpy()[:, :, :, :, 3], g["bk_g_plane"]) < TOL
            gd = ctx.bk_gmaps(diff, data=True)
            q = ctx.bk_q3(gd)
            ctx.bk_q1_accumulate(gd, gA, tgA, 2, q)
            ctx.bk_q1_accumulate(gd, gB, tgB, 2, q)
            D = ctx.bk_delta().cpu().numpy()
            nt = ctx.n_tri
            q = q.cpu().numpy()
            qa = np.concatenate([q[:, l] / D[l * nt:(l + 1) * nt] for l in range(mesh.n_l)])
            p = np.matmul(np.linalg.inv(g["bk_fish_comb"]), qa)
            assert relerr(p, g["bk_p"][:, 3:].T.ravel()) < 2e-8
            assert relerr(p, g["bk_p_f64"]) < 1e-9
        ctx.close()


def test_include_pix_mirrors(world):
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import covariances_pk as cp
    from spectra_without_windows_b200 import opt_utilities as ou
    w = world("toyA")
    box, grid = np.asarray(w.box), np.asarray(w.grid)
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, include_pix=True)
    density = ou.MeshInfo(box, grid, w.box_center)
    k_grids, r_grids = ou.load_coord_grids(box, grid, density)
    k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
    k_grids /= (1e-12 + k_norm)
    r_grids /= (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
    MAS = ou.load_MAS(box, grid)
    Y = ou.compute_spherical_harmonic_functions(lmax)
    filt = ou.compute_filters(kmin, kmax, dk)
    x = np.random.default_rng(1).normal(size=tuple(grid))
    got = cp.applyCinv_fkp(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=True)
    assert relerr(np.real(got), np.real(orc.applyCinv_fkp(x, S.nbar, S.v_cell, w.shot_fac, MAS=S.MAS))) < 1e-12
    got = cp.applyN(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=True)
    assert relerr(got, np.real(orc.applyN(x, S.nbar, S.v_cell, w.shot_fac, MAS=S.MAS))) < 1e-12
    ref = S.applyC_alpha(x, data=True)
    one = cp.applyC_alpha_single(x, S.nbar, MAS, Y, k_grids, r_grids, S.v_cell, filt, k_norm, 1, 3, include_pix=True, data=True)
    assert relerr(np.real(one), np.real(ref[1 * S.n_k + 3])) < 1e-12


def test_scripts_with_include_pix(world, golden, tmp_path):
    from spectra_without_windows_b200 import synthetic as syn
    w, g = world("toyA"), golden("toyA_pix")
    p = syn.write_world_files(w, str(tmp_path), include_pix=True)
    mc, od = str(tmp_path) + "/mc/", str(tmp_path) + "/out/"
    for script, arg in (("pk/compute_pk_randoms.py", 1), ("pk/compute_pk_randoms.py", 2), ("pk/compute_pk_data.py", 1),
                        ("bk/compute_bk_randoms.py", 1), ("bk/compute_bk_data.py", 1)):
        # FP64 goldens: complex128 bias files (the default, complex64 like the reference, costs ~1e-7 on B_l)
        r = subprocess.run([sys.executable, os.path.join(ROOT, script), str(arg), p], capture_output=True, text=True,
                           env=dict(os.environ, SWW_BIAS_DTYPE="complex128"))
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    f = glob.glob(od + "pk_*.txt")[0]
    assert open(f).read().split("\n")[:15] == str(g["pk_txt"]).split("\n")[:15]     # header says pixellation : 1
    assert relerr(np.loadtxt(f), g["pk_p"]) < 2e-8
    assert relerr(np.loadtxt(glob.glob(od + "bk_*.txt")[0]), g["bk_p"]) < 2e-8


def test_include_pix_on_own_fft_passes(world):
    """include_pix = True through the hand-written FFT backend (64x64x128) against the live oracle."""
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import engine
    w = world("toyD")
    kmin, kmax, dk, lmax = w.pk_bins
    mesh = engine.Mesh(w.box, w.grid, w.box_center, kmin, kmax, dk, lmax)
    ctx = engine.Context(mesh, 0, ("ops", "pk_fisher", "pk_data"))
    assert ctx.fft_backend == 1
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    ctx.set_include_pix(True)
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax, include_pix=True)
    a = S.grf(4)
    F, qbar = ctx.pk_fisher(a)
    qb_o, F_o = orc.pk_fisher_realisation(S, a)
    assert relerr(F.cpu().numpy(), F_o) < TOL and relerr(qbar.cpu().numpy(), qb_o) < TOL
    d = np.random.default_rng(2).normal(size=mesh.shape) * 1e-4
    assert relerr(ctx.pk_data_q(d).cpu().numpy(), orc.pk_data_q(S, d)) < TOL
    ctx.close()
