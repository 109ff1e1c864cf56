"""Maximum-likelihood weights on the CUDA path (csrc/ml.cu) against golden vectors produced by the reference's
own functions (oracle/make_golden.py toyA+ml) and against the CPU oracle on the power-of-two world."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from spectra_without_windows_b200 import engine
    return engine


def _planes(m):
    m = np.asarray(m)
    nz = m.shape[2]
    return {"p3": m[:, :, 3], "pny": m[:, :, nz // 2], "sums": np.asarray([m.sum(), (m * m).sum(), np.abs(m).max()])}


def _check(m, g, key, tol=TOL):
    for kk, v in _planes(m.cpu().numpy()).items():
        assert relerr(v, g[key + "_" + kk]) < tol, (key, kk)


def make_ctx(eng, w, table, pix=False):
    mesh = eng.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = eng.Context(mesh, 0, ("ops", "pk_ml"))
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    ctx.set_include_pix(pix)
    ctx.set_fiducial_pk(table)
    return mesh, ctx


@pytest.mark.parametrize("pix", [False, True])
def test_ml_operators_against_reference_functions(eng, world, golden, pix):
    w, g = world("toyA"), golden("toyA_ml")
    mesh, ctx = make_ctx(eng, w, g["ml_pk_table"], pix)
    tag = "pix" if pix else "nopix"
    x = np.random.default_rng(5).normal(size=tuple(w.grid))
    _check(ctx.ml_apply("S", x), g, "ml_S_" + tag)
    _check(ctx.ml_apply("C", x), g, "ml_C_" + tag)
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    z, it, why = ctx.ml_apply_cinv(diff, max_it=30, rel_tol=1e-6)
    assert why == 1 and "after %d iterations" % it in str(g["ml_cg_log_" + tag])
    _check(z, g, "ml_Cinv_d_" + tag)
    z4, it4, why4 = ctx.ml_apply_cinv(diff, max_it=4, abs_tol=1e-3)
    assert it4 == 4 and why4 == 0
    _check(z4, g, "ml_Cinv_d_it4_" + tag)
    if not pix:
        q = ctx.pk_q_from_cinv(z).cpu().numpy()
        assert relerr(q, g["ml_pk_q"]) < TOL
    ctx.close()


def test_ml_fisher_realisation(eng, world, golden):
    from oracle import sww_oracle as orc
    w, g = world("toyA"), golden("toyA_ml")
    mesh, ctx = make_ctx(eng, w, g["ml_pk_table"])
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
    for r in (1, 2):
        F, qbar, its = ctx.pk_fisher_ml(S.grf(r))
        assert its > mesh.n_k * mesh.n_l
        assert relerr(F.cpu().numpy(), g["ml_pk_fish_%d" % r]) < TOL
        assert relerr(qbar.cpu().numpy(), g["ml_pk_qbar_%d" % r]) < TOL
    ctx.set_include_pix(True)          # MAS forward-modelled in S, N, the preconditioner and C_,alpha
    F, qbar, _ = ctx.pk_fisher_ml(S.grf(1))
    assert relerr(F.cpu().numpy(), g["ml_pk_fish_pix_1"]) < TOL
    assert relerr(qbar.cpu().numpy(), g["ml_pk_qbar_pix_1"]) < TOL
    ctx.close()


def test_ml_on_power_of_two_mesh_against_oracle(eng, world, golden):
    """64 x 64 x 128: every transform of the ML operators -- real and complex -- runs on the hand-written passes (complex z
    pass + column passes over the full box; no cuFFT plan is involved on power-of-two meshes with nz <= 512)."""
    from oracle import sww_oracle as orc
    w = world("toyD")
    table = golden("toyA_ml")["ml_pk_table"]
    table = np.concatenate([table, [[10.0, table[-1, 1], table[-1, 2], table[-1, 3]]]])   # cover this mesh's |k|
    mesh, ctx = make_ctx(eng, w, table)
    assert ctx.fft_backend == 1
    kmin, kmax, dk, lmax = w.pk_bins
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
    pk_map = orc.pk_map_from_table(S, table)
    diff = ctx.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box_center)
    d = diff.cpu().numpy()
    info = {}
    z_o = orc.applyCinv(S, d, pk_map, rel_tol=1e-6, max_it=30, info=info)
    z, it, _ = ctx.ml_apply_cinv(diff, max_it=30, rel_tol=1e-6)
    assert it == info["iterations"]
    assert relerr(z.cpu().numpy(), z_o) < TOL
    q, _ = ctx.pk_data_q_ml(diff)
    C_a = S.applyC_alpha(z_o, data=True)
    q_o = np.asarray([np.real(np.sum(z_o * c)) / 2.0 for c in C_a])
    assert relerr(q.cpu().numpy(), q_o) < TOL
    ctx.close()


def test_ml_loud_errors(eng, world, golden):
    w = world("toyA")
    mesh = eng.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    ctx = eng.Context(mesh, 0, ("ops", "pk_ml"))
    ctx.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    with pytest.raises(Exception, match="fiducial"):
        ctx.ml_apply("S", np.ones(tuple(w.grid)))
    with pytest.raises(ValueError, match="interpolation range"):
        ctx.set_fiducial_pk(np.asarray([[0.0, 1.0, 1.0, 1.0], [0.01, 1.0, 1.0, 1.0]]))
    small = eng.Context(mesh, 0, ("ops",))
    small.set_fields_from_randoms_density(w.nbar_r, w.alpha, w.shot_fac)
    small.set_fiducial_pk(golden("toyA_ml")["ml_pk_table"])
    with pytest.raises(Exception, match="workspace too small"):
        small.ml_apply("S", np.ones(tuple(w.grid)))


def test_ml_mirrors_keep_the_reference_call_surface(world, golden):
    """covariances_pk.applyS / applyC / applyCinv with the reference's positional arguments (src/covariances_pk.py:10,45,222)."""
    from scipy.interpolate import interp1d
    from spectra_without_windows_b200 import covariances_pk as cp
    from spectra_without_windows_b200 import opt_utilities as ou
    w, g = world("toyA"), golden("toyA_ml")
    box, grid = np.asarray(w.box), np.asarray(w.grid)
    kmin, kmax, dk, lmax = w.pk_bins
    data = ou.Catalog(w.data["pos"], w.data["w"], box_center=w.box_center)
    rand = ou.Catalog(w.randoms["pos"], w.randoms["w"], box_center=w.box_center)
    diff, density = ou.grid_data(data, rand, box, grid, MAS="TSC", return_randoms=False)
    k_grids, r_grids = ou.load_coord_grids(box, grid, density)
    k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
    k_grids /= (1e-12 + k_norm)
    r_grids /= (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
    MAS = ou.load_MAS(box, grid)
    Y = ou.compute_spherical_harmonic_functions(lmax)
    nbar = w.nbar_r * w.alpha
    v_cell = 1.0 * box.prod() / (1.0 * grid.prod())
    table = g["ml_pk_table"]
    pk_map = interp1d(table[:, 0], table[:, 1:].T)(np.asarray(k_norm))[:lmax // 2 + 1]   # pk/compute_pk_randoms.py:192-193
    x = np.random.default_rng(5).normal(size=tuple(grid))
    Sx = cp.applyS(x, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, include_pix=False)
    assert Sx.dtype == np.complex128
    for kk, v in _planes(Sx).items():
        assert relerr(v, g["ml_S_nopix_" + kk]) < TOL
    Cx = cp.applyC(x, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, w.shot_fac, include_pix=True)
    for kk, v in _planes(Cx).items():
        assert relerr(v, g["ml_C_pix_" + kk]) < TOL
    Ci = cp.applyCinv(diff, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, w.shot_fac, rel_tol=1e-6, verb=0, max_it=30,
                      include_pix=False)
    for kk, v in _planes(Ci).items():
        assert relerr(v, g["ml_Cinv_d_nopix_" + kk]) < TOL
    with pytest.raises(Exception, match="hooks"):
        cp.applyCinv(diff, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, w.shot_fac, applyC=lambda *a, **k: 0)


def test_ml_scripts_end_to_end(world, golden, tmp_path):
    """weights = ML through the drop-in CLIs: per-seed Fisher / bias files, then the P_l(k) text, against the files the
    reference's own functions + its unmodified pk/compute_pk_data.py produce."""
    import glob
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from spectra_without_windows_b200 import synthetic as syn
    w, g = world("toyA"), golden("toyA_ml")
    p = syn.write_world_files(w, str(tmp_path), weights="ML")
    mc, od = str(tmp_path) + "/mc/", str(tmp_path) + "/out/"

    def run(script, arg):
        r = subprocess.run([sys.executable, os.path.join(ROOT, script), str(arg), p], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return r.stdout

    for r in (1, 2):
        out = run("pk/compute_pk_randoms.py", r)
        assert "conjugate-gradient solves" in out
        F = np.load(glob.glob(mc + "*%d_ML_pk_fish_a_*.npy" % r)[0])
        qb = np.load(glob.glob(mc + "*%d_ML_pk_q-bar_a_*.npy" % r)[0])
        assert relerr(F, g["ml_pk_fish_%d" % r]) < TOL and relerr(qb, g["ml_pk_qbar_%d" % r]) < TOL
    run("pk/compute_pk_data.py", 1)
    f = glob.glob(od + "pk_*.txt")[0]
    assert os.path.basename(f) == str(g["ml_pk_txt_name"])
    assert open(f).read().split("\n")[:15] == str(g["ml_pk_txt"]).split("\n")[:15]
    assert relerr(np.loadtxt(f), g["ml_pk_p"]) < 2e-8
    # the sharded Monte-Carlo driver produces the same combined Fisher / bias files from scratch
    for f in glob.glob(od + "fisher-pk_*") + glob.glob(od + "bias-pk_*"):
        os.remove(f)
    r = subprocess.run([sys.executable, "-m", "spectra_without_windows_b200.mc_driver", "pk", p], capture_output=True, text=True,
                       cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    fish = 0.5 * (g["ml_pk_fish_1"] + g["ml_pk_fish_2"])
    bias = 0.5 * (g["ml_pk_qbar_1"] + g["ml_pk_qbar_2"])
    assert relerr(np.load(glob.glob(od + "fisher-pk_*_ML_*.npy")[0]), fish) < TOL
    assert relerr(np.load(glob.glob(od + "bias-pk_*_ML_*.npy")[0]), bias) < TOL
    # the bispectrum scripts refuse ML weights loudly
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bk/compute_bk_randoms.py"), "1", p], capture_output=True, text=True)
    assert r.returncode != 0 and "ML" in r.stderr
