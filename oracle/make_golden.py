"""TEST INFRASTRUCTURE -- generator of tests/golden/*.npz.

Runs the UNMODIFIED reference scripts (FP64-promoted, see oracle/ref_harness.py) on the
deterministic toy worlds of spectra_without_windows_b200/synthetic.py and stores their
outputs.  Must be run in the build container (needs /root/reference):

    python -m oracle.make_golden [world ...]

A second, shipped-precision (complex64) run of the same scripts is stored under the
`c64_` prefix to document the ~1e-7 gap between the reference as shipped and its FP64
promotion (SURVEY.md finding 1).
"""
from __future__ import annotations

import glob
import os
import shutil
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from spectra_without_windows_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def parse_txt(path):
    return np.loadtxt(path, comments="#")


def run_world(name, fp64=True, include_pix=False, rand_nbar=False):
    w = syn.make_world(name)
    root = tempfile.mkdtemp(prefix="sww_gold_")
    out = {}
    try:
        p = syn.write_world_files(w, root, include_pix=include_pix, rand_nbar=rand_nbar)
        ou, cp = rh.bind_world(w, fp64=fp64)
        mc, od = root + "/mc/", root + "/out/"
        # ---- pk Fisher realisations (pk/compute_pk_randoms.py), then data (pk/compute_pk_data.py)
        for r in range(1, w.N_mc + 1):
            cnt = rh.run_script("pk/compute_pk_randoms.py", r, p)
            out["pk_nfft"] = np.asarray(cnt)
            out["pk_qbar_%d" % r] = np.load(glob.glob(mc + "*%d_FKP_pk_q-bar_a_*.npy" % r)[0])
            out["pk_fish_%d" % r] = np.load(glob.glob(mc + "*%d_FKP_pk_fish_a_*.npy" % r)[0])
        cap = []
        rh.run_script("pk/compute_pk_data.py", 1, p, capture=cap)
        out["pk_p_f64"] = cap[-1]                      # p_alpha in full precision (pk/compute_pk_data.py:250)
        out["pk_bias_comb"] = np.load(glob.glob(od + "bias-pk_*.npy")[0])
        out["pk_fish_comb"] = np.load(glob.glob(od + "fisher-pk_*.npy")[0])
        f = glob.glob(od + "pk_*.txt")[0]
        out["pk_txt_name"] = np.asarray(os.path.basename(f))
        out["pk_txt"] = np.asarray(open(f).read())
        out["pk_p"] = parse_txt(f)
        # ---- function-level q_alpha (the script does not save it): pk/compute_pk_data.py:137-206
        if fp64 and not include_pix and not rand_nbar:
            out.update(function_level(w, ou, cp))
        # ---- bk Fisher pair r=1 (seeds 2,3) then bk data
        cnt = rh.run_script("bk/compute_bk_randoms.py", 1, p)
        out["bk_nfft"] = np.asarray(cnt)
        out["bk_fish_1"] = np.load(glob.glob(mc + "*1_FKP_bk_fish_alpha_beta_*.npy")[0])
        z = np.load(glob.glob(mc + "*2_FKP_bk_bias_alpha_*.npz")[0])
        ga, tga = z["g_a"], z["tilde_g_a"]
        out["bk_g_dtype"] = np.asarray(str(ga.dtype))
        out["bk_g_shape"] = np.asarray(ga.shape)
        # a thin sample of the saved maps: the z=3 plane of every (l, bin)
        out["bk_g_plane"] = np.real(ga[:, :, :, :, 3])
        out["bk_tg_plane"] = np.real(tga[:, :, :, :, 3])
        out["bk_g_maximag"] = np.asarray(np.abs(ga.imag).max() / np.abs(ga.real).max())
        cap = []
        rh.run_script("bk/compute_bk_data.py", 1, p, capture=cap)
        out["bk_p_f64"] = cap[-1]                      # p_alpha in full precision (bk/compute_bk_data.py:417)
        out["bk_fish_comb"] = np.load(glob.glob(od + "fisher-bk_*.npy")[0])
        f = glob.glob(od + "bk_*.txt")[0]
        out["bk_txt_name"] = np.asarray(os.path.basename(f))
        out["bk_txt"] = np.asarray(open(f).read())
        out["bk_p"] = parse_txt(f)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return out


def function_level(w, ou, cp):
    """Reference src/ functions called directly, following pk/compute_pk_data.py:137-206 and
    bk/compute_bk_data.py:250-283,350-367, to pin q_alpha (never written to disk by the scripts)
    and single operator outputs."""
    out = {}
    box, grid = np.asarray(w.box, float), np.asarray(w.grid, int)
    data = ou.load_data(1, None, None)
    rand = ou.load_randoms(None, None)
    diff, density = ou.grid_data(data, rand, box, grid, MAS="TSC", return_randoms=False)
    alpha, shot = w.alpha, w.shot_fac
    nbar = w.nbar_r * alpha
    k_grids, r_grids = ou.load_coord_grids(box, grid, density)
    k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
    k_grids /= (1e-12 + k_norm)
    r_grids /= (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
    MAS = ou.load_MAS(box, grid)
    v_cell = 1.0 * box.prod() / (1.0 * grid.prod())
    out["diff_sum"] = np.asarray([diff.sum(), (diff ** 2).sum(), diff[1, 2, 3], diff[-1, 0, 5]])
    out["diff_plane"] = diff[:, :, 2]
    out["MAS_plane"] = MAS[:, :, 3]
    for tag, (kmin, kmax, dk, lmax) in (("pk", w.pk_bins), ("bk", w.bk_bins)):
        Y = ou.compute_spherical_harmonic_functions(lmax)
        filt = ou.compute_filters(kmin, kmax, dk)
        n_k = int((kmax - kmin) / dk)
        Cinv_d = cp.applyCinv_fkp(diff, nbar, MAS, v_cell, shot, include_pix=False)
        if tag == "pk":
            C_a = cp.applyC_alpha(Cinv_d, nbar, MAS, Y, k_grids, r_grids, v_cell, filt, k_norm, n_k, lmax,
                                  include_pix=False, data=True)
            out["pk_q"] = np.asarray([np.real_if_close(np.sum(Cinv_d * c)) / 2.0 for c in C_a]).real
            # one derivative map of a GRF through applyC_alpha_single, a plane of it
            np.random.seed(7)
            a = np.random.normal(size=tuple(grid))
            m = cp.applyC_alpha_single(a, nbar, MAS, Y, k_grids, r_grids, v_cell, filt, k_norm, 1, 2,
                                       include_pix=False, data=False)
            out["pk_Calpha_l2_a2_plane"] = np.real(m[:, :, 1])
    return out


def _planes(m):
    """Thin sample of a complex map: two z-planes (one of them the Nyquist plane when nz is even) + checksums."""
    m = np.asarray(m, dtype=np.complex128)
    nz = m.shape[2]
    return {"p3": m[:, :, 3].copy(), "pny": m[:, :, nz // 2].copy(),
            "sums": np.asarray([m.sum(), (m * m).sum(), np.abs(m).max()])}


def ml_fisher_function_level(cp, a, S):
    """pk/compute_pk_randoms.py:205-281 with weights = ML, composed from the reference's own functions along its
    non-low_mem branch (:229-241, :247-281).  The script itself cannot run this configuration: the low_mem branch
    passes an undefined name (`Yr_lm`, :261) and the other one deletes `nbar` twice (:241, :278)."""
    kw = dict(include_pix=S["include_pix"])
    args_c = (S["nbar"], S["MAS"], S["pk_map"], S["Y"], S["k_grids"], S["r_grids"], S["v_cell"], S["shot"])
    Cinv_a = np.real(cp.applyCinv(a, *args_c, rel_tol=1e-6, verb=0, max_it=30, **kw))
    Ainv_a = a / S["nbar_A"]
    N_Ainv_a = cp.applyN(Ainv_a, S["nbar"], S["MAS"], S["v_cell"], S["shot"], **kw)
    ca = (S["nbar"], S["MAS"], S["Y"], S["k_grids"], S["r_grids"], S["v_cell"], S["filt"], S["k_norm"], S["n_k"], S["lmax"])
    C_a_Cinv = cp.applyC_alpha(Cinv_a, *ca, data=False, **kw)
    C_a_Ainv = cp.applyC_alpha(Ainv_a, *ca, data=False, **kw)
    n_bins = len(C_a_Cinv)
    U = [cp.applyCinv(C_a_Cinv[al], *args_c, rel_tol=1e-6, verb=0, max_it=30, **kw) for al in range(n_bins)]
    fish = np.zeros((n_bins, n_bins))
    qb = np.zeros(n_bins)
    for al in range(n_bins):
        qb[al] = 0.5 * np.real(np.sum(U[al] * N_Ainv_a))
        for be in range(n_bins):
            fish[al, be] = 0.5 * np.real(np.sum(U[al] * C_a_Ainv[be]))
    return qb, 0.5 * (fish + fish.T)


def run_ml(name):
    """Maximum-likelihood weights (the other value of the `weights` key): reference functions applyS / applyC /
    applyCinv (src/covariances_pk.py:10-95,222-327) at function level, the data branch of pk/compute_pk_data.py:194-206,
    the Fisher matrix along pk/compute_pk_randoms.py:205-281, and the unmodified pk/compute_pk_data.py script run
    with weights = ML on Fisher / bias files produced by that composition."""
    import contextlib
    import io
    import warnings
    from scipy.interpolate import interp1d
    w = syn.make_world(name)
    ou, cp = rh.bind_world(w, fp64=True)
    out = {}
    box, grid = np.asarray(w.box, float), np.asarray(w.grid, int)
    data = ou.load_data(1, None, None)
    rand = ou.load_randoms(None, None)
    diff, density = ou.grid_data(data, rand, box, grid, MAS="TSC", return_randoms=False)
    alpha, shot = w.alpha, w.shot_fac
    nbar = w.nbar_r * alpha
    nbar_A = w.nbar_r.copy()
    nbar_A[nbar_A == 0] = min(nbar_A[nbar_A != 0]) * 0.01
    k_grids, r_grids = ou.load_coord_grids(box, grid, density)
    k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
    k_grids /= (1e-12 + k_norm)
    r_grids /= (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
    MAS = ou.load_MAS(box, grid)
    v_cell = 1.0 * box.prod() / (1.0 * grid.prod())
    kmin, kmax, dk, lmax = w.pk_bins
    Y = ou.compute_spherical_harmonic_functions(lmax)
    filt = ou.compute_filters(kmin, kmax, dk)
    n_k = int((kmax - kmin) / dk)
    k_top = 1.01 * float(np.sqrt(np.sum((np.pi * grid / box) ** 2)))
    table = syn.fiducial_pk_table(k_top)
    out["ml_pk_table"] = table
    pk_map = interp1d(table[:, 0], table[:, 1:].T)(k_norm)[:lmax // 2 + 1]      # pk/compute_pk_randoms.py:192-193
    x = np.random.default_rng(5).normal(size=tuple(grid))
    for pix in (False, True):
        tag = "pix" if pix else "nopix"
        Sx = cp.applyS(x, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, include_pix=pix)
        Cx = cp.applyC(x, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, shot, include_pix=pix)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Ci = cp.applyCinv(diff, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, shot, rel_tol=1e-6, verb=1,
                              max_it=30, include_pix=pix)
        txt = buf.getvalue()
        out["ml_cg_log_" + tag] = np.asarray(txt)
        for key, m in (("S", Sx), ("C", Cx), ("Cinv_d", Ci)):
            for kk, v in _planes(m).items():
                out["ml_%s_%s_%s" % (key, tag, kk)] = v
        # absolute-tolerance stopping rule, few iterations
        Ci2 = cp.applyCinv(diff, nbar, MAS, pk_map, Y, k_grids, r_grids, v_cell, shot, abs_tol=1e-3, verb=0, max_it=4,
                           include_pix=pix)
        for kk, v in _planes(Ci2).items():
            out["ml_Cinv_d_it4_%s_%s" % (tag, kk)] = v
        if not pix:
            C_a = cp.applyC_alpha(Ci, nbar, MAS, Y, k_grids, r_grids, v_cell, filt, k_norm, n_k, lmax, include_pix=False,
                                  data=True)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out["ml_pk_q"] = np.asarray([np.real(np.sum(Ci * c)) / 2.0 for c in C_a])
    # ---- Fisher matrix and bias with ML weights for seeds 1..N_mc (function-level composition)
    S = dict(nbar=nbar, nbar_A=nbar_A, MAS=MAS, pk_map=pk_map, Y=Y, k_grids=k_grids, r_grids=r_grids, v_cell=v_cell,
             shot=shot, filt=filt, k_norm=k_norm, n_k=n_k, lmax=lmax, include_pix=False)
    root = tempfile.mkdtemp(prefix="sww_gold_ml_")
    try:
        p = syn.write_world_files(w, root, weights="ML")
        mc, od = root + "/mc/", root + "/out/"
        for r in range(1, w.N_mc + 1):
            np.random.seed(r)                                                  # pk/compute_pk_randoms.py:139-140
            a = np.random.normal(loc=0.0, scale=np.sqrt(nbar_A), size=nbar_A.shape)
            qb, fish = ml_fisher_function_level(cp, a, S)
            out["ml_pk_qbar_%d" % r], out["ml_pk_fish_%d" % r] = qb, fish
            kk = (kmin, kmax, dk, lmax)
            np.save(mc + "%s%d_ML_pk_q-bar_a_k%.3f_%.3f_%.3f_l%d.npy" % ((w.type, r) + kk), qb)   # :116-117
            np.save(mc + "%s%d_ML_pk_fish_a_k%.3f_%.3f_%.3f_l%d.npy" % ((w.type, r) + kk), fish)
        # include_pix = True, one seed (function level only)
        np.random.seed(1)
        a = np.random.normal(loc=0.0, scale=np.sqrt(nbar_A), size=nbar_A.shape)
        out["ml_pk_qbar_pix_1"], out["ml_pk_fish_pix_1"] = ml_fisher_function_level(cp, a, dict(S, include_pix=True))
        rh.run_script("pk/compute_pk_data.py", 1, p)
        f = glob.glob(od + "pk_*_ML_*.txt")[0]
        out["ml_pk_txt_name"] = np.asarray(os.path.basename(f))
        out["ml_pk_txt"] = np.asarray(open(f).read())
        out["ml_pk_p"] = parse_txt(f)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return out


def main(argv):
    names = argv[1:] or ["toyA", "toyB", "toyC"]
    os.makedirs(GOLD, exist_ok=True)
    for n in [x for x in names if x.endswith("+ml")]:
        g = run_ml(n.split("+")[0])
        np.savez_compressed(os.path.join(GOLD, "%s_ml.npz" % n.split("+")[0]), **g)
        print(n, str(g["ml_cg_log_nopix"]).strip().splitlines()[:2], g["ml_pk_fish_1"].shape)
    names = [x for x in names if not x.endswith("+ml")]
    for n in [x for x in names if "+" in x]:
        # variants: toyA+pix (include_pix = True), toyA+rand (rand_nbar = True)
        base, var = n.split("+")
        g = run_world(base, fp64=True, include_pix=(var == "pix"), rand_nbar=(var == "rand"))
        np.savez_compressed(os.path.join(GOLD, "%s_%s.npz" % (base, var)), **g)
        print(n, "pk", g["pk_fish_1"].shape, "bk", g["bk_fish_1"].shape)
    for n in [x for x in names if "+" not in x]:
        g = run_world(n, fp64=True)
        c = run_world(n, fp64=False)
        for k in ("pk_fish_1", "pk_qbar_1", "pk_p", "bk_fish_1", "bk_p"):
            g["c64_" + k] = c[k]
        np.savez_compressed(os.path.join(GOLD, n + ".npz"), **g)
        print(n, "pk nfft", g["pk_nfft"], "bk nfft", g["bk_nfft"], "n_b pk", g["pk_fish_1"].shape,
              "bk", g["bk_fish_1"].shape, "g dtype", g["bk_g_dtype"], "im/re", g["bk_g_maximag"])


if __name__ == "__main__":
    main(sys.argv)
