"""The opt_utilities / covariances_pk mirrors: same call surface as the reference, GPU-backed."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def test_mirror_call_surface(world):
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import covariances_pk as cp
    from spectra_without_windows_b200 import opt_utilities as ou
    w = world("toyA")
    box, grid = np.asarray(w.box), np.asarray(w.grid)
    kmin, kmax, dk, lmax = w.pk_bins
    data = ou.Catalog(w.data["pos"], w.data["w"], box_center=w.box_center)
    rand = ou.Catalog(w.randoms["pos"], w.randoms["w"], box_center=w.box_center)
    diff, rnd, density = ou.grid_data(data, rand, box, grid, MAS="TSC", return_randoms=True)
    dref, rref = orc.grid_data(w.data["pos"], w.data["w"], w.randoms["pos"], w.randoms["w"], w.box, w.grid, w.box_center)
    assert relerr(diff, dref) < 1e-12 and relerr(rnd, rref) < 1e-12
    k_grids, r_grids = ou.load_coord_grids(box, grid, density)
    k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
    k_grids /= (1e-12 + k_norm)
    r_grids /= (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
    MAS = ou.load_MAS(box, grid)
    assert np.array_equal(MAS, orc.load_MAS(box, grid))
    Y = ou.compute_spherical_harmonic_functions(lmax)
    Yo = orc.compute_spherical_harmonic_functions(lmax)
    for a, b in zip(Y, Yo):
        for f, fo in zip(a, b):
            assert (f.l, f.m) == (fo.l, fo.m)
            assert np.max(np.abs(f(*r_grids) - fo(*r_grids))) < 1e-14
    filt = ou.compute_filters(kmin, kmax, dk)
    n_k = int((kmax - kmin) / dk)
    S = orc.Setup(w.box, w.grid, w.box_center, w.nbar_r, w.alpha, w.shot_fac, kmin, kmax, dk, lmax)
    x = np.random.default_rng(0).normal(size=tuple(grid))
    assert relerr(ou.ft(x), orc.ft(x)) < 1e-13 and relerr(ou.ift(x), orc.ift(x)) < 1e-13
    got = cp.applyCinv_fkp(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=False)
    assert relerr(got, orc.applyCinv_fkp(x, S.nbar, S.v_cell, w.shot_fac)) < 1e-14
    got = cp.applyN(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=False)
    assert relerr(got, orc.applyN(x, S.nbar, S.v_cell, w.shot_fac)) < 1e-14
    maps = cp.applyC_alpha(x, S.nbar, MAS, Y, k_grids, r_grids, S.v_cell, filt, k_norm, n_k, lmax, include_pix=False, data=True)
    ref = S.applyC_alpha(x, data=True)
    assert len(maps) == len(ref)
    for i in (0, 7, len(ref) - 1):
        assert relerr(maps[i], np.real(ref[i])) < 1e-12
    one = cp.applyC_alpha_single(x, S.nbar, MAS, Y, k_grids, r_grids, S.v_cell, filt, k_norm, 2, 3, include_pix=False)
    assert relerr(one, np.real(S.applyC_alpha(x)[2 * n_k + 3])) < 1e-12
    # include_pix = True branches (src/covariances_pk.py:122, :155-162)
    got = cp.applyCinv_fkp(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=True)
    assert relerr(got, orc.applyCinv_fkp(x, S.nbar, S.v_cell, w.shot_fac, MAS=MAS)) < 1e-12
    got = cp.applyN(x, S.nbar, MAS, S.v_cell, w.shot_fac, include_pix=True)
    assert relerr(got, orc.applyN(x, S.nbar, S.v_cell, w.shot_fac, MAS=MAS)) < 1e-12
    # ML weights are not built (dead code in the shipped reference): loud failure, never a CPU fallback
    with pytest.raises(Exception):
        cp.applyCinv(x)
