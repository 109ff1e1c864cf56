"""Drop-in mirror of the reference's src/covariances_pk.py FKP operators, backed by libsww.so.

Same names, argument order and meaning as the reference; ndarray in -> ndarray out.  Only the
branch every shipped paramfile uses is implemented on the GPU (`include_pix=False`, FKP weights);
the other branches raise (SURVEY.md section 8f rows N1/N2 -- not yet built, never a CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import engine


def _need_no_pix(include_pix):
    if include_pix:
        raise Exception("include_pix=True (forward-modelled pixellation) is not implemented on the B200 path yet")


def _info(k_grids, r_grids):
    info = getattr(k_grids, "info", None) or getattr(r_grids, "info", None)
    if info is None:
        raise Exception("k_grids/r_grids must come from spectra_without_windows_b200.opt_utilities.load_coord_grids "
                        "(they carry the mesh geometry the CUDA kernels rebuild k-hat and r-hat from)")
    return info


def _context(k_grids, r_grids, k_filters, lmax):
    info = _info(k_grids, r_grids)
    for att in ("kmin", "kmax", "dk"):
        if not hasattr(k_filters, att):
            raise Exception("k_filters must come from spectra_without_windows_b200.opt_utilities.compute_filters")
    mesh = engine.Mesh(info["box"], info["grid"], info["box_center"], k_filters.kmin, k_filters.kmax, k_filters.dk, lmax)
    return engine.get_context(mesh, 0, ("ops",))


def _any_context(input_map):
    from .opt_utilities import _fft_ctx
    return _fft_ctx(np.asarray(input_map).shape)


def applyN(input_map, nbar, MAS_mat, v_cell, shot_fac, include_pix=True):
    """src/covariances_pk.py:97-124 (include_pix=False branch): shot_fac*nbar*input_map/v_cell."""
    _need_no_pix(include_pix)
    ctx = _any_context(input_map)
    return ctx.apply_n(np.real(input_map), nbar, float(v_cell), float(shot_fac)).cpu().numpy()


def applyCinv_fkp(input_map, nbar, MAS_mat, v_cell, shot_fac, P_fkp=1e4, include_pix=True):
    """src/covariances_pk.py:126-164 (include_pix=False branch).  Returns a complex128 map like the
    FP64-promoted reference (the reference's buffer is complex64, :154)."""
    _need_no_pix(include_pix)
    ctx = _any_context(input_map)
    x = np.asarray(input_map)
    out = ctx.apply_cinv_fkp(np.real(x), nbar, float(v_cell), float(shot_fac), float(P_fkp)).cpu().numpy().astype(np.complex128)
    if np.iscomplexobj(x):
        out = out + 1j * ctx.apply_cinv_fkp(np.imag(x), nbar, float(v_cell), float(shot_fac), float(P_fkp)).cpu().numpy()
    return out


def applyC_alpha_single(input_map, nbar, MAS_mat, Y_lms, k_grids, r_grids, v_cell, k_filters, k_norm, i, a,
                        include_pix=True, data=False):
    """src/covariances_pk.py:404-466 (include_pix=False): one derivative map C_,alpha x for multipole
    index i and k-bin a."""
    _need_no_pix(include_pix)
    lmax = 2 * (len(Y_lms) - 1)
    ctx = _context(k_grids, r_grids, k_filters, lmax)
    return ctx.apply_c_alpha_single(np.real(input_map), nbar, float(v_cell), int(i), int(a), data).cpu().numpy().astype(np.complex128)


def applyC_alpha(input_map, nbar, MAS_mat, Y_lms, k_grids, r_grids, v_cell, k_filters, k_norm, n_k, lmax,
                 include_pix=True, data=False):
    """src/covariances_pk.py:329-401 (include_pix=False): list of the n_b derivative maps."""
    _need_no_pix(include_pix)
    ctx = _context(k_grids, r_grids, k_filters, lmax)
    x, n = ctx.dev(np.real(input_map)), ctx.dev(nbar)
    out = []
    for l_i in range(lmax // 2 + 1):
        for a in range(n_k):
            out.append(ctx.apply_c_alpha_single(x, n, float(v_cell), l_i, a, data).cpu().numpy().astype(np.complex128))
    return out


def applyCinv(*a, **k):
    raise Exception("ML weights (preconditioned-CG applyCinv) are not implemented on the B200 path yet "
                    "(SURVEY.md section 8f row N1); use weights = FKP")


applyC = applyS = applyCinv_approx = apply_inv_preconditoner = applyCinv
