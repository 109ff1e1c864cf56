"""Drop-in mirror of the reference's src/covariances_pk.py FKP operators, backed by libsww.so.

Same names, argument order and meaning as the reference; ndarray in -> ndarray out.  FKP weights, both values of `include_pix`;
`include_pix=True` is composed from the same kernels plus the k-space MAS filter; ML weights raise
(SURVEY.md section 8f row N1 -- not yet built, never a CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import engine


def _mesh_context(input_map, MAS_mat):
    """Context whose MAS tables match the caller's mesh (the MAS window only depends on the grid size)."""
    from .opt_utilities import _fft_ctx
    return _fft_ctx(np.asarray(input_map).shape)


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
    ctx = _any_context(input_map)
    x = np.real(input_map)
    if include_pix:   # shot*ift(1/MAS*ft(nbar*ift(ft(x)/MAS)))/v_cell   (:122)
        y = ctx.apply_n(ctx.mas_filter(x, -1), nbar, float(v_cell), float(shot_fac))
        return ctx.mas_filter(y, -1).cpu().numpy()
    return ctx.apply_n(x, nbar, float(v_cell), float(shot_fac)).cpu().numpy()


def applyCinv_fkp(input_map, nbar, MAS_mat, v_cell, shot_fac, P_fkp=1e4, include_pix=True):
    """src/covariances_pk.py:126-164 (include_pix=False branch).  Returns a complex128 map like the
    FP64-promoted reference (the reference's buffer is complex64, :154)."""
    ctx = _any_context(input_map)
    x = np.asarray(input_map)

    def one(part):
        if include_pix:   # ift(MAS*ft(ratio))*v_cell with ratio from ift(ft(x)*MAS)   (:155-162)
            t = ctx.apply_cinv_fkp(ctx.mas_filter(part, +1), nbar, float(v_cell), float(shot_fac), float(P_fkp))
            return ctx.mas_filter(t, +1).cpu().numpy()
        return ctx.apply_cinv_fkp(part, nbar, float(v_cell), float(shot_fac), float(P_fkp)).cpu().numpy()

    out = one(np.real(x)).astype(np.complex128)
    if np.iscomplexobj(x):
        out = out + 1j * one(np.imag(x))
    return out


def applyC_alpha_single(input_map, nbar, MAS_mat, Y_lms, k_grids, r_grids, v_cell, k_filters, k_norm, i, a,
                        include_pix=True, data=False):
    """src/covariances_pk.py:404-466 (include_pix=False): one derivative map C_,alpha x for multipole
    index i and k-bin a."""
    lmax = 2 * (len(Y_lms) - 1)
    ctx = _context(k_grids, r_grids, k_filters, lmax)
    if include_pix:   # ift(ft(nbar*ift(tmp2))/MAS)/v_cell with tmp_map = ift(ft(x)/MAS)*nbar, no MAS^2   (:444-464)
        y = ctx.apply_c_alpha_single(ctx.mas_filter(np.real(input_map), -1), nbar, float(v_cell), int(i), int(a), False)
        return ctx.mas_filter(y, -1).cpu().numpy().astype(np.complex128)
    return ctx.apply_c_alpha_single(np.real(input_map), nbar, float(v_cell), int(i), int(a), data).cpu().numpy().astype(np.complex128)


def applyC_alpha(input_map, nbar, MAS_mat, Y_lms, k_grids, r_grids, v_cell, k_filters, k_norm, n_k, lmax,
                 include_pix=True, data=False):
    """src/covariances_pk.py:329-401 (include_pix=False): list of the n_b derivative maps."""
    ctx = _context(k_grids, r_grids, k_filters, lmax)
    x, n = ctx.dev(np.real(input_map)), ctx.dev(nbar)
    if include_pix:
        x = ctx.mas_filter(x, -1)
    out = []
    for l_i in range(lmax // 2 + 1):
        for a in range(n_k):
            y = ctx.apply_c_alpha_single(x, n, float(v_cell), l_i, a, data and not include_pix)
            if include_pix:
                y = ctx.mas_filter(y, -1)
            out.append(y.cpu().numpy().astype(np.complex128))
    return out


def applyCinv(*a, **k):
    raise Exception("ML weights (preconditioned-CG applyCinv) are not implemented on the B200 path yet "
                    "(SURVEY.md section 8f row N1); use weights = FKP")


applyC = applyS = applyCinv_approx = apply_inv_preconditoner = applyCinv
