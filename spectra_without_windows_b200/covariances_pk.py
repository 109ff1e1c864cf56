"""Drop-in mirror of the reference's src/covariances_pk.py FKP operators, backed by libsww.so.

Same names, argument order and meaning as the reference; ndarray in -> ndarray out.  FKP and ML weights, both values of
`include_pix` (`include_pix=True` is composed from the same kernels plus the k-space MAS filter).  The ML operators
(applyS, applyC, applyCinv) keep full complex128 maps like the reference -- see csrc/ml.cu.  Never a CPU fallback.
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


def _ml_context(input_map, nbar, pk_map, Y_lms, k_grids, r_grids, v_cell, shot_fac, P_fkp, include_pix):
    """Context with the "pk_ml" workspace for the caller's mesh; n(r), shot factor and the fiducial P_l(k) maps are
    (re)bound on every call because the reference passes them as arguments."""
    info = _info(k_grids, r_grids)
    lmax = 2 * (len(Y_lms) - 1)
    # the binning of the context is irrelevant for these operators: any valid one below the Nyquist frequency
    kny = float(np.min(np.pi * np.asarray(info["grid"]) / np.asarray(info["box"], float)))
    mesh = engine.Mesh(info["box"], info["grid"], info["box_center"], 0.0, 0.5 * kny, 0.25 * kny, lmax)
    ctx = engine.get_context(mesh, 0, ("ops", "pk_ml"))
    ctx.set_fields(nbar, nbar, np.where(np.asarray(nbar) > 0, nbar, 1.0), float(shot_fac), float(P_fkp))
    ctx.set_include_pix(include_pix)
    if pk_map is not None:
        pk = np.ascontiguousarray(np.real(pk_map), dtype=np.float64)
        if pk.shape != (mesh.n_l,) + mesh.shape:
            raise Exception("pk_map must hold lmax//2+1 multipoles on the mesh")
        ctx._pk_fid = ctx.dev(pk)
        engine._lib.check(ctx.lib.sww_set_fiducial_pk(ctx._h, ctx._pk_fid.data_ptr()))
    if abs(ctx.mesh.v_cell - float(v_cell)) > 1e-12 * abs(float(v_cell)):
        raise Exception("v_cell does not match the mesh of k_grids/r_grids")
    return ctx


def applyS(input_map, nbar, MAS_mat, pk_map, Y_lms, k_grids, r_grids, v_cell, include_pix=True):
    """src/covariances_pk.py:45-95: signal covariance S x with the fiducial multipoles `pk_map`; complex128 out."""
    ctx = _ml_context(input_map, nbar, pk_map, Y_lms, k_grids, r_grids, v_cell, 1.0, 1e4, include_pix)
    return ctx.ml_apply("S", input_map).cpu().numpy()


def applyC(input_map, nbar, MAS_mat, pk_map, Y_lms, k_grids, r_grids, v_cell, shot_fac, include_pix=True):
    """src/covariances_pk.py:10-43: C x = S x + N x; complex128 out."""
    ctx = _ml_context(input_map, nbar, pk_map, Y_lms, k_grids, r_grids, v_cell, shot_fac, 1e4, include_pix)
    return ctx.ml_apply("C", input_map).cpu().numpy()


def applyCinv_approx(input_map, nbar, MAS_mat, v_cell, shot_fac, P_fkp=1e4, include_pix=True):
    """src/covariances_pk.py:166-192: the FKP form, used as first guess of the conjugate-gradient inversion."""
    return applyCinv_fkp(input_map, nbar, MAS_mat, v_cell, shot_fac, P_fkp=P_fkp, include_pix=include_pix)


def apply_inv_preconditoner(pix, nbar, MAS_mat, v_cell, shot_fac, P_fkp=1e4, include_pix=True):
    """src/covariances_pk.py:194-220 (sic): the preconditioner of the conjugate-gradient inversion."""
    return applyCinv_approx(pix, nbar, MAS_mat, v_cell, shot_fac, P_fkp=P_fkp, include_pix=include_pix)


def applyCinv(pix, nbar, MAS_mat, P3D, Y_lms, k_grids, r_grids, v_cell, shot_fac, applyC=applyC, applyCinv_approx=applyCinv_approx,
              max_it=100, abs_tol=1e-8, rel_tol=None, verb=1, apply_inv_preconditoner=apply_inv_preconditoner, P_fkp=1e4,
              include_pix=True):
    """src/covariances_pk.py:222-327: x = C^-1 pix by preconditioned conjugate gradients, on the device.  Same
    stopping rules and messages; the operator hooks must be this module's own (a custom applyC cannot run on the GPU)."""
    import time
    g = globals()
    if applyC is not g["applyC"] or applyCinv_approx is not g["applyCinv_approx"] or apply_inv_preconditoner is not g["apply_inv_preconditoner"]:
        raise Exception("custom applyC / applyCinv_approx / apply_inv_preconditoner hooks are not supported on the B200 path")
    start = time.time()
    ctx = _ml_context(pix, nbar, P3D, Y_lms, k_grids, r_grids, v_cell, shot_fac, P_fkp, include_pix)
    assert max_it < np.asarray(pix).size
    x, it, why = ctx.ml_apply_cinv(pix, max_it=max_it, abs_tol=abs_tol, rel_tol=rel_tol)
    if why == 2 and verb:
        print("Inversion stalled after step %d; exiting" % (it + 1))
    if why == 1 and verb:
        print("Inversion stopped early after %d iterations" % it)
    if why == 0 and it == max_it:
        print("CGD did not stop early: is this converged?")
    if verb:
        print("\nInversion took %d seconds" % (time.time() - start))
    return x.cpu().numpy()
