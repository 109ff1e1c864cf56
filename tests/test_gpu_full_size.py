"""BASELINE-size checks (C2: 512^3, n_b = 246; C4: 256^3, n_b = 573) through properties that do not need the CPU oracle:
the fused sm_100a chains against the cuFFT-backed pipeline, an independent torch.fft (FP64) evaluation of the l = 0
block of the Fisher matrix and of q-bar, and the quadratic scaling in the field."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALPHA, SHOT, P_FKP = 0.02, 1.02, 1e4


def survey_nbar_r(rax, t):
    X, Y, Z = t.meshgrid(*rax, indexing="ij")
    r = t.sqrt(X * X + Y * Y + Z * Z)
    lon = t.arctan2(Y, X)
    lat = t.arcsin(Z / (r + 1e-300))
    W = (abs(lon) < np.deg2rad(70.0)) * (abs(lat) < np.deg2rad(30.0)) * (0.9 + 0.1 * t.cos(3.0 * lon))
    return 4e-4 * t.exp(-0.5 * ((r - 1400.0) / 250.0) ** 2) * (r > 1150.0) * (r < 1750.0) * W / ALPHA


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def ylm_torch(t, l, m, x, y, z):
    """Y_lm from the sympy-generated monomial table (the text of the reference's lambdified expressions), evaluated with
    torch on broadcastable unit-vector components -- independent of the CUDA evaluation (generated C, Horner-in-z rows)."""
    from spectra_without_windows_b200._ylm_table import YLM_TABLE
    out = 0.0
    for c, a, b, d in YLM_TABLE[(l, m)][0]:
        out = out + c * x ** a * y ** b * z ** d
    return out


def torch_multipole_spectrum(t, mesh, f, l, dev):
    """4 pi/(2l+1) sum_m Y_lm(k^) FT[f Y_lm(r^)] with torch.fft (src/covariances_pk.py:376-384,391), half spectrum."""
    rx, ry, rz = (t.from_numpy(a).to(dev) for a in mesh.rax)
    kx, ky = (t.from_numpy(a).to(dev) for a in mesh.kax[:2])
    kz = t.from_numpy(mesh.kax[2][:mesh.nzh].copy()).to(dev)
    kz = kz.abs()                                           # rfft keeps +k_z (the Nyquist entry of kax is negative)
    X, Y, Z = rx[:, None, None], ry[None, :, None], rz[None, None, :]
    ir = 1.0 / (1e-12 + t.sqrt(X * X + Y * Y + Z * Z))
    KX, KY, KZ = kx[:, None, None], ky[None, :, None], kz[None, None, :]
    ik = 1.0 / (1e-12 + t.sqrt(KX * KX + KY * KY + KZ * KZ))
    out = None
    for m in range(-l, l + 1):
        yr = ylm_torch(t, l, m, X * ir, Y * ir, Z * ir)
        yk = ylm_torch(t, l, m, KX * ik, KY * ik, KZ * ik)
        term = t.fft.rfftn(f * yr) * yk
        out = term if out is None else out + term
        del term, yr, yk
    return out * (4.0 * np.pi / (2.0 * l + 1.0))


def test_pk_fisher_c2_full_size():
    import torch as t
    from spectra_without_windows_b200 import engine
    mesh = engine.Mesh((1800.0, 3450.0, 1900.0), (512, 512, 512), (1200.0, 0.0, 0.0), 0.0, 0.41, 0.005, 4)
    ctx = engine.Context(mesh, 0, ("pk_fisher",))
    assert ctx.fft_backend == 1 and ctx.n_b == 246
    dev = ctx.device
    nbar_r = survey_nbar_r([t.from_numpy(a).to(dev) for a in mesh.rax], t).contiguous()
    n = (nbar_r * ALPHA).contiguous()
    nA = nbar_r.clone()
    nA[nA == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(n, n, nA, SHOT, P_FKP)
    gen = t.Generator(device=dev)
    gen.manual_seed(2024)
    a = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(nA)
    F1, q1 = (x.clone() for x in ctx.pk_fisher(a))
    assert bool(t.isfinite(F1).all()) and t.equal(F1, F1.T)
    # (1) quadratic in the field
    F2, q2 = ctx.pk_fisher(2.0 * a)
    assert rel(F2, 4.0 * F1) < 1e-13 and rel(q2, 4.0 * q1) < 1e-13
    # (2) independent torch.fft evaluation of the l = l' = 0 rows of three k bins and of q-bar there
    N, v = float(mesh.N), mesh.v_cell
    binmap = ctx.binmap().long()
    wk = t.full((mesh.nzh,), 2.0, dtype=t.float64, device=dev)
    wk[0] = 1.0
    wk[-1] = 1.0
    den = n * (n * P_FKP + SHOT)
    cinv_a = t.where(n > 0, a * v / den, t.zeros_like(a))
    W = t.where(n > 0, n / den, t.zeros_like(a))
    FC0 = t.fft.rfftn(n * cinv_a)
    G0 = t.fft.rfftn(n * a / nA)
    NAa = t.fft.rfftn(W * (SHOT * n * (a / nA) / v))
    bins = [3, 40, 81]
    rows = {}
    for ka in bins:
        u = t.fft.irfftn(t.where(binmap == ka, FC0, t.zeros_like(FC0)), s=mesh.shape)
        T = t.fft.rfftn(n * W * u)
        prod = (T.real * G0.real + T.imag * G0.imag) * wk
        rows[ka] = {kb: 0.5 / (N * v) * float(prod[binmap == kb].sum()) for kb in bins}
        qb = (FC0.real * NAa.real + FC0.imag * NAa.imag) * wk
        assert abs(0.5 / N * float(qb[binmap == ka].sum()) - float(q1[ka])) < 1e-9 * float(q1[:82].abs().max())
    scale = float(F1[:82, :82].abs().max())
    for ka in bins:
        for kb in bins:
            ref = 0.5 * (rows[ka][kb] + rows[kb][ka])
            assert abs(ref - float(F1[ka, kb])) < 1e-9 * scale, (ka, kb)
    del FC0, G0, NAa, u, T, prod, qb
    # (2b) the same for l > 0: element [(l=2, ka), (l'=4, kb)] needs the rows of both multipoles (the matrix is symmetrised),
    #      with every Y_lm(r^), Y_lm(k^) evaluated in torch from the sympy table
    FC = {l: torch_multipole_spectrum(t, mesh, n * cinv_a, l, dev) for l in (2, 4)}
    G = {l: torch_multipole_spectrum(t, mesh, n * a / nA, l, dev) for l in (2, 4)}
    n_k = mesh.n_k

    def row_vs(l, ka, lp, kb):
        u = t.fft.irfftn(t.where(binmap == ka, FC[l], t.zeros_like(FC[l])), s=mesh.shape)
        T = t.fft.rfftn(n * W * u)
        prod = (T.real * G[lp].real + T.imag * G[lp].imag) * wk
        return 0.5 / (N * v) * float(prod[binmap == kb].sum())

    scale_all = float(F1.abs().max())
    for (l, ka, lp, kb) in ((2, 12, 4, 30), (4, 55, 2, 55), (2, 70, 2, 9)):
        ref = 0.5 * (row_vs(l, ka, lp, kb) + row_vs(lp, kb, l, ka))
        got = float(F1[(l // 2) * n_k + ka, (lp // 2) * n_k + kb])
        assert abs(ref - got) < 1e-9 * scale_all, (l, ka, lp, kb, ref, got)
    qb4 = (FC[4].real * t.fft.rfftn(W * (SHOT * n * (a / nA) / v)).real + FC[4].imag * t.fft.rfftn(W * (SHOT * n * (a / nA) / v)).imag) * wk
    assert abs(0.5 / N * float(qb4[binmap == 33].sum()) - float(q1[2 * n_k + 33])) < 1e-9 * float(q1.abs().max())
    del FC, G, qb4
    # (3) the cuFFT-backed pipeline (separate kernels, no fusion, no pruning of the transforms) agrees on every element
    ctx.set_fft_backend(0)
    F0, q0 = ctx.pk_fisher(a)
    assert rel(F1, F0) < 1e-10 and rel(q1, q0) < 1e-10
    ctx.close()


def test_bk_fisher_c4_full_size():
    import torch as t
    from spectra_without_windows_b200 import engine
    mesh = engine.Mesh((1750.0, 3220.0, 1830.0), (256, 256, 256), (1200.0, 0.0, 0.0), 0.0, 0.11, 0.01, 4)
    ctx = engine.Context(mesh, 0, ("bk_fisher",))
    assert ctx.fft_backend == 1 and ctx.n_b_bk == 573
    dev = ctx.device
    nbar_r = survey_nbar_r([t.from_numpy(a).to(dev) for a in mesh.rax], t).contiguous()
    n = (nbar_r * ALPHA).contiguous()
    nA = nbar_r.clone()
    nA[nA == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(n, n, nA, SHOT, P_FKP)
    gen = t.Generator(device=dev)
    gen.manual_seed(7)
    aA = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(nA)
    aB = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(nA)
    F1, maps = ctx.bk_fisher_pair(aA, aB)
    F1 = F1.clone()
    g1 = maps[0][1, 5].clone()
    assert bool(t.isfinite(F1).all()) and rel(F1, F1.T) < 1e-14
    # F is a quartic form in the fields (two filtered maps per phi, two phi per element)
    F2, maps = ctx.bk_fisher_pair(2.0 * aA, 2.0 * aB, maps)
    assert rel(F2, 16.0 * F1) < 1e-12
    ctx.set_fft_backend(0)
    F0, maps = ctx.bk_fisher_pair(aA, aB, maps)
    assert rel(F1, F0) < 1e-10
    assert rel(g1, maps[0][1, 5]) < 1e-11
    ctx.close()


def test_pk_fisher_c5_full_size():
    """BASELINE config 5 (1024^3, k <= 0.4, n_b = 240): the rows run on the 512 x 1024 x 512 row mesh (csrc/rowmesh.cu).
    Checked against (1) the quadratic scaling in the field, (2) the same realisation with the rows on the FULL 1024^3 mesh
    (SWW_ROW_MESH=0: generic nz = 1024 passes, no band-limited weight map -- a different set of kernels and a different
    formulation), and (3) an independent torch.fft evaluation of l = 0 elements and q-bar at full size."""
    import gc
    import os
    import torch as t
    from spectra_without_windows_b200 import engine
    os.environ["SWW_TIGHT_WS"] = "1"
    try:
        mesh = engine.Mesh((1800.0, 3450.0, 1900.0), (1024, 1024, 1024), (1200.0, 0.0, 0.0), 0.0, 0.40, 0.005, 4)
        ctx = engine.Context(mesh, 0, ("pk_fisher",))
        assert ctx.fft_backend == 1 and ctx.n_b == 240 and ctx.row_mesh_shape() == (512, 1024, 512)
        dev = ctx.device
        nbar_r = survey_nbar_r([t.from_numpy(a).to(dev) for a in mesh.rax], t).contiguous()
        n = (nbar_r * ALPHA).contiguous()
        nA = nbar_r
        nA[nA == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
        ctx.set_fields(n, n, nA, SHOT, P_FKP)
        gen = t.Generator(device=dev)
        gen.manual_seed(99)
        a = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64).mul_(t.sqrt(nA))
        F1, q1 = (x.clone() for x in ctx.pk_fisher(a))
        assert bool(t.isfinite(F1).all()) and t.equal(F1, F1.T)
        a.mul_(2.0)
        F2, q2 = ctx.pk_fisher(a)
        assert rel(F2, 4.0 * F1) < 1e-13 and rel(q2, 4.0 * q1) < 1e-13
        a.mul_(0.5)
        binmap = ctx.binmap()
        ctx.close()
        del ctx
        gc.collect()
        t.cuda.empty_cache()
        # (2) rows on the full mesh
        os.environ["SWW_ROW_MESH"] = "0"
        try:
            cf = engine.Context(mesh, 0, ("pk_fisher",))
        finally:
            del os.environ["SWW_ROW_MESH"]
        assert cf.row_mesh_shape() is None
        cf.set_fields(n, n, nA, SHOT, P_FKP)
        F0, q0 = (x.clone() for x in cf.pk_fisher(a))
        cf.close()
        del cf
        gc.collect()
        t.cuda.empty_cache()
        assert rel(F1, F0) < 1e-10 and rel(q1, q0) < 1e-10
        # (3) torch.fft, l = l' = 0, two bins
        N, v = float(mesh.N), mesh.v_cell
        wk = t.full((mesh.nzh,), 2.0, dtype=t.float64, device=dev)
        wk[0] = 1.0
        wk[-1] = 1.0
        den = n * (n * P_FKP + SHOT)
        W = t.where(n > 0, n / den, t.zeros_like(n))
        FC0 = t.fft.rfftn(n * t.where(n > 0, a * v / den, t.zeros_like(a)))
        G0 = t.fft.rfftn(n * a / nA)
        NAa = t.fft.rfftn(W * (SHOT * n * (a / nA) / v))
        del den
        bins = [5, 61]
        rows = {}
        for ka in bins:
            sel = binmap == ka
            qb = ((FC0.real * NAa.real + FC0.imag * NAa.imag) * wk)[sel].sum()
            assert abs(0.5 / N * float(qb) - float(q1[ka])) < 1e-9 * float(q1[:80].abs().max())
            u = t.fft.irfftn(t.where(sel, FC0, t.zeros_like(FC0)), s=mesh.shape)
            u.mul_(n).mul_(W)
            T = t.fft.rfftn(u)
            del u
            prod = (T.real * G0.real + T.imag * G0.imag) * wk
            del T
            rows[ka] = {kb: 0.5 / (N * v) * float(prod[binmap == kb].sum()) for kb in bins}
            del prod
        scale = float(F1[:80, :80].abs().max())
        for ka in bins:
            for kb in bins:
                assert abs(0.5 * (rows[ka][kb] + rows[kb][ka]) - float(F1[ka, kb])) < 1e-9 * scale, (ka, kb)
    finally:
        os.environ.pop("SWW_TIGHT_WS", None)


def test_bk_data_c3_full_size():
    """BASELINE config 3 (256^3, k in [0, 0.11], l = 0, 2, 191 triangles): g-maps of a catalogue-like field, the 3-field sums
    and the 1-field term against torch evaluations on the same maps (three triangles, all multipoles), the independent
    torch.fft construction of one g-map, and the cubic / bilinear scalings."""
    import torch as t
    from spectra_without_windows_b200 import engine
    mesh = engine.Mesh((1000.0, 1000.0, 1000.0), (256, 256, 256), (0.0, 0.0, 2000.0), 0.0, 0.11, 0.01, 2)
    ctx = engine.Context(mesh, 0, ("bk_fisher", "bk_data"))
    assert ctx.n_tri == 191 and ctx.n_b_bk == 382
    dev = ctx.device
    n = t.full(mesh.shape, 3e-4, dtype=t.float64, device=dev)
    ctx.set_fields(n, n, n / ALPHA, SHOT, P_FKP)
    gen = t.Generator(device=dev)
    gen.manual_seed(5)
    d = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * 1e-2
    g = ctx.bk_gmaps(d, data=True)
    q3 = ctx.bk_q3(g).clone()
    tris = mesh.triangles()
    for it in (0, 77, 190):
        a_, b_, c_ = tris[it]
        for l in range(mesh.n_l):
            ref = float((g[0, a_] * g[0, b_] * g[l, c_]).sum())
            assert abs(ref - float(q3[it, l])) < 1e-11 * float(q3.abs().max()), (it, l)
    assert rel(ctx.bk_q3(2.0 * g), 8.0 * q3) < 1e-13
    # one g-map from scratch with torch.fft: g_i^l = ift(Theta_i 4pi/(2l+1) sum_m ft(n C^-1 d Y_lm(r^)) Y_lm(k^) MAS), data variant
    v = mesh.v_cell
    cinv = d * v / (n * (n * P_FKP + SHOT))
    F2 = torch_multipole_spectrum(t, mesh, n * cinv, 2, dev)
    mas = (t.from_numpy(mesh.mas[0]).to(dev)[:, None, None] * t.from_numpy(mesh.mas[1]).to(dev)[None, :, None]
           * t.from_numpy(mesh.mas[2][:mesh.nzh].copy()).to(dev)[None, None, :])
    bm = ctx.binmap()
    g_ref = t.fft.irfftn(t.where(bm == 6, F2 * mas, t.zeros_like(F2)), s=mesh.shape)
    assert rel(g[1, 6], g_ref) < 1e-9
    # 1-field term with two Monte-Carlo map sets: bilinear in (g, g~) and linear in the data maps
    aA = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(n / ALPHA)
    gm, tgm = ctx.bk_gmaps(aA)
    q1 = ctx.zeros((ctx.n_tri, mesh.n_l))
    ctx.bk_q1_accumulate(g, gm, tgm, 2, q1)
    it = 101
    a_, b_, c_ = tris[it]
    for l in range(mesh.n_l):
        ref = (g[0, a_] * (gm[0, b_] * tgm[l, c_] + tgm[0, b_] * gm[l, c_])).sum() \
            + (g[0, b_] * (gm[l, c_] * tgm[0, a_] + tgm[l, c_] * gm[0, a_])).sum() \
            + (g[l, c_] * (gm[0, a_] * tgm[0, b_] + tgm[0, a_] * gm[0, b_])).sum()
        assert abs(-0.25 * float(ref) - float(q1[it, l])) < 1e-10 * float(q1.abs().max()), l
    q1b = ctx.zeros((ctx.n_tri, mesh.n_l))
    ctx.bk_q1_accumulate(3.0 * g, 2.0 * gm, tgm, 2, q1b)
    assert rel(q1b, 6.0 * q1) < 1e-12
    ctx.close()


def test_bk_fisher_nseries_mesh():
    """The mesh of the reference's published Nseries bispectrum outputs, 167 x 307 x 175 (two prime axes: the any-length
    passes of csrc/fft_any_passes.cuh): symmetric, finite, quartic in the fields, one g-map against an independent torch.fft
    construction, and the whole Fisher matrix against the cuFFT backend."""
    import torch as t
    from spectra_without_windows_b200 import engine
    mesh = engine.Mesh((1750.0, 3220.0, 1830.0), (167, 307, 175), (1200.0, 0.0, 0.0), 0.0, 0.11, 0.01, 4)
    ctx = engine.Context(mesh, 0, ("bk_fisher",))
    assert ctx.fft_backend == 1 and ctx.n_b_bk == 573
    dev = ctx.device
    nbar_r = survey_nbar_r([t.from_numpy(a).to(dev) for a in mesh.rax], t).contiguous()
    n = (nbar_r * ALPHA).contiguous()
    nA = nbar_r.clone()
    nA[nA == 0] = float(nbar_r[nbar_r != 0].min()) * 0.01
    ctx.set_fields(n, n, nA, SHOT, P_FKP)
    gen = t.Generator(device=dev)
    gen.manual_seed(11)
    aA = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(nA)
    aB = t.randn(mesh.shape, generator=gen, device=dev, dtype=t.float64) * t.sqrt(nA)
    F1, maps = ctx.bk_fisher_pair(aA, aB)
    F1 = F1.clone()
    g_l4 = maps[0][2, 7].clone()
    assert bool(t.isfinite(F1).all()) and rel(F1, F1.T) < 1e-14
    # g_i^l = ift(Theta_i 4pi/(2l+1) sum_m ft(n C^-1 a Y_lm(r^)) Y_lm(k^)) (bk/compute_bk_randoms.py:222-267), l = 4, bin 7
    den = n * (n * P_FKP + SHOT)
    cinv = t.where(n > 0, aA * mesh.v_cell / den, t.zeros_like(aA))
    F4 = torch_multipole_spectrum(t, mesh, n * cinv, 4, dev)
    g_ref = t.fft.irfftn(t.where(ctx.binmap() == 7, F4, t.zeros_like(F4)), s=mesh.shape)
    assert rel(g_l4, g_ref) < 1e-9
    F2, maps = ctx.bk_fisher_pair(2.0 * aA, 2.0 * aB, maps)
    assert rel(F2, 16.0 * F1) < 1e-12
    ctx.set_fft_backend(0)
    F0, maps = ctx.bk_fisher_pair(aA, aB, maps)
    assert rel(F1, F0) < 1e-11
    ctx.close()
