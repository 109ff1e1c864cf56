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
