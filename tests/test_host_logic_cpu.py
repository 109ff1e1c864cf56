"""Host-side logic on CPU: .param parsing (also the reference's own paramfiles when present), file
naming, text writers, triangle lists and the Y_lm table against sympy."""
import glob
import os

import numpy as np
import pytest

from spectra_without_windows_b200 import drivers, engine, synthetic
from spectra_without_windows_b200._ylm_table import YLM_TABLE

REF = "/root/reference/paramfiles"


def test_params_roundtrip(tmp_path, world):
    w = world("toyA")
    p = synthetic.write_world_files(w, str(tmp_path), include_pix=True, rand_nbar=True)
    P = drivers.Params(p, "pk")
    assert (P.k_min, P.k_max, P.dk, P.lmax) == w.pk_bins
    assert tuple(P.grid_3d) == w.grid and tuple(P.boxsize_grid) == w.box
    assert P.include_pix and P.rand_nbar and P.use_qbar and P.N_mc == w.N_mc and P.wtype == 0
    assert P.ktag() == "k0.000_0.120_0.020_l4"
    B = drivers.Params(p, "bk")
    assert (B.k_min, B.k_max, B.dk, B.lmax) == w.bk_bins
    with pytest.raises(Exception, match="Parameter file does not exist!"):
        drivers.Params(str(tmp_path) + "/missing.param", "pk")


def test_params_ml_weights(tmp_path, world):
    """weights = ML: the fiducial_pk table is located like pk/compute_pk_randoms.py:69-71 and covers every |k| of the mesh;
    the bispectrum scripts refuse ML loudly; anything but FKP / ML raises the reference's message."""
    w = world("toyA")
    p = synthetic.write_world_files(w, str(tmp_path), weights="ML")
    P = drivers.Params(p, "pk")
    assert P.wtype == 1 and P.weight_type == "ML" and os.path.exists(P.pk_input_file)
    P.unsupported()                                     # P_l(k): supported
    tab = np.loadtxt(P.pk_input_file)
    assert tab.shape[1] == 4 and tab[0, 0] == 0.0
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    k_top = np.sqrt(sum(np.max(np.abs(a)) ** 2 for a in mesh.kax))
    assert tab[-1, 0] > k_top                           # interp1d would raise otherwise
    with pytest.raises(Exception, match="ML"):
        drivers.Params(p, "bk").unsupported()
    bad = str(tmp_path) + "/bad.param"
    open(bad, "w").write(open(p).read().replace("weights = ML", "weights = XYZ"))
    with pytest.raises(Exception, match="weights must be 'FKP' or 'ML'"):
        drivers.Params(bad, "pk")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference paramfiles only exist in the build container")
def test_reference_paramfiles_parse_unchanged():
    files = sorted(glob.glob(REF + "/*.param"))
    assert len(files) == 8
    for f in files:
        for spec in ("pk", "bk"):
            P = drivers.Params(f, spec)
            assert P.weight_type == "FKP" and P.N_mc == 100 and not P.include_pix and not P.rand_nbar
            assert P.lmax == 4 and len(P.grid_3d) == 3 and P.k_max > P.k_min
            mesh = engine.Mesh(P.boxsize_grid, P.grid_3d, (0.0, 0.0, 0.0), P.k_min, P.k_max, P.dk, P.lmax)
            assert mesh.n_k == int((P.k_max - P.k_min) / P.dk)
            assert mesh.k_hi[-1] < (np.pi * P.grid_3d / P.boxsize_grid).min()     # every shipped binning is below Nyquist


def test_bk_triangle_count_matches_survey():
    # k in [0, 0.11], dk = 0.01 -> 191 triangles (C3/C4)
    mesh = engine.Mesh((1000.0,) * 3, (64,) * 3, (0.0, 0.0, 0.0), 0.0, 0.11, 0.01, 4)
    assert mesh.n_k == 11 and len(mesh.triangles()) == 191


def test_text_writers_layout(tmp_path, world):
    w = world("toyC")
    p = synthetic.write_world_files(w, str(tmp_path))
    P = drivers.Params(p, "pk")
    n_k = int((P.k_max - P.k_min) / P.dk)
    vals = np.arange(n_k * (P.lmax // 2 + 1)) * 1.5 - 2.0
    f = str(tmp_path) + "/pk.txt"
    drivers.write_pk_text(f, P, 3, vals, n_k)
    lines = open(f).read().split("\n")
    assert lines[0] == "####### Power Spectrum of toyC_ Simulation 3 #############"
    assert lines[-n_k].startswith("%.4f\t%.8e" % (P.k_min + 0.5 * P.dk, vals[0]))
    assert np.allclose(np.loadtxt(f)[:, 1:].T.ravel(), vals)
    B = drivers.Params(p, "bk")
    mesh = engine.Mesh(B.boxsize_grid, B.grid_3d, (0, 0, 0), B.k_min, B.k_max, B.dk, B.lmax)
    tris = mesh.triangles()
    vb = np.linspace(-1, 1, len(tris) * (B.lmax // 2 + 1))
    fb = str(tmp_path) + "/bk.txt"
    drivers.write_bk_text(fb, B, -1, vb, tris)
    txt = open(fb).read()
    assert txt.startswith("####### Bispectrum of toyC_ #############") and "# Format: k1 | k2 | k3 | B0(k1,k2,k3) | B2(k1,k2,k3)" in txt
    assert np.allclose(np.loadtxt(fb)[:, 3:].T.ravel(), vb)


def test_ylm_table_matches_sympy():
    from oracle import sww_oracle as orc
    Y = orc.compute_spherical_harmonic_functions(4)
    v = np.random.default_rng(0).normal(size=(3, 40))
    v /= np.linalg.norm(v, axis=0)
    for row in Y:
        for f in row:
            mono, _ = YLM_TABLE[(f.l, f.m)]
            mine = sum(c * v[0] ** a * v[1] ** b * v[2] ** d for c, a, b, d in mono)
            assert np.max(np.abs(mine - f(*v))) < 1e-14


def test_oracle_painters_conserve_mass_and_agree_on_cell_centres():
    """TSC and CIC restatements (src/opt_utilities.py:219-267 call site): mass conservation, periodic wrap, and a particle
    on a cell centre deposits (1/8, 3/4, 1/8) per axis with TSC and everything into one cell with CIC."""
    from oracle import sww_oracle as orc
    grid, box, bc = (12, 10, 8), (120.0, 100.0, 80.0), (5.0, -3.0, 1.0)
    rng = np.random.default_rng(2)
    pos = np.asarray(bc) + rng.uniform(-1.5, 1.5, size=(500, 3)) * np.asarray(box)
    w = rng.uniform(0.5, 1.5, 500)
    for f in (orc.tsc_paint, orc.cic_paint):
        m = f(pos, w, box, grid, bc)
        assert m.shape == grid and abs(m.sum() - w.sum()) < 1e-10
    one = np.asarray(bc)[None, :] + np.array([[30.0, 20.0, 10.0]])
    t = orc.tsc_paint(one, np.ones(1), box, grid, bc)
    assert abs(t[3, 2, 1] - 0.75 ** 3) < 1e-15 and abs(t[2, 2, 1] - 0.125 * 0.75 ** 2) < 1e-15
    c = orc.cic_paint(one, np.ones(1), box, grid, bc)
    assert c[3, 2, 1] == 1.0 and c.sum() == 1.0
