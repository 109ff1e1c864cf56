"""Catalogue ingestion and the n-bar(r) builder (spectra_without_windows_b200/ingest.py, generate_mask.py): the contracts
of src/opt_utilities.py:12-143 and generate_mask.py:77,96-146, checked on synthetic files (CPU only)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from spectra_without_windows_b200 import ingest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_fits_table(path, cols):
    """Minimal FITS file: empty primary HDU + one BINTABLE of float64 / int32 scalar columns (big-endian)."""
    def card(k, v):
        if isinstance(v, str):
            s = "%-8s= '%-8s'" % (k, v)
        elif isinstance(v, bool):
            s = "%-8s= %20s" % (k, "T" if v else "F")
        else:
            s = "%-8s= %20s" % (k, v)
        return s.ljust(80).encode("ascii")

    def header(cards):
        h = b"".join(cards) + b"END".ljust(80)
        return h + b" " * (-len(h) % 2880)

    n = len(next(iter(cols.values())))
    fields = [(k, ">f8" if v.dtype.kind == "f" else ">i4") for k, v in cols.items()]
    tab = np.zeros(n, dtype=np.dtype(fields))
    for k, v in cols.items():
        tab[k] = v
    body = tab.tobytes()
    body += b"\0" * (-len(body) % 2880)
    prim = header([card("SIMPLE", True), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", True)])
    cards = [card("XTENSION", "BINTABLE"), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", tab.dtype.itemsize),
             card("NAXIS2", n), card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", len(fields))]
    for i, (k, f) in enumerate(fields):
        cards += [card("TTYPE%d" % (i + 1), k), card("TFORM%d" % (i + 1), "D" if f == ">f8" else "J")]
    with open(path, "wb") as f:
        f.write(prim + header(cards) + body)


def test_cosmology_distance_and_round_trip():
    from scipy.integrate import quad
    c = ingest.Cosmology(h=0.676).match(Omega0_m=0.31)
    for z in (0.1, 0.43, 0.7, 1.5):
        ref = quad(lambda x: 1.0 / c.efunc(x), 0.0, z, epsabs=1e-13, epsrel=1e-13)[0] * ingest.C_KMS / 100.0
        assert abs(c.comoving_distance(z) - ref) < 1e-7 * ref
    rng = np.random.default_rng(3)
    ra, dec, z = rng.uniform(0, 360, 1000), rng.uniform(-80, 80, 1000), rng.uniform(0.2, 0.9, 1000)
    pos = ingest.sky_to_cartesian(ra, dec, z, c)
    ra2, dec2, z2 = ingest.cartesian_to_sky(pos, c)
    assert np.max(np.abs(((ra2 - ra + 180) % 360) - 180)) < 1e-9 and np.max(np.abs(dec2 - dec)) < 1e-9
    assert np.max(np.abs(z2 - z)) < 1e-7
    # BOSS CMASS sanity: chi(z = 0.57) ~ 1.48 Gpc/h for h = 0.676, Omega_m = 0.31
    assert 1450.0 < c.comoving_distance(0.57) < 1520.0


def test_fits_and_text_catalogues_follow_the_reference_column_logic(tmp_path):
    rng = np.random.default_rng(5)
    n = 500
    cols = {"RA": rng.uniform(100, 260, n), "DEC": rng.uniform(-5, 60, n), "Z": rng.uniform(0.3, 0.8, n),
            "WEIGHT_SYSTOT": rng.uniform(0.9, 1.1, n), "WEIGHT_NOZ": rng.uniform(1.0, 1.2, n),
            "WEIGHT_CP": rng.integers(1, 3, n).astype(np.int32), "WEIGHT_FKP": rng.uniform(0.2, 0.9, n)}
    fits = str(tmp_path / "gal.fits")
    write_fits_table(fits, cols)
    back = ingest.read_fits_table(fits)
    for k, v in cols.items():
        assert np.array_equal(back[k], v)
    cosmo = ingest.Cosmology(h=0.676).match(Omega0_m=0.31)
    cat = ingest.apply_reference_columns(back, 0.43, 0.70, cosmo, True, -1, fkp_weights=True)
    keep = (cols["Z"] > 0.43) & (cols["Z"] < 0.70)
    w = (cols["WEIGHT_SYSTOT"] * (cols["WEIGHT_NOZ"] + cols["WEIGHT_CP"] - 1.0))[keep]
    assert np.allclose(cat["WEIGHT"], w, rtol=0, atol=0)
    nbar = (1.0 / cols["WEIGHT_FKP"][keep] - 1.0) / 1e4
    assert np.allclose(cat["NBAR"], nbar, rtol=1e-15) and np.allclose(cat["WEIGHT_FKP"], 1.0 / (1.0 + 1e4 * nbar), rtol=1e-15)
    assert cat["Position"].shape == (keep.sum(), 3)
    # text catalogue with named columns (mock convention: VETO FLAG x FIBER COLLISION)
    txt = str(tmp_path / "mock_0007.dat")
    arr = np.stack([cols["RA"], cols["DEC"], cols["Z"], np.ones(n), rng.integers(0, 2, n).astype(float), cols["WEIGHT_FKP"]], 1)
    np.savetxt(txt, arr)
    names = ["RA", "DEC", "Z", "VETO FLAG", "FIBER COLLISION", "WEIGHT_FKP"]
    c2 = ingest.apply_reference_columns(ingest.read_catalog(txt, names), 0.43, 0.70, cosmo, True, 7, fkp_weights=False)
    assert np.array_equal(c2["WEIGHT"], arr[keep, 3] * arr[keep, 4]) and np.all(c2["WEIGHT_FKP"] == 1.0)


def test_mangle_polygons(tmp_path):
    # two caps: polygon 0 = cap of 30 deg around the north pole minus nothing; polygon 1 = band 30..60 deg from the pole
    ply = tmp_path / "mask.ply"
    cm30, cm60 = 1 - np.cos(np.deg2rad(30.0)), 1 - np.cos(np.deg2rad(60.0))
    ply.write_text("2 polygons\n"
                   "polygon 0 ( 1 caps, 0.9 weight, 0 pixel, 0.8 str):\n 0 0 1 %.17g\n"
                   "polygon 1 ( 2 caps, 0.5 weight, 0 pixel, 2.3 str):\n 0 0 1 %.17g\n 0 0 1 %.17g\n" % (cm30, cm60, -cm30))
    m = ingest.MangleMask(str(ply))
    dec = np.array([85.0, 61.0, 59.0, 31.0, 29.0, -10.0])
    ids = m.get_polyids(np.full(6, 123.0), dec)
    assert list(ids) == [0, 0, 1, 1, -1, -1]
    assert list(m.weights) == [0.9, 0.5]


def test_generate_mask_cli_contract(tmp_path):
    """File name (generate_mask.py:77), normalisation sum(nbar) V_cell = sum(w_r) (:140-142), early exit when present."""
    rng = np.random.default_rng(9)
    cosmo = ingest.Cosmology(h=0.676).match(Omega0_m=0.31)
    n = 20000
    ra, dec, z = rng.uniform(150, 210, n), np.rad2deg(np.arcsin(rng.uniform(-0.3, 0.3, n))), rng.uniform(0.40, 0.75, n)
    w = rng.uniform(0.5, 1.5, n)
    rfile = str(tmp_path / "randoms.dat")
    np.savetxt(rfile, np.stack([ra, dec, z, w, rng.uniform(0.2, 0.9, n)], 1))
    out = str(tmp_path / "out") + "/"
    par = tmp_path / "toy.param"
    par.write_text("[catalogs]\ndata_file = none\ndata_columns =\nrandoms_file = %s\nrandoms_columns = RA, DEC, Z, WEIGHT, WEIGHT_FKP\n"
                   "mask_file = none\n[pk-binning]\nk_min = 0.0\nk_max = 0.1\ndk = 0.02\nlmax = 2\ngrid = 24, 32, 20\n"
                   "[bk-binning]\nk_min = 0.0\nk_max = 0.1\ndk = 0.02\nlmax = 2\ngrid = 12, 16, 10\n"
                   "[sample]\ntype = toy_\nz_min = 0.43\nz_max = 0.70\nbox = 1200, 2400, 1300\n"
                   "[parameters]\nh_fid = 0.676\nOmegaM_fid = 0.31\n[directories]\ntemporary = %s\nmonte_carlo = %s\noutput = %s\n"
                   % (rfile, out, out, out))
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "generate_mask.py"), "pk", str(par)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    f = out + "nbar_toy_z0.430_0.700_g24_32_20.npy"
    assert os.path.exists(f)
    nbar = np.load(f)
    keep = (z > 0.43) & (z < 0.70)
    v_cell = 1200.0 * 2400.0 * 1300.0 / (24 * 32 * 20)
    assert nbar.shape == (24, 32, 20) and np.all(nbar >= 0)
    assert abs(nbar.sum() * v_cell - w[keep].sum()) < 1e-9 * w[keep].sum()
    # n(r) follows the radial selection: empty outside the redshift shell
    pos = ingest.sky_to_cartesian(ra[keep], dec[keep], z[keep], cosmo)
    bc = 0.5 * (pos.min(0) + pos.max(0))
    assert nbar.max() > 0 and (nbar == 0).any()
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "generate_mask.py"), "pk", str(par)], capture_output=True, text=True, env=env)
    assert r2.returncode == 0 and "Mask already completed; exiting!" in r2.stdout
    with pytest.raises(Exception):
        ingest.read_catalog(str(tmp_path / "missing.fits"))
