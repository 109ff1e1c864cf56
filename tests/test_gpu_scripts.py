"""Script-level parity on the GPU: the four drop-in CLIs produce the reference's files, and the
text outputs agree with the golden files written by the unmodified reference scripts."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu


# the goldens come from the FP64-promoted reference: the file-based bispectrum path is pinned with complex128 bias files
# (the default, like the reference's shipped scripts, is complex64 -- exercised by the hand-off test below)
ENV128 = dict(os.environ, SWW_BIAS_DTYPE="complex128")


def run(script, arg, param, env=ENV128):
    r = subprocess.run([sys.executable, os.path.join(ROOT, script), str(arg), param], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("name", ["toyA", "toyC"])
def test_four_scripts_end_to_end(name, world, golden, tmp_path):
    from spectra_without_windows_b200 import synthetic as syn
    w, g = world(name), golden(name)
    p = syn.write_world_files(w, str(tmp_path))
    mc, od = str(tmp_path) + "/mc/", str(tmp_path) + "/out/"
    for r in (1, 2):
        out = run("pk/compute_pk_randoms.py", r, p)
        F = np.load(glob.glob(mc + "*%d_FKP_pk_fish_a_*.npy" % r)[0])
        qb = np.load(glob.glob(mc + "*%d_FKP_pk_q-bar_a_*.npy" % r)[0])
        assert relerr(F, g["pk_fish_%d" % r]) < 1e-9 and relerr(qb, g["pk_qbar_%d" % r]) < 1e-9
    assert "already completed" in run("pk/compute_pk_randoms.py", 1, p)      # idempotent re-entry
    run("pk/compute_pk_data.py", 1, p)
    f = glob.glob(od + "pk_*.txt")[0]
    assert os.path.basename(f) == str(g["pk_txt_name"])
    txt = open(f).read()
    ref = str(g["pk_txt"])
    assert txt.split("\n")[:15] == ref.split("\n")[:15]                      # header lines identical
    assert relerr(np.loadtxt(f), g["pk_p"]) < 2e-8
    assert not glob.glob(mc + "*_pk_fish_a_*")                               # per-seed files folded + removed
    assert relerr(np.load(glob.glob(od + "fisher-pk_*.npy")[0]), g["pk_fish_comb"]) < 1e-9
    # bispectrum
    run("bk/compute_bk_randoms.py", 1, p)
    F = np.load(glob.glob(mc + "*1_FKP_bk_fish_alpha_beta_*.npy")[0])
    assert relerr(F, g["bk_fish_1"]) < 1e-9
    z = np.load(glob.glob(mc + "*2_FKP_bk_bias_alpha_*.npz")[0])
    assert tuple(z["g_a"].shape) == tuple(g["bk_g_shape"])
    run("bk/compute_bk_data.py", 1, p)
    f = glob.glob(od + "bk_*.txt")[0]
    assert os.path.basename(f) == str(g["bk_txt_name"])
    assert open(f).read().split("\n")[:13] == str(g["bk_txt"]).split("\n")[:13]
    assert relerr(np.loadtxt(f), g["bk_p"]) < 2e-8


def test_error_behaviour(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "pk/compute_pk_randoms.py")], capture_output=True, text=True)
    assert r.returncode != 0 and "Need to specify random iteration and parameter file!" in r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "pk/compute_pk_data.py"), "1", str(tmp_path) + "/nope.param"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Parameter file does not exist!" in r.stderr


def test_mc_driver_writes_combined_files(world, golden, tmp_path):
    """The sharded MC driver (here 1 rank) reproduces the combined Fisher / bias files that
    compute_pk_data.py would fold from the per-seed files, and the bk combined Fisher."""
    from spectra_without_windows_b200 import synthetic as syn
    w, g = world("toyD"), golden("toyD")
    p = syn.write_world_files(w, str(tmp_path))
    od = str(tmp_path) + "/out/"
    for spec in ("pk", "bk"):
        r = subprocess.run([sys.executable, "-m", "spectra_without_windows_b200.mc_driver", spec, p], capture_output=True,
                           text=True, cwd=ROOT, env=ENV128)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert relerr(np.load(glob.glob(od + "fisher-pk_*.npy")[0]), g["pk_fish_comb"]) < 1e-9
    assert relerr(np.load(glob.glob(od + "bias-pk_*.npy")[0]), g["pk_bias_comb"]) < 1e-9
    assert relerr(np.load(glob.glob(od + "fisher-bk_*.npy")[0]), g["bk_fish_comb"]) < 1e-9
    # and the data scripts pick the combined files up
    run("pk/compute_pk_data.py", 1, p)
    run("bk/compute_bk_data.py", 1, p)
    assert relerr(np.loadtxt(glob.glob(od + "pk_*.txt")[0]), g["pk_p"]) < 2e-8
    assert relerr(np.loadtxt(glob.glob(od + "bk_*.txt")[0]), g["bk_p"]) < 2e-8


def test_bk_in_memory_handoff_to_the_data_step(world, golden, tmp_path):
    """`mc_driver bk --data 1`: the Monte-Carlo g / g~ maps feed the catalogue's 1-field sums on the device (no .npz
    round trip) and rank 0 writes the B_l text compute_bk_data.py would write; the contract files still appear
    (asynchronous complex64 writer), and the classic two-step path on those files gives the same numbers."""
    from spectra_without_windows_b200 import synthetic as syn
    w, g = world("toyD"), golden("toyD")
    p = syn.write_world_files(w, str(tmp_path))
    mc, od = str(tmp_path) + "/mc/", str(tmp_path) + "/out/"
    r = subprocess.run([sys.executable, "-m", "spectra_without_windows_b200.mc_driver", "bk", p, "--data", "1"], capture_output=True,
                       text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    f = glob.glob(od + "bk_*.txt")[0]
    assert os.path.basename(f) == str(g["bk_txt_name"])
    assert open(f).read().split("\n")[:13] == str(g["bk_txt"]).split("\n")[:13]
    fused = np.loadtxt(f)
    assert relerr(fused, g["bk_p"]) < 2e-8
    z = np.load(sorted(glob.glob(mc + "*_FKP_bk_bias_alpha_*.npz"))[0])
    assert z["g_a"].dtype == np.complex64 and tuple(z["g_a"].shape) == tuple(g["bk_g_shape"])
    os.remove(f)
    run("bk/compute_bk_data.py", 1, p, env=None)              # two-step path on the written files (complex64 maps)
    two = np.loadtxt(glob.glob(od + "bk_*.txt")[0])
    assert relerr(two[:, 3:], fused[:, 3:]) < 1e-5            # single-precision files, as in the reference
