"""Golden vectors on the reference's PRODUCTION meshes (paramfiles/boss_nC.param:23,33 and the Nseries bispectrum mesh), from the
numpy restatement of the reference (oracle/sww_oracle.py, itself pinned to the unmodified scripts on the toy worlds): q-bar and the
Fisher matrix of one P_l(k) realisation at full mesh size with a coarse k binning (n_b = 15), so that the oracle finishes in
minutes and the fixture stays small.  Test infrastructure only.

    python oracle/make_golden_production.py        ->  tests/golden/production_meshes.npz  (~25 GB of host memory, ~15 min)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import sww_oracle as orc                                   # noqa: E402
from spectra_without_windows_b200 import synthetic as syn              # noqa: E402

BOX = (1800.0, 3450.0, 1900.0)
CASES = [((258, 496, 274), 0.41, 0.08), ((167, 307, 175), 0.11, 0.02), ((172, 330, 182), 0.16, 0.04)]
ALPHA, SHOT, SEED = 0.05, 1.05, 2


def inputs(shape):
    """The inputs of tests/test_gpu_fft_any.py::test_production_meshes_*: masked survey n(r), one Gaussian field."""
    bc = (2.0 * BOX[0], 30.0, -20.0)
    n_gal = syn.survey_nbar(BOX, shape, bc, n0=3e-4)
    return bc, n_gal / ALPHA


def main():
    out = {}
    for shape, kmax, dk in CASES:
        t0 = time.time()
        bc, nbar_r = inputs(shape)
        S = orc.Setup(BOX, shape, bc, nbar_r, ALPHA, SHOT, 0.0, kmax, dk, 4)
        a = np.random.default_rng(SEED).normal(size=shape) * np.sqrt(S.nbar_A)
        qbar, F = orc.pk_fisher_realisation(S, a)
        tag = "%dx%dx%d" % shape
        out["F_" + tag] = F
        out["qbar_" + tag] = qbar
        print(tag, "n_b", F.shape[0], "%.0f s" % (time.time() - t0), flush=True)
        del S, a
    np.savez(os.path.join(ROOT, "tests", "golden", "production_meshes.npz"), **out)


if __name__ == "__main__":
    main()
