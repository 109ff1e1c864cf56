"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, exports every
symbol include/sww.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spectra_without_windows_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from spectra_without_windows_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sww.h")).read()
    declared = set(re.findall(r"\b(sww_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    # and the ctypes prototypes cover the header
    assert declared - {"sww_last_error"} == set(_lib.SIGNATURES)
    assert lib.sww_abi_version() == 1


def test_compiled_for_sm100a_only():
    so = os.path.join(ROOT, "spectra_without_windows_b200", "libsww.so")
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spectra_without_windows_b200 import engine, synthetic
    w = synthetic.make_world("toyC")
    mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
    with pytest.raises(Exception, match="no CUDA device"):
        engine.Context(mesh, 0)


def test_mesh_matches_oracle_grids():
    """Host geometry tables are the reference expressions (bit-exact vs the oracle's grids)."""
    from oracle import sww_oracle as orc
    from spectra_without_windows_b200 import engine, synthetic
    for name in ("toyA", "toyC"):
        w = synthetic.make_world(name)
        mesh = engine.Mesh(w.box, w.grid, w.box_center, *w.pk_bins)
        k, r = orc.load_coord_grids(w.box, w.grid, w.box_center)
        for d in range(3):
            sl = [0, 0, 0]
            sl[d] = slice(None)
            assert np.array_equal(k[d][tuple(sl)], mesh.kax[d])
            assert np.array_equal(r[d][tuple(sl)], mesh.rax[d])
        M = orc.load_MAS(w.box, w.grid)
        assert np.array_equal(M, mesh.mas[0][:, None, None] * mesh.mas[1][None, :, None] * mesh.mas[2][None, None, :])
        lo, hi = orc.bin_edges(*w.pk_bins[:3])
        assert np.array_equal(lo[:mesh.n_k], mesh.k_lo) and np.array_equal(hi[:mesh.n_k], mesh.k_hi)
        assert mesh.triangles() == orc.triangle_bins(mesh.k_min, mesh.k_max, mesh.dk, mesh.n_k)
