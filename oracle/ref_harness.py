"""TEST INFRASTRUCTURE -- not part of the product path.

Runs the UNMODIFIED reference scripts under /root/reference (pk/compute_pk_randoms.py,
pk/compute_pk_data.py, bk/compute_bk_randoms.py, bk/compute_bk_data.py and the src/
functions they import) in this container, to pin the numpy restatement in
oracle/sww_oracle.py and to generate the golden vectors committed under tests/golden/
(generator: oracle/make_golden.py).  /root/reference does not exist on the GPU box, so
nothing imported by `-m gpu` tests, smoke() or bench.py may import this module.

Recipe:
  * sys.modules stubs for the absent third-party modules: pyfftw (-> scipy.fft),
    matplotlib, nbodykit.lab (exports `numpy` and a dummy `cosmology`);
  * sympy.lambdify(..., 'numexpr') redirected to 'numpy' (src/opt_utilities.py:383);
  * load_data / load_randoms / grid_data replaced by synthetic-catalogue versions BEFORE
    the script binds them by name (pk/compute_pk_randoms.py:12); the painter is the plain
    TSC restatement of src/opt_utilities.py:219-267 (pmesh is absent -> painting parity is
    "unpinned": our definition, stated in DESIGN.md);
  * FP64-promoted mode: the pyfftw stub allocates complex128 whatever dtype is requested
    and covariances_pk sees np.complex64 == np.complex128 (src/covariances_pk.py:154).
"""
from __future__ import annotations

import os
import runpy
import sys
import types

import numpy as np
import scipy.fft

REF = "/root/reference"
_state = {"fp64": True, "workers": os.cpu_count(), "world": None, "nfft": [0, 0]}


class Lazy:
    """Minimal stand-in for a dask column: arithmetic + .sum/.mean + .compute()."""

    def __init__(self, v):
        self.v = np.asarray(v)

    def compute(self):
        return self.v[()] if self.v.ndim == 0 else self.v

    def sum(self):
        return Lazy(self.v.sum())

    def mean(self):
        return Lazy(self.v.mean())

    def __pow__(self, p):
        return Lazy(self.v ** p)

    def _o(self, o):
        return o.v if isinstance(o, Lazy) else o

    def __truediv__(self, o):
        return Lazy(self.v / self._o(o))

    def __mul__(self, o):
        return Lazy(self.v * self._o(o))

    def __add__(self, o):
        return Lazy(self.v + self._o(o))

    def __len__(self):
        return len(self.v)


def tsc_paint(pos, w, box, grid, box_center):
    """Plain TSC deposit: cell = rint((x-BoxCenter)/H) mod N, weights 1/2(1/2-d)^2, 3/4-d^2,
    1/2(1/2+d)^2 with d = x/H - rint(x/H).  Restates what FKPCatalog.to_mesh(window='tsc')
    with compensation disabled does (src/opt_utilities.py:219-260)."""
    grid = np.asarray(grid, int)
    H = np.asarray(box, float) / grid
    s = (pos - np.asarray(box_center, float)) / H
    c = np.rint(s)
    d = s - c
    c = c.astype(np.int64)
    wts = [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]
    out = np.zeros(grid, dtype=np.float64)
    for ix in range(3):
        i = (c[:, 0] + ix - 1) % grid[0]
        for iy in range(3):
            j = (c[:, 1] + iy - 1) % grid[1]
            for iz in range(3):
                k = (c[:, 2] + iz - 1) % grid[2]
                np.add.at(out, (i, j, k), w * wts[ix][:, 0] * wts[iy][:, 1] * wts[iz][:, 2])
    return out


class _Density:
    """Stub of the nbodykit mesh object consumed by load_coord_grids
    (src/opt_utilities.py:428-429): attrs['BoxCenter'] and slabs.optx."""

    def __init__(self, box, grid, box_center):
        from spectra_without_windows_b200.synthetic import fft_index
        H = np.asarray(box, float) / np.asarray(grid, int)
        self.attrs = {"BoxCenter": np.asarray(box_center, float)}
        optx = []
        for i in range(3):
            shp = [1, 1, 1]
            shp[i] = grid[i]
            optx.append((fft_index(grid[i]) * H[i]).reshape(shp))
        self.slabs = types.SimpleNamespace(optx=optx)


def install_stubs():
    """Install sys.modules stubs; idempotent."""
    if "pyfftw" in sys.modules and getattr(sys.modules["pyfftw"], "_sww_stub", False):
        return
    pyfftw = types.ModuleType("pyfftw")
    pyfftw._sww_stub = True

    def empty_aligned(shape, dtype="complex64"):
        return np.empty(shape, dtype=("complex128" if _state["fp64"] else dtype))

    class FFTW:
        def __init__(self, a_in, a_out, axes=(0, 1, 2), flags=(), direction="FFTW_FORWARD", threads=1):
            self.d = direction

        def __call__(self, a_in, a_out):
            W = _state["workers"]
            if self.d == "FFTW_FORWARD":
                _state["nfft"][0] += 1
                a_out[:] = scipy.fft.fftn(a_in, workers=W)
            else:  # pyfftw normalises the inverse by 1/N by default
                _state["nfft"][1] += 1
                a_out[:] = scipy.fft.ifftn(a_in, workers=W)
            return a_out

    pyfftw.empty_aligned = empty_aligned
    pyfftw.FFTW = FFTW
    sys.modules["pyfftw"] = pyfftw
    m = types.ModuleType("matplotlib")
    m.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = m
    sys.modules["matplotlib.pyplot"] = m.pyplot
    lab = types.ModuleType("nbodykit.lab")
    lab.numpy = np
    lab.cosmology = types.SimpleNamespace(
        Cosmology=lambda **k: types.SimpleNamespace(match=lambda **k: None))
    lab.__all__ = ["numpy", "cosmology"]
    sys.modules["nbodykit"] = types.ModuleType("nbodykit")
    sys.modules["nbodykit.lab"] = lab
    import sympy as sp
    if not getattr(sp, "_sww_patched", False):
        _orig = sp.lambdify

        def lambdify(args, expr, modules=None, **kw):
            if modules == "numexpr":
                modules = "numpy"
            return _orig(args, expr, modules, **kw)

        sp.lambdify = lambdify
        sp._sww_patched = True
    if REF + "/src" not in sys.path:
        sys.path.insert(0, REF + "/src")


def ref_modules(fp64=True):
    """Import the reference's src modules under the stubs and set the precision mode."""
    install_stubs()
    _state["fp64"] = fp64
    import opt_utilities as ou
    import covariances_pk as cp

    class _NP:
        def __getattr__(self, k):
            if k == "complex64" and _state["fp64"]:
                return np.complex128
            return getattr(np, k)

    cp.np = _NP()
    return ou, cp


def bind_world(world, fp64=True):
    """Replace the nbodykit-dependent entry points of opt_utilities with synthetic versions."""
    ou, cp = ref_modules(fp64)
    _state["world"] = world

    def load_data(sim_no, config, cosmo, **k):
        return {"WEIGHT": Lazy(world.data["w"]), "pos": world.data["pos"]}

    def load_randoms(config, cosmo, **k):
        return {"WEIGHT": Lazy(world.randoms["w"]), "pos": world.randoms["pos"]}

    def grid_data(data, randoms, boxsize_grid, grid_3d, MAS="TSC", return_randoms=True, return_norm=False):
        assert MAS == "TSC" and not return_norm
        wd, wr = data["WEIGHT"].v, randoms["WEIGHT"].v
        alpha = wd.sum() / wr.sum()
        v_cell = np.prod(boxsize_grid) / np.prod(grid_3d)
        ng = tsc_paint(data["pos"], wd, boxsize_grid, grid_3d, world.box_center)
        nr = tsc_paint(randoms["pos"], wr, boxsize_grid, grid_3d, world.box_center)
        diff = (ng - alpha * nr) / v_cell
        dens = _Density(boxsize_grid, grid_3d, world.box_center)
        if return_randoms:
            return diff, alpha * nr / v_cell, dens
        return diff, dens

    ou.load_data, ou.load_randoms, ou.grid_data = load_data, load_randoms, grid_data
    return ou, cp


def run_script(script_rel, int_arg, paramfile, quiet=True, capture=None):
    """runpy one of the four reference scripts, verbatim.  Returns (n_fwd, n_inv) FFT counts.
    `capture` (a list) receives every 1-D result of np.inner / np.matmul during the run: the FP64 p_alpha vectors of
    pk/compute_pk_data.py:250 and bk/compute_bk_data.py:417, which the scripts only write as '%.8e' text."""
    _state["nfft"] = [0, 0]
    argv = sys.argv
    sys.argv = [REF + "/" + script_rel, str(int_arg), paramfile]
    out = sys.stdout
    import numpy as _np
    saved = (_np.inner, _np.matmul)
    if capture is not None:
        def _wrap(fn):
            def f(*a, **k):
                r = fn(*a, **k)
                if getattr(r, "ndim", 0) == 1:
                    capture.append(_np.array(r, dtype=_np.float64, copy=True))
                return r
            return f
        _np.inner, _np.matmul = _wrap(saved[0]), _wrap(saved[1])
    try:
        if quiet:
            sys.stdout = open(os.devnull, "w")
        try:
            runpy.run_path(REF + "/" + script_rel, run_name="__main__")
        except SystemExit:
            pass
    finally:
        _np.inner, _np.matmul = saved
        if quiet:
            sys.stdout.close()
        sys.stdout = out
        sys.argv = argv
    return tuple(_state["nfft"])
