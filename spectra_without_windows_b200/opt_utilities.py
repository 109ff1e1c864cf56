"""Drop-in mirror of the reference's src/opt_utilities.py call surface (same names, argument
meaning and error behaviour: bare `Exception`s), backed by libsww.so on the GPU.

ndarray in -> ndarray out, like the reference, so the parity tests read like the reference's
own call sites.  The four driver scripts in pk/ and bk/ do NOT round-trip maps through these
mirrors: they call the fused device pipelines of `engine.Context` directly.

`load_data` / `load_randoms` read FITS binary tables, text catalogues (through `ingest.py`: the reference's column logic
and a restated sky -> Cartesian transform) and Cartesian .npz catalogues.  Out of scope: `plotter`.
"""
from __future__ import annotations

import os
import types

import numpy as np

from . import engine
from ._ylm_table import YLM_TABLE


# ----------------------------------------------------------------------------- catalogues
class Column:
    """Stand-in for a dask column: arithmetic, .sum(), .mean(), .compute() (pk/compute_pk_randoms.py:146-147)."""

    def __init__(self, v):
        self.v = np.asarray(v)

    def compute(self):
        return self.v[()] if self.v.ndim == 0 else self.v

    def sum(self):
        return Column(self.v.sum())

    def mean(self):
        return Column(self.v.mean())

    def __pow__(self, p):
        return Column(self.v ** p)

    @staticmethod
    def _o(o):
        return o.v if isinstance(o, Column) else o

    def __truediv__(self, o):
        return Column(self.v / self._o(o))

    def __mul__(self, o):
        return Column(self.v * self._o(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return Column(self.v + self._o(o))

    def __len__(self):
        return len(self.v)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.v, dtype=dtype)


class Catalog(dict):
    """Minimal catalogue: columns 'Position' [n,3] (Cartesian, Mpc/h), 'WEIGHT', 'WEIGHT_FKP', 'NBAR'."""

    def __init__(self, pos, w, nbar=None, box_center=None):
        super().__init__()
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        w = np.ascontiguousarray(w, dtype=np.float64)
        self["Position"] = Column(pos)
        self["WEIGHT"] = Column(w)
        self["WEIGHT_FKP"] = Column(np.ones(len(w)))
        self["NBAR"] = Column(np.ones(len(w)) if nbar is None else np.asarray(nbar, dtype=np.float64))
        self.attrs = {}
        if box_center is not None:
            self.attrs["BoxCenter"] = np.asarray(box_center, dtype=np.float64)


def _catalog_from_columns(cat):
    c = Catalog(cat["Position"], cat["WEIGHT"], cat.get("NBAR"))
    c["WEIGHT_FKP"] = Column(np.asarray(cat["WEIGHT_FKP"], dtype=np.float64))
    for k in ("Z", "RA", "DEC"):
        if k in cat:
            c[k] = Column(cat[k])
    return c


def _load_cartesian(path):
    """Cartesian .npz catalogue (pos, w[, nbar, box_center]): positions already in Mpc/h, no redshift cut."""
    if not os.path.exists(path):
        raise Exception("Catalogue file '%s' does not exist!" % path)
    z = np.load(path)
    return Catalog(z["pos"], z["w"], z["nbar"] if "nbar" in z else None, z["box_center"] if "box_center" in z else None)


def _getlist(config, sec, key):
    try:
        return [c for c in config.getlist(sec, key) if c]
    except Exception:
        return [c.strip() for c in str(config[sec][key]).split(",") if c.strip()]


def load_data(sim_no, config, cosmo, fkp_weights=False, weight_only=False, P_fkp=1e4):
    """Mirror of src/opt_utilities.py:12-82.  `data_file` may contain '%d' (replaced by sim_no, with the zero-padded
    fall-back of :38-40); `.fits` tables and whitespace-separated text files go through the reference's column logic
    (redshift cut, WEIGHT conventions, SkyToCartesian with `cosmo`, NBAR / WEIGHT_FKP) in `ingest.py`; Cartesian `.npz`
    catalogues (pos, w[, nbar, box_center]) are taken as they are."""
    from . import ingest
    name = str(config["catalogs"]["data_file"])
    if "%d" in name:
        if sim_no != -1:
            datfile = name % sim_no
            if not os.path.exists(datfile):
                datfile = datfile.replace("_%d" % sim_no, "_%s" % (str(sim_no).zfill(4)))
            name = datfile
        else:
            sim_no = -1
    else:
        sim_no = -1
    if name.endswith(".npz"):
        cat = _load_cartesian(name)
        if weight_only:
            return cat["WEIGHT"]
        if fkp_weights:
            cat["WEIGHT_FKP"] = Column(1.0 / (1.0 + cat["NBAR"].v * P_fkp))
        return cat
    if cosmo is None:
        raise Exception("a fiducial cosmology is needed to convert '%s' to Cartesian co-ordinates" % name)
    cols = ingest.read_catalog(name, _getlist(config, "catalogs", "data_columns"))
    cols = ingest.apply_reference_columns(cols, float(config["sample"]["z_min"]), float(config["sample"]["z_max"]), cosmo,
                                          True, sim_no, fkp_weights, P_fkp)
    print("Loaded %d galaxies\n" % len(cols["WEIGHT"]))
    if weight_only:
        return Column(cols["WEIGHT"])
    return _catalog_from_columns(cols)


def load_randoms(config, cosmo, fkp_weights=False, weight_only=False, P_fkp=1e4):
    """Mirror of src/opt_utilities.py:84-143 (see load_data)."""
    from . import ingest
    name = str(config["catalogs"]["randoms_file"])
    if name.endswith(".npz"):
        cat = _load_cartesian(name)
        if weight_only:
            return cat["WEIGHT"]
        if fkp_weights:
            cat["WEIGHT_FKP"] = Column(1.0 / (1.0 + cat["NBAR"].v * P_fkp))
        return cat
    if cosmo is None:
        raise Exception("a fiducial cosmology is needed to convert '%s' to Cartesian co-ordinates" % name)
    cols = ingest.read_catalog(name, _getlist(config, "catalogs", "randoms_columns"))
    cols = ingest.apply_reference_columns(cols, float(config["sample"]["z_min"]), float(config["sample"]["z_max"]), cosmo,
                                          False, -1, fkp_weights, P_fkp)
    print("Loaded %d randoms\n" % len(cols["WEIGHT"]))
    if weight_only:
        return Column(cols["WEIGHT"])
    return _catalog_from_columns(cols)


def load_nbar(config, spec_type, alpha_ran):
    """src/opt_utilities.py:145-181, unchanged contract: np.load(nbar file) * alpha_ran."""
    outdir = str(config["directories"]["output"])
    if spec_type == "pk":
        grid_3d = np.array(config.getlist("pk-binning", "grid"), dtype=int)
    elif spec_type == "bk":
        grid_3d = np.array(config.getlist("bk-binning", "grid"), dtype=int)
    else:
        raise Exception("Spectrum type must be 'pk' or 'bk'!")
    string = str(config["sample"]["type"])
    ZMIN, ZMAX = float(config["sample"]["z_min"]), float(config["sample"]["z_max"])
    nbar_file = outdir + "nbar_%sz%.3f_%.3f_g%d_%d_%d.npy" % (string, ZMIN, ZMAX, grid_3d[0], grid_3d[1], grid_3d[2])
    if not os.path.exists(nbar_file):
        raise Exception("n_bar file '%s' has not been computed!" % nbar_file)
    return np.load(nbar_file) * alpha_ran


# ----------------------------------------------------------------------------- meshes
class MeshInfo:
    """What the scripts use of nbodykit's `density` object: attrs['BoxCenter'] and slabs.optx
    (src/opt_utilities.py:428-429)."""

    def __init__(self, boxsize_grid, grid_3d, box_center):
        self.attrs = {"BoxCenter": np.asarray(box_center, dtype=np.float64), "BoxSize": np.asarray(boxsize_grid, float),
                      "Nmesh": np.asarray(grid_3d, int)}
        H = np.asarray(boxsize_grid, float) / np.asarray(grid_3d, int)
        optx = []
        for i in range(3):
            shp = [1, 1, 1]
            shp[i] = int(grid_3d[i])
            optx.append((engine.signed_index(int(grid_3d[i])) * H[i]).reshape(shp))
        self.slabs = types.SimpleNamespace(optx=optx)


def box_center_of(randoms):
    """nbodykit's FKPCatalog centres the box on the randoms' extent; a catalogue may pin it explicitly."""
    if getattr(randoms, "attrs", None) and "BoxCenter" in randoms.attrs:
        return np.asarray(randoms.attrs["BoxCenter"], dtype=np.float64)
    pos = randoms["Position"].v
    return 0.5 * (pos.min(axis=0) + pos.max(axis=0))


_BARE = {}


def _bare_context(boxsize_grid, grid_3d, box_center=(0.0, 0.0, 0.0)):
    """A context that only needs the mesh (painting, plain FFTs): dummy one-bin binning below Nyquist."""
    box = np.asarray(boxsize_grid, float)
    grid = np.asarray(grid_3d, int)
    key = (tuple(box), tuple(grid), tuple(np.asarray(box_center, float)))
    if key not in _BARE:
        kny = (np.pi * grid / box).min()
        mesh = engine.Mesh(box, grid, box_center, 0.0, 0.5 * kny, 0.5 * kny, 0)
        _BARE[key] = engine.Context(mesh, 0, ("ops",))
    return _BARE[key]


def grid_data(data, randoms, boxsize_grid, grid_3d, MAS="TSC", return_randoms=True, return_norm=False):
    """Mirror of src/opt_utilities.py:184-276: TSC mesh of (n_g - alpha n_r)/V_cell, optionally the
    randoms mesh alpha n_r / V_cell and the normalisation.  Returns ndarrays + a MeshInfo `density`."""
    assert MAS == "TSC"
    bc = box_center_of(randoms)
    ctx = _bare_context(boxsize_grid, grid_3d, bc)
    dw = data["WEIGHT"].v * data["WEIGHT_FKP"].v
    rw = randoms["WEIGHT"].v * randoms["WEIGHT_FKP"].v
    alpha_norm = data["WEIGHT"].v.sum() / randoms["WEIGHT"].v.sum()
    v_cell = ctx.mesh.v_cell
    rpos_d, rw_d = ctx.dev(randoms["Position"].v), ctx.dev(rw)
    diff = ctx.paint(data["Position"].v, dw, bc, scale=1.0 / v_cell)
    ctx.paint(rpos_d, rw_d, bc, scale=-alpha_norm / v_cell, out=diff)
    density = MeshInfo(boxsize_grid, grid_3d, bc)
    density.attrs["alpha"] = alpha_norm
    out = [diff.cpu().numpy()]
    if return_randoms:
        out.append(ctx.paint(rpos_d, rw_d, bc, scale=alpha_norm / v_cell).cpu().numpy())
    out.append(density)
    if return_norm:
        out.append(1.0 / np.asarray(alpha_norm * (randoms["NBAR"].v * randoms["WEIGHT"].v * randoms["WEIGHT_FKP"].v ** 2.0).sum()))
    return tuple(out)


class GridArray(np.ndarray):
    """ndarray that remembers which mesh it was built for (so applyC_alpha can find the geometry)."""

    def __new__(cls, arr, info=None):
        obj = np.asarray(arr).view(cls)
        obj.info = info
        return obj

    def __array_finalize__(self, obj):
        self.info = getattr(obj, "info", None)


def load_coord_grids(boxsize_grid, grid_3d, density):
    """src/opt_utilities.py:400-431: (k_grids[3,nx,ny,nz], r_grids[3,nx,ny,nz])."""
    box = np.asarray(boxsize_grid, float)
    grid = np.asarray(grid_3d, int)
    kF = 2.0 * np.pi / (1.0 * box)
    kax = [engine.signed_index(int(grid[i])) * kF[i] for i in range(3)]
    kNy = kF * grid / 2.0
    print("k_Nyquist = %.2f h/Mpc" % kNy.mean())
    offset = density.attrs["BoxCenter"] + 0.5 * box / grid
    rax = [xx.real.astype("f8").ravel() + offset[ii].real for ii, xx in enumerate(density.slabs.optx)]
    info = {"box": box, "grid": grid, "box_center": np.asarray(density.attrs["BoxCenter"], float)}
    k = GridArray(np.asarray(np.meshgrid(*kax, indexing="ij")), info)
    r = GridArray(np.asarray(np.meshgrid(*rax, indexing="ij")), info)
    return k, r


def load_MAS(boxsize_grid, grid_3d):
    """src/opt_utilities.py:433-465 (float32 pi/N prefactor kept)."""
    box = np.asarray(boxsize_grid, float)
    grid = np.asarray(grid_3d, int)
    kF = 2.0 * np.pi / (1.0 * box)
    prefact = np.pi / np.asarray(grid, dtype=np.float32)
    one = []
    for i in range(3):
        kk = engine.signed_index(int(grid[i])) * kF[i] / kF[i]
        s = np.sin(prefact[i] * kk) ** 2.0
        one.append(1.0 / (1.0 - s + 2.0 / 15 * s ** 2.0) ** 0.5)
    return one[0][:, None, None] * one[1][None, :, None] * one[2][None, None, :]


class Ylm:
    """Callable real spherical harmonic with the reference's meta-data (.l, .m, .expr)."""

    def __init__(self, l, m, monomials, expr):
        self.l, self.m, self.monomials, self.expr = l, m, monomials, expr

    def __call__(self, xhat, yhat, zhat):
        out = 0.0
        for c, a, b, d in self.monomials:
            out = out + c * xhat ** a * yhat ** b * zhat ** d
        return out


def compute_spherical_harmonic_functions(lmax):
    """src/opt_utilities.py:323-398; polynomials come from the same generated table as the CUDA code."""
    out = []
    for l in np.arange(0, lmax + 1, 2):
        if int(l) > max(k[0] for k in YLM_TABLE):
            raise Exception("Y_lm table was generated up to l=%d (tools/gen_ylm.py)" % max(k[0] for k in YLM_TABLE))
        out.append(np.asarray([Ylm(int(l), int(m), *YLM_TABLE[(int(l), int(m))]) for m in np.arange(-l, l + 1, 1)]))
    return out


class Filters:
    """k-shell predicate (src/opt_utilities.py:467-488) that also exposes its binning."""

    def __init__(self, kmin, kmax, dk):
        self.kmin, self.kmax, self.dk = kmin, kmax, dk
        k_all = np.arange(kmin, kmax + dk, dk)
        self.k_lo, self.k_hi = k_all[:-1], k_all[1:]

    def __call__(self, i, k_norm):
        return np.logical_and(k_norm >= self.k_lo[i], k_norm < self.k_hi[i])


def compute_filters(kmin, kmax, dk):
    return Filters(kmin, kmax, dk)


# ----------------------------------------------------------------------------- FFTs
def _fft_ctx(shape):
    return _bare_context(np.asarray(shape, float), np.asarray(shape, int))


def ft(pix, threads=4):
    """src/opt_utilities.py:502-528 -- forward 3-D DFT (here in complex128 on the GPU)."""
    pix = np.asarray(pix)
    return _fft_ctx(pix.shape).fft_c2c(pix.astype(np.complex128), inverse=False).cpu().numpy()


def ift(pix, threads=4):
    """src/opt_utilities.py:530-555 -- inverse 3-D DFT normalised by 1/N."""
    pix = np.asarray(pix)
    return _fft_ctx(pix.shape).fft_c2c(pix.astype(np.complex128), inverse=True).cpu().numpy()


def plotter(*a, **k):
    raise Exception("plotter is out of scope of the B200 hot path (SURVEY.md section 2 row 5)")
