"""Deterministic synthetic survey worlds (meshes, n(r) maps, catalogues, .param files).

These are the inputs shared byte-for-byte by the CUDA path, the oracle and the
verbatim-reference harness.  Nothing here is on the hot
path: it only builds numpy inputs and writes the files the four driver scripts
read (`.param` in the reference's INI layout, paramfiles/boss_nC.param:1-74, and
the `nbar_*.npy` map whose name is fixed by src/opt_utilities.py:175).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


def fft_index(n: int) -> np.ndarray:
    """Signed FFT-ordered integer index, as built in src/opt_utilities.py:422
    (`kkk-grid if kkk>=grid/2 else kkk`)."""
    i = np.arange(n)
    mid = n / 2
    return np.where(i >= mid, i - n, i)


def mesh_axes(box, grid, box_center):
    """1-D real-space cell coordinates per axis: fft-ordered mesh coordinate +
    BoxCenter + half a cell (src/opt_utilities.py:428-430)."""
    box = np.asarray(box, float)
    grid = np.asarray(grid, int)
    H = box / grid
    return [fft_index(grid[i]) * H[i] + box_center[i] + 0.5 * H[i] for i in range(3)]


@dataclass
class World:
    name: str
    grid: tuple
    box: tuple
    box_center: tuple
    type: str = "syn_"
    zmin: float = 0.43
    zmax: float = 0.70
    pk_bins: tuple = (0.0, 0.12, 0.02, 4)      # k_min, k_max, dk, lmax
    bk_bins: tuple = (0.0, 0.08, 0.02, 2)
    N_mc: int = 2
    h_fid: float = 0.676
    OmegaM_fid: float = 0.31
    nbar_r: np.ndarray = None                  # randoms-density map n(r)/alpha  [nx,ny,nz]
    data: dict = field(default_factory=dict)     # {'pos': [Ng,3], 'w': [Ng]}
    randoms: dict = field(default_factory=dict)  # {'pos': [Nr,3], 'w': [Nr]}

    @property
    def alpha(self):
        return self.data["w"].sum() / self.randoms["w"].sum()

    @property
    def shot_fac(self):
        # pk/compute_pk_randoms.py:147
        a = self.alpha
        return ((self.data["w"] ** 2.0).mean() + a * (self.randoms["w"] ** 2.0).mean()) / self.randoms["w"].mean()

    @property
    def v_cell(self):
        return float(np.prod(self.box)) / float(np.prod(self.grid))

    def nbar_file(self, outdir):
        g = self.grid
        return outdir + "nbar_%sz%.3f_%.3f_g%d_%d_%d.npy" % (self.type, self.zmin, self.zmax, g[0], g[1], g[2])


def survey_nbar(box, grid, box_center, n0=4e-4, r0=None, sig=None, rmin=None, rmax=None,
                lon_deg=70.0, lat_deg=30.0):
    """BOSS-like synthetic n(r): Gaussian shell in |r| times an angular footprint with a
    smooth modulation.  Returns the *galaxy* density n(r)."""
    x, y, z = mesh_axes(box, grid, box_center)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    cx = float(np.linalg.norm(box_center))
    if r0 is None:
        r0 = cx + 0.1 * box[0]
    if sig is None:
        sig = 0.14 * box[0]
    if rmin is None:
        rmin = r0 - 0.14 * box[0] * 1.8
    if rmax is None:
        rmax = r0 + 0.14 * box[0] * 2.5
    lon = np.arctan2(Y, X)
    lat = np.arcsin(Z / np.maximum(r, 1e-30))
    W = (np.abs(lon) < np.deg2rad(lon_deg)) * (np.abs(lat) < np.deg2rad(lat_deg)) * (0.9 + 0.1 * np.cos(3.0 * lon))
    n = n0 * np.exp(-0.5 * ((r - r0) / sig) ** 2) * (r > rmin) * (r < rmax) * W
    return n


def sample_catalog(rng, dens, box, grid, box_center, n_target):
    """Draw ~n_target points with density proportional to `dens` (cell picked with
    probability ~ dens, uniform inside the cell)."""
    grid = np.asarray(grid, int)
    H = np.asarray(box, float) / grid
    p = dens.ravel() / dens.sum()
    cells = rng.choice(p.size, size=n_target, p=p)
    ijk = np.stack(np.unravel_index(cells, grid), axis=1)
    ax = mesh_axes(box, grid, box_center)
    pos = np.stack([ax[d][ijk[:, d]] for d in range(3)], axis=1)
    pos += (rng.random((n_target, 3)) - 0.5) * H
    return pos


def make_world(name: str) -> World:
    """Named deterministic worlds.

    toyA : 32x24x20 non-cubic mesh, masked n(r) with empty cells, weighted catalogues.
    toyB : 32^3 cubic power-of-two mesh, uniform n(r)   (exercises the custom FFT path)
    toyC : 20x18x15 mesh with an odd last axis (no Nyquist plane on z)
    toyD : 64x64x128 power-of-two mesh, masked n(r)   (exercises the hand-written FFT passes)
    """
    if name == "toyA":
        grid, box, bc = (32, 24, 20), (640.0, 480.0, 400.0), (900.0, 40.0, -30.0)
        pk_bins, bk_bins = (0.0, 0.12, 0.02, 4), (0.0, 0.08, 0.02, 2)
        ng, seed = 4000, 11
    elif name == "toyB":
        grid, box, bc = (32, 32, 32), (640.0, 640.0, 640.0), (0.0, 0.0, 1500.0)
        pk_bins, bk_bins = (0.0, 0.12, 0.02, 4), (0.0, 0.10, 0.02, 4)
        ng, seed = 6000, 21
    elif name == "toyC":
        grid, box, bc = (20, 18, 15), (400.0, 360.0, 300.0), (700.0, -20.0, 35.0)
        pk_bins, bk_bins = (0.01, 0.13, 0.03, 2), (0.01, 0.10, 0.03, 2)
        ng, seed = 2500, 31
    elif name == "toyD":   # smallest mesh the hand-written sm_100a FFT passes accept (64 x 64 x 128)
        grid, box, bc = (64, 64, 128), (1280.0, 1408.0, 2560.0), (1700.0, 60.0, -45.0)
        pk_bins, bk_bins = (0.0, 0.12, 0.02, 4), (0.0, 0.08, 0.02, 2)
        ng, seed = 30000, 41
    else:
        raise Exception("unknown world '%s'" % name)
    rng = np.random.default_rng(seed)
    if name == "toyB":
        n_gal = np.full(grid, ng / float(np.prod(box)))
    else:
        n_gal = survey_nbar(box, grid, bc, n0=1.0)
        n_gal *= ng / (n_gal.sum() * np.prod(box) / np.prod(grid))
    nr = 20 * ng
    dpos = sample_catalog(rng, n_gal, box, grid, bc, ng)
    rpos = sample_catalog(rng, n_gal, box, grid, bc, nr)
    if name == "toyA":
        dw = 0.8 + 0.4 * rng.random(ng)
        rw = 0.9 + 0.2 * rng.random(nr)
    else:
        dw, rw = np.ones(ng), np.ones(nr)
    w = World(name=name, grid=grid, box=box, box_center=bc, type=name + "_",
              pk_bins=pk_bins, bk_bins=bk_bins, N_mc=2)
    w.data = {"pos": dpos, "w": dw}
    w.randoms = {"pos": rpos, "w": rw}
    # randoms-density map, normalised like generate_mask.py:140-142: sum(n V_cell) = sum(w_randoms)
    nbar_r = n_gal.copy()
    nbar_r *= rw.sum() / (nbar_r.sum() * w.v_cell)
    w.nbar_r = nbar_r
    return w


PARAM_TEMPLATE = """### synthetic parameter file ({name})
[catalogs]
data_file = {data_file}
data_columns =
randoms_file = {randoms_file}
randoms_columns =
mask_file = none

[pk-binning]
k_min = {pk[0]}
k_max = {pk[1]}
dk = {pk[2]}
lmax = {pk[3]}
grid = {g[0]}, {g[1]}, {g[2]}

[bk-binning]
k_min = {bk[0]}
k_max = {bk[1]}
dk = {bk[2]}
lmax = {bk[3]}
grid = {g[0]}, {g[1]}, {g[2]}

[sample]
type = {type}
z_min = {zmin}
z_max = {zmax}
box = {b[0]}, {b[1]}, {b[2]}

[parameters]
h_fid = {h}
OmegaM_fid = {om}

[directories]
temporary = {tmp}
monte_carlo = {mc}
output = {out}

[settings]
weights = {weights}
N_mc = {nmc}
include_pix = {include_pix}
rand_nbar = {rand_nbar}
fiducial_pk = {fiducial_pk}
use_qbar = {use_qbar}
nbar_unif = 1e-3
low_mem = True
"""


def fiducial_pk_table(k_top, n=257):
    """Synthetic fiducial multipoles for the ML weights: columns k, P_0, P_2, P_4 like the `fiducial_pk` text file
    of the reference (pk/compute_pk_randoms.py:188-193).  Amplitude comparable to the shot noise of the toy
    worlds, so that the conjugate-gradient inversion of C = S + N takes several iterations."""
    k = np.linspace(0.0, float(k_top), n)
    P0 = 3.0e4 / (1.0 + (k / 0.02) ** 2) ** 0.6
    P2 = 0.5 * P0 * k / (k + 0.05)
    P4 = 0.1 * P0 * (k / (k + 0.1)) ** 2
    return np.stack([k, P0, P2, P4], axis=1)


def write_world_files(w: World, root: str, use_qbar=True, N_mc=None, include_pix=False, rand_nbar=False,
                      weights="FKP"):
    """Write the .param file, the nbar map and .npz Cartesian catalogues under `root`.
    Returns the param-file path."""
    out, mc, tmp = root + "/out/", root + "/mc/", root + "/tmp/"
    for d in (out, mc, tmp):
        os.makedirs(d, exist_ok=True)
    fid = "None"
    if weights == "ML":
        fid = root + "/fiducial_pk.txt"
        k_top = 1.01 * float(np.sqrt(np.sum((np.pi * np.asarray(w.grid) / np.asarray(w.box, float)) ** 2)))
        np.savetxt(fid, fiducial_pk_table(k_top))
    np.save(w.nbar_file(out), w.nbar_r)
    dfile, rfile = root + "/data_%d.npz", root + "/randoms.npz"
    np.savez(dfile % 1, pos=w.data["pos"], w=w.data["w"], box_center=np.asarray(w.box_center, float))
    np.savez(rfile, pos=w.randoms["pos"], w=w.randoms["w"], box_center=np.asarray(w.box_center, float))
    pfile = root + "/%s.param" % w.name
    with open(pfile, "w") as f:
        f.write(PARAM_TEMPLATE.format(name=w.name, data_file=dfile, randoms_file=rfile, pk=w.pk_bins,
                                      bk=w.bk_bins, g=w.grid, type=w.type, zmin=w.zmin, zmax=w.zmax,
                                      b=w.box, h=w.h_fid, om=w.OmegaM_fid, tmp=tmp + w.type, mc=mc, out=out,
                                      nmc=(w.N_mc if N_mc is None else N_mc), use_qbar=use_qbar,
                                      include_pix=include_pix, rand_nbar=rand_nbar, weights=weights,
                                      fiducial_pk=fid))
    return pfile
