"""Catalogue ingestion and the n-bar(r) builder: the step BEFORE the hot path.

Mirrors, without nbodykit / mangle / fitsio (none of them is installable here):
  * `load_data` / `load_randoms` of src/opt_utilities.py:12-143 -- FITS binary tables and whitespace-separated text
    catalogues, redshift cut, the weight conventions, sky -> Cartesian, NBAR / WEIGHT_FKP columns;
  * `generate_mask.py:96-146` -- n(z) histogram of the randoms -> spline of the volume density, angular mask
    evaluated at the cell centres, product normalised so that  sum(nbar) * V_cell = sum(w_randoms)  (:140-142),
    written to `{output}nbar_{type}z%.3f_%.3f_g%d_%d_%d.npy` (:77).

Third-party pieces restated here (parity UNPINNED: the originals are absent from the reference tree and this image):
  * nbodykit `cosmology.Cosmology(h=).match(Omega0_m=)` + `transform.SkyToCartesian / CartesianToSky`: flat LCDM
    comoving distance in Mpc/h with photons (T_cmb = 2.7255 K) and N_ur = 2.0328 massless neutrinos; the single 0.06 eV
    neutrino of nbodykit's default is counted inside Omega0_m (it is non-relativistic over the survey's redshifts);
    expected agreement with CLASS distances ~1e-4.  x = chi cos(dec) cos(ra), y = chi cos(dec) sin(ra), z = chi sin(dec).
  * mangle polygon files (`.ply` / `.pol`): polygons as lists of caps (x, y, z, cm); a point is inside a polygon iff
    1 - r.n < cm for every cap with cm > 0 and 1 - r.n > |cm| for cm < 0.  `get_polyids` tests all polygons (chunked);
    the pixelised lookup of mangle (`pixelization` header) is NOT used as an accelerator -- large masks are slow, not wrong.
"""
from __future__ import annotations

import os

import numpy as np

C_KMS = 299792.458


# ----------------------------------------------------------------------------- cosmology
class Cosmology:
    """Flat LCDM background; distances in Mpc/h.  `match(Omega0_m=)` keeps the nbodykit call shape."""

    def __init__(self, h=0.6774, Omega0_m=0.3089, T0_cmb=2.7255, N_ur=2.0328):
        self.h, self.Omega0_m = float(h), float(Omega0_m)
        og = 2.472979e-5 * (T0_cmb / 2.7255) ** 4 / self.h ** 2          # photons
        self.Omega0_r = og * (1.0 + N_ur * (7.0 / 8.0) * (4.0 / 11.0) ** (4.0 / 3.0))
        self._tab = None

    def match(self, Omega0_m):
        return Cosmology(self.h, Omega0_m)

    def efunc(self, z):
        z = np.asarray(z, dtype=np.float64)
        ol = 1.0 - self.Omega0_m - self.Omega0_r
        return np.sqrt(self.Omega0_r * (1 + z) ** 4 + self.Omega0_m * (1 + z) ** 3 + ol)

    def _table(self, zmax):
        if self._tab is None or self._tab[0][-1] < zmax:
            zt = np.linspace(0.0, max(zmax * 1.05, 2.0), 20001)
            f = 1.0 / self.efunc(zt)
            # composite Simpson on the uniform table, cumulative at even nodes, cubic-accurate fill of odd nodes
            dz = zt[1] - zt[0]
            chi = np.zeros_like(zt)
            chi[2::2] = np.cumsum(dz / 3.0 * (f[0:-2:2] + 4.0 * f[1:-1:2] + f[2::2]))
            chi[1::2] = chi[0:-1:2] + dz / 12.0 * (5.0 * f[0:-1:2] + 8.0 * f[1::2] - f[2::2])
            self._tab = (zt, chi * C_KMS / 100.0)
        return self._tab

    def comoving_distance(self, z):
        z = np.asarray(z, dtype=np.float64)
        zt, chi = self._table(float(np.max(z)) if z.size else 1.0)
        return np.interp(z, zt, chi)

    def redshift_at_distance(self, chi):
        chi = np.asarray(chi, dtype=np.float64)
        zmax = 2.0
        while self.comoving_distance(zmax) < (np.max(chi) if chi.size else 0.0):
            zmax *= 2.0
        zt, ct = self._table(zmax)
        return np.interp(chi, ct, zt)


def sky_to_cartesian(ra, dec, z, cosmo):
    """nbodykit transform.SkyToCartesian (degrees in, Mpc/h out) -- src/opt_utilities.py:67,128."""
    ra, dec = np.deg2rad(np.asarray(ra, float)), np.deg2rad(np.asarray(dec, float))
    chi = cosmo.comoving_distance(z)
    return np.stack([chi * np.cos(dec) * np.cos(ra), chi * np.cos(dec) * np.sin(ra), chi * np.sin(dec)], axis=-1)


def cartesian_to_sky(pos, cosmo):
    """nbodykit transform.CartesianToSky: (ra in [0,360), dec, z) -- generate_mask.py:103."""
    pos = np.asarray(pos, float)
    r = np.sqrt((pos ** 2).sum(-1))
    dec = np.rad2deg(np.arcsin(np.clip(pos[..., 2] / np.where(r > 0, r, 1.0), -1.0, 1.0)))
    ra = np.rad2deg(np.arctan2(pos[..., 1], pos[..., 0])) % 360.0
    return ra, dec, cosmo.redshift_at_distance(r)


# ----------------------------------------------------------------------------- file readers
_TFORM = {"L": "i1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8", "A": "S"}


def read_fits_table(path, ext=1):
    """Minimal FITS BINTABLE reader (what `FITSCatalog` is used for at src/opt_utilities.py:46,106): scalar and fixed-length
    vector columns of types L B I J K E D A.  Returns {column name: ndarray}."""
    with open(path, "rb") as f:
        raw = f.read()
    off, hdu = 0, 0
    while off < len(raw):
        cards = {}
        while True:
            block = raw[off:off + 2880]
            if len(block) < 2880:
                raise Exception("truncated FITS header in '%s'" % path)
            off += 2880
            end = False
            for i in range(36):
                card = block[80 * i:80 * (i + 1)].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    end = True
                    break
                if card[8:10] == "= ":
                    val = card[10:].split("/")[0].strip()
                    if val.startswith("'"):
                        val = card[10:].strip()
                        val = val[1:val.index("'", 1)].strip()
                    cards[key] = val
            if end:
                break
        naxis = int(cards.get("NAXIS", 0))
        size = abs(int(cards.get("BITPIX", 8))) // 8
        for i in range(naxis):
            size *= int(cards["NAXIS%d" % (i + 1)])
        size = 0 if naxis == 0 else size + int(cards.get("PCOUNT", 0))
        if hdu == ext:
            if cards.get("XTENSION", "").strip() != "BINTABLE":
                raise Exception("HDU %d of '%s' is not a binary table" % (ext, path))
            nrow, rowlen = int(cards["NAXIS2"]), int(cards["NAXIS1"])
            fields = []
            for i in range(1, int(cards["TFIELDS"]) + 1):
                tf = cards["TFORM%d" % i].strip()
                j = 0
                while j < len(tf) and tf[j].isdigit():
                    j += 1
                rep, code = int(tf[:j] or 1), tf[j]
                if code not in _TFORM:
                    raise Exception("unsupported FITS column format '%s'" % tf)
                name = cards.get("TTYPE%d" % i, "COL%d" % i).strip()
                if code == "A":
                    fields.append((name, "S%d" % rep))
                elif rep == 1:
                    fields.append((name, _TFORM[code]))
                else:
                    fields.append((name, _TFORM[code], (rep,)))
            dt = np.dtype(fields)
            if dt.itemsize != rowlen:
                raise Exception("FITS row length mismatch in '%s' (%d != %d)" % (path, dt.itemsize, rowlen))
            tab = np.frombuffer(raw, dtype=dt, count=nrow, offset=off)
            return {n: np.ascontiguousarray(tab[n]).astype(tab[n].dtype.newbyteorder("="), copy=False) for n in dt.names}
        off += (size + 2879) // 2880 * 2880
        hdu += 1
    raise Exception("no HDU %d in '%s'" % (ext, path))


def read_catalog(path, columns=None):
    """{column: array} from a .fits table, a Cartesian .npz (pos, w[, nbar]) or a whitespace-separated text file whose
    column names come from the .param file (`data_columns` / `randoms_columns`, like nbodykit's CSVCatalog)."""
    if not os.path.exists(path):
        raise Exception("Catalogue file '%s' does not exist!" % path)
    ext = path.split(".")[-1]
    if ext == "fits":
        return read_fits_table(path)
    if ext == "npz":
        z = np.load(path)
        return {k: z[k] for k in z.files}
    if not columns:
        raise Exception("column names are needed for the text catalogue '%s'" % path)
    arr = np.loadtxt(path, ndmin=2)
    if arr.shape[1] < len(columns):
        raise Exception("'%s' has %d columns, %d names given" % (path, arr.shape[1], len(columns)))
    return {c: arr[:, i] for i, c in enumerate(columns)}


def apply_reference_columns(cat, z_min, z_max, cosmo, is_data, sim_no=-1, fkp_weights=False, P_fkp=1e4):
    """The column logic of load_data (src/opt_utilities.py:52-82) / load_randoms (:111-143) on a dict of arrays."""
    valid = (cat["Z"] > z_min) & (cat["Z"] < z_max)
    cat = {k: v[valid] for k, v in cat.items()}
    if "WEIGHT" not in cat:
        if is_data:
            if sim_no == -1:
                cat["WEIGHT"] = cat["WEIGHT_SYSTOT"] * (cat["WEIGHT_NOZ"] + cat["WEIGHT_CP"] - 1.0)
            else:
                cat["WEIGHT"] = 1.0 * cat["VETO FLAG"] * cat["FIBER COLLISION"]
        elif "VETO FLAG" in cat:
            cat["WEIGHT"] = 1.0 * cat["VETO FLAG"] * cat["FIBER COLLISION"]
        else:
            cat["WEIGHT"] = 1.0 * cat["Weight"]
    cat["Position"] = sky_to_cartesian(cat["RA"], cat["DEC"], cat["Z"], cosmo)
    if "WEIGHT_FKP" in cat:
        cat["NBAR"] = (1.0 / cat["WEIGHT_FKP"] - 1.0) / P_fkp
    if fkp_weights:
        if "NBAR" in cat:
            cat["WEIGHT_FKP"] = 1.0 / (1.0 + P_fkp * cat["NBAR"])
    else:
        cat["WEIGHT_FKP"] = np.ones(len(cat["NBAR"]))     # KeyError without NBAR, like the reference (:81,:142)
    return cat


# ----------------------------------------------------------------------------- mangle polygons
class MangleMask:
    """Polygon mask in mangle's `polygon` text format: `polygon <id> ( <n> caps, <w> weight, ... ):` followed by n lines
    `x y z cm`.  `weights[get_polyids(ra, dec)]` is what generate_mask.py:129-131 uses; points outside get id -1."""

    def __init__(self, path):
        self.caps, self.weights, self.ids = [], [], []
        with open(path) as f:
            lines = f.read().split("\n")
        i = 0
        while i < len(lines):
            t = lines[i].split()
            if t and t[0] == "polygon":
                head = lines[i]
                ncap = int(head[head.index("(") + 1:].split()[0])
                w = float(head.split("caps,")[1].split("weight")[0])
                caps = np.array([[float(v) for v in lines[i + 1 + j].split()[:4]] for j in range(ncap)]).reshape(ncap, 4)
                self.caps.append(caps)
                self.weights.append(w)
                self.ids.append(int(t[1]))
                i += ncap
            i += 1
        self.weights = np.asarray(self.weights, dtype=np.float64)
        if len(self.caps) == 0:
            raise Exception("no polygons found in mask file '%s'" % path)

    def get_polyids(self, ra, dec, chunk=1 << 20):
        ra, dec = np.deg2rad(np.ravel(ra)), np.deg2rad(np.ravel(dec))
        out = np.full(ra.shape, -1, dtype=np.int64)
        for s in range(0, ra.size, chunk):
            sl = slice(s, s + chunk)
            r = np.stack([np.cos(dec[sl]) * np.cos(ra[sl]), np.cos(dec[sl]) * np.sin(ra[sl]), np.sin(dec[sl])], axis=1)
            todo = np.ones(r.shape[0], dtype=bool)
            for p, caps in enumerate(self.caps):
                if not todo.any():
                    break
                inside = todo.copy()
                for x, y, z, cm in caps:
                    cd = 1.0 - (r[:, 0] * x + r[:, 1] * y + r[:, 2] * z)
                    inside &= (cd < cm) if cm >= 0 else (cd > -cm)
                idx = np.nonzero(inside)[0]
                out[s + idx] = p
                todo[idx] = False
        return out


# ----------------------------------------------------------------------------- n-bar(r)
def nbar_file_name(outdir, string, zmin, zmax, grid):
    return outdir + "nbar_%sz%.3f_%.3f_g%d_%d_%d.npy" % (string, zmin, zmax, grid[0], grid[1], grid[2])


def build_nbar_mask(r_axes, box, grid, random_z, random_w, zmin, zmax, cosmo, mask=None, chunk_planes=8):
    """generate_mask.py:96-146.  `r_axes`: the three 1-D cell-centre coordinate axes of load_coord_grids
    (src/opt_utilities.py:428-430); `mask`: a MangleMask, a callable (ra, dec) -> angular weight, or None (full sky)."""
    from scipy.interpolate import UnivariateSpline
    grid = np.asarray(grid, int)
    box = np.asarray(box, float)
    nz, zed = np.histogram(random_z, bins=100, weights=random_w, range=[zmin, zmax])
    z_av = 0.5 * (zed[:-1] + zed[1:])
    rz = cosmo.comoving_distance(zed)
    volz = 4.0 * np.pi / 3.0 * (rz[1:] ** 3.0 - rz[:-1] ** 3.0)
    nbar_z = UnivariateSpline(z_av, nz / volz, s=0.0000001, ext="zeros")
    out = np.empty(tuple(grid), dtype=np.float64)
    X, Y, Z = (np.asarray(a, float) for a in r_axes)
    for i0 in range(0, grid[0], chunk_planes):
        xs = X[i0:i0 + chunk_planes]
        pos = np.stack(np.meshgrid(xs, Y, Z, indexing="ij"), axis=-1).reshape(-1, 3)
        ra, dec, zz = cartesian_to_sky(pos, cosmo)
        n = nbar_z(zz)
        if mask is not None:
            if isinstance(mask, MangleMask):
                ids = mask.get_polyids(ra, dec)
                ang = mask.weights[ids]
                ang[ids == -1] = 0.0
            else:
                ang = np.asarray(mask(ra, dec), dtype=np.float64)
            n = n * ang
        out[i0:i0 + chunk_planes] = n.reshape(len(xs), grid[1], grid[2])
    v_cell = 1.0 * grid.prod() / (1.0 * box.prod())          # generate_mask.py:140 (inverse cell volume, as written there)
    rescale = np.sum(random_w) * v_cell / np.sum(out)
    return out * rescale
