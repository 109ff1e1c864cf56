"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

A numpy (FP64, complex128 C2C) restatement of the reference's algorithm for the
Monte-Carlo Fisher / quadratic + cubic estimator path.  It follows the reference's
*map-based real-space* formulation (every derivative map is materialised and the
Fisher matrix is a double loop of full-map dot products), i.e. it is deliberately NOT
the Fourier-space shell-binned formulation the CUDA path uses, so the two are
independent.  Each function cites the reference file:line it restates.

Pinned against the UNMODIFIED reference scripts run under oracle/ref_harness.py
(FP64-promoted mode) through the golden vectors in tests/golden/ (generator:
oracle/make_golden.py; check: tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.fft

WORKERS = os.cpu_count()


# ----------------------------------------------------------------------------- FFTs
def ft(pix):
    """Forward unnormalised 3-D C2C DFT (src/opt_utilities.py:502-528), in complex128."""
    return scipy.fft.fftn(np.asarray(pix, dtype=np.complex128), workers=WORKERS)


def ift(pix):
    """Inverse 3-D C2C DFT normalised by 1/N (src/opt_utilities.py:530-555; pyfftw default)."""
    return scipy.fft.ifftn(np.asarray(pix, dtype=np.complex128), workers=WORKERS)


# ----------------------------------------------------------------------------- Y_lm
_YLM_CACHE = {}


def compute_spherical_harmonic_functions(lmax):
    """Real Y_lm for even l<=lmax as Cartesian polynomials built with sympy exactly like
    src/opt_utilities.py:323-398 (lambdified with numpy instead of numexpr).  The
    lambdified source carries 15-significant-digit coefficients; that is what the
    reference evaluates, and what tools/gen_ylm.py bakes into the CUDA table."""
    if lmax in _YLM_CACHE:
        return _YLM_CACHE[lmax]
    import sympy as sp

    def real_ylm(l, m):
        x, y, z, r = sp.symbols("x y z r", real=True, positive=True)
        xh, yh, zh = sp.symbols("xhat yhat zhat", real=True, positive=True)
        phi, theta = sp.symbols("phi theta")
        subs = [(sp.sin(phi), y / sp.sqrt(x ** 2 + y ** 2)),
                (sp.cos(phi), x / sp.sqrt(x ** 2 + y ** 2)),
                (sp.cos(theta), z / sp.sqrt(x ** 2 + y ** 2 + z ** 2))]
        if m == 0:
            amp = sp.sqrt((2 * l + 1) / (4 * np.pi))
        else:
            amp = sp.sqrt(2 * (2 * l + 1) / (4 * np.pi) * sp.factorial(l - abs(m)) / sp.factorial(l + abs(m)))
        e = (-1) ** m * sp.assoc_legendre(l, abs(m), sp.cos(theta))
        if m < 0:
            e *= sp.expand_trig(sp.sin(abs(m) * phi))
        elif m > 0:
            e *= sp.expand_trig(sp.cos(m * phi))
        e = sp.together(e.subs(subs)).subs(x ** 2 + y ** 2 + z ** 2, r ** 2)
        e = amp * e.expand().subs([(x / r, xh), (y / r, yh), (z / r, zh)])
        f = sp.lambdify((xh, yh, zh), e, "numpy")
        f.expr, f.l, f.m = e, l, m
        return f

    out = [[real_ylm(int(l), int(m)) for m in range(-l, l + 1)] for l in range(0, lmax + 1, 2)]
    _YLM_CACHE[lmax] = out
    return out


# ----------------------------------------------------------------------------- grids
def _signed_index(n):
    """src/opt_utilities.py:422: kkk-grid if kkk>=grid/2 else kkk."""
    mid = n / 2
    return np.asarray([i - n if i >= mid else i for i in range(n)])


def load_coord_grids(box, grid, box_center):
    """k_grids[3,nx,ny,nz] and r_grids[3,nx,ny,nz] (src/opt_utilities.py:400-431).  The real-space
    axes are the fft-ordered mesh coordinate + BoxCenter + half a cell (:428-430)."""
    box = np.asarray(box, float)
    grid = np.asarray(grid, int)
    kF = 2.0 * np.pi / (1.0 * box)
    kax = [_signed_index(grid[i]) * kF[i] for i in range(3)]
    k3 = np.meshgrid(*kax, indexing="ij")
    off = np.asarray(box_center, float) + 0.5 * box / grid
    rax = [(_signed_index(grid[i]) * (box[i] / grid[i])).astype("f8") + off[i] for i in range(3)]
    r3 = np.meshgrid(*rax, indexing="ij")
    return np.asarray(k3), np.asarray(r3)


def load_MAS(box, grid):
    """Separable TSC compensation window with the float32 pi/N prefactor
    (src/opt_utilities.py:433-465; the float32 rounding at :451 matters at the 1e-8 level)."""
    grid = np.asarray(grid, int)
    box = np.asarray(box, float)
    kF = 2.0 * np.pi / (1.0 * box)
    pref = np.pi / np.asarray(grid, dtype=np.float32)
    one = []
    for i in range(3):
        k = _signed_index(grid[i]) * kF[i]
        s = np.sin(pref[i] * (k / kF[i])) ** 2.0
        one.append(1.0 / (1.0 - s + 2.0 / 15 * s ** 2.0) ** 0.5)
    return one[0][:, None, None] * one[1][None, :, None] * one[2][None, None, :]


def bin_edges(kmin, kmax, dk):
    """src/opt_utilities.py:485-487."""
    k_all = np.arange(kmin, kmax + dk, dk)
    return k_all[:-1], k_all[1:]


def compute_filters(kmin, kmax, dk):
    """Theta_i(k) = [k_lo[i] <= |k| < k_hi[i]]  (src/opt_utilities.py:467-488)."""
    lo, hi = bin_edges(kmin, kmax, dk)
    return lambda i, k_norm: np.logical_and(k_norm >= lo[i], k_norm < hi[i])


# ----------------------------------------------------------------------------- operators
def applyCinv_fkp(x, nbar, v_cell, shot_fac, P_fkp=1e4, MAS=None):
    """FKP C^-1 (src/covariances_pk.py:126-164), complex128 buffer.  MAS given = include_pix=True
    branch (:155-156, :161-162)."""
    w = nbar * (nbar * P_fkp + shot_fac)
    out = np.zeros(x.shape, dtype=np.complex128)
    tmp = ift(ft(x) * MAS) if MAS is not None else x
    f = nbar > 0
    out[f] = tmp[f] / w[f]
    if MAS is not None:
        return ift(MAS * ft(out)) * v_cell
    return out * v_cell


def applyN(x, nbar, v_cell, shot_fac, MAS=None):
    """src/covariances_pk.py:97-124 (MAS given = include_pix=True branch, :122)."""
    if MAS is not None:
        return shot_fac * ift(1.0 / MAS * ft(nbar * ift(ft(x) / MAS))) / v_cell
    return shot_fac * nbar * x / v_cell


class Setup:
    """Everything the four scripts build between 'LOAD DATA' and the estimator proper
    (pk/compute_pk_randoms.py:135-202; identical blocks in the other three scripts)."""

    def __init__(self, box, grid, box_center, nbar_r, alpha, shot_fac, kmin, kmax, dk, lmax, include_pix=False,
                 nbar_override=None):
        self.include_pix = include_pix
        self.box = np.asarray(box, float)
        self.grid = np.asarray(grid, int)
        self.v_cell = 1.0 * self.box.prod() / (1.0 * self.grid.prod())
        self.shot_fac = shot_fac
        self.kmin, self.kmax, self.dk, self.lmax = kmin, kmax, dk, lmax
        self.n_k = int((kmax - kmin) / dk)
        self.n_l = lmax // 2 + 1
        # nbar_A: randoms-density map with empty cells filled (pk/compute_pk_randoms.py:135-136)
        nbar_A = nbar_r * 1.0
        nbar_A[nbar_A == 0] = min(nbar_A[nbar_A != 0]) * 0.01
        self.nbar_A = nbar_A
        self.nbar = nbar_r * alpha            # load_nbar(config, ., alpha_ran), opt_utilities.py:180
        self.nbar_weight = self.nbar.copy()
        if nbar_override is not None:         # rand_nbar=True: n from painted randoms (pk/compute_pk_randoms.py:174-176)
            self.nbar = nbar_override.copy()
        k_grids, r_grids = load_coord_grids(box, grid, box_center)
        self.k_norm = np.sqrt(np.sum(k_grids ** 2.0, axis=0))
        self.k_hat = k_grids / (1e-12 + self.k_norm)
        self.r_hat = r_grids / (1e-12 + np.sqrt(np.sum(r_grids ** 2.0, axis=0)))
        self.MAS = load_MAS(box, grid)
        self.pix = self.MAS if include_pix else None      # passed to the operators as their MAS argument
        self.Y = compute_spherical_harmonic_functions(lmax)
        self.filt = compute_filters(kmin, kmax, dk)

    # -- src/covariances_pk.py:376-389
    def ylm_sum(self, x_times_n, l_i, mas_power=0):
        f = 0.0
        for Y in self.Y[l_i]:
            f = f + ft(x_times_n * Y(*self.r_hat)) * Y(*self.k_hat)
        if mas_power:
            f = f * self.MAS ** float(mas_power)
        return f

    def applyC_alpha(self, x, data=False):
        """All n_b derivative maps C_,alpha x (src/covariances_pk.py:329-401, include_pix=False)."""
        out = []
        pix = self.include_pix
        tmp = ift(ft(x) / self.MAS) * self.nbar if pix else x * self.nbar          # :371-374
        for l_i in range(self.n_l):
            f = self.ylm_sum(tmp, l_i, 2 if (data and not pix) else 0)             # :386-389
            for a in range(self.n_k):
                t2 = 4.0 * np.pi / (4.0 * l_i + 1.0) * self.filt(a, self.k_norm) * f
                if pix:
                    out.append(ift(ft(self.nbar * ift(t2)) / self.MAS) / self.v_cell)  # :396-397
                else:
                    out.append(self.nbar * ift(t2) / self.v_cell)
        return out

    def grf(self, seed):
        """pk/compute_pk_randoms.py:139-140 -- numpy legacy stream."""
        np.random.seed(seed)
        return np.random.normal(loc=0.0, scale=np.sqrt(self.nbar_A), size=self.nbar_A.shape)


# ----------------------------------------------------------------------------- pk
def pk_fisher_realisation(S: Setup, a):
    """q-bar_alpha [n_b] and symmetrised F_ab [n_b,n_b] of one Gaussian realisation `a`
    (pk/compute_pk_randoms.py:206-281, the non-spilling flow :229-241, :264-274)."""
    Cinv_a = np.real(applyCinv_fkp(a, S.nbar_weight, S.v_cell, S.shot_fac, MAS=S.pix))
    Ainv_a = a / S.nbar_A
    N_Ainv_a = applyN(Ainv_a, S.nbar, S.v_cell, S.shot_fac, MAS=S.pix)
    C_a_Cinv = S.applyC_alpha(Cinv_a)
    C_a_Ainv = S.applyC_alpha(Ainv_a)
    nb = len(C_a_Cinv)
    U = [applyCinv_fkp(C_a_Cinv[i], S.nbar_weight, S.v_cell, S.shot_fac, MAS=S.pix) for i in range(nb)]
    fish = np.zeros((nb, nb))
    qbar = np.zeros(nb)
    if S.include_pix:      # complex maps with O(1e-17) imaginary parts: the reference sums complex products (:267-274)
        V = np.asarray([v.ravel() for v in C_a_Ainv])
        for i in range(nb):
            qbar[i] = 0.5 * np.real(np.sum(U[i] * N_Ainv_a))
            fish[i] = 0.5 * np.real(V @ U[i].ravel())
        return qbar, 0.5 * (fish + fish.T)
    V = np.asarray([np.real(v).ravel() for v in C_a_Ainv])
    for i in range(nb):
        qbar[i] = 0.5 * np.real(np.sum(U[i] * N_Ainv_a))
        fish[i] = 0.5 * (V @ np.real(U[i]).ravel())
    return qbar, 0.5 * (fish + fish.T)


def pk_data_q(S: Setup, diff):
    """q_alpha of a painted data-minus-randoms map (pk/compute_pk_data.py:191-206)."""
    Cinv_d = applyCinv_fkp(diff, S.nbar_weight, S.v_cell, S.shot_fac, MAS=S.pix)
    C_a = S.applyC_alpha(Cinv_d, data=True)
    return np.asarray([np.real(np.sum(Cinv_d * c)) / 2.0 for c in C_a])


def pk_solve(fish, q, qbar):
    """pk/compute_pk_data.py:250."""
    return np.inner(np.linalg.inv(fish), q - qbar)


# ----------------------------------------------------------------------------- ML weights
def pk_map_from_table(S: Setup, table):
    """Fiducial multipoles on the mesh (pk/compute_pk_randoms.py:188-193): linear interpolation of the columns
    P_0, P_2, P_4 of the `fiducial_pk` table at |k| (scipy interp1d, default kind = linear)."""
    t = np.asarray(table, float)
    return np.asarray([np.interp(S.k_norm, t[:, 0], t[:, 1 + i]) for i in range(S.n_l)])


def applyS(S: Setup, x, pk_map):
    """Signal covariance S x (src/covariances_pk.py:45-95), complex128 throughout."""
    tmp = ift(ft(x) / S.MAS) * S.nbar if S.include_pix else x * S.nbar                 # :76-79
    f = 0.0
    for l_i in range(len(pk_map)):
        f = f + 4.0 * np.pi / (4.0 * l_i + 1.0) * S.ylm_sum(tmp, l_i) * pk_map[l_i]     # :82-89
    if S.include_pix:
        return ift(ft(S.nbar * ift(f)) / S.MAS) / S.v_cell                             # :92-93
    return S.nbar * ift(f) / S.v_cell


def applyC(S: Setup, x, pk_map):
    """C = S + N (src/covariances_pk.py:10-43)."""
    return applyS(S, x, pk_map) + applyN(x, S.nbar, S.v_cell, S.shot_fac, MAS=S.pix)


def applyCinv(S: Setup, pix, pk_map, max_it=100, abs_tol=1e-8, rel_tol=None, P_fkp=1e4, info=None):
    """Preconditioned conjugate gradients for C x = pix with the FKP inverse as first guess and preconditioner
    (src/covariances_pk.py:222-327).  The products are complex and bilinear (np.sum(r*pre_r), no conjugation), the
    stall test runs every 10 iterations, the relative test compares sqrt(new_sum/init_sum) with rel_tol."""
    pre = lambda m: applyCinv_fkp(m, S.nbar_weight, S.v_cell, S.shot_fac, P_fkp=P_fkp, MAS=S.pix)
    x = pre(pix)
    r = pix - applyC(S, x, pk_map)
    p = pre(r)
    C_p = applyC(S, p, pk_map)
    pre_r = pre(r)
    old_sum = np.sum(r * pre_r)
    init_sum = old_sum.copy()
    alpha = old_sum / np.sum(p * C_p)
    save_sum = old_sum.copy()
    n_it = 0
    for i in range(max_it):
        if i % 10 == 0 and i > 0:
            if np.abs((save_sum - old_sum) / old_sum) < 0.05:
                break
            save_sum = old_sum
        x = x + alpha * p
        r = r - alpha * C_p
        pre_r = pre(r)
        new_sum = np.sum(r * pre_r)
        p = pre_r + (new_sum / old_sum) * p
        n_it = i + 1
        if rel_tol is not None:
            if np.sqrt(new_sum / init_sum) < rel_tol:
                break
        elif np.sqrt(new_sum) < abs_tol:
            break
        old_sum = new_sum
        C_p = applyC(S, p, pk_map)
        alpha = old_sum / np.sum(p * C_p)
    if info is not None:
        info["iterations"] = n_it
    return x


def pk_data_q_ml(S: Setup, diff, pk_map):
    """q_alpha with ML weights (pk/compute_pk_data.py:194-206)."""
    Cinv_d = applyCinv(S, diff, pk_map, rel_tol=1e-6, max_it=30)
    C_a = S.applyC_alpha(Cinv_d, data=True)
    return np.asarray([np.real(np.sum(Cinv_d * c)) / 2.0 for c in C_a])


def pk_fisher_realisation_ml(S: Setup, a, pk_map):
    """q-bar_alpha and F_ab with ML weights: pk/compute_pk_randoms.py:205-281 along its non-low_mem branch
    (:229-241, :247-281).  The reference script itself cannot run this configuration (undefined name at :261,
    double `del` at :278), so the golden values come from its own functions composed the same way."""
    Cinv_a = np.real(applyCinv(S, a, pk_map, rel_tol=1e-6, max_it=30))
    Ainv_a = a / S.nbar_A
    N_Ainv_a = applyN(Ainv_a, S.nbar, S.v_cell, S.shot_fac, MAS=S.pix)
    C_a_Cinv = S.applyC_alpha(Cinv_a)
    C_a_Ainv = S.applyC_alpha(Ainv_a)
    nb = len(C_a_Cinv)
    fish = np.zeros((nb, nb))
    qbar = np.zeros(nb)
    V = np.asarray([v.ravel() for v in C_a_Ainv])
    for i in range(nb):
        U = applyCinv(S, C_a_Cinv[i], pk_map, rel_tol=1e-6, max_it=30)
        qbar[i] = 0.5 * np.real(np.sum(U * N_Ainv_a))
        fish[i] = 0.5 * np.real(V @ U.ravel())
    return qbar, 0.5 * (fish + fish.T)


# ----------------------------------------------------------------------------- painting
def tsc_paint(pos, w, box, grid, box_center):
    """TSC deposit without compensation/interlacing; restates FKPCatalog.to_mesh(window='tsc')
    as used at src/opt_utilities.py:219-260.  pmesh itself is not available: parity of the
    painter is 'unpinned' (our definition -- see DESIGN.md)."""
    grid = np.asarray(grid, int)
    H = np.asarray(box, float) / grid
    s = (np.asarray(pos, float) - np.asarray(box_center, float)) / H
    c = np.rint(s)
    d = s - c
    c = c.astype(np.int64)
    wt = [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]
    out = np.zeros(int(grid.prod()))
    for ix in range(3):
        i = (c[:, 0] + ix - 1) % grid[0]
        for iy in range(3):
            j = (c[:, 1] + iy - 1) % grid[1]
            for iz in range(3):
                k = (c[:, 2] + iz - 1) % grid[2]
                out += np.bincount((i * grid[1] + j) * grid[2] + k,
                                   weights=w * wt[ix][:, 0] * wt[iy][:, 1] * wt[iz][:, 2],
                                   minlength=out.size)
    return out.reshape(grid)


def cic_paint(pos, w, box, grid, box_center):
    """CIC deposit (window='cic' of the same to_mesh call): base cell floor((x - BoxCenter)/H), weights 1-d and d."""
    grid = np.asarray(grid, int)
    H = np.asarray(box, float) / grid
    s = (np.asarray(pos, float) - np.asarray(box_center, float)) / H
    c = np.floor(s)
    d = s - c
    c = c.astype(np.int64)
    wt = [1.0 - d, d]
    out = np.zeros(int(grid.prod()))
    for ix in range(2):
        i = (c[:, 0] + ix) % grid[0]
        for iy in range(2):
            j = (c[:, 1] + iy) % grid[1]
            for iz in range(2):
                k = (c[:, 2] + iz) % grid[2]
                out += np.bincount((i * grid[1] + j) * grid[2] + k,
                                   weights=w * wt[ix][:, 0] * wt[iy][:, 1] * wt[iz][:, 2], minlength=out.size)
    return out.reshape(grid)


def grid_data(dpos, dw, rpos, rw, box, grid, box_center):
    """(n_g - alpha n_r)/V_cell and alpha n_r/V_cell (src/opt_utilities.py:228,257-267)."""
    alpha = dw.sum() / rw.sum()
    v_cell = np.prod(np.asarray(box, float)) / np.prod(np.asarray(grid, int))
    ng = tsc_paint(dpos, dw, box, grid, box_center)
    nr = tsc_paint(rpos, rw, box, grid, box_center)
    return (ng - alpha * nr) / v_cell, alpha * nr / v_cell


# ----------------------------------------------------------------------------- bk
def triangle_bins(kmin, kmax, dk, n_k):
    """Allowed (a<=b<=c) bin triplets (bk/compute_bk_randoms.py:186-211)."""
    k_lo = np.arange(kmin, kmax, dk)
    k_hi = k_lo + dk
    tol = 1e-6
    out = []
    for a in range(n_k):
        for b in range(a, n_k):
            for c in range(b, n_k):
                k_up = k_hi[a] + k_hi[b]
                k_do = 0.0 if a == b else k_lo[b] - k_hi[a]
                if k_lo[c] + tol >= k_up or k_hi[c] - tol <= k_do:
                    continue
                out.append((a, b, c))
    return out


def compute_g_a(S: Setup, a_map, data=False):
    """g_i^l and tilde-g_i^l maps (bk/compute_bk_randoms.py:222-267).  With data=True only g is
    returned and the extra MAS factor of bk/compute_bk_data.py:259-283 is applied."""
    Cinv = applyCinv_fkp(a_map, S.nbar_weight, S.v_cell, S.shot_fac, MAS=S.pix)
    pix = S.include_pix
    nC = ift(ft(Cinv) / S.MAS) * S.nbar if pix else S.nbar * Cinv                      # :235-240 / data :267-270
    g, tg = [], []
    if not data:
        Ainv = a_map / S.nbar_A
        nA = ift(ft(Ainv) / S.MAS) * S.nbar if pix else S.nbar * Ainv
    for l_i in range(S.n_l):
        fC = S.ylm_sum(nC, l_i, 1 if (data and not pix) else 0) * (4.0 * np.pi / (4.0 * l_i + 1.0))
        g.append([ift(S.filt(i, S.k_norm) * fC) for i in range(S.n_k)])
        if not data:
            fA = S.ylm_sum(nA, l_i) * (4.0 * np.pi / (4.0 * l_i + 1.0))
            tg.append([ift(S.filt(i, S.k_norm) * fA) for i in range(S.n_k)])
    return (g, tg) if not data else g


def _phi1(S, g, a, b, c, l_i):
    """n IFT[Theta_c FT[g_a^0 g_b^l]]/v_cell (bk/compute_bk_randoms.py:278-311)."""
    gabc = ift(ft(g[0][a] * g[l_i][b]) * S.filt(c, S.k_norm))
    if S.include_pix:
        return np.real(ift(ft(gabc * S.nbar) / S.MAS) / S.v_cell)                      # :304-306
    return np.real(gabc * S.nbar / S.v_cell)


def _phi2(S, g, a, b, c, l_i):
    """4pi/(2l+1) sum_m n Y_lm(r) IFT[Y_lm(k) Theta_c FT[g_a^0 g_b^0]]/v_cell (:313-356)."""
    fg = ft(g[0][a] * g[0][b]) * S.filt(c, S.k_norm)
    out = 0.0
    for Y in S.Y[l_i]:
        if S.include_pix:                                                               # :344-346
            out = out + np.real(ift(ft(ift(Y(*S.k_hat) * fg) * S.nbar) / S.MAS) / S.v_cell * Y(*S.r_hat))
        else:
            out = out + np.real(ift(Y(*S.k_hat) * fg) * (Y(*S.r_hat) * S.nbar / S.v_cell))
    return out * (4.0 * np.pi / (4.0 * l_i + 1.0))


def phi_alpha(S, g, a, b, c, l_i):
    """Symmetrised phi_alpha (bk/compute_bk_randoms.py:363-403)."""
    p_abc = _phi1(S, g, a, b, c, l_i) if l_i == 0 else _phi2(S, g, a, b, c, l_i)
    p_acb = p_abc if (b == c and l_i == 0) else _phi1(S, g, a, c, b, l_i)
    if a == b:
        p_bca = p_acb
    elif c == a and l_i == 0:
        p_bca = p_abc
    else:
        p_bca = _phi1(S, g, b, c, a, l_i)
    return p_abc + p_acb + p_bca


def delta_abc(S: Setup, tris):
    """Degeneracy factors Delta_abc (bk/compute_bk_randoms.py:427-473)."""
    nt = len(tris)
    D = np.zeros(nt * S.n_l)
    for i, (a, b, c) in enumerate(tris):
        D[i] = 6.0 if (a == b and b == c) else (2.0 if (a == b or a == c or b == c) else 1.0)
    if S.n_l > 1:
        bins = [ift(S.filt(a, S.k_norm)) for a in range(S.n_k)]
        for l_i in range(1, S.n_l):
            for i, (a, b, c) in enumerate(tris):
                j = i + l_i * nt
                if a != b and b != c:
                    D[j] = 1.0
                elif a == b and b != c:
                    D[j] = 2.0
                else:
                    Dl = [ift(S.filt(c, S.k_norm) * Y(*S.k_hat)) for Y in S.Y[l_i]]
                    Nl = np.sum([np.sum(bins[a] * d * d).real for d in Dl]) * 4.0 * np.pi / (4.0 * l_i + 1.0)
                    if a != b:
                        N0 = np.sum(bins[a] * bins[c] ** 2.0).real
                        D[j] = 1.0 + Nl / N0
                    else:
                        N0 = np.sum(bins[c] ** 3.0).real
                        D[j] = 2.0 * (1.0 + 2.0 * Nl / N0)
    return D


def bk_fisher_pair(S: Setup, aA, aB, return_g=False):
    """Bispectrum Fisher contribution of one pair of realisations
    (bk/compute_bk_randoms.py:269-517)."""
    tris = triangle_bins(S.kmin, S.kmax, S.dk, S.n_k)
    nt = len(tris)
    nb = nt * S.n_l
    gA, tgA = compute_g_a(S, aA)
    gB, tgB = compute_g_a(S, aB)
    tphi = {"A": [], "B": []}
    cphi = {"A": [], "B": []}
    for X, g, tg in (("A", gA, tgA), ("B", gB, tgB)):
        for idx in range(nb):
            l_i, (a, b, c) = idx // nt, tris[idx % nt]
            tphi[X].append(phi_alpha(S, tg, a, b, c, l_i).ravel())
            p = phi_alpha(S, g, a, b, c, l_i)
            cphi[X].append(np.real(applyCinv_fkp(p, S.nbar_weight, S.v_cell, S.shot_fac, MAS=S.pix)).ravel())
    D = delta_abc(S, tris)
    F = np.zeros((nb, nb))
    TA, TB = np.asarray(tphi["A"]), np.asarray(tphi["B"])
    CA, CB = np.asarray(cphi["A"]), np.asarray(cphi["B"])
    full = (TA @ CA.T + TB @ CB.T - TA @ CB.T - TB @ CA.T) / 12.0 / 2.0
    iu = np.triu_indices(nb)
    F[iu] = full[iu]
    F *= 4.0 / np.outer(D, D)
    F = np.triu(F) + np.triu(F, 1).T
    if return_g:
        return F, (gA, tgA), (gB, tgB)
    return F


def bk_data_q(S: Setup, diff, mc_maps=None, N_mc=0):
    """Cubic-estimator numerator q_alpha/Delta (bk/compute_bk_data.py:250-283, :350-410).
    mc_maps: iterable of (g, tilde_g) per MC simulation for the 1-field term."""
    tris = triangle_bins(S.kmin, S.kmax, S.dk, S.n_k)
    nt = len(tris)
    g = compute_g_a(S, diff, data=True)
    q = np.zeros((nt, S.n_l))
    for i, (a, b, c) in enumerate(tris):
        for l_i in range(S.n_l):
            q[i, l_i] = np.real(np.sum(g[0][a] * g[0][b] * g[l_i][c]))
    if mc_maps is not None:
        for gm, tgm in mc_maps:
            for i, (a, b, c) in enumerate(tris):
                for l_i in range(S.n_l):
                    t1 = g[0][a] * gm[0][b] * tgm[l_i][c] + g[0][b] * gm[l_i][c] * tgm[0][a] + g[l_i][c] * gm[0][a] * tgm[0][b]
                    t2 = g[0][a] * tgm[0][b] * gm[l_i][c] + g[0][b] * tgm[l_i][c] * gm[0][a] + g[l_i][c] * tgm[0][a] * gm[0][b]
                    q[i, l_i] -= 0.5 * np.real(np.sum(t1)) / N_mc
                    q[i, l_i] -= 0.5 * np.real(np.sum(t2)) / N_mc
    D = delta_abc(S, tris)
    return np.concatenate([q[:, l_i] / D[l_i * nt:(l_i + 1) * nt] for l_i in range(S.n_l)])
