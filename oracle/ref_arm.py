"""TEST / BENCH INFRASTRUCTURE -- the reference's own CPU algorithm for one P_l(k) Fisher realisation, timed on the
host cores (bench.py `--impl reference` and the `cpu_baseline` leg; nothing in the product path imports this).

What runs is the `low_mem = True` branch of pk/compute_pk_randoms.py:200-281 -- the only branch the reference can
execute (the other one dies at :278) -- in the reference's shipped precision (complex64 transforms,
src/opt_utilities.py:502-555; complex64 C^-1 buffer, src/covariances_pk.py:154), with every full-map operation on ALL
host threads: torch CPU tensors stand in for pyfftw (MKL FFT, multi-threaded like FFTW with threads=N) and for numexpr
(multi-threaded element-wise evaluation of the Y_lm polynomials, src/opt_utilities.py:383).  pyfftw / numexpr cannot be
installed in this image; the torch CPU kernels are at least as fast, so the baseline is not handicapped.

One realisation of the reference costs (n_b = n_k (lmax/2+1) bins)

    prologue : GRF, C^-1 a, A^-1 a, N A^-1 a                                  (:139-140, :206-217)   once
    loop 1   : n_b x [ applyC_alpha_single(A^-1 a) ; np.save ]                 (:220-226)
    loop 2   : n_b x [ applyC_alpha_single(C^-1 a) ; applyCinv_fkp ; q-bar dot (:244-265)
                       ; n_b x ( np.load ; full-map product-sum ) ]            (:268-272)

and cannot be run in full at 512^3 (n_b^2 = 6e4 loads of 2 GB maps): a STEP of this arm is a bounded sample of exactly
that code path on the full-size mesh --

    alternately one loop-1 row and one loop-2 row at multipole l = 2 (2l+1 = 5 forward transforms = the mean over
    l = 0,2,4 of the rows of a realisation), the latter with `n_dots` of its np.load + product-sum terms against the saved
    loop-1 map (the prologue is timed once, outside the steps)

-- and the time of one realisation follows from the reference's own loop counts and the means over the steps:

    T_real = t_prologue + n_b (t_row1 + t_row2) + n_b^2 t_dot

`step_fraction` = (time of the step's sample) / T_real is the share of one realisation a step covers, so the line's
value (realisations/s) = step_fraction / step time.
"""
from __future__ import annotations

import os
import tempfile
import time

import numpy as np

from . import sww_oracle as orc


class PkFisherRefArm:
    def __init__(self, box, grid, box_center, nbar_r, alpha, shot_fac, k_min, k_max, dk, lmax, threads=None, n_dots=3,
                 tmpdir=None):
        import torch
        self.t = torch
        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.box, self.grid = np.asarray(box, float), np.asarray(grid, int)
        self.v_cell = float(self.box.prod() / self.grid.prod())
        self.lmax, self.n_l = int(lmax), int(lmax) // 2 + 1
        self.n_k = int((k_max - k_min) / dk)
        self.n_b = self.n_k * self.n_l
        self.shot, self.P_fkp = float(shot_fac), 1e4
        self.n_dots = int(n_dots)
        # geometry with the oracle's restatement of load_coord_grids (src/opt_utilities.py:400-431), unit vectors as in
        # pk/compute_pk_randoms.py:164-169
        k, r = orc.load_coord_grids(box, grid, box_center)
        k, r = torch.from_numpy(k), torch.from_numpy(r)
        self.k_norm = torch.sqrt((k ** 2.0).sum(0))
        k /= (1e-12 + self.k_norm)
        r /= (1e-12 + torch.sqrt((r ** 2.0).sum(0)))
        self.k_hat, self.r_hat = k, r
        self.Y = orc.compute_spherical_harmonic_functions(self.lmax)   # polynomials: evaluate on torch tensors as well
        lo, hi = orc.bin_edges(k_min, k_max, dk)
        self.k_lo, self.k_hi = lo, hi
        nbar_r = torch.as_tensor(np.asarray(nbar_r, dtype=np.float64))
        self.nbar = nbar_r * float(alpha)                     # load_nbar(..., alpha_ran)
        nA = nbar_r.clone()                                    # pk/compute_pk_randoms.py:135-136
        nA[nA == 0] = float(nA[nA != 0].min()) * 0.01
        self.nbar_A = nA
        self.tmpdir = tempfile.mkdtemp(prefix="sww_refarm_", dir=tmpdir)
        self._fn = None
        self.l_i = 1 if self.n_l > 1 else 0

    # ---- reference operators, shipped precision -------------------------------------------------------------
    def ft(self, x):
        """src/opt_utilities.py:502-528: complex64 in, complex64 out, unnormalised."""
        t = self.t
        return t.fft.fftn(x.to(t.complex64))

    def ift(self, x):
        t = self.t
        return t.fft.ifftn(x.to(t.complex64))

    def applyCinv_fkp(self, x):
        """src/covariances_pk.py:126-164 with include_pix = False."""
        t = self.t
        w = self.nbar * (self.nbar * self.P_fkp + self.shot)
        ratio = t.zeros(x.shape, dtype=t.complex64)
        f = self.nbar > 0
        ratio[f] = (x[f] / w[f]).to(t.complex64)
        return ratio * self.v_cell

    def applyN(self, x):
        """src/covariances_pk.py:97-124 with include_pix = False."""
        return self.shot * self.nbar * x / self.v_cell

    def applyC_alpha_single(self, x, i, a):
        """src/covariances_pk.py:404-466 with include_pix = False, data = False."""
        tmp = x * self.nbar
        f = 0.0
        for Ylm in self.Y[i]:
            yr, yk = Ylm(*self.r_hat), Ylm(*self.k_hat)       # scalars for l = 0, full maps otherwise
            f = f + self.ft(tmp * yr) * yk
        filt = (self.k_norm >= self.k_lo[a]) & (self.k_norm < self.k_hi[a])
        tmp2 = 4.0 * np.pi / (4.0 * i + 1.0) * filt * f
        return self.nbar * self.ift(tmp2) / self.v_cell

    # ---- bounded samples -------------------------------------------------------------------------------------
    def prologue(self):
        """GRF, C^-1 a, A^-1 a, N A^-1 a (pk/compute_pk_randoms.py:139-140, 206-217): once per realisation; timed once."""
        t = self.t
        T0 = time.perf_counter()
        np.random.seed(1)
        a = t.from_numpy(np.random.normal(loc=0.0, scale=1.0, size=tuple(self.grid))) * t.sqrt(self.nbar_A)
        self.Cinv = self.applyCinv_fkp(a).real.contiguous()
        self.Ainv = a / self.nbar_A
        self.NAinv = self.applyN(self.Ainv)
        self.t_prologue = time.perf_counter() - T0
        self.times = {"row1": [], "row2": [], "dot": []}
        self._n = 0

    def _row1(self, ka):
        """one row of loop 1 (:220-226): C_,alpha A^-1 a at l = 2, saved like the reference does"""
        T0 = time.perf_counter()
        m1 = self.applyC_alpha_single(self.Ainv, self.l_i, ka)
        fn = os.path.join(self.tmpdir, "C_a_Ai_%d.npy" % ka)
        np.save(fn, m1.numpy())
        self.times["row1"].append(time.perf_counter() - T0)
        return fn

    def _row2(self, ka, fn):
        """one row of loop 2 (:244-272): C^-1 C_,alpha C^-1 a, the q-bar dot, and n_dots of its n_b np.load + product-sums"""
        t = self.t
        T0 = time.perf_counter()
        tmp = self.applyC_alpha_single(self.Cinv, self.l_i, ka)
        row = self.applyCinv_fkp(tmp)
        del tmp
        qb = 0.5 * float(t.sum(row * self.NAinv).real)
        T1 = time.perf_counter()
        acc = 0.0
        for _ in range(self.n_dots):
            acc += 0.5 * float(t.sum(row * t.from_numpy(np.load(fn))).real)
        T2 = time.perf_counter()
        self.times["row2"].append(T1 - T0)
        self.times["dot"].append((T2 - T1) / max(1, self.n_dots))
        return qb, acc

    def step(self):
        """One bounded sample: even steps run a loop-1 row, odd steps the loop-2 row of the same bin with n_dots of its
        product-sums.  Returns (seconds of this step, T_real from the means so far -- None until both kinds of row have
        been sampled --, parts)."""
        if not hasattr(self, "times"):
            self.prologue()
        T0 = time.perf_counter()
        ka = (self.n_k // 2 + self._n // 2) % self.n_k
        if self._n % 2 == 0 or self._fn is None:
            self._fn = self._row1(ka)
        else:
            self._row2(ka, self._fn)
            os.remove(self._fn)
            self._fn = None
        self._n += 1
        dt = time.perf_counter() - T0
        both = self.times["row1"] and self.times["row2"]
        return dt, (self.T_real() if both else None), self.parts()

    def parts(self):
        mean = lambda v: float(np.mean(v)) if v else float("nan")
        return {"t_prologue": self.t_prologue, "t_row1": mean(self.times["row1"]), "t_row2": mean(self.times["row2"]),
                "t_dot": mean(self.times["dot"]), "n_row1": len(self.times["row1"]), "n_row2": len(self.times["row2"])}

    def T_real(self):
        """seconds per realisation with the reference's loop counts; a missing kind of row is sampled first"""
        if not self.times["row1"] or not self.times["row2"]:
            ka = self.n_k // 2
            fn = self._fn if self._fn else self._row1(ka)
            if not self.times["row2"]:
                self._row2(ka, fn)
            if fn != self._fn:
                os.remove(fn)
        p = self.parts()
        return p["t_prologue"] + self.n_b * (p["t_row1"] + p["t_row2"]) + self.n_b ** 2 * p["t_dot"]

    def describe(self):
        return ("reference low_mem branch (pk/compute_pk_randoms.py:200-281) in shipped precision (complex64) on torch-CPU "
                "kernels with %d threads, full mesh %s: steps alternate between one loop-1 row and one loop-2 row at l=2 (5 forward "
                "transforms = the mean row), the latter with %d of its np.load+product-sum terms; T_real = t_prologue + n_b "
                "(t_row1+t_row2) + n_b^2 t_dot with n_b=%d from the means over the steps (prologue timed once, outside the steps)"
                % (self.threads, "x".join(str(int(g)) for g in self.grid), self.n_dots, self.n_b))

    def close(self):
        if self._fn and os.path.exists(self._fn):
            os.remove(self._fn)
        try:
            os.rmdir(self.tmpdir)
        except OSError:
            pass


def bk_pair_sample(grid, n_b, n_l, n_k, threads=None, n_rep=3):
    """CPU baseline of one bispectrum Fisher PAIR (bk/compute_bk_randoms.py:222-517), bounded sample: the three operations
    that dominate the reference -- a complex64 3-D FFT (pyfftw at src/opt_utilities.py:502-555; here torch CPU / MKL on all
    threads), an np.save of a phi map (:396-412) and an np.load + product-sum of two maps (load_row, :477-501) -- timed on the
    full mesh and multiplied by the reference's own counts for n_k = 11, lmax = 4: 22 535 transforms,
    4 n_b saved maps, n_b (n_b + 1) loaded pairs.  Element-wise work between the transforms is NOT counted, so this is an
    upper bound on the reference's speed."""
    import torch
    threads = int(threads or os.cpu_count() or 1)
    torch.set_num_threads(threads)
    x = torch.randn(tuple(grid), dtype=torch.float64)
    xc = x.to(torch.complex64)
    torch.fft.fftn(xc)
    t0 = time.perf_counter()
    for _ in range(n_rep):
        torch.fft.fftn(xc)
    t_fft = (time.perf_counter() - t0) / n_rep
    d = tempfile.mkdtemp(prefix="sww_bkarm_")
    fn = os.path.join(d, "phi.npy")
    t0 = time.perf_counter()
    for _ in range(n_rep):
        np.save(fn, x.numpy())
    t_save = (time.perf_counter() - t0) / n_rep
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(n_rep):
        acc += float(torch.sum(torch.from_numpy(np.load(fn)) * x))
    t_dot = (time.perf_counter() - t0) / n_rep
    os.remove(fn)
    os.rmdir(d)
    # counts of SURVEY.md 3.3 scale with the number of bins: 22 535 transforms at n_b = 573 (n_k = 11, lmax = 4)
    n_fft = 22535.0 * n_b / 573.0
    T = n_fft * t_fft + 4.0 * n_b * t_save + n_b * (n_b + 1.0) * t_dot
    return {"value": 2.0 / T, "unit": "realisations/s", "cores": threads, "kind": "port",
            "sample": "one complex64 FFT, one np.save and one np.load + product-sum of full %s maps (torch CPU, %d threads), times the "
                      "reference's counts per pair: %.0f transforms, %d saved maps, %d loaded pairs; element-wise work not counted "
                      "(upper bound on the reference's speed)" % ("x".join(str(g) for g in grid), threads, n_fft, 4 * n_b, n_b * (n_b + 1)),
            "parts_s": {"t_fft": t_fft, "t_save": t_save, "t_load_dot": t_dot, "seconds_per_pair": T}}
