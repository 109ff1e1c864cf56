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

    one loop-1 row and one loop-2 row at multipole l = 2 (2l+1 = 5 forward transforms = the mean over l = 0,2,4 of the
    rows of a realisation), with `n_dots` of the loop-2 row's np.load + product-sum terms against the saved loop-1 map

-- and the time of one realisation follows from the reference's own loop counts, every term measured inside the step:

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
        self._seed = 0

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

    # ---- one bounded sample ------------------------------------------------------------------------------------
    def step(self):
        """Returns (seconds of the step, T_real extrapolated with the reference's loop counts, parts)."""
        t = self.t
        self._seed += 1
        T0 = time.perf_counter()
        # prologue (pk/compute_pk_randoms.py:139-140, 206-217)
        np.random.seed(self._seed)
        a = t.from_numpy(np.random.normal(loc=0.0, scale=1.0, size=tuple(self.grid))) * t.sqrt(self.nbar_A)
        Cinv = self.applyCinv_fkp(a).real
        Ainv = a / self.nbar_A
        NAinv = self.applyN(Ainv)
        del a
        T1 = time.perf_counter()
        i = 1 if self.n_l > 1 else 0
        ka = (self.n_k // 2 + self._seed) % self.n_k
        # one row of loop 1 (:220-226)
        m1 = self.applyC_alpha_single(Ainv, i, ka)
        fn = os.path.join(self.tmpdir, "C_a_Ai_%d.npy" % ka)
        np.save(fn, m1.numpy())
        del m1
        T2 = time.perf_counter()
        # one row of loop 2 without its beta loop (:244-265)
        tmp = self.applyC_alpha_single(Cinv, i, ka)
        row = self.applyCinv_fkp(tmp)
        del tmp
        qb = 0.5 * float(t.sum(row * NAinv).real)
        T3 = time.perf_counter()
        # n_dots terms of that row's beta loop (:268-272): np.load + product-sum
        acc = 0.0
        for _ in range(self.n_dots):
            acc += 0.5 * float(t.sum(row * t.from_numpy(np.load(fn))).real)
        T4 = time.perf_counter()
        os.remove(fn)
        parts = {"t_prologue": T1 - T0, "t_row1": T2 - T1, "t_row2": T3 - T2, "t_dot": (T4 - T3) / max(1, self.n_dots),
                 "check": [qb, acc]}
        T_real = parts["t_prologue"] + self.n_b * (parts["t_row1"] + parts["t_row2"]) + self.n_b ** 2 * parts["t_dot"]
        return T4 - T0, T_real, parts

    def describe(self):
        return ("reference low_mem branch (pk/compute_pk_randoms.py:200-281) in shipped precision (complex64) on torch-CPU "
                "kernels with %d threads, full mesh %s: a step = prologue + one loop-1 row + one loop-2 row at l=2 (5 forward "
                "transforms = the mean row) + %d np.load+product-sum terms; T_real = t_prologue + n_b (t_row1+t_row2) + "
                "n_b^2 t_dot with n_b=%d, every term measured inside the step"
                % (self.threads, "x".join(str(int(g)) for g in self.grid), self.n_dots, self.n_b))

    def close(self):
        try:
            os.rmdir(self.tmpdir)
        except OSError:
            pass
