"""TEST / BENCH INFRASTRUCTURE -- CPU baseline of the pk Fisher realisation.

Times a BOUNDED sample of the reference algorithm (the numpy port in oracle/sww_oracle.py,
which follows pk/compute_pk_randoms.py:206-281 in its non-spilling flow) on the host cores
and extrapolates to one full realisation with the reference's own operation counts:

    T_real = 2 M t_fwd  +  2 n_b t_deriv  +  n_b t_cinv  +  (n_b^2 + n_b) t_dot

  t_fwd   : one term  ft(x n Y_lm(r^)) Y_lm(k^)                      (src/covariances_pk.py:384)
  t_deriv : one map   n ift(4pi/(2l+1) Theta_a f_l) / v_cell         (src/covariances_pk.py:393-399)
  t_cinv  : one applyCinv_fkp of a derivative map                    (pk/compute_pk_randoms.py:238)
  t_dot   : one full-map product-sum                                 (pk/compute_pk_randoms.py:267-274)

The reference uses pyfftw (FFTW, 4 threads, complex64); pyfftw cannot be installed in this image,
so FFTs run through scipy.fft in complex128 with `workers` host threads.  Only bench.py
(cpu_baseline leg and --impl reference) may import this module.
"""
from __future__ import annotations

import os
import time

import numpy as np
import scipy.fft

from . import sww_oracle as orc


class PkFisherSample:
    def __init__(self, box, grid, box_center, nbar, k_min, k_max, dk, lmax, workers=None):
        self.workers = workers or os.cpu_count()
        orc.WORKERS = self.workers
        self.box, self.grid = np.asarray(box, float), np.asarray(grid, int)
        self.v_cell = self.box.prod() / self.grid.prod()
        self.nbar = nbar
        self.lmax, self.n_l = lmax, lmax // 2 + 1
        self.n_k = int((k_max - k_min) / dk)
        self.n_b = self.n_k * self.n_l
        self.M = sum(2 * l + 1 for l in range(0, lmax + 1, 2))
        k, r = orc.load_coord_grids(box, grid, box_center)
        self.k_norm = np.sqrt(np.sum(k ** 2.0, axis=0))
        k /= (1e-12 + self.k_norm)
        r /= (1e-12 + np.sqrt(np.sum(r ** 2.0, axis=0)))
        self.k_hat, self.r_hat = k, r
        self.Y = orc.compute_spherical_harmonic_functions(lmax)
        self.filt = orc.compute_filters(k_min, k_max, dk)
        rng = np.random.default_rng(1)
        self.x = rng.standard_normal(tuple(self.grid))

    def sample(self):
        """One bounded sample; returns (seconds_extrapolated_per_realisation, parts)."""
        Y = self.Y[1][2] if self.n_l > 1 else self.Y[0][0]
        t0 = time.perf_counter()
        f = orc.ft(self.x * self.nbar * Y(*self.r_hat)) * Y(*self.k_hat)
        t1 = time.perf_counter()
        a = self.n_k // 2
        m = self.nbar * orc.ift(4.0 * np.pi / 5.0 * self.filt(a, self.k_norm) * f) / self.v_cell
        t2 = time.perf_counter()
        u = orc.applyCinv_fkp(m, self.nbar, self.v_cell, 1.02)
        t3 = time.perf_counter()
        s = np.real(np.sum(u * m))
        t4 = time.perf_counter()
        parts = {"t_fwd": t1 - t0, "t_deriv": t2 - t1, "t_cinv": t3 - t2, "t_dot": t4 - t3, "check": float(s)}
        T = (2 * self.M * parts["t_fwd"] + 2 * self.n_b * parts["t_deriv"] + self.n_b * parts["t_cinv"]
             + (self.n_b ** 2 + self.n_b) * parts["t_dot"])
        return T, parts

    def describe(self):
        return ("1 Y_lm forward term + 1 derivative map + 1 C^-1 + 1 map dot at full mesh %s, complex128 scipy.fft "
                "(%d workers), extrapolated with 2M=%d, n_b=%d: T = 2M t_fwd + 2n_b t_deriv + n_b t_cinv + (n_b^2+n_b) t_dot"
                % ("x".join(str(int(g)) for g in self.grid), self.workers, 2 * self.M, self.n_b))
