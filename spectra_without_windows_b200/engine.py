"""Host-side engine: mesh geometry built with the reference's own expressions, and a thin
object wrapper over the C ABI (libsww.so).  PyTorch is used only to own device buffers,
pinned host buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_TORCH = None


def torch():
    global _TORCH
    if _TORCH is None:
        import torch as t
        _TORCH = t
    return _TORCH


def signed_index(n):
    """src/opt_utilities.py:422 -- `kkk-grid if kkk>=grid/2 else kkk`."""
    i = np.arange(n)
    return np.where(i >= n / 2, i - n, i)


class Mesh:
    """Geometry + binning of one (box, grid, BoxCenter, k-binning, lmax) configuration.

    Everything that decides bin membership or enters at >1e-9 is evaluated here on the host
    with the reference's expressions and uploaded as small 1-D tables:
      kax  : src/opt_utilities.py:420-422         rax : src/opt_utilities.py:428-430
      mas  : src/opt_utilities.py:448-463 (float32 pi/N prefactor)
      edges: src/opt_utilities.py:485-487, n_k = int((k_max-k_min)/dk) (pk/compute_pk_randoms.py:197)
    """

    def __init__(self, box, grid, box_center, k_min, k_max, dk, lmax):
        self.box = np.asarray(box, dtype=float)
        self.grid = np.asarray(grid, dtype=int)
        self.box_center = np.asarray(box_center, dtype=float)
        self.k_min, self.k_max, self.dk, self.lmax = float(k_min), float(k_max), float(dk), int(lmax)
        assert self.k_max > self.k_min and self.dk > 0 and self.k_min >= 0
        assert (self.lmax // 2) * 2 == self.lmax, "l-max must be even!"
        self.n_l = self.lmax // 2 + 1
        self.n_k = int((self.k_max - self.k_min) / self.dk)
        k_all = np.arange(self.k_min, self.k_max + self.dk, self.dk)
        if len(k_all) - 1 < self.n_k:
            raise Exception("k-binning yields fewer filters than n_k")
        self.k_lo = np.ascontiguousarray(k_all[:-1][:self.n_k])
        self.k_hi = np.ascontiguousarray(k_all[1:][:self.n_k])
        kF = 2.0 * np.pi / (1.0 * self.box)
        self.kF = kF
        self.kax = [np.ascontiguousarray(signed_index(self.grid[i]) * kF[i], dtype=np.float64) for i in range(3)]
        offset = self.box_center + 0.5 * self.box / self.grid
        H = self.box / self.grid
        self.rax = [np.ascontiguousarray((signed_index(self.grid[i]) * H[i]).astype("f8") + offset[i]) for i in range(3)]
        prefact = np.pi / np.asarray(self.grid, dtype=np.float32)
        self.mas = []
        for i in range(3):
            s = np.sin(prefact[i] * (self.kax[i] / kF[i])) ** 2.0
            self.mas.append(np.ascontiguousarray(1.0 / (1.0 - s + 2.0 / 15 * s ** 2.0) ** 0.5, dtype=np.float64))
        self.v_cell = 1.0 * self.box.prod() / (1.0 * self.grid.prod())
        self.N = int(self.grid.prod())
        self.nzh = int(self.grid[2]) // 2 + 1
        self.Nc = int(self.grid[0]) * int(self.grid[1]) * self.nzh

    @property
    def shape(self):
        return tuple(int(g) for g in self.grid)

    @property
    def cshape(self):
        return (int(self.grid[0]), int(self.grid[1]), self.nzh)

    def key(self):
        return (tuple(self.box), tuple(self.grid), tuple(self.box_center), self.k_min, self.k_max, self.dk, self.lmax)

    def triangles(self):
        """Allowed (a<=b<=c) bin triplets -- test_bin of bk/compute_bk_randoms.py:186-211."""
        k_lo = np.arange(self.k_min, self.k_max, self.dk)
        k_hi = k_lo + self.dk
        tol = 1e-6
        out = []
        for a in range(self.n_k):
            for b in range(a, self.n_k):
                for c in range(b, self.n_k):
                    k_up = k_hi[a] + k_hi[b]
                    k_do = 0.0 if a == b else k_lo[b] - k_hi[a]
                    if k_lo[c] + tol >= k_up or k_hi[c] - tol <= k_do:
                        continue
                    out.append((a, b, c))
        return out


PIPELINES = {"ops": _lib.WS_OPS, "pk_fisher": _lib.WS_PK_FISHER, "pk_data": _lib.WS_PK_DATA,
             "bk_fisher": _lib.WS_BK_FISHER, "bk_data": _lib.WS_BK_DATA, "pk_ml": _lib.WS_PK_ML}


class Context:
    """One libsww context on one GPU.  Owns (through torch) the single workspace the library
    carves its buffers from, and keeps the n(r) maps alive while the library borrows them."""

    def __init__(self, mesh: Mesh, device: int = 0, pipelines=("ops", "pk_fisher", "pk_data")):
        t = torch()
        self.lib = _lib.load()
        if not t.cuda.is_available():
            raise Exception("no CUDA device: spectra_without_windows_b200 has no CPU path")
        self.mesh = mesh
        self.device = t.device("cuda", device)
        t.cuda.set_device(self.device)
        g = _lib.Geometry()
        g.grid[:] = [int(x) for x in mesh.grid]
        g.box[:] = [float(x) for x in mesh.box]
        g.lmax, g.n_k, g.device = mesh.lmax, mesh.n_k, device
        as_p = lambda a: a.ctypes.data_as(_lib.pd)
        g.k_lo, g.k_hi = as_p(mesh.k_lo), as_p(mesh.k_hi)
        for i in range(3):
            g.rax[i], g.kax[i], g.mas[i] = as_p(mesh.rax[i]), as_p(mesh.kax[i]), as_p(mesh.mas[i])
        self._h = _lib.ctxp()
        _lib.check(self.lib.sww_ctx_create(C.byref(g), C.byref(self._h)))
        self.n_b = mesh.n_k * mesh.n_l
        self._tris = None
        if any(p.startswith("bk") for p in pipelines):
            self.set_triangles(mesh.triangles())
        # FFT backend: the hand-written sm_100a passes wherever the mesh allows (power-of-two sizes),
        # cuFFT plans otherwise.  SWW_FFT=0/1 forces one.  Chosen before the workspace is sized.
        want = os.environ.get("SWW_FFT", "auto")
        self.fft_backend = 0
        if want != "0":
            if self.lib.sww_set_fft_backend(self._h, 1) == 0:
                self.fft_backend = 1
            elif want == "1":
                _lib.check(-1)
        need = 0
        for p in pipelines:
            b = _lib.u64(0)
            _lib.check(self.lib.sww_workspace_bytes(self._h, PIPELINES[p], C.byref(b)))
            need = max(need, b.value)
        self.workspace = t.empty(int(need), dtype=t.uint8, device=self.device)
        self.use_stream(t.cuda.current_stream(self.device))
        _lib.check(self.lib.sww_bind_workspace(self._h, self.workspace.data_ptr(), int(need)))
        self._fields = None
        # optional cache of the Y_lm(r^) maps (l > 0) for the bispectrum pipelines, when it costs < 15 % of the device memory:
        # their batched inverse z pass multiplies every plane by a harmonic (340 ms per C4 pair with the cache, 453 ms
        # evaluating the row polynomials on the fly).  The P_l(k) forward transforms evaluate Y_lm(r^) on the fly (a degree-l
        # Horner polynomial in z per mesh row, csrc/ylm_gen.cuh): as fast as reading the cached map at 512^3, and no memory.
        self._ylm_cache = None
        want_cache = os.environ.get("SWW_YLM_CACHE", "auto")
        if self.fft_backend == 1 and want_cache != "0" and (want_cache == "1" or any(p.startswith("bk") for p in pipelines)):
            b = _lib.u64(0)
            _lib.check(self.lib.sww_ylm_cache_bytes(self._h, C.byref(b)))
            total = t.cuda.get_device_properties(self.device).total_memory
            if 0 < b.value <= 0.15 * total:
                self._ylm_cache = t.empty(int(b.value), dtype=t.uint8, device=self.device)
                _lib.check(self.lib.sww_set_ylm_cache(self._h, self._ylm_cache.data_ptr()))

    # -- plumbing
    def use_stream(self, stream):
        self.stream = stream
        _lib.check(self.lib.sww_ctx_set_stream(self._h, stream.cuda_stream))

    def close(self):
        if self._h:
            self.lib.sww_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dev(self, x, dtype=None):
        """numpy / torch -> contiguous device tensor (float64 unless dtype given)."""
        t = torch()
        dtype = dtype or t.float64
        if isinstance(x, t.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        npd = {t.float64: np.float64, t.complex128: np.complex128, t.int32: np.int32}[dtype]
        return t.from_numpy(np.ascontiguousarray(x, dtype=npd)).to(self.device)

    def empty(self, shape, dtype=None):
        t = torch()
        return t.empty(shape, dtype=dtype or t.float64, device=self.device)

    def zeros(self, shape, dtype=None):
        t = torch()
        return t.zeros(shape, dtype=dtype or t.float64, device=self.device)

    def launch_counts(self):
        a, b = _lib.u64(0), _lib.u64(0)
        _lib.check(self.lib.sww_launch_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def row_mesh_shape(self):
        """Grid of the coarse row mesh the Fisher rows run on (csrc/rowmesh.cu), or None on the full mesh."""
        g = (_lib.i32 * 3)()
        _lib.check(self.lib.sww_row_mesh(self._h, g))
        return tuple(int(x) for x in g) if g[0] > 0 else None

    def profile(self, enable=True):
        """Start (or stop) CUDA-event profiling of the library's own launches."""
        _lib.check(self.lib.sww_profile_enable(self._h, int(bool(enable))))

    def profile_report(self):
        """{kernel family: (launches, total ms)} since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check(self.lib.sww_profile_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            parts = line.rsplit(None, 2)     # "<name> <launches> <ms>"
            if len(parts) == 3:
                out[parts[0].strip()] = (int(parts[1]), float(parts[2]))
        return out

    def set_fft_backend(self, backend: int):
        _lib.check(self.lib.sww_set_fft_backend(self._h, int(backend)))
        self.fft_backend = int(backend)

    # -- fields
    def set_fields(self, nbar, nbar_weight, nbar_A, shot_fac, P_fkp=1e4):
        n, nw, nA = self.dev(nbar), self.dev(nbar_weight), self.dev(nbar_A)
        self._fields = (n, nw, nA)
        self.shot_fac, self.P_fkp = float(shot_fac), float(P_fkp)
        _lib.check(self.lib.sww_set_fields(self._h, n.data_ptr(), nw.data_ptr(), nA.data_ptr(), float(shot_fac), float(P_fkp)))

    def set_fields_from_randoms_density(self, nbar_r, alpha, shot_fac, P_fkp=1e4):
        """The n(r) maps of the four scripts from the randoms-density map (load_nbar(...,1.)):
        nbar = nbar_weight = alpha * nbar_r; nbar_A = nbar_r with empty cells set to
        0.01 * min(nbar_r > 0)  (pk/compute_pk_randoms.py:135-136,160,173-179)."""
        nbar_A = np.array(nbar_r, dtype=np.float64, copy=True)
        nbar_A[nbar_A == 0] = nbar_A[nbar_A != 0].min() * 0.01
        nbar = nbar_r * alpha
        self.set_fields(nbar, nbar.copy(), nbar_A, shot_fac, P_fkp)
        return nbar_A

    def set_include_pix(self, flag):
        """The `include_pix` .param boolean: forward-model the pixel window (MAS) in the covariance."""
        _lib.check(self.lib.sww_set_include_pix(self._h, int(bool(flag))))
        self.include_pix = bool(flag)

    def mas_filter(self, x, power):
        """IFT[MAS^power FT[x]] (power = +1 / -1)."""
        x = self.dev(x)
        out = self.empty(self.mesh.shape)
        _lib.check(self.lib.sww_mas_filter(self._h, x.data_ptr(), int(power), out.data_ptr()))
        return out

    # -- Gaussian fields
    def grf_legacy(self, seed, sigma, out=None, stream=None):
        """Device-side `np.random.seed(seed); np.random.normal(0, sigma, sigma.shape)` (numpy legacy stream).
        With `stream` (a torch stream) the generator runs there and overlaps with work on the context stream."""
        sigma = self.dev(sigma)
        if out is None:
            out = self.empty(tuple(sigma.shape))
        _lib.check(self.lib.sww_grf_legacy(self._h, int(seed) & 0xFFFFFFFF, sigma.data_ptr(), out.data_ptr(), int(sigma.numel()),
                                           0 if stream is None else stream.cuda_stream))
        return out

    def grf_legacy_batch(self, seeds, sigma, outs=None, stream=None):
        """Fields of up to 8 seeds at once (independent MT19937 recurrences side by side: the cost of one field)."""
        sigma = self.dev(sigma)
        seeds = [int(s) & 0xFFFFFFFF for s in seeds]
        if outs is None:
            outs = [self.empty(tuple(sigma.shape)) for _ in seeds]
        arr = (C.c_uint32 * len(seeds))(*seeds)
        ptrs = (_lib.u64 * len(seeds))(*[o.data_ptr() for o in outs[:len(seeds)]])
        _lib.check(self.lib.sww_grf_legacy_batch(self._h, arr, len(seeds), sigma.data_ptr(), ptrs, int(sigma.numel()),
                                                 0 if stream is None else stream.cuda_stream))
        return outs[:len(seeds)]

    def field_pipeline(self, seeds, sigma, batch=None, consumers=None):
        """Iterate device-generated legacy fields.  Batch j+1 (`batch` seeds, generated side by side) is produced on a
        side stream while the caller's work on the fields of batch j runs on the `consumers` streams (default: the
        context stream; several when realisations are kept in flight on several contexts).  Yields (seed, field)."""
        t = torch()
        sigma = self.dev(sigma)
        seeds = list(seeds)
        if not seeds:
            return
        if batch is None:
            # two sets of `batch` fields stay resident: keep them below ~8 % of the device memory
            total = t.cuda.get_device_properties(self.device).total_memory
            batch = int(max(1, min(4, (0.08 * total) // (2 * 8 * sigma.numel()))))
        batch = max(1, min(int(batch), 8, len(seeds)))
        groups = [seeds[i:i + batch] for i in range(0, len(seeds), batch)]
        consumers = list(consumers) if consumers else [self.stream]
        side = t.cuda.Stream(device=self.device)
        bufs = [[self.empty(tuple(sigma.shape)) for _ in range(batch)] for _ in range(2)]
        for st in consumers:
            side.wait_stream(st)
        self.grf_legacy_batch(groups[0], sigma, bufs[0], stream=side)
        done = t.cuda.Event()
        done.record(side)
        used = []                       # used[j]: the consumers' work on batch j has been enqueued up to these events
        for j, grp in enumerate(groups):
            for st in consumers:
                st.wait_event(done)
            if j + 1 < len(groups):
                if j >= 1:
                    for ev in used[j - 1]:
                        side.wait_event(ev)     # buffer set (j+1)%2 was last read by the work on batch j-1
                # enqueue-only on the side stream: overlaps the consumers' work on batch j
                self.grf_legacy_batch(groups[j + 1], sigma, bufs[(j + 1) % 2], stream=side)
                done = t.cuda.Event()
                done.record(side)
            for i, s in enumerate(grp):
                yield s, bufs[j % 2][i]
            evs = []
            for st in consumers:
                ev = t.cuda.Event()
                ev.record(st)
                evs.append(ev)
            used.append(evs)
        _lib.check(self.lib.sww_grf_check(self._h))

    def clone_for_stream(self, stream, pipelines=("pk_fisher",)):
        """A second context on the same mesh, fields and settings that works on `stream`: independent realisations can
        then be kept in flight side by side (the tails of one realisation's kernels overlap the other's)."""
        t = torch()
        with t.cuda.stream(stream):
            c = Context(self.mesh, self.device.index, pipelines)
            if c.fft_backend != self.fft_backend:
                c.set_fft_backend(self.fft_backend)
            n, nw, nA = self._fields
            c.set_fields(n, nw, nA, self.shot_fac, self.P_fkp)
            c.set_include_pix(getattr(self, "include_pix", False))
            if getattr(self, "_pk_fid", None) is not None:      # ML weights: share the fiducial P_l(k) maps
                c._pk_fid = self._pk_fid
                _lib.check(c.lib.sww_set_fiducial_pk(c._h, c._pk_fid.data_ptr()))
        return c

    # -- P_l(k)
    def pk_fisher(self, a, fish_out=None, qbar_out=None):
        """One realisation -> (F [n_b,n_b] symmetrised, qbar [n_b]) on the device."""
        a = self.dev(a)
        F = fish_out if fish_out is not None else self.empty((self.n_b, self.n_b))
        q = qbar_out if qbar_out is not None else self.empty((self.n_b,))
        _lib.check(self.lib.sww_pk_fisher_realisation(self._h, a.data_ptr(), F.data_ptr(), q.data_ptr()))
        return F, q

    def mc_finalize(self, F_sum, q_sum, n_realisations):
        """Mean over realisations + symmetrisation of the summed Fisher / bias (after the all-reduce), in place."""
        _lib.check(self.lib.sww_mc_finalize(self._h, F_sum.data_ptr(), 0 if q_sum is None else q_sum.data_ptr(), int(F_sum.shape[0]),
                                            int(n_realisations)))
        return F_sum, q_sum

    def pk_data_q(self, diff):
        d = self.dev(diff)
        q = self.empty((self.n_b,))
        _lib.check(self.lib.sww_pk_data_q(self._h, d.data_ptr(), q.data_ptr()))
        return q

    # -- maximum-likelihood weights (context created with the "pk_ml" pipeline)
    def set_fiducial_pk(self, table):
        """Fiducial multipoles on the mesh from the `fiducial_pk` table (columns k, P_0, P_2, ...):
        `interp1d(pk_input[:,0], pk_input[:,1:].T)(k_norm)[:lmax//2+1]`, pk/compute_pk_randoms.py:188-193
        (linear interpolation; a |k| outside the table raises like interp1d does)."""
        tab = np.asarray(table, dtype=np.float64)
        m = self.mesh
        kn = np.sqrt(m.kax[0][:, None, None] ** 2.0 + m.kax[1][None, :, None] ** 2.0 + m.kax[2][None, None, :] ** 2.0)
        if kn.max() > tab[-1, 0] or kn.min() < tab[0, 0]:
            raise ValueError("A value in x_new is outside the interpolation range.")
        if tab.shape[1] - 1 < m.n_l:
            raise Exception("fiducial_pk table has fewer multipoles than lmax//2+1")
        pk = np.stack([np.interp(kn, tab[:, 0], tab[:, 1 + i]) for i in range(m.n_l)])
        self._pk_fid = self.dev(pk)
        _lib.check(self.lib.sww_set_fiducial_pk(self._h, self._pk_fid.data_ptr()))

    def _cdev(self, x):
        t = torch()
        return self.dev(x, t.complex128)

    def ml_apply(self, op, x):
        """applyS / applyN / applyC / applyCinv_fkp on a complex map (src/covariances_pk.py:10-164)."""
        t = torch()
        x = self._cdev(x)
        out = t.empty_like(x)
        _lib.check(self.lib.sww_ml_apply(self._h, {"S": 0, "N": 1, "C": 2, "Cinv_fkp": 3}[op], x.data_ptr(), out.data_ptr()))
        return out

    def ml_apply_cinv(self, pix, max_it=100, abs_tol=1e-8, rel_tol=None):
        """applyCinv (src/covariances_pk.py:222-327) -> (complex map, iterations, stop reason)."""
        t = torch()
        pix = self._cdev(pix)
        out = t.empty_like(pix)
        it, why = _lib.i32(0), _lib.i32(0)
        _lib.check(self.lib.sww_ml_apply_cinv(self._h, pix.data_ptr(), int(max_it), float(abs_tol),
                                              -1.0 if rel_tol is None else float(rel_tol), out.data_ptr(),
                                              C.byref(it), C.byref(why)))
        return out, it.value, why.value

    def pk_q_from_cinv(self, z):
        """q_alpha = 1/2 sum z C_,alpha z for a C^-1-weighted map; complex z: q[Re z] - q[Im z]."""
        t = torch()
        z = z if isinstance(z, t.Tensor) else t.from_numpy(np.ascontiguousarray(z)).to(self.device)
        parts = [z.real.contiguous(), z.imag.contiguous()] if z.is_complex() else [z.to(t.float64).contiguous()]
        q = None
        for i, p in enumerate(parts):
            qi = self.empty((self.n_b,))
            _lib.check(self.lib.sww_pk_q_from_cinv(self._h, p.data_ptr(), qi.data_ptr()))
            q = qi if q is None else q - qi
        return q

    def pk_data_q_ml(self, diff, max_it=30, rel_tol=1e-6):
        """q_alpha with ML weights (pk/compute_pk_data.py:194-206) -> (q, CG iterations)."""
        z, it, _ = self.ml_apply_cinv(diff, max_it=max_it, rel_tol=rel_tol)
        return self.pk_q_from_cinv(z), it

    def pk_fisher_ml(self, a, max_it=30, rel_tol=1e-6, fish_out=None, qbar_out=None):
        """One realisation with ML weights -> (F, qbar, total CG iterations)."""
        a = self.dev(a)
        F = fish_out if fish_out is not None else self.empty((self.n_b, self.n_b))
        q = qbar_out if qbar_out is not None else self.empty((self.n_b,))
        its = _lib.i32(0)
        _lib.check(self.lib.sww_pk_fisher_realisation_ml(self._h, a.data_ptr(), int(max_it), float(rel_tol), F.data_ptr(),
                                                         q.data_ptr(), C.byref(its)))
        return F, q, its.value

    # -- painting
    def paint(self, pos, w, box_center, scheme="TSC", scale=1.0, out=None):
        t = torch()
        pos, w = self.dev(pos), self.dev(w)
        if out is None:
            out = self.zeros(self.mesh.shape)
        bc = (C.c_double * 3)(*[float(x) for x in box_center])
        _lib.check(self.lib.sww_paint(self._h, pos.data_ptr(), w.data_ptr(), int(w.numel()), bc,
                                      3 if scheme.upper() == "TSC" else 2, float(scale), out.data_ptr()))
        return out

    def grid_data(self, dpos, dw, rpos, rw, box_center, return_randoms=False):
        """(n_g - alpha n_r)/V_cell [and alpha n_r/V_cell]  (src/opt_utilities.py:228,257-267)."""
        dw_h = np.asarray(dw, dtype=np.float64)
        rw_h = np.asarray(rw, dtype=np.float64)
        alpha = dw_h.sum() / rw_h.sum()
        v = self.mesh.v_cell
        rpos_d, rw_d = self.dev(rpos), self.dev(rw_h)
        diff = self.paint(dpos, dw_h, box_center, scale=1.0 / v)
        self.paint(rpos_d, rw_d, box_center, scale=-alpha / v, out=diff)
        if return_randoms:
            return diff, self.paint(rpos_d, rw_d, box_center, scale=alpha / v)
        return diff

    # -- operators (mirrors)
    def fft_c2c(self, x, inverse=False):
        t = torch()
        x = self.dev(x, t.complex128)
        out = t.empty_like(x)
        _lib.check(self.lib.sww_fft_c2c(self._h, x.data_ptr(), out.data_ptr(), int(bool(inverse))))
        return out

    def fft_r2c(self, x):
        t = torch()
        x = self.dev(x)
        out = t.empty(self.mesh.cshape, dtype=t.complex128, device=self.device)
        _lib.check(self.lib.sww_fft_r2c(self._h, x.data_ptr(), out.data_ptr()))
        return out

    def fft_c2r(self, x):
        t = torch()
        x = self.dev(x, t.complex128).clone()
        out = self.empty(self.mesh.shape)
        _lib.check(self.lib.sww_fft_c2r(self._h, x.data_ptr(), out.data_ptr()))
        return out

    def apply_cinv_fkp(self, x, nbar_w, v_cell, shot_fac, P_fkp=1e4):
        x, nw = self.dev(x), self.dev(nbar_w)
        out = self.empty(self.mesh.shape)
        _lib.check(self.lib.sww_apply_cinv_fkp(self._h, x.data_ptr(), nw.data_ptr(), v_cell, shot_fac, P_fkp, out.data_ptr()))
        return out

    def apply_n(self, x, nbar, v_cell, shot_fac):
        x, n = self.dev(x), self.dev(nbar)
        out = self.empty(self.mesh.shape)
        _lib.check(self.lib.sww_apply_n(self._h, x.data_ptr(), n.data_ptr(), v_cell, shot_fac, out.data_ptr()))
        return out

    def ylm_project(self, x, mas_power=0):
        t = torch()
        x = self.dev(x)
        out = t.empty((self.mesh.n_l,) + self.mesh.cshape, dtype=t.complex128, device=self.device)
        _lib.check(self.lib.sww_ylm_project(self._h, x.data_ptr(), int(mas_power), out.data_ptr()))
        return out

    def apply_c_alpha_single(self, x, nbar, v_cell, l_i, a, data=False):
        x, n = self.dev(x), self.dev(nbar)
        out = self.empty(self.mesh.shape)
        _lib.check(self.lib.sww_apply_c_alpha_single(self._h, x.data_ptr(), n.data_ptr(), float(v_cell), int(l_i),
                                                     int(a), int(bool(data)), out.data_ptr()))
        return out

    def binmap(self):
        t = torch()
        out = t.empty(self.mesh.cshape, dtype=t.uint8, device=self.device)
        _lib.check(self.lib.sww_get_binmap(self._h, out.data_ptr()))
        return out

    # -- B_l
    def set_triangles(self, tris):
        arr = np.ascontiguousarray(np.asarray(tris, dtype=np.int32).reshape(-1, 3))
        self._tris = arr
        self.n_tri = arr.shape[0]
        self.n_b_bk = self.n_tri * self.mesh.n_l
        _lib.check(self.lib.sww_bk_set_triangles(self._h, arr.ctypes.data_as(C.POINTER(_lib.i32)), arr.shape[0]))


_CTX_CACHE = {}


def get_context(mesh: Mesh, device=0, pipelines=("ops", "pk_fisher", "pk_data")) -> Context:
    """Cached context per (geometry, device, pipelines)."""
    k = (mesh.key(), device, tuple(pipelines))
    if k not in _CTX_CACHE:
        _CTX_CACHE[k] = Context(mesh, device, pipelines)
    return _CTX_CACHE[k]


def _bk_methods():
    """B_l(k1,k2,k3) methods of Context (kept together; attached below)."""
    t = torch

    def bk_delta(self):
        """Delta_abc [n_tri*n_l] (bk/compute_bk_randoms.py:427-473); cached per context."""
        if getattr(self, "_delta", None) is None:
            d = self.empty((self.n_b_bk,))
            _lib.check(self.lib.sww_bk_delta(self._h, d.data_ptr()))
            self._delta = d
        return self._delta

    def bk_gmaps(self, a, data=False):
        """g_i^l (and tilde-g_i^l unless data) maps [n_l, n_k, nx, ny, nz] of one field."""
        a = self.dev(a)
        shp = (self.mesh.n_l, self.mesh.n_k) + self.mesh.shape
        g = self.empty(shp)
        tg = None if data else self.empty(shp)
        _lib.check(self.lib.sww_bk_gmaps(self._h, a.data_ptr(), int(bool(data)), g.data_ptr(),
                                         0 if data else tg.data_ptr()))
        return g if data else (g, tg)

    def bk_fisher_pair(self, aA, aB, maps=None):
        """Fisher contribution of one pair of realisations -> (F [n_b,n_b], (gA,tgA,gB,tgB))."""
        aA, aB = self.dev(aA), self.dev(aB)
        shp = (self.mesh.n_l, self.mesh.n_k) + self.mesh.shape
        if maps is None:
            maps = tuple(self.empty(shp) for _ in range(4))
        F = self.empty((self.n_b_bk, self.n_b_bk))
        d = self.bk_delta()
        _lib.check(self.lib.sww_bk_fisher_pair(self._h, aA.data_ptr(), aB.data_ptr(), d.data_ptr(), maps[0].data_ptr(),
                                               maps[1].data_ptr(), maps[2].data_ptr(), maps[3].data_ptr(), F.data_ptr()))
        return F, maps

    def bk_q3(self, g):
        q = self.empty((self.n_tri, self.mesh.n_l))
        _lib.check(self.lib.sww_bk_q3(self._h, g.data_ptr(), q.data_ptr()))
        return q

    def bk_q1_accumulate(self, g, g_mc, tg_mc, N_mc, q):
        g_mc, tg_mc = self.dev(g_mc), self.dev(tg_mc)
        _lib.check(self.lib.sww_bk_q1_accumulate(self._h, g.data_ptr(), g_mc.data_ptr(), tg_mc.data_ptr(),
                                                 0.5 / float(N_mc), q.data_ptr()))
        return q

    def gram(self, A, B, out=None, accumulate=False):
        """C[i][j] = sum_x A[i][x] B[j][x] on the FP64 tensor cores."""
        A, B = self.dev(A), self.dev(B)
        m, n = A.shape[0], B.shape[0]
        K = A[0].numel()
        assert B[0].numel() == K
        if out is None:
            out = self.zeros((m, n))
        _lib.check(self.lib.sww_gram_f64(self._h, A.data_ptr(), B.data_ptr(), m, n, K, int(bool(accumulate)), out.data_ptr()))
        return out

    for f in (bk_delta, bk_gmaps, bk_fisher_pair, bk_q3, bk_q1_accumulate, gram):
        setattr(Context, f.__name__, f)


_bk_methods()
