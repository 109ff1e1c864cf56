// Passes of the 3-D transforms for axis lengths that are not powers of two (included by fft_sm100.cu after the
// power-of-two kernels, whose argument structures, box / sphere pruning predicates and epilogues they share).  One kernel
// per pass with the length as a run-time radix plan (fft_any.cuh):
//   gpass_i1 / gpass_i2 / gpass_f2 / gpass_f3<MODE>   column passes, tile = [n][8 z-columns], same items and buffers as
//                                                      pass_i1 / pass_i2 / pass_f2 / pass_f3
//   gpass_z<INV, FWD>                                  z pass, tile = 16 real rows packed as 8 complex rows (z = x_A + i x_B:
//                                                      two real transforms for one complex transform of length nz, valid
//                                                      for odd nz too -- the half-length trick of pass_zw is not)
//   gpass_zc<DIR>                                      complex z pass (C2C transforms)
// Every mesh of the reference's parameter files runs on these (258 x 496 x 274, 172 x 330 x 182, 167 x 307 x 175) as do
// the toy worlds of the goldens (32 x 24 x 20, 20 x 18 x 15); a mesh may mix power-of-two and other axes.
#pragma once

#define ANY_SMEM_TILES(n) ((size_t)(n) * (2 * FFT_LDT + 1) * sizeof(double2))

// 16-byte asynchronous copy global -> shared (LDGSTS); ok == false writes zeros (src-size 0).  All copies of a tile are in
// flight at once, so the load phase costs one memory latency, not one per element (ncu on the first version: 28 % of the
// stall samples sat on the shared-memory store behind each dependent global load).
__device__ __forceinline__ void any_cp16(double2* dst, const double2* src, bool ok) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void any_cp_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

struct AnyTile {
  double2 *b0, *b1, *tw;
};
__device__ __forceinline__ AnyTile any_tile_setup(unsigned char* smem_raw, int n, const double2* __restrict__ twg, int tid, int nthr) {
  AnyTile t;
  t.b0 = reinterpret_cast<double2*>(smem_raw);
  t.b1 = t.b0 + (size_t)n * FFT_LDT;
  t.tw = t.b1 + (size_t)n * FFT_LDT;
  for (int i = tid; i < n; i += nthr) t.tw[i] = twg[i];
  return t;
}

// ---- first inverse pass (gather) -------------------------------------------------------------------------------------
template <int AX>
__global__ void __maxnreg__(96)
gpass_i1_kernel(Geo g, AnyPlan pl, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* __restrict__ src, int shell,
                const double2* __restrict__ twg, double2* __restrict__ Pbase, GatherBatch gb, BoxAxis cbx, BoxAxis cby,
                long long pstride, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, N = pl.n;
  AnyTile T = any_tile_setup(smem_raw, N, twg, tid, nthr);
  const int ztiles = (Kzp + FFT_TZ - 1) / FFT_TZ;   // a ragged last tile is allowed here (complex transforms: pitch nz)
  const unsigned per = (unsigned)bO.K * ztiles, items = per * (gb.nb > 0 ? gb.nb : 1);
  for (unsigned it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const unsigned bat = it0 / per, it = it0 - bat * per;
    const GatherArgs& ga = gb.ga[bat];
    double2* __restrict__ P = Pbase + (long long)bat * pstride;
    const int jo = (int)(it / ztiles), zt = (int)(it - (unsigned)jo * ztiles);
    const int io = bO.idx(jo), iz0 = zt * FFT_TZ;
    if (klim2 > 0.0) {
      const int n_o = AX == 0 ? g.ny : g.nx;
      if (sidx_abs(io, n_o) > sphere_lim(klim2, g.kax[2][iz0], g.kax[AX == 0 ? 1 : 0][1], n_o)) continue;
    }
    __syncthreads();
    for (int e = tid; e < N * FFT_TZ; e += nthr) {
      const int ia = e >> 3, cc = e & 7;
      double2 v = make_double2(0.0, 0.0);
      const int iz = iz0 + cc;
      if (bA.has(ia) && iz < Kz) {
        long long cell = row_of<AX>(ia, io, g.ny) * g.nzh + iz;
        if (ga.n_terms == 0) {
          const double2* __restrict__ sp = gb.nb > 0 ? ga.src : src;
          const int sh = gb.nb > 0 ? ga.shell : shell;
          if (sh < 0 || g.binmap[cell] == sh) v = sp[cell];
        } else {
          const int b = g.binmap[cell];
          const int ix = AX == 0 ? ia : io, iy = AX == 0 ? io : ia;
          const long long ci = ((long long)cbx.pos(ix) * cby.K + cby.pos(iy)) * g.Kz + iz;
          bool any = false;
#pragma unroll
          for (int t = 0; t < 3; ++t)
            if (t < ga.n_terms && b == ga.s[t]) {
              double2 q = ga.p[t][ci];
              v.x += q.x;
              v.y += q.y;
              any = true;
            }
          if (any) {
            double f = ga.coef;
            if (ga.lm >= 0) {
              double kx = g.kax[0][ix], ky = g.kax[1][iy], kz = g.kax[2][iz];
              double inv = inv_norm_eps(kx * kx + ky * ky + kz * kz);
              f *= sww_ylm_dyn(ga.lm, kx * inv, ky * inv, kz * inv);
            }
            v.x *= f;
            v.y *= f;
          }
        }
      }
      T.b0[ia * FFT_LDT + cc] = v;
    }
    __syncthreads();
    const double2* res = any_fft_tile<-1>(T.b0, T.b1, T.tw, pl, tid, nthr);
    for (int e = tid; e < N * FFT_TZ; e += nthr) {
      const int ia = e >> 3, cc = e & 7;
      if (iz0 + cc < Kzp) P[((long long)jo * N + ia) * Kzp + iz0 + cc] = res[ia * FFT_LDT + cc];
    }
  }
}

// ---- second inverse pass ---------------------------------------------------------------------------------------------
template <int AX>
__global__ void __maxnreg__(96)
gpass_i2_kernel(AnyPlan pl, int nF, int ny, BoxAxis bA, int Kzp, const double2* __restrict__ Pbase, const double2* __restrict__ twg,
                double2* __restrict__ Qbase, int nb, long long pstride, long long qstride, const double* __restrict__ kax_a,
                const double* __restrict__ kax_z, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, N = pl.n;
  AnyTile T = any_tile_setup(smem_raw, N, twg, tid, nthr);
  const int ztiles = (Kzp + FFT_TZ - 1) / FFT_TZ;   // a ragged last tile is allowed here (complex transforms: pitch nz)
  const unsigned per = (unsigned)nF * ztiles, items = per * nb;   // 32-bit item arithmetic (checked by the launcher)
  const int cc = tid & 7, ia0 = tid >> 3, ias = nthr >> 3;        // a thread keeps its z column; its rows advance by nthr / 8
  for (unsigned it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const unsigned bat = it0 / per, it = it0 - bat * per;
    const double2* __restrict__ P = Pbase + (long long)bat * pstride;
    double2* __restrict__ Q = Qbase + (long long)bat * qstride;
    const int f = (int)(it / ztiles), zt = (int)(it - (unsigned)f * ztiles);
    const int iz0 = zt * FFT_TZ;
    const bool colok = iz0 + cc < Kzp;
    __syncthreads();
    const int lim = klim2 > 0.0 ? sphere_lim(klim2, kax_z[iz0], kax_a[1], N) : N;
    {
      // box rows: asynchronous copies from P (one pointer increment per row); the rows outside the box are zero
      const double2* src = P + ((long long)ia0 * nF + f) * Kzp + iz0 + cc;
      const long long sstep = (long long)ias * nF * Kzp;
      for (int ja = ia0; ja < bA.K; ja += ias, src += sstep) {
        const int ia = bA.idx(ja);
        const bool ok = colok && sidx_abs(ia, N) <= lim;
        any_cp16(T.b0 + ia * FFT_LDT + cc, ok ? src : P, ok);
      }
      for (int ia = bA.b + 1 + ia0; ia < N - (bA.K - bA.b - 1); ia += ias) T.b0[ia * FFT_LDT + cc] = make_double2(0.0, 0.0);
    }
    any_cp_wait();
    __syncthreads();
    const double2* res = any_fft_tile<-1>(T.b0, T.b1, T.tw, pl, tid, nthr);
    if (colok) {
      double2* dst = Q + row_of<AX>(ia0, f, ny) * Kzp + iz0 + cc;
      const long long dstep = (AX == 0 ? (long long)ias * ny : (long long)ias) * Kzp;
      for (int ia = ia0; ia < N; ia += ias, dst += dstep) *dst = res[ia * FFT_LDT + cc];
    }
  }
}

// ---- first forward pass ----------------------------------------------------------------------------------------------
template <int AX>
__global__ void __maxnreg__(96)
gpass_f2_kernel(AnyPlan pl, int nC, int ny, BoxAxis bA, int Kzp, const double2* __restrict__ Rbase, const double2* __restrict__ twg,
                double2* __restrict__ Pbase, int nb, long long rstride, long long pstride, const double* __restrict__ kax_a,
                const double* __restrict__ kax_z, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, N = pl.n;
  AnyTile T = any_tile_setup(smem_raw, N, twg, tid, nthr);
  const int ztiles = (Kzp + FFT_TZ - 1) / FFT_TZ;   // a ragged last tile is allowed here (complex transforms: pitch nz)
  const unsigned per = (unsigned)nC * ztiles, items = per * nb;
  const int cc = tid & 7, ia0 = tid >> 3, ias = nthr >> 3;
  for (unsigned it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const unsigned bat = it0 / per, it = it0 - bat * per;
    const double2* __restrict__ R = Rbase + (long long)bat * rstride;
    double2* __restrict__ P = Pbase + (long long)bat * pstride;
    const int ic = (int)(it / ztiles), zt = (int)(it - (unsigned)ic * ztiles);
    const int iz0 = zt * FFT_TZ;
    const bool colok = iz0 + cc < Kzp;
    const int lim = klim2 > 0.0 ? sphere_lim(klim2, kax_z[iz0], kax_a[1], N) : N;
    __syncthreads();
    {
      const double2* src = R + row_of<AX>(ia0, ic, ny) * Kzp + iz0 + cc;
      const long long sstep = (AX == 0 ? (long long)ias * ny : (long long)ias) * Kzp;
      for (int ia = ia0; ia < N; ia += ias, src += sstep) any_cp16(T.b0 + ia * FFT_LDT + cc, colok ? src : R, colok);
    }
    any_cp_wait();
    __syncthreads();
    const double2* res = any_fft_tile<1>(T.b0, T.b1, T.tw, pl, tid, nthr);
    if (colok) {
      double2* dst = P + ((long long)ia0 * nC + ic) * Kzp + iz0 + cc;
      const long long dstep = (long long)ias * nC * Kzp;
      for (int ja = ia0; ja < bA.K; ja += ias, dst += dstep) {
        const int ia = bA.idx(ja);
        if (sidx_abs(ia, N) <= lim) *dst = res[ia * FFT_LDT + cc];
      }
    }
  }
}

// ---- last forward pass + epilogue (modes of pass_f3_kernel) ----------------------------------------------------------------
template <int AX, int MODE>
__global__ void __maxnreg__(96)
gpass_f3_kernel(Geo g, AnyPlan pl, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* __restrict__ P,
                const double2* __restrict__ twg, F3Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ bool is_last;
  const int tid = threadIdx.x, nthr = blockDim.x, N = pl.n;
  AnyTile T = any_tile_setup(smem_raw, N, twg, tid, nthr);
  double* hist = reinterpret_cast<double*>(T.tw + N);  // MODE 2: [nwarps][n_g*n_k]
  const int NW = nthr >> 5;
  const int warp = tid >> 5, lane = tid & 31;
  const int nb = A.n_g * g.n_k;
  if (MODE == 2)
    for (int i = tid; i < NW * nb; i += nthr) hist[i] = 0.0;
  const int ztiles = (Kzp + FFT_TZ - 1) / FFT_TZ;   // a ragged last tile is allowed here (complex transforms: pitch nz)
  const unsigned items = (unsigned)bO.K * ztiles;
  const bool nz_even = (g.nz % 2) == 0;
  const int nrows = (MODE == 2 && A.nrows > 1) ? A.nrows : 1;
  const int row = (int)(blockIdx.x % (unsigned)nrows), blk = (int)(blockIdx.x / (unsigned)nrows);
  const int nblk = (int)(gridDim.x / (unsigned)nrows);
  if (MODE == 2 && nrows > 1) P += (long long)row * A.row_pstride;
  for (unsigned it = blk; it < items; it += nblk) {
    const int jo = (int)(it / ztiles), zt = (int)(it - (unsigned)jo * ztiles);
    const int io = bO.idx(jo), iz0 = zt * FFT_TZ;
    if (A.klim2 > 0.0) {
      const int n_o = AX == 0 ? g.ny : g.nx;
      if (sidx_abs(io, n_o) > sphere_lim(A.klim2, g.kax[2][iz0], g.kax[AX == 0 ? 1 : 0][1], n_o)) continue;
    }
    __syncthreads();
    const int cc0 = tid & 7, ia0 = tid >> 3, ias = nthr >> 3;
    {
      const bool colok = iz0 + cc0 < Kzp;
      const double2* src = P + ((long long)jo * N + ia0) * Kzp + iz0 + cc0;
      const long long sstep = (long long)ias * Kzp;
      for (int ia = ia0; ia < N; ia += ias, src += sstep) any_cp16(T.b0 + ia * FFT_LDT + cc0, colok ? src : P, colok);
    }
    any_cp_wait();
    __syncthreads();
    const double2* res = any_fft_tile<1>(T.b0, T.b1, T.tw, pl, tid, nthr);
    if (MODE == 3) {
      if (iz0 + cc0 < Kz) {
        const int* pp = A.perm + ((long long)jo * bA.K + ia0) * Kzp + iz0 + cc0;
        const long long pstep = (long long)ias * Kzp;
        for (int ja = ia0; ja < bA.K; ja += ias, pp += pstep) {
          const int p = __ldg(pp);
          if (p >= 0) A.sorted_dst[p] = res[bA.idx(ja) * FFT_LDT + cc0];
        }
      }
    } else if (MODE != 2) {
      // four box rows per round; MODE 1 loads the previous values of the round before its first store (see pass_f3_kernel)
      const int iz = iz0 + cc0;
      if (iz < Kz) {
        constexpr int U = 4;
        for (int jb = ia0; jb < bA.K; jb += U * ias) {
          long long cells[U];
          double2 old[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int ja = jb + u * ias;
            cells[u] = -1;
            old[u] = make_double2(0.0, 0.0);
            if (ja < bA.K) {
              const int ia = bA.idx(ja);
              cells[u] = remap_cell(g, A, AX == 0 ? ia : io, AX == 0 ? io : ia, iz);
              if (MODE == 1 && !A.first) old[u] = A.dst[cells[u]];
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (cells[u] < 0) continue;
            const int ia = bA.idx(jb + u * ias);
            const int ix = AX == 0 ? ia : io, iy = AX == 0 ? io : ia;
            const double2 t = res[ia * FFT_LDT + cc0];
            if (MODE == 0) {
              A.dst[cells[u]] = t;
            } else {
              double kx = g.kax[0][ix], ky = g.kax[1][iy], kz = g.kax[2][iz];
              double inv = inv_norm_eps(kx * kx + ky * ky + kz * kz);
              double yk = sww_ylm_dyn(A.lm, kx * inv, ky * inv, kz * inv);
              double2 o = old[u];
              o.x += t.x * yk;
              o.y += t.y * yk;
              if (A.coef != 0.0) {
                if (A.mas_power) {
                  double m = g.mas[0][ix] * g.mas[1][iy] * g.mas[2][iz];
                  if (A.mas_power == 2) m = m * m;
                  o.x *= m;
                  o.y *= m;
                }
                o.x *= A.coef;
                o.y *= A.coef;
              }
              A.dst[cells[u]] = o;
            }
          }
        }
      }
    } else {
      // shell binning from the natural-order tile: the scheme of pass_f3_kernel (pairs +j / -j share a bin, segmented
      // shuffle reduction along j, warp-private histograms, fixed order -> deterministic), for any N
      constexpr int GRP = 2;
      const bool full = (bA.K == N);
      const int npos = full ? N / 2 + 1 : bA.b + 1;
      const int nsteps = (npos + 15) >> 4;
      const int nitems = nsteps * 4;
      const int par = lane & 1, sub = lane >> 1;
      for (int q0 = warp; q0 < nitems; q0 += GRP * NW) {
        int bq[GRP], jq[GRP], ccq[GRP];
        long long cellp[GRP], gcp[GRP], gcm[GRP];
        bool mir[GRP];
#pragma unroll
        for (int j = 0; j < GRP; ++j) {
          const int q = q0 + j * NW;
          const int stp = q >> 2, cc = ((q & 3) << 1) + par, iz = iz0 + cc;
          const int ja = (stp << 4) + sub;
          const bool inb = (q < nitems) && ja < npos && iz < Kz;
          jq[j] = inb ? ja : 0;
          ccq[j] = cc;
          mir[j] = inb && ja > 0 && !(full && 2 * ja == N);
          cellp[j] = row_of<AX>(jq[j], io, g.ny) * g.nzh + (inb ? iz : 0);
          {
            const int jm = mir[j] ? N - jq[j] : 0;
            const int izc = inb ? iz : 0;
            gcp[j] = AX == 0 ? remap_cell(g, A, jq[j], io, izc) : remap_cell(g, A, io, jq[j], izc);
            gcm[j] = AX == 0 ? remap_cell(g, A, jm, io, izc) : remap_cell(g, A, io, jm, izc);
          }
          bq[j] = inb ? (int)g.binmap[cellp[j]] : SWW_NO_BIN;
        }
        double2 Gp[GRP][3], Gm[GRP][3];
#pragma unroll
        for (int j = 0; j < GRP; ++j)
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < A.n_g && bq[j] != SWW_NO_BIN) {
              Gp[j][q] = A.G[q][gcp[j]];
              if (mir[j]) Gm[j][q] = A.G[q][gcm[j]];
            }
#pragma unroll
        for (int j = 0; j < GRP; ++j) {
          const int bn = bq[j];
          double v[3] = {0.0, 0.0, 0.0};
          if (bn != SWW_NO_BIN) {
            const int iz = iz0 + ccq[j];
            const double wq = (iz == 0 || (nz_even && iz == g.nz / 2)) ? 1.0 : 2.0;
            const double2 t = res[jq[j] * FFT_LDT + ccq[j]];
            double2 tm = make_double2(0.0, 0.0);
            if (mir[j]) tm = res[(N - jq[j]) * FFT_LDT + ccq[j]];
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < A.n_g) {
                double sacc = t.x * Gp[j][q].x + t.y * Gp[j][q].y;
                if (mir[j]) sacc += tm.x * Gm[j][q].x + tm.y * Gm[j][q].y;
                v[q] = wq * sacc;
              }
          }
#pragma unroll
          for (int d = 2; d < 32; d <<= 1) {
            const int nbk = __shfl_down_sync(0xffffffffu, bn, d);
            const bool take = (lane + d < 32) && (nbk == bn);
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < A.n_g) {
                const double nv = __shfl_down_sync(0xffffffffu, v[q], d);
                if (take) v[q] += nv;
              }
          }
          const int pbk = __shfl_up_sync(0xffffffffu, bn, 2);
          const bool head = (bn != SWW_NO_BIN) && (lane < 2 || pbk != bn);
#pragma unroll
          for (int rq = 0; rq < 2; ++rq) {
            if (head && par == rq) {
#pragma unroll
              for (int q = 0; q < 3; ++q)
                if (q < A.n_g) hist[warp * nb + q * g.n_k + bn] += v[q];
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (MODE == 2) {
    const int row_ = (int)(blockIdx.x % (unsigned)nrows);
    double* out_row = nrows > 1 ? A.out_rows[row_] : A.out_row;
    double* partial = A.partial + (long long)row_ * nblk * nb;
    unsigned int* counter = A.counter + row_;
    __syncthreads();
    for (int j = tid; j < nb; j += nthr) {
      double sm = 0.0;
      for (int w = 0; w < NW; ++w) sm += hist[w * nb + j];
      partial[(long long)blk * nb + j] = sm;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      unsigned t = atomicAdd(counter, 1u);
      is_last = (t == (unsigned)nblk - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int j = tid; j < nb; j += nthr) {
        double sm = 0.0;
        for (int bl = 0; bl < nblk; ++bl) sm += partial[(long long)bl * nb + j];
        out_row[j] = sm * A.scale;
      }
      if (tid == 0) *counter = 0;
    }
  }
}

// ---- z pass: 16 real rows = 8 complex rows per tile ----------------------------------------------------------------------
// Column c of the tile carries z = x_A + i x_B with A = row0 + 2c, B = A + 1.  Inverse: Z[k] = X_A[k] + i X_B[k] over the
// Hermitian extension of the two half-spectra (the imaginary parts of the DC and Nyquist modes are dropped, like every
// complex-to-real transform does); forward: X_A[k] = (Z[k] + conj Z[n-k]) / 2, X_B[k] = -i (Z[k] - conj Z[n-k]) / 2.
// Arguments as in pass_zw_kernel (ZArgs); the real-space step between the two transforms is done in shared memory.
template <bool INV, bool FWD>
__global__ void __maxnreg__(96)
gpass_z_kernel(Geo g, AnyPlan pl, ZArgs A, const double2* __restrict__ twg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, n = pl.n;
  AnyTile T = any_tile_setup(smem_raw, n, twg, tid, nthr);
  const long long nrows = (long long)g.nx * g.ny;
  const long long ntiles = (nrows + 15) / 16;
  const int c = tid & 7, m0 = tid >> 3, mstep = nthr >> 3;
  const int nzh = g.nzh;
  const int npl = (INV && A.nb > 1) ? A.nb : 1;
  const bool sum_planes = INV && !FWD && A.nb > 1;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long rowA = tile * 16 + 2 * c, rowB = rowA + 1;
    const bool hasA = rowA < nrows, hasB = rowB < nrows;
    double XA = 0.0, YA = 0.0, XB = 0.0, YB = 0.0;   // r^ coordinates of the two rows
    if (hasA) {
      const int ix = (int)(rowA / g.ny), iy = (int)(rowA - (long long)ix * g.ny);
      XA = g.rax[0][ix];
      YA = g.rax[1][iy];
    }
    if (hasB) {
      const int ix = (int)(rowB / g.ny), iy = (int)(rowB - (long long)ix * g.ny);
      XB = g.rax[0][ix];
      YB = g.rax[1][iy];
    }
    for (int pb = 0; pb < npl; ++pb) {
      __syncthreads();
      double2* cur = T.b0;
      double2* oth = T.b1;
      if (INV) {
        const double2* __restrict__ Qp = A.Q + (long long)pb * A.q_stride;
        for (int kb = m0; kb < nzh; kb += 4 * mstep) {
          double2 xa[4], xb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = kb + u * mstep;
            xa[u] = xb[u] = make_double2(0.0, 0.0);
            if (k < A.Kz_in) {
              if (hasA) xa[u] = Qp[rowA * A.Kzp_in + k];
              if (hasB) xb[u] = Qp[rowB * A.Kzp_in + k];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = kb + u * mstep;
            if (k >= nzh) continue;
            const bool self = (k == 0) || (2 * k == n);
            if (self) {
              xa[u].y = 0.0;
              xb[u].y = 0.0;
            }
            T.b0[k * FFT_LDT + c] = make_double2(xa[u].x - xb[u].y, xa[u].y + xb[u].x);
            if (!self) T.b0[(n - k) * FFT_LDT + c] = make_double2(xa[u].x + xb[u].y, xb[u].x - xa[u].y);
          }
        }
        __syncthreads();
        double2* res = any_fft_tile<-1>(T.b0, T.b1, T.tw, pl, tid, nthr);
        cur = res;
        oth = (res == T.b0) ? T.b1 : T.b0;
      }
      // real-space step
      const int lmc = sum_planes ? A.lms[pb] : A.lm;
      const bool ymap = lmc >= 1 && A.ymaps != nullptr;
      const bool yfly = lmc >= 0 && !ymap;
      double AcA[3] = {0.0, 0.0, 0.0}, AcB[3] = {0.0, 0.0, 0.0};
      bool oddA = false, oddB = false;
      const int deg = yfly ? SWW_YLM_ZDEG(lmc) : 0;
      if (yfly) {
        sww_ylm_zpack(lmc, XA, YA, AcA, &oddA);
        sww_ylm_zpack(lmc, XB, YB, AcB, &oddB);
      }
      const double xy2A = XA * XA + YA * YA, xy2B = XB * XB + YB * YB;
      // four cells per round: every global load of the round is issued before the first dependent use
      constexpr int U = 4;
      for (int mb = m0; mb < n; mb += U * mstep) {
        double2 v[U];
        double fa[U], fb[U], ya[U], yb[U], oa[U], ob[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int m = mb + u * mstep;
          const bool in = m < n;
          fa[u] = fb[u] = A.scale;
          ya[u] = yb[u] = 1.0;
          oa[u] = ob[u] = 0.0;
          v[u] = make_double2(0.0, 0.0);
          if (!INV) {
            if (in && hasA) v[u].x = A.real_in[rowA * n + m];
            if (in && hasB) v[u].y = A.real_in[rowB * n + m];
          }
          if (A.wmap) {
            fa[u] *= (in && hasA) ? A.wmap[rowA * n + m] : 0.0;
            fb[u] *= (in && hasB) ? A.wmap[rowB * n + m] : 0.0;
          }
          if (ymap) {
            const double* ym = A.ymaps + (long long)(lmc - 1) * A.N;
            ya[u] = (in && hasA) ? ym[rowA * n + m] : 0.0;
            yb[u] = (in && hasB) ? ym[rowB * n + m] : 0.0;
          }
          if (INV && !FWD && (A.accumulate || (sum_planes && pb > 0))) {
            if (in && hasA) oa[u] = A.real_out[rowA * n + m];
            if (in && hasB) ob[u] = A.real_out[rowB * n + m];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int m = mb + u * mstep;
          if (m >= n) continue;
          if (INV) v[u] = cur[m * FFT_LDT + c];
          double ga = fa[u] * ya[u], gb = fb[u] * yb[u];
          if (yfly) {
            const double Z = g.rax[2][m], w = Z * Z;
            ga *= sww_ylm_zpacked(deg, AcA, oddA, Z, w, deg ? sww_rcp_pos(xy2A + w) : 1.0);
            gb *= sww_ylm_zpacked(deg, AcB, oddB, Z, w, deg ? sww_rcp_pos(xy2B + w) : 1.0);
          }
          v[u].x *= ga;
          v[u].y *= gb;
          if (INV && !FWD) {
            if (hasA) A.real_out[rowA * n + m] = oa[u] + v[u].x;
            if (hasB) A.real_out[rowB * n + m] = ob[u] + v[u].y;
          } else {
            cur[m * FFT_LDT + c] = v[u];
          }
        }
      }
      if (FWD) {
        __syncthreads();
        const double2* res = any_fft_tile<1>(cur, oth, T.tw, pl, tid, nthr);
        double2* __restrict__ Rp = A.R + (long long)pb * A.r_stride;
        for (int k = m0; k < A.Kz_out; k += mstep) {
          const double2 a = res[k * FFT_LDT + c];
          const double2 b = res[(k == 0 ? 0 : n - k) * FFT_LDT + c];
          if (hasA) Rp[rowA * A.Kzp_out + k] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
          if (hasB) Rp[rowB * A.Kzp_out + k] = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));
        }
      }
    }
  }
}

// ---- complex z pass: in-place transforms of 8 contiguous rows of length n ------------------------------------------------
template <int DIR>
__global__ void __maxnreg__(96)
gpass_zc_kernel(AnyPlan pl, long long nrows, double2* __restrict__ data, const double2* __restrict__ twg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, n = pl.n;
  AnyTile T = any_tile_setup(smem_raw, n, twg, tid, nthr);
  const long long ntiles = (nrows + 7) / 8;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    for (int e = tid; e < 8 * n; e += nthr) {
      const int c = e / n, k = e - c * n;
      const long long row = tile * 8 + c;
      any_cp16(T.b0 + k * FFT_LDT + c, row < nrows ? data + row * n + k : data, row < nrows);
    }
    any_cp_wait();
    __syncthreads();
    const double2* res = any_fft_tile<DIR>(T.b0, T.b1, T.tw, pl, tid, nthr);
    for (int e = tid; e < 8 * n; e += nthr) {
      const int c = e / n, k = e - c * n;
      const long long row = tile * 8 + c;
      if (row < nrows) data[row * n + k] = res[k * FFT_LDT + c];
    }
  }
}
