// Element-wise, shell-binning and painting kernels of libsww (sm_100a, FP64 throughout).
// All kernels are HBM-bound streaming kernels: consecutive threads touch consecutive z cells
// (coalesced 8/16-byte accesses), grids are sized in multiples of the 148 SMs and loop.
#include "sww_internal.cuh"
#include "ylm_gen.cuh"


// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void unit_vec(double X, double Y, double Z, double& x, double& y, double& z) {
  // v / (1e-12 + |v|)   (pk/compute_pk_randoms.py:164-166)
  double r = sqrt(X * X + Y * Y + Z * Z);
  double inv = 1.0 / (1e-12 + r);
  x = X * inv;
  y = Y * inv;
  z = Z * inv;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// bin-index map: Theta_i(k) = [k_lo[i] <= |k| < k_hi[i]]  (src/opt_utilities.py:485-488).
// |k| is formed with explicitly rounded operations in numpy's order ((kx^2+ky^2)+kz^2, sqrt)
// so that bin membership is bit-identical to the reference's k_norm comparison.
// ------------------------------------------------------------------------------------------
__global__ void binmap_kernel(Geo g, const double* __restrict__ klo, const double* __restrict__ khi,
                              uint8_t* __restrict__ out) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double kx = g.kax[0][ix], ky = g.kax[1][iy];
    double kxy = __dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky));
    for (int iz = threadIdx.x; iz < g.nzh; iz += blockDim.x) {
      double kz = g.kax[2][iz];
      double k = __dsqrt_rn(__dadd_rn(kxy, __dmul_rn(kz, kz)));
      int b = SWW_NO_BIN;
      for (int i = 0; i < g.n_k; ++i)
        if (k >= klo[i] && k < khi[i]) {
          b = i;
          break;
        }
      out[row * g.nzh + iz] = (uint8_t)b;
    }
  }
}

int k_build_binmap(sww_ctx* c) {
  binmap_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, c->d_klo, c->d_khi, c->d_binmap);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// FKP / noise weights (src/covariances_pk.py:124,153-164; pk/compute_pk_randoms.py:200-217)
// ------------------------------------------------------------------------------------------
__global__ void prep_static_kernel(long long N, const double* __restrict__ n, const double* __restrict__ nw,
                                   double shot, double P, double* __restrict__ nW) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double w = nw[i], nn = n[i];
    double fkp = w * (w * P + shot);
    nW[i] = (w > 0) ? nn * (nn / fkp) : 0.0;
  }
}

int k_prep_static(sww_ctx* c) {
  prep_static_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, c->nbar, c->nbar_w, c->shot, c->P_fkp, c->d_nW);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// nCa = n C^-1 a ; nAa = n A^-1 a ; WS = W * (N A^-1 a)
// include_pix: a_c = M a replaces a in the C^-1 branch, ainv_in = M^-1 (a / n_A) replaces a / n_A
// a_c / ainv_in may alias nCa / WS-side scratch (the include_pix path prepares them in place): no __restrict__ on those
__global__ void prep_pk_kernel(long long N, const double* __restrict__ a, const double* __restrict__ n,
                               const double* __restrict__ nw, const double* __restrict__ nA, double shot, double P,
                               double v_cell, double* nCa, double* nAa, double* WS, const double* a_c, const double* ainv_in) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double x = a_c ? a_c[i] : a[i], nn = n[i], w = nw[i];
    double fkp = w * (w * P + shot);
    bool on = w > 0;
    double cinv = on ? (x / fkp) * v_cell : 0.0;
    double ainv = ainv_in ? ainv_in[i] : a[i] / nA[i];
    nCa[i] = cinv * nn;
    nAa[i] = ainv * nn;
    double NAa = shot * nn * ainv / v_cell;
    WS[i] = on ? (nn / fkp) * NAa : 0.0;
  }
}

int k_prep_pk(sww_ctx* c, const double* a, double* nCa, double* nAa, double* WS, const double* a_c, const double* ainv) {
  prep_pk_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, a, c->nbar, c->nbar_w, c->nbar_A, c->shot, c->P_fkp,
                                                c->v_cell, nCa, nAa, WS, a_c, ainv);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

__global__ void cinv_fkp_kernel(long long N, const double* __restrict__ x, const double* __restrict__ nw, double v_cell,
                                double shot, double P, const double* __restrict__ post, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double w = nw[i];
    double r = (w > 0) ? (x[i] / (w * (w * P + shot))) * v_cell : 0.0;
    out[i] = post ? r * post[i] : r;
  }
}

int k_cinv_fkp(sww_ctx* c, const double* x, const double* nw, double v_cell, double shot, double P, double* out,
               const double* post_mul) {
  cinv_fkp_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, x, nw, v_cell, shot, P, post_mul, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

__global__ void apply_n_kernel(long long N, const double* __restrict__ x, const double* __restrict__ n, double v_cell,
                               double shot, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = shot * n[i] * x[i] / v_cell;
}

int k_apply_n(sww_ctx* c, const double* x, const double* nbar, double v_cell, double shot, double* out) {
  apply_n_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, x, nbar, v_cell, shot, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// `out` may alias `a` or `b` (k_weight, sww_fft_c2r scale in place): element i is read before it is written, no __restrict__
__global__ void mul_kernel(long long N, const double* a, const double* b, double s, double* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = b ? a[i] * b[i] * s : a[i] * s;
}

int k_mul(sww_ctx* c, const double* a, const double* b, double s, double* out) {
  mul_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, a, b, s, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

int k_weight(sww_ctx* c, double* u, const double* w, double s) { return k_mul(c, u, w, s, u); }

__global__ void scale_c_kernel(long long n, double2* __restrict__ x, double s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 v = x[i];
    v.x *= s;
    v.y *= s;
    x[i] = v;
  }
}

int k_scale_c(sww_ctx* c, double2* x, double s, long long n) {
  scale_c_kernel<<<SMS * 8, 256, 0, c->stream>>>(n, x, s);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Y_lm(r^) multiply in real space and Y_lm(k^) accumulate in Fourier space
// (src/covariances_pk.py:383-384)
// ------------------------------------------------------------------------------------------
__global__ void mul_ylm_r_kernel(Geo g, const double* __restrict__ f, int lm, double* __restrict__ out) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double X = g.rax[0][ix], Y = g.rax[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      double x, y, z;
      unit_vec(X, Y, g.rax[2][iz], x, y, z);
      long long i = row * g.nz + iz;
      out[i] = f[i] * sww_ylm_dyn(lm, x, y, z);
    }
  }
}

int k_mul_ylm_r(sww_ctx* c, const double* f, int lm, double* out) {
  mul_ylm_r_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, f, lm, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// out = (first ? 0 : out) + Y_lm(k^) spec ; on the last m: out *= MAS^p, out *= coef.
// pruned: only the box |ix|<=bx, |iy|<=by, iz<=bz that carries binned modes is touched.
__global__ void acc_ylm_k_kernel(Geo g, const double2* __restrict__ spec, int lm, double coef, int mas_power, int first,
                                 int pruned, double2* __restrict__ out) {
  int RX = pruned ? g.Kx : g.nx, RY = pruned ? g.Ky : g.ny, RZ = pruned ? g.Kz : g.nzh;
  long long rows = (long long)RX * RY;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int jx = (int)(row / RY), jy = (int)(row - (long long)jx * RY);
    int ix = (!pruned || jx <= g.bx) ? jx : g.nx - g.Kx + jx;
    int iy = (!pruned || jy <= g.by) ? jy : g.ny - g.Ky + jy;
    double kx = g.kax[0][ix], ky = g.kax[1][iy];
    double mxy = g.mas[0][ix] * g.mas[1][iy];
    long long base = ((long long)ix * g.ny + iy) * g.nzh;
    for (int iz = threadIdx.x; iz < RZ; iz += blockDim.x) {
      double x, y, z;
      unit_vec(kx, ky, g.kax[2][iz], x, y, z);
      double yk = sww_ylm_dyn(lm, x, y, z);
      double2 s = spec[base + iz];
      double2 o = first ? make_double2(0.0, 0.0) : out[base + iz];
      o.x += s.x * yk;
      o.y += s.y * yk;
      if (coef != 0.0) {  // last m of this multipole
        if (mas_power) {
          double m = mxy * g.mas[2][iz];
          if (mas_power == 2) m = m * m;
          o.x *= m;
          o.y *= m;
        }
        o.x *= coef;
        o.y *= coef;
      }
      out[base + iz] = o;
    }
  }
}

int k_acc_ylm_k(sww_ctx* c, const double2* spec, int lm, double coef, int mas_power, int first, int pruned,
                double2* out) {
  acc_ylm_k_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, spec, lm, coef, mas_power, first, pruned, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// X = Theta_a(k) * coef * F(k) inside the pruning box (X was zeroed by the caller).
__global__ void shell_fill_kernel(Geo g, const double2* __restrict__ F, int a, double coef, double2* __restrict__ X) {
  long long rows = (long long)g.Kx * g.Ky;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int jx = (int)(row / g.Ky), jy = (int)(row - (long long)jx * g.Ky);
    int ix = (jx <= g.bx) ? jx : g.nx - g.Kx + jx;
    int iy = (jy <= g.by) ? jy : g.ny - g.Ky + jy;
    long long base = ((long long)ix * g.ny + iy) * g.nzh;
    for (int iz = threadIdx.x; iz < g.Kz; iz += blockDim.x) {
      if (g.binmap[base + iz] == a) {
        double2 v = F[base + iz];
        v.x *= coef;
        v.y *= coef;
        X[base + iz] = v;
      }
    }
  }
}

int k_shell_fill(sww_ctx* c, const double2* F, int a, double coef, double2* X) {
  SWW_CUDA(cudaMemsetAsync(X, 0, sizeof(double2) * c->g.Nc, c->stream));
  shell_fill_kernel<<<SMS * 4, 128, 0, c->stream>>>(c->g, F, a, coef, X);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// k-shell binning:  out[g*n_k + b] = scale * sum_{k in shell b} w_herm(k) Re[conj(T(k)) G_g(k)]
// Block-privatised shared-memory histogram (one private copy per warp, conflicts inside a warp
// resolved by a shuffle reduction over equal bins), fixed-order second stage -> deterministic.
// Replaces the n_b^2 full-map dot products of pk/compute_pk_randoms.py:267-274.
// ------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(256)
shell_bin_kernel(Geo g, const double2* __restrict__ T, const double2* __restrict__ G0, const double2* __restrict__ G1,
                 const double2* __restrict__ G2, double scale, double* __restrict__ partial,
                 unsigned int* __restrict__ counter, double* __restrict__ out) {
  extern __shared__ double hist[];  // [nwarps][NG*n_k]
  __shared__ bool is_last;
  const int nb = NG * g.n_k;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nwarps * nb; i += blockDim.x) hist[i] = 0.0;
  __syncthreads();
  double* myh = hist + warp * nb;
  const long long total = (long long)g.Kx * g.Ky * g.Kz;
  const bool nz_even = (g.nz % 2) == 0;
  const double2* Gs[3] = {G0, G1, G2};
  for (long long base = ((long long)blockIdx.x * nwarps + warp) * 32; base < total;
       base += (long long)gridDim.x * nwarps * 32) {
    long long idx = base + lane;
    int b = SWW_NO_BIN;
    double v[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) v[q] = 0.0;
    if (idx < total) {
      int iz = (int)(idx % g.Kz);
      long long r = idx / g.Kz;
      int jy = (int)(r % g.Ky), jx = (int)(r / g.Ky);
      int ix = (jx <= g.bx) ? jx : g.nx - g.Kx + jx;
      int iy = (jy <= g.by) ? jy : g.ny - g.Ky + jy;
      long long cell = ((long long)ix * g.ny + iy) * g.nzh + iz;
      b = g.binmap[cell];
      if (b != SWW_NO_BIN) {
        double2 t = T[cell];
        double w = (iz == 0 || (nz_even && iz == g.nz / 2)) ? 1.0 : 2.0;
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          double2 G = Gs[q][cell];
          v[q] = w * (t.x * G.x + t.y * G.y);
        }
      }
    }
    unsigned rem = __ballot_sync(0xffffffffu, b != SWW_NO_BIN);
    while (rem) {
      int src = __ffs(rem) - 1;
      int cur = __shfl_sync(0xffffffffu, b, src);
      bool mine = (b == cur);
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        double s = warp_sum(mine ? v[q] : 0.0);
        if (lane == 0) myh[q * g.n_k + cur] += s;
      }
      rem &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += hist[w * nb + j];
    partial[(long long)blockIdx.x * nb + j] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      double s = 0.0;
      for (unsigned bl = 0; bl < gridDim.x; ++bl) s += partial[(long long)bl * nb + j];
      out[j] = s * scale;
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

int k_shell_bin(sww_ctx* c, const double2* T, const double2* const* G, int n_g, double scale, double* out_row) {
  const int block = 256, nwarps = block / 32;
  int nb = n_g * c->g.n_k;
  long long total = (long long)c->g.Kx * c->g.Ky * c->g.Kz;
  int grid = (int)((total + block - 1) / block);
  if (grid > SMS * 4) grid = SMS * 4;
  if (grid < 1) grid = 1;
  SWW_REQUIRE((size_t)grid * nb <= c->partial_elems, "shell_bin: partial buffer too small");
  size_t smem = sizeof(double) * nwarps * nb;
  SWW_REQUIRE(smem <= 200 * 1024, "shell_bin: histogram too large for shared memory");
  const double2* g0 = G[0];
  const double2* g1 = n_g > 1 ? G[1] : G[0];
  const double2* g2 = n_g > 2 ? G[2] : G[0];
#define LAUNCH(NG)                                                                                        \
  do {                                                                                                    \
    if (smem > 48 * 1024)                                                                                 \
      SWW_CUDA(cudaFuncSetAttribute(shell_bin_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                    (int)smem));                                                          \
    shell_bin_kernel<NG><<<grid, block, smem, c->stream>>>(c->g, T, g0, g1, g2, scale, c->d_partial,       \
                                                           c->d_counter, out_row);                         \
  } while (0)
  if (n_g == 1)
    LAUNCH(1);
  else if (n_g == 2)
    LAUNCH(2);
  else if (n_g == 3)
    LAUNCH(3);
  else
    SWW_REQUIRE(false, "shell_bin: at most 3 multipoles");
#undef LAUNCH
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

__global__ void symmetrise_kernel(double* F, int n) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    int i = idx / n, j = idx - i * n;
    if (i < j) {
      double s = 0.5 * (F[i * n + j] + F[j * n + i]);
      F[i * n + j] = s;
      F[j * n + i] = s;
    }
  }
}

int k_symmetrise(sww_ctx* c, double* F, int n) {
  symmetrise_kernel<<<sww_grid_for((long long)n * n, 256), 256, 0, c->stream>>>(F, n);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// painting: paint.cu

// ------------------------------------------------------------------------------------------
// optional cache of Y_lm(r^) maps, lm >= 1
// ------------------------------------------------------------------------------------------
__global__ void ylm_fill_kernel(Geo g, int lm, double* __restrict__ out) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double X = g.rax[0][ix], Y = g.rax[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      double x, y, z;
      unit_vec(X, Y, g.rax[2][iz], x, y, z);
      out[row * g.nz + iz] = sww_ylm_dyn(lm, x, y, z);
    }
  }
}

extern "C" int sww_ylm_cache_bytes(sww_ctx* c, uint64_t* bytes) {
  SWW_REQUIRE(c && bytes, "null argument");
  *bytes = (uint64_t)(c->n_lm - 1) * c->g.N * sizeof(double);
  return 0;
}

extern "C" int sww_set_ylm_cache(sww_ctx* c, sww_devptr maps) {
  SWW_REQUIRE(c, "null context");
  c->ylm_maps = nullptr;
  if (!maps || c->n_lm <= 1) return 0;
  double* m = (double*)maps;
  for (int lm = 1; lm < c->n_lm; ++lm) {
    ylm_fill_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, lm, m + (size_t)(lm - 1) * c->g.N);
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
  }
  c->ylm_maps = m;
  return 0;
}

// ------------------------------------------------------------------------------------------
// include_pix building blocks: element-wise division and the k-space MAS filter
// ------------------------------------------------------------------------------------------
__global__ void div_kernel(long long N, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] / b[i];
}

int k_div(sww_ctx* c, const double* a, const double* b, double* out) {
  div_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, a, b, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

__global__ void mas_mul_kernel(Geo g, double2* __restrict__ spec, int power, double scale) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double mxy = g.mas[0][ix] * g.mas[1][iy];
    for (int iz = threadIdx.x; iz < g.nzh; iz += blockDim.x) {
      double m = mxy * g.mas[2][iz];
      double f = (power > 0 ? m : 1.0 / m) * scale;
      double2 v = spec[row * g.nzh + iz];
      v.x *= f;
      v.y *= f;
      spec[row * g.nzh + iz] = v;
    }
  }
}

int k_mas_mul(sww_ctx* c, double2* spec, int power, double scale) {
  mas_mul_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, spec, power, scale);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// out = IFT[MAS^power FT[x]]; spec: complex scratch [Nc] (x may alias out)
int sww_mas_filter_internal(sww_ctx* c, const double* x, int power, double2* spec, double* out) {
  SWW_TRY(sww_r2c(c, x, spec));
  SWW_TRY(k_mas_mul(c, spec, power, 1.0 / (double)c->g.N));
  return sww_c2r(c, spec, out);
}
