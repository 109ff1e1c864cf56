// Maximum-likelihood weights (`weights = ML`): the signal + noise covariance C = S + N of
// src/covariances_pk.py:10-95 and its preconditioned conjugate-gradient inverse (:222-327), plus the P_l(k)
// Fisher / data flows built on them (pk/compute_pk_randoms.py:205-281, pk/compute_pk_data.py:194-206).
//
// Unlike the FKP path these operators keep FULL COMPLEX maps, double2[nx][ny][nz]: the fiducial spectrum
// P_l(|k|) covers every mode of the mesh, and on the Nyquist planes Y_lm(k^) P_l(k) is not Hermitian, so the
// reference's complex128 maps carry imaginary parts of O(1e-2) of the Nyquist-plane power which feed back
// through the bilinear CG products (np.sum(r*pre_r), no conjugation).  Matching it to 1e-9 needs the same
// complex arithmetic; transforms are C2C (cuFFT Z2Z plan, every element-wise step is a kernel of this file).
#include <cmath>
#include <complex>

#include "sww_internal.cuh"
#include "ylm_gen.cuh"

typedef std::complex<double> cplx;

static inline double lcoef(int l_i) { return 4.0 * M_PI / (4.0 * l_i + 1.0); }

__device__ __forceinline__ void unit_vec_ml(double X, double Y, double Z, double& x, double& y, double& z) {
  double r = sqrt(X * X + Y * Y + Z * Z);
  double inv = 1.0 / (1e-12 + r);
  x = X * inv;
  y = Y * inv;
  z = Z * inv;
}

// ---- element-wise kernels on complex maps (rows = (ix, iy), threads along z: coalesced 16-byte accesses) ----
// out = in * m * s   (m: optional real map)
__global__ void c_scale_kernel(long long N, const double2* __restrict__ in, const double* __restrict__ m, double s,
                               double2* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double f = m ? m[i] * s : s;
    double2 v = in[i];
    out[i] = make_double2(v.x * f, v.y * f);
  }
}

// out = in * Y_lm(r^)
__global__ void c_mul_ylm_r_kernel(Geo g, const double2* __restrict__ in, int lm, double2* __restrict__ out) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double X = g.rax[0][ix], Y = g.rax[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      double x, y, z;
      unit_vec_ml(X, Y, g.rax[2][iz], x, y, z);
      double f = sww_ylm_dyn(lm, x, y, z);
      long long i = row * g.nz + iz;
      double2 v = in[i];
      out[i] = make_double2(v.x * f, v.y * f);
    }
  }
}

// acc = (first ? 0 : acc) + spec * Y_lm(k^); with pk != null (last m of a multipole):
// F = (first_l ? 0 : F) + coef * pk * acc      (src/covariances_pk.py:82-89)
__global__ void c_acc_ylm_k_kernel(Geo g, const double2* __restrict__ spec, int lm, int first, double2* __restrict__ acc,
                                   const double* __restrict__ pk, double coef, int first_l, double2* __restrict__ F) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double kx = g.kax[0][ix], ky = g.kax[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      double x, y, z;
      unit_vec_ml(kx, ky, g.kax[2][iz], x, y, z);
      double yk = sww_ylm_dyn(lm, x, y, z);
      long long i = row * g.nz + iz;
      double2 s = spec[i];
      double2 a = first ? make_double2(0.0, 0.0) : acc[i];
      a.x += s.x * yk;
      a.y += s.y * yk;
      if (pk) {
        double w = coef * pk[i];
        double2 f = first_l ? make_double2(0.0, 0.0) : F[i];
        f.x += w * a.x;
        f.y += w * a.y;
        F[i] = f;
      } else {
        acc[i] = a;
      }
    }
  }
}

// spec *= MAS^power (power = +1 / -1) on the full grid, times s
__global__ void c_mas_kernel(Geo g, double2* __restrict__ spec, int power, double s) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double mxy = g.mas[0][ix] * g.mas[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      double m = mxy * g.mas[2][iz];
      double f = (power > 0 ? m : 1.0 / m) * s;
      long long i = row * g.nz + iz;
      double2 v = spec[i];
      spec[i] = make_double2(v.x * f, v.y * f);
    }
  }
}

// FKP ratio map: out = in / (w (w P + s)) * scale on w > 0, else 0   (src/covariances_pk.py:153-160)
__global__ void c_fkp_kernel(long long N, const double2* __restrict__ in, const double* __restrict__ nw, double P, double shot,
                             double scale, double2* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double w = nw[i];
    double2 v = in[i];
    if (w > 0) {
      double d = w * (w * P + shot);
      out[i] = make_double2(v.x / d * scale, v.y / d * scale);
    } else {
      out[i] = make_double2(0.0, 0.0);
    }
  }
}

// y = a + s * b   (complex scalar s); a may be null (treated as 0)
__global__ void c_xpsy_kernel(long long N, const double2* __restrict__ a, double2 s, const double2* __restrict__ b,
                              double2* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double2 u = a ? a[i] : make_double2(0.0, 0.0), v = b[i];
    y[i] = make_double2(u.x + s.x * v.x - s.y * v.y, u.y + s.x * v.y + s.y * v.x);
  }
}

// real <-> complex
__global__ void c_from_real_kernel(long long N, const double* __restrict__ in, double2* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = make_double2(in[i], 0.0);
}
// out = part(in) * m  (part: 0 real, 1 imaginary; m optional)
__global__ void c_part_kernel(long long N, const double2* __restrict__ in, int part, const double* __restrict__ m,
                              double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double v = part ? in[i].y : in[i].x;
    out[i] = m ? v * m[i] : v;
  }
}

// bilinear product sum_i a_i b_i (no conjugation), b complex or real; fixed-order two-stage reduction
#define DOT_BLOCKS (SMS * 4)
template <bool BREAL>
__global__ void __launch_bounds__(256) c_dot_kernel(long long N, const double2* __restrict__ a, const void* __restrict__ b_,
                                                    double2* __restrict__ partial) {
  __shared__ double2 red[8];
  double sx = 0.0, sy = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double2 u = a[i];
    if (BREAL) {
      double v = ((const double*)b_)[i];
      sx += u.x * v;
      sy += u.y * v;
    } else {
      double2 v = ((const double2*)b_)[i];
      sx += u.x * v.x - u.y * v.y;
      sy += u.x * v.y + u.y * v.x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_double2(sx, sy);
  __syncthreads();
  if (threadIdx.x == 0) {
    double2 t = make_double2(0.0, 0.0);
    for (int w = 0; w < 8; ++w) {
      t.x += red[w].x;
      t.y += red[w].y;
    }
    partial[blockIdx.x] = t;
  }
}
__global__ void c_dot_final_kernel(const double2* __restrict__ partial, int n, double2* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double2 t = make_double2(0.0, 0.0);
    for (int i = 0; i < n; ++i) {
      t.x += partial[i].x;
      t.y += partial[i].y;
    }
    *out = t;
  }
}

// ------------------------------------------------------------------------------------------
// buffers
// ------------------------------------------------------------------------------------------
struct MlBufs {
  // P_l(k) flow (real maps and half-spectra, as in the FKP pipelines)
  double2 *FC[SWW_MAX_L], *G[SWW_MAX_L], *X;
  double *tmp, *nCa, *nAa, *NAa, *u, *row;
  // complex maps of the operators and of the conjugate-gradient solver
  double2 *cx, *cr, *cp, *cCp, *cz, *ctmp, *ct, *cacc, *cF, *cin;
  double2* dots;   // [DOT_BLOCKS + 4]
};

static void ml_layout(sww_ctx* c, Arena& ar, MlBufs& b, bool with_pk) {
  size_t N = c->g.N, Nc = c->g.Nc;
  ar.reset(c->persist_bytes);
  b.cx = ar.alloc<double2>(N);
  b.cr = ar.alloc<double2>(N);
  b.cp = ar.alloc<double2>(N);
  b.cCp = ar.alloc<double2>(N);
  b.cz = ar.alloc<double2>(N);
  b.ctmp = ar.alloc<double2>(N);
  b.ct = ar.alloc<double2>(N);
  b.cacc = ar.alloc<double2>(N);
  b.cF = ar.alloc<double2>(N);
  b.cin = ar.alloc<double2>(N);
  b.dots = ar.alloc<double2>(DOT_BLOCKS + 4);
  if (with_pk) {
    for (int l = 0; l < c->n_l; ++l) b.FC[l] = ar.alloc<double2>(Nc);
    for (int l = 0; l < c->n_l; ++l) b.G[l] = ar.alloc<double2>(Nc);
    b.X = ar.alloc<double2>(Nc);
    b.tmp = ar.alloc<double>(N);
    b.nCa = ar.alloc<double>(N);
    b.nAa = ar.alloc<double>(N);
    b.NAa = ar.alloc<double>(N);
    b.u = ar.alloc<double>(N);
    b.row = ar.alloc<double>((size_t)c->n_l * c->g.n_k);
  }
}

size_t sww_ml_pipeline_bytes(sww_ctx* c) {
  Arena dry;
  dry.dry = true;
  size_t save = c->persist_bytes;
  c->persist_bytes = 0;
  MlBufs b;
  ml_layout(c, dry, b, true);
  c->persist_bytes = save;
  return dry.peak;
}

#define ML_READY(c)                                                                   \
  SWW_REQUIRE(c, "null context");                                                     \
  SWW_REQUIRE((c)->ar.base, "no workspace bound (sww_bind_workspace)");               \
  SWW_REQUIRE(!(c)->fields_dirty && (c)->nbar, "fields not set (sww_set_fields)")
#define ML_ARENA(c) \
  SWW_REQUIRE((c)->ar.ok(), "workspace too small for the ML pipeline: needs %zu bytes, %zu bound (create the context with the 'pk_ml' pipeline)", (c)->ar.off, (c)->ar.cap)

static inline int ew_grid() { return SMS * 8; }
#define LAUNCHED(c)             \
  do {                          \
    (c)->n_own++;               \
    SWW_CUDA(cudaGetLastError()); \
  } while (0)

static int c_scale(sww_ctx* c, const double2* in, const double* m, double s, double2* out) {
  c_scale_kernel<<<ew_grid(), 256, 0, c->stream>>>(c->g.N, in, m, s, out);
  LAUNCHED(c);
  return 0;
}
static int c_xpsy(sww_ctx* c, const double2* a, cplx s, const double2* b, double2* y) {
  c_xpsy_kernel<<<ew_grid(), 256, 0, c->stream>>>(c->g.N, a, make_double2(s.real(), s.imag()), b, y);
  LAUNCHED(c);
  return 0;
}
static int c_fkp(sww_ctx* c, const double2* in, double scale, double2* out) {
  c_fkp_kernel<<<ew_grid(), 256, 0, c->stream>>>(c->g.N, in, c->nbar_w, c->P_fkp, c->shot, scale, out);
  LAUNCHED(c);
  return 0;
}
// out = IFT[ MAS^power FT[in] ]  (in == out allowed)
static int c_mas_filter(sww_ctx* c, const double2* in, int power, double2* out) {
  SWW_TRY(sww_c2c(c, in, out, 0));
  c_mas_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, out, power, 1.0);
  LAUNCHED(c);
  return sww_c2c(c, out, out, 1);
}
// complex bilinear dot -> host value (synchronises the stream)
static int c_dot(sww_ctx* c, MlBufs& b, const double2* x, const void* y, bool y_real, cplx* out) {
  if (y_real)
    c_dot_kernel<true><<<DOT_BLOCKS, 256, 0, c->stream>>>(c->g.N, x, y, b.dots);
  else
    c_dot_kernel<false><<<DOT_BLOCKS, 256, 0, c->stream>>>(c->g.N, x, y, b.dots);
  LAUNCHED(c);
  c_dot_final_kernel<<<1, 32, 0, c->stream>>>(b.dots, DOT_BLOCKS, b.dots + DOT_BLOCKS);
  LAUNCHED(c);
  double2 h;
  SWW_CUDA(cudaMemcpyAsync(&h, b.dots + DOT_BLOCKS, sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
  SWW_CUDA(cudaStreamSynchronize(c->stream));
  *out = cplx(h.x, h.y);
  return 0;
}

// ---- operators on complex maps; `out` must not alias the scratch buffers ctmp / ct / cacc / cF ----
// FKP inverse (first guess and preconditioner), src/covariances_pk.py:126-164,166-220
static int ml_cinv_fkp(sww_ctx* c, MlBufs& b, const double2* in, double2* out) {
  if (c->include_pix) {
    SWW_TRY(c_mas_filter(c, in, +1, b.ct));
    SWW_TRY(c_fkp(c, b.ct, 1.0, b.ct));
    SWW_TRY(c_mas_filter(c, b.ct, +1, out));
    return c_scale(c, out, nullptr, c->v_cell, out);
  }
  return c_fkp(c, in, c->v_cell, out);
}

// N x, src/covariances_pk.py:97-124
static int ml_apply_n(sww_ctx* c, MlBufs& b, const double2* in, double2* out) {
  if (c->include_pix) {
    SWW_TRY(c_mas_filter(c, in, -1, b.ct));
    SWW_TRY(c_scale(c, b.ct, c->nbar, 1.0, b.ct));
    SWW_TRY(c_mas_filter(c, b.ct, -1, out));
    return c_scale(c, out, nullptr, c->shot / c->v_cell, out);
  }
  return c_scale(c, in, c->nbar, c->shot / c->v_cell, out);
}

// S x, src/covariances_pk.py:45-95
static int ml_apply_s(sww_ctx* c, MlBufs& b, const double2* in, double2* out) {
  SWW_REQUIRE(c->pk_fid, "fiducial P_l(k) maps not set (sww_set_fiducial_pk)");
  if (c->include_pix) {
    SWW_TRY(c_mas_filter(c, in, -1, b.ctmp));
    SWW_TRY(c_scale(c, b.ctmp, c->nbar, 1.0, b.ctmp));
  } else {
    SWW_TRY(c_scale(c, in, c->nbar, 1.0, b.ctmp));
  }
  int lm = 0;
  for (int l_i = 0; l_i < c->n_l; ++l_i) {
    const int nm = 4 * l_i + 1;
    for (int m = 0; m < nm; ++m, ++lm) {
      c_mul_ylm_r_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, b.ctmp, lm, b.ct);
      LAUNCHED(c);
      SWW_TRY(sww_c2c(c, b.ct, b.ct, 0));
      const bool last = (m == nm - 1);
      c_acc_ylm_k_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, b.ct, lm, m == 0, b.cacc,
                                                        last ? c->pk_fid + (size_t)l_i * c->g.N : nullptr, lcoef(l_i),
                                                        l_i == 0, b.cF);
      LAUNCHED(c);
    }
  }
  SWW_TRY(sww_c2c(c, b.cF, b.cF, 1));
  if (c->include_pix) {
    SWW_TRY(c_scale(c, b.cF, c->nbar, 1.0, b.cF));
    SWW_TRY(c_mas_filter(c, b.cF, -1, out));
    return c_scale(c, out, nullptr, 1.0 / c->v_cell, out);
  }
  return c_scale(c, b.cF, c->nbar, 1.0 / c->v_cell, out);
}

// C x = S x + N x, src/covariances_pk.py:10-43.  Uses cin as a second output buffer.
static int ml_apply_c(sww_ctx* c, MlBufs& b, const double2* in, double2* out) {
  SWW_TRY(ml_apply_s(c, b, in, out));
  SWW_TRY(ml_apply_n(c, b, in, b.cacc));
  return c_xpsy(c, out, cplx(1.0, 0.0), b.cacc, out);
}

// Preconditioned conjugate gradients, src/covariances_pk.py:276-327.  pix -> b.cx.  stop: 1 relative / absolute
// tolerance met, 2 stalled, 0 iteration limit.
static int ml_cg(sww_ctx* c, MlBufs& b, const double2* pix, int max_it, double abs_tol, double rel_tol, int* iterations,
                 int* stop) {
  SWW_REQUIRE((long long)max_it < c->g.N, "max_it must be smaller than the number of cells");
  SWW_TRY(ml_cinv_fkp(c, b, pix, b.cx));                         // x = C~^-1 pix
  SWW_TRY(ml_apply_c(c, b, b.cx, b.cr));
  SWW_TRY(c_xpsy(c, pix, cplx(-1.0, 0.0), b.cr, b.cr));          // r = pix - C x
  SWW_TRY(ml_cinv_fkp(c, b, b.cr, b.cp));                        // p = C~^-1 r
  SWW_TRY(ml_apply_c(c, b, b.cp, b.cCp));
  cplx old_sum, pCp, new_sum(0.0, 0.0);
  SWW_TRY(c_dot(c, b, b.cr, b.cp, false, &old_sum));             // pre_r == p at this point
  const cplx init_sum = old_sum;
  SWW_TRY(c_dot(c, b, b.cp, b.cCp, false, &pCp));
  cplx alpha = old_sum / pCp;
  cplx save_sum = old_sum;
  int it = 0, why = 0;
  for (int i = 0; i < max_it; ++i) {
    if (i % 10 == 0 && i > 0) {
      if (std::abs((save_sum - old_sum) / old_sum) < 0.05) {
        why = 2;
        break;
      }
      save_sum = old_sum;
    }
    SWW_TRY(c_xpsy(c, b.cx, alpha, b.cp, b.cx));                 // x += alpha p
    SWW_TRY(c_xpsy(c, b.cr, -alpha, b.cCp, b.cr));               // r -= alpha C p
    SWW_TRY(ml_cinv_fkp(c, b, b.cr, b.cz));                      // pre_r
    SWW_TRY(c_dot(c, b, b.cr, b.cz, false, &new_sum));
    SWW_TRY(c_xpsy(c, b.cz, new_sum / old_sum, b.cp, b.cp));     // p = pre_r + (new/old) p
    it = i + 1;
    // numpy orders complex numbers lexicographically: the real parts decide (src/covariances_pk.py:309-315)
    const cplx crit = rel_tol >= 0.0 ? std::sqrt(new_sum / init_sum) : std::sqrt(new_sum);
    const double tol = rel_tol >= 0.0 ? rel_tol : abs_tol;
    if (crit.real() < tol || (crit.real() == tol && crit.imag() < 0.0)) {
      why = 1;
      break;
    }
    old_sum = new_sum;
    SWW_TRY(ml_apply_c(c, b, b.cp, b.cCp));
    SWW_TRY(c_dot(c, b, b.cp, b.cCp, false, &pCp));
    alpha = old_sum / pCp;
  }
  if (iterations) *iterations = it;
  if (stop) *stop = why;
  return 0;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int sww_set_fiducial_pk(sww_ctx* c, sww_devptr pk_maps) {
  SWW_REQUIRE(c, "null context");
  c->pk_fid = (const double*)pk_maps;
  return 0;
}

extern "C" int sww_ml_apply(sww_ctx* c, int op, sww_devptr x, sww_devptr out) {
  ML_READY(c);
  SWW_REQUIRE(x && out && x != out, "ml_apply: null or aliased pointers");
  MlBufs b;
  ml_layout(c, c->ar, b, false);
  ML_ARENA(c);
  switch (op) {
    case 0: return ml_apply_s(c, b, (const double2*)x, (double2*)out);
    case 1: return ml_apply_n(c, b, (const double2*)x, (double2*)out);
    case 2: return ml_apply_c(c, b, (const double2*)x, (double2*)out);
    case 3: return ml_cinv_fkp(c, b, (const double2*)x, (double2*)out);
    default: SWW_REQUIRE(false, "ml_apply: unknown operator %d", op);
  }
  return 0;
}

extern "C" int sww_ml_apply_cinv(sww_ctx* c, sww_devptr pix, int max_it, double abs_tol, double rel_tol, sww_devptr out,
                                 int32_t* iterations, int32_t* stop_reason) {
  ML_READY(c);
  SWW_REQUIRE(pix && out, "ml_apply_cinv: null pointer");
  MlBufs b;
  ml_layout(c, c->ar, b, false);
  ML_ARENA(c);
  int it = 0, why = 0;
  SWW_TRY(ml_cg(c, b, (const double2*)pix, max_it, abs_tol, rel_tol, &it, &why));
  SWW_CUDA(cudaMemcpyAsync((void*)out, b.cx, sizeof(double2) * c->g.N, cudaMemcpyDeviceToDevice, c->stream));
  if (iterations) *iterations = it;
  if (stop_reason) *stop_reason = why;
  return 0;
}

int sww_ylm_project_internal(sww_ctx* c, const double* f, int mas_power, bool with_coef, bool pruned, double* tmp,
                             double2* spec, double2* const* out, const sww_ctx* out_mesh = nullptr);   // pipelines.cu
int sww_mas_filter_internal(sww_ctx* c, const double* x, int power, double2* spec, double* out);

// shell row: row[q*n_k + b] = scale * sum_{k in b} w Re[conj(FT[t]) G_q]
static int bin_against(sww_ctx* c, MlBufs& b, const double* t, double2* const* G, double scale, double* out_row) {
  if (c->fft_backend == 1) {
    F3Args A = {};
    for (int q = 0; q < c->n_l; ++q) A.G[q] = G[q];
    A.n_g = c->n_l;
    A.scale = scale;
    A.out_row = out_row;
    c->n_fft++;
    return sww_sm100_forward_box(c, t, nullptr, -1, 2, A);
  }
  SWW_TRY(sww_r2c(c, t, b.X));
  return k_shell_bin(c, b.X, (const double2* const*)G, c->n_l, scale, out_row);
}

// q_alpha of an already inverse-covariance-weighted REAL map z (pk/compute_pk_data.py:198-206):
// q_alpha = 1/2 sum_x z C_,alpha z.  Complex C^-1 d: q[Re z] - q[Im z] (the products are bilinear).
extern "C" int sww_pk_q_from_cinv(sww_ctx* c, sww_devptr z_, sww_devptr q_out) {
  ML_READY(c);
  SWW_REQUIRE(z_ && q_out, "null device pointer");
  MlBufs b;
  ml_layout(c, c->ar, b, true);
  ML_ARENA(c);
  const double N = (double)c->g.N;
  const double* z = (const double*)z_;
  if (c->include_pix) {   // C_,alpha = M^-1 n ... n M^-1 and M is symmetric: weight M^-1 z, no MAS^2 factor
    SWW_TRY(sww_mas_filter_internal(c, z, -1, b.X, b.tmp));
    SWW_TRY(k_mul(c, b.tmp, c->nbar, 1.0, b.nCa));
  } else {
    SWW_TRY(k_mul(c, z, c->nbar, 1.0, b.nCa));
  }
  SWW_TRY(sww_ylm_project_internal(c, b.nCa, c->include_pix ? 0 : 2, true, true, b.tmp, b.X, b.FC));
  return bin_against(c, b, b.nCa, b.FC, 0.5 / (N * c->v_cell), (double*)q_out);
}

// One Monte-Carlo realisation with ML weights: q-bar_alpha and symmetrised F_ab
// (pk/compute_pk_randoms.py:205-281 along the non-low_mem branch; n_b + 1 conjugate-gradient solves).
extern "C" int sww_pk_fisher_realisation_ml(sww_ctx* c, sww_devptr a_, int max_it, double rel_tol, sww_devptr fish_out,
                                            sww_devptr qbar_out, int32_t* total_iterations) {
  ML_READY(c);
  SWW_REQUIRE(a_ && fish_out && qbar_out, "null device pointer");
  // include_pix = True: C_,alpha = M^-1 n IFT[...] n M^-1 / v (src/covariances_pk.py:371-372,396-397), N = M^-1 n M^-1,
  // and M^-1 (a real, even filter) is self-adjoint, so it is moved onto the map each product is taken with
  const bool pix = c->include_pix;
  MlBufs b;
  ml_layout(c, c->ar, b, true);
  ML_ARENA(c);
  const double* a = (const double*)a_;
  double* F = (double*)fish_out;
  double* qbar = (double*)qbar_out;
  const int n_k = c->g.n_k, n_l = c->n_l, n_b = n_k * n_l;
  const long long Nl = c->g.N;
  const double N = (double)Nl;
  int it = 0, its = 0;
  // C^-1 a (real part, :211), A^-1 a, N A^-1 a
  c_from_real_kernel<<<ew_grid(), 256, 0, c->stream>>>(Nl, a, b.cin);
  LAUNCHED(c);
  SWW_TRY(ml_cg(c, b, b.cin, max_it, 1e-8, rel_tol, &it, nullptr));
  its += it;
  c_part_kernel<<<ew_grid(), 256, 0, c->stream>>>(Nl, b.cx, 0, pix ? nullptr : c->nbar, b.nCa);   // [n *] Re C^-1 a
  LAUNCHED(c);
  SWW_TRY(k_div(c, a, c->nbar_A, b.tmp));                                            // A^-1 a
  if (pix) {
    SWW_TRY(sww_mas_filter_internal(c, b.nCa, -1, b.X, b.u));
    SWW_TRY(k_mul(c, b.u, c->nbar, 1.0, b.nCa));                                     // n M^-1 Re C^-1 a
    SWW_TRY(sww_mas_filter_internal(c, b.tmp, -1, b.X, b.u));
    SWW_TRY(k_mul(c, b.u, c->nbar, 1.0, b.nAa));                                     // n M^-1 A^-1 a
    SWW_TRY(sww_mas_filter_internal(c, b.nAa, -1, b.X, b.NAa));
    SWW_TRY(k_mul(c, b.NAa, nullptr, c->shot / c->v_cell, b.NAa));                   // N A^-1 a = s M^-1 n M^-1 A^-1 a / v
  } else {
    SWW_TRY(k_mul(c, b.tmp, c->nbar, 1.0, b.nAa));                                   // n A^-1 a
    SWW_TRY(k_mul(c, b.tmp, c->nbar, c->shot / c->v_cell, b.NAa));                   // N A^-1 a
  }
  SWW_TRY(sww_ylm_project_internal(c, b.nCa, 0, true, true, b.tmp, b.X, b.FC));
  SWW_TRY(sww_ylm_project_internal(c, b.nAa, 0, true, true, b.tmp, b.X, b.G));
  for (int alpha = 0; alpha < n_b; ++alpha) {
    const int l_i = alpha / n_k, ka = alpha % n_k;
    // C_,alpha C^-1 a = n IFT[Theta_a F^C_l] / v_cell
    if (c->fft_backend == 1) {
      c->n_fft++;
      SWW_TRY(sww_sm100_inverse_box(c, b.FC[l_i], ka, -1, c->nbar, -1, 1.0 / (N * c->v_cell), 0, b.u));
    } else {
      SWW_TRY(k_shell_fill(c, b.FC[l_i], ka, 1.0, b.X));
      SWW_TRY(sww_c2r(c, b.X, b.u));
      SWW_TRY(k_mul(c, b.u, c->nbar, 1.0 / (N * c->v_cell), b.u));
    }
    if (pix) {
      SWW_TRY(sww_mas_filter_internal(c, b.u, -1, b.X, b.tmp));
      SWW_TRY(k_mul(c, b.tmp, nullptr, 1.0, b.u));
    }
    // U_alpha = C^-1 [C_,alpha C^-1 a]
    c_from_real_kernel<<<ew_grid(), 256, 0, c->stream>>>(Nl, b.u, b.cin);
    LAUNCHED(c);
    SWW_TRY(ml_cg(c, b, b.cin, max_it, 1e-8, rel_tol, &it, nullptr));
    its += it;
    // q-bar_alpha = 1/2 Re sum U_alpha N A^-1 a
    cplx d;
    SWW_TRY(c_dot(c, b, b.cx, b.NAa, true, &d));
    const double qb = 0.5 * d.real();
    SWW_CUDA(cudaMemcpyAsync(qbar + alpha, &qb, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SWW_CUDA(cudaStreamSynchronize(c->stream));
    // F_alpha,beta = 1/2 sum Re(U_alpha) n IFT[Theta_b G_l'] / v_cell, binned in Fourier space
    if (pix) {
      c_part_kernel<<<ew_grid(), 256, 0, c->stream>>>(Nl, b.cx, 0, nullptr, b.tmp);
      LAUNCHED(c);
      SWW_TRY(sww_mas_filter_internal(c, b.tmp, -1, b.X, b.u));
      SWW_TRY(k_mul(c, b.u, c->nbar, 1.0, b.u));
    } else {
      c_part_kernel<<<ew_grid(), 256, 0, c->stream>>>(Nl, b.cx, 0, c->nbar, b.u);
      LAUNCHED(c);
    }
    SWW_TRY(bin_against(c, b, b.u, b.G, 0.5 / (N * c->v_cell), F + (size_t)alpha * n_b));
  }
  SWW_TRY(k_symmetrise(c, F, n_b));
  if (total_iterations) *total_iterations = its;
  return 0;
}
