// B_l(k1,k2,k3) pipelines of libsww: filtered g-maps, phi maps, Delta_abc, triple sums and the
// FP64 tensor-core Gram contraction.
//
// Formulation (checked against bk/compute_bk_randoms.py + bk/compute_bk_data.py run verbatim):
//   g_i^l = IFT[Theta_i F^C_l],  tilde-g_i^l = IFT[Theta_i G_l]           (bk/compute_bk_randoms.py:222-267)
//   P(i,j,l) = g_i^0 g_j^l (field A) - g_i^0 g_j^l (field B)      -- the A-B difference is taken at the
//              product level because everything downstream is linear; FT[P] is computed ONCE per unique
//              (i<=j, l) and kept only inside the pruning box |k|<k_max (a few hundred KB each);
//   phi_alpha(A)-phi_alpha(B) = (n/v) { IFT[Theta_c FT P(a,b,0) + Theta_b FT P(a,c,0) + Theta_a FT P(b,c,0)] }   l = 0
//                             = (n/v) { IFT[Theta_b FT P(a,c,l) + Theta_a FT P(b,c,l)]
//                                       + 4pi/(2l+1) sum_m Y_lm(r^) IFT[Y_lm(k^) Theta_c FT P(a,b,0)] }        l > 0
//     (the three "rotations" of bk/compute_bk_randoms.py:379-403 are summed in Fourier space, so the
//      l=0 map costs one inverse transform instead of three);
//   F_ab = sum_x (tphi^A-tphi^B)_a (C^-1 phi^A - C^-1 phi^B)_b / 24 * 4/(Delta_a Delta_b), b >= a, mirrored
//     -- the one dense contraction of the path, done on the FP64 tensor cores (gram.cu).
#include <cmath>
#include <map>

#include "sww_internal.cuh"
#include "ylm_gen.cuh"


int sww_ylm_project_internal(sww_ctx* c, const double* f, int mas_power, bool with_coef, bool pruned, double* tmp,
                             double2* spec, double2* const* out, const sww_ctx* out_mesh = nullptr);
size_t sww_sm100_qstack_bytes(const sww_ctx* c, int nb);
int sww_gram3(sww_ctx* c, const double* X, const double* Y, const double* Z, int ni, int nj, int nz, long long K, double* C);
int sww_gram_upper(sww_ctx* c, const double* A, const double* B, int m, int n, long long K, int row_limit,
                   double* C, int ldc, int col0);

static inline double lcoef(int l_i) { return 4.0 * M_PI / (4.0 * l_i + 1.0); }

__device__ __forceinline__ void unit_vec_bk(double X, double Y, double Z, double& x, double& y, double& z) {
  double r = sqrt(X * X + Y * Y + Z * Z);
  double inv = 1.0 / (1e-12 + r);
  x = X * inv;
  y = Y * inv;
  z = Z * inv;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// bk weight maps: Wt = n / v_cell ; Wc = n / (n_w (n_w P + s)) on n_w > 0   (src/covariances_pk.py:153-164)
__global__ void bk_weights_kernel(long long N, const double* __restrict__ n, const double* __restrict__ nw, double shot,
                                  double P, double v_cell, double* __restrict__ Wt, double* __restrict__ Wc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double nn = n[i], w = nw[i];
    Wt[i] = nn / v_cell;
    Wc[i] = (w > 0) ? nn / (w * (w * P + shot)) : 0.0;
  }
}

// out = a1*b1 - a2*b2   (a2 == nullptr: plain product)
__global__ void prod_diff_kernel(long long N, const double* __restrict__ a1, const double* __restrict__ b1,
                                 const double* __restrict__ a2, const double* __restrict__ b2, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = a2 ? a1[i] * b1[i] - a2[i] * b2[i] : a1[i] * b1[i];
}

// copy the pruning box of a half-spectrum into compact storage [Kx][Ky][Kz]
__global__ void extract_box_kernel(Geo g, const double2* __restrict__ spec, double2* __restrict__ out) {
  long long rows = (long long)g.Kx * g.Ky;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int jx = (int)(row / g.Ky), jy = (int)(row - (long long)jx * g.Ky);
    int ix = (jx <= g.bx) ? jx : g.nx - g.Kx + jx;
    int iy = (jy <= g.by) ? jy : g.ny - g.Ky + jy;
    long long base = ((long long)ix * g.ny + iy) * g.nzh;
    for (int iz = threadIdx.x; iz < g.Kz; iz += blockDim.x) out[row * g.Kz + iz] = spec[base + iz];
  }
}

// X[k in box] = coef * Y_lm(k^) * sum_t [bin(k)==s_t] P_t[k]   (lm < 0: no harmonic); X zeroed by the caller
__global__ void assemble_kernel(Geo g, const double2* __restrict__ p0, int s0, const double2* __restrict__ p1, int s1,
                                const double2* __restrict__ p2, int s2, int lm, double coef, double2* __restrict__ X) {
  long long rows = (long long)g.Kx * g.Ky;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int jx = (int)(row / g.Ky), jy = (int)(row - (long long)jx * g.Ky);
    int ix = (jx <= g.bx) ? jx : g.nx - g.Kx + jx;
    int iy = (jy <= g.by) ? jy : g.ny - g.Ky + jy;
    long long base = ((long long)ix * g.ny + iy) * g.nzh;
    double kx = g.kax[0][ix], ky = g.kax[1][iy];
    for (int iz = threadIdx.x; iz < g.Kz; iz += blockDim.x) {
      int b = g.binmap[base + iz];
      double2 acc = make_double2(0.0, 0.0);
      bool any = false;
      long long ci = row * g.Kz + iz;
      if (p0 && b == s0) {
        double2 v = p0[ci];
        acc.x += v.x;
        acc.y += v.y;
        any = true;
      }
      if (p1 && b == s1) {
        double2 v = p1[ci];
        acc.x += v.x;
        acc.y += v.y;
        any = true;
      }
      if (p2 && b == s2) {
        double2 v = p2[ci];
        acc.x += v.x;
        acc.y += v.y;
        any = true;
      }
      if (!any) {   // every box cell is written, so the box never holds stale data
        X[base + iz] = acc;
        continue;
      }
      double f = coef;
      if (lm >= 0) {
        double x, y, z;
        unit_vec_bk(kx, ky, g.kax[2][iz], x, y, z);
        f *= sww_ylm_dyn(lm, x, y, z);
      }
      acc.x *= f;
      acc.y *= f;
      X[base + iz] = acc;
    }
  }
}

// out = (first ? 0 : out) + scale * w(x) * [Y_lm(r^)] * u(x)
__global__ void acc_real_kernel(Geo g, const double* __restrict__ u, const double* __restrict__ w, int lm, double scale,
                                int first, double* __restrict__ out) {
  long long rows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    double X = g.rax[0][ix], Y = g.rax[1][iy];
    for (int iz = threadIdx.x; iz < g.nz; iz += blockDim.x) {
      long long i = row * g.nz + iz;
      double f = scale * (w ? w[i] : 1.0);
      if (lm >= 0) {
        double x, y, z;
        unit_vec_bk(X, Y, g.rax[2][iz], x, y, z);
        f *= sww_ylm_dyn(lm, x, y, z);
      }
      double v = f * u[i];
      out[i] = first ? v : out[i] + v;
    }
  }
}

__global__ void fill_complex_kernel(double2* x, long long n, double re, double im) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = make_double2(re, im);
}

// Triple-product sums over a list of terms:  part[block][t] = sum_{x in block's cells} m1 m2 m3
#define TS_MAX_TERMS 16
struct TripleTerms {
  const double* m1[TS_MAX_TERMS];
  const double* m2[TS_MAX_TERMS];
  const double* m3[TS_MAX_TERMS];
  int out_idx[TS_MAX_TERMS];
  double coef[TS_MAX_TERMS];
  int n;
};

__device__ __forceinline__ double warp_sum_bk(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
triple_sum_kernel(long long N, TripleTerms T, double* __restrict__ partial, unsigned int* __restrict__ counter,
                  double* __restrict__ out) {
  __shared__ double red[8][TS_MAX_TERMS];
  __shared__ bool is_last;
  double acc[TS_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < TS_MAX_TERMS; ++t) acc[t] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int t = 0; t < TS_MAX_TERMS; ++t)
      if (t < T.n) acc[t] += T.m1[t][i] * T.m2[t][i] * T.m3[t][i];
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < TS_MAX_TERMS; ++t) {
    double s = warp_sum_bk(acc[t]);
    if (lane == 0) red[warp][t] = s;
  }
  __syncthreads();
  if (threadIdx.x < TS_MAX_TERMS) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(long long)blockIdx.x * TS_MAX_TERMS + threadIdx.x] = s;
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
    if (threadIdx.x < TS_MAX_TERMS) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += partial[(long long)b * TS_MAX_TERMS + threadIdx.x];
      red[0][threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 0; t < T.n; ++t) out[T.out_idx[t]] += T.coef[t] * red[0][t];  // fixed order: deterministic
      *counter = 0;
    }
  }
}

static int run_triple_terms(sww_ctx* c, const TripleTerms& T, double* out) {
  SWW_PROF(c, "bk.triple_sum");
  int grid = SMS * 4;
  SWW_REQUIRE((size_t)grid * TS_MAX_TERMS <= c->partial_elems, "triple_sum: partial buffer too small");
  triple_sum_kernel<<<grid, 256, 0, c->stream>>>(c->g.N, T, c->d_partial, c->d_counter, out);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

struct TermBatcher {
  sww_ctx* c;
  double* out;
  TripleTerms T;
  TermBatcher(sww_ctx* c_, double* out_) : c(c_), out(out_) { T.n = 0; }
  int add(const double* a, const double* b, const double* d, int idx, double coef) {
    T.m1[T.n] = a;
    T.m2[T.n] = b;
    T.m3[T.n] = d;
    T.out_idx[T.n] = idx;
    T.coef[T.n] = coef;
    if (++T.n == TS_MAX_TERMS) return flush();
    return 0;
  }
  int flush() {
    if (T.n == 0) return 0;
    int r = run_triple_terms(c, T, out);
    T.n = 0;
    return r;
  }
};

// ------------------------------------------------------------------------------------------
// host-side pipelines
// ------------------------------------------------------------------------------------------
struct BkBufs {
  double2 *FC[SWW_MAX_L], *G[SWW_MAX_L], *X;
  double *tmp, *nCa, *nAa, *WS, *Wt, *Wc, *u;
  double2 *PS, *PST;  // compact product spectra [slot][box]
  double *Dt, *CD;    // tilde maps [n_b][N], C maps block [blk][N]
  double2* Qstack;    // backend 1: intermediate planes of a batched inverse (up to 2 lmax + 2)
  int blk;
};

static size_t box_elems(const sww_ctx* c) { return (size_t)c->g.Kx * c->g.Ky * c->g.Kz; }

// unique (i<=j, l) products used by the triangle list
struct PairTable {
  std::map<long long, int> slot;
  std::vector<int> i, j, l;
  int get(int a, int b, int l_) const {
    auto it = slot.find(((long long)a * 1024 + b) * 8 + l_);
    return it == slot.end() ? -1 : it->second;
  }
  void add(int a, int b, int l_) {
    long long key = ((long long)a * 1024 + b) * 8 + l_;
    if (slot.count(key)) return;
    slot[key] = (int)i.size();
    i.push_back(a);
    j.push_back(b);
    l.push_back(l_);
  }
};

static PairTable build_pairs(const sww_ctx* c) {
  PairTable P;
  for (int t = 0; t < c->n_tri; ++t) {
    int a = c->tris[3 * t], b = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
    P.add(a, b, 0);
    for (int l = 0; l < c->n_l; ++l) {
      P.add(a, cc, l);
      P.add(b, cc, l);
    }
  }
  return P;
}

static int choose_block(const sww_ctx* c) {
  int nb = c->n_tri * c->n_l;
  int blk = 64;
  if (blk > nb) blk = nb;
  return blk;
}

static void bk_fisher_layout(sww_ctx* c, Arena& ar, BkBufs& b, int n_slots) {
  size_t N = c->g.N, Nc = c->g.Nc;
  ar.reset(c->persist_bytes);
  for (int l = 0; l < c->n_l; ++l) b.FC[l] = ar.alloc<double2>(Nc);
  for (int l = 0; l < c->n_l; ++l) b.G[l] = ar.alloc<double2>(Nc);
  b.X = ar.alloc<double2>(Nc);
  b.tmp = ar.alloc<double>(N);
  b.nCa = ar.alloc<double>(N);
  b.nAa = ar.alloc<double>(N);
  b.WS = ar.alloc<double>(N);
  b.Wt = ar.alloc<double>(N);
  b.Wc = ar.alloc<double>(N);
  b.u = ar.alloc<double>(N);
  b.PS = ar.alloc<double2>((size_t)n_slots * box_elems(c));
  b.PST = ar.alloc<double2>((size_t)n_slots * box_elems(c));
  int nb = c->n_tri * c->n_l;
  b.blk = choose_block(c);
  b.Dt = ar.alloc<double>((size_t)nb * N);
  b.CD = ar.alloc<double>((size_t)b.blk * N);
  b.Qstack = ar.alloc<double2>(sww_sm100_qstack_bytes(c, SWW_MAX_ZBATCH) / sizeof(double2));
}

static void bk_small_layout(sww_ctx* c, Arena& ar, BkBufs& b) {
  // g-maps / Delta / data: spectra + a few real maps
  size_t N = c->g.N, Nc = c->g.Nc;
  ar.reset(c->persist_bytes);
  for (int l = 0; l < c->n_l; ++l) b.FC[l] = ar.alloc<double2>(Nc);
  for (int l = 0; l < c->n_l; ++l) b.G[l] = ar.alloc<double2>(Nc);
  b.X = ar.alloc<double2>(Nc);
  b.tmp = ar.alloc<double>(N);
  b.nCa = ar.alloc<double>(N);
  b.nAa = ar.alloc<double>(N);
  b.WS = ar.alloc<double>(N);
  // Delta_abc scratch: n_k bin maps + up to 9 harmonic maps, and a compact unit spectrum
  b.Dt = ar.alloc<double>((size_t)(c->g.n_k + 9) * N);
  b.PS = ar.alloc<double2>(box_elems(c));
  b.u = ar.alloc<double>(N);
  b.Qstack = ar.alloc<double2>(sww_sm100_qstack_bytes(c, SWW_MAX_ZBATCH) / sizeof(double2));
}

size_t sww_bk_pipeline_bytes(sww_ctx* c, int what) {
  Arena dry;
  dry.dry = true;
  size_t save = c->persist_bytes;
  c->persist_bytes = 0;
  BkBufs b;
  if (what == SWW_WS_BK_FISHER) {
    PairTable P = build_pairs(c);
    bk_fisher_layout(c, dry, b, (int)P.i.size());
    size_t r1 = dry.peak;
    bk_small_layout(c, dry, b);
    c->persist_bytes = save;
    return r1 > dry.peak ? r1 : dry.peak;
  }
  bk_small_layout(c, dry, b);
  c->persist_bytes = save;
  return dry.peak;
}

#define CHECK_READY(c)            \
  SWW_REQUIRE(c, "null context"); \
  SWW_REQUIRE((c)->ar.base, "no workspace bound (sww_bind_workspace)");
#define CHECK_FIELDS(c) SWW_REQUIRE(!(c)->fields_dirty && (c)->nbar, "fields not set (sww_set_fields)")
#define CHECK_ARENA(c) \
  SWW_REQUIRE((c)->ar.ok(), "workspace too small: pipeline needs %zu bytes, %zu bound", (c)->ar.off, (c)->ar.cap)
#define CHECK_TRIS(c) SWW_REQUIRE((c)->n_tri > 0, "triangle list not set (sww_bk_set_triangles)")

static int launch_ok(sww_ctx* c) {
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// g / tilde-g maps of one field into caller buffers [n_l][n_k][N]
static int gmaps(sww_ctx* c, BkBufs& b, const double* a, int data, double* g_out, double* tg_out) {
  const int n_k = c->g.n_k;
  const size_t N = c->g.N;
  const double invN = 1.0 / (double)c->g.N;
  if (data) {
    // f = n C^-1 d, spectra carry one MAS factor (bk/compute_bk_data.py:259-276); include_pix: f = n v D_w M d
    // and no MAS factor (:267-268)
    const double* d_in = a;
    if (c->include_pix) {
      SWW_TRY(sww_mas_filter_internal(c, a, +1, b.X, b.tmp));
      d_in = b.tmp;
    }
    SWW_TRY(k_cinv_fkp(c, d_in, c->nbar_w, c->v_cell, c->shot, c->P_fkp, b.nCa, c->nbar));
    SWW_TRY(sww_ylm_project_internal(c, b.nCa, c->include_pix ? 0 : 1, true, true, b.tmp, b.X, b.FC));
  } else {
    if (c->include_pix) {
      SWW_TRY(sww_mas_filter_internal(c, a, +1, b.X, b.nCa));
      SWW_TRY(k_div(c, a, c->nbar_A, b.nAa));
      SWW_TRY(sww_mas_filter_internal(c, b.nAa, -1, b.X, b.tmp));
      SWW_TRY(k_prep_pk(c, a, b.nCa, b.nAa, b.WS, b.nCa, b.tmp));
    } else {
      SWW_TRY(k_prep_pk(c, a, b.nCa, b.nAa, b.WS));
    }
    SWW_TRY(sww_ylm_project_internal(c, b.nCa, 0, true, true, b.tmp, b.X, b.FC));
    SWW_TRY(sww_ylm_project_internal(c, b.nAa, 0, true, true, b.tmp, b.X, b.G));
  }
  if (c->fft_backend == 1) {
    for (int l = 0; l < c->n_l; ++l)
      for (int i = 0; i < n_k; ++i) {
        SWW_TRY(sww_sm100_inverse_box(c, b.FC[l], i, -1, nullptr, -1, invN, 0, g_out + ((size_t)l * n_k + i) * N));
        c->n_fft++;
        if (!data) {
          SWW_TRY(sww_sm100_inverse_box(c, b.G[l], i, -1, nullptr, -1, invN, 0, tg_out + ((size_t)l * n_k + i) * N));
          c->n_fft++;
        }
      }
    return 0;
  }
  for (int l = 0; l < c->n_l; ++l)
    for (int i = 0; i < n_k; ++i) {
      SWW_TRY(k_shell_fill(c, b.FC[l], i, 1.0, b.X));
      SWW_TRY(sww_c2r(c, b.X, b.tmp));
      SWW_TRY(k_mul(c, b.tmp, nullptr, invN, g_out + ((size_t)l * n_k + i) * N));
      if (!data) {
        SWW_TRY(k_shell_fill(c, b.G[l], i, 1.0, b.X));
        SWW_TRY(sww_c2r(c, b.X, b.tmp));
        SWW_TRY(k_mul(c, b.tmp, nullptr, invN, tg_out + ((size_t)l * n_k + i) * N));
      }
    }
  return 0;
}

extern "C" int sww_bk_gmaps(sww_ctx* c, sww_devptr a, int data, sww_devptr g_out, sww_devptr tg_out) {
  CHECK_READY(c);
  CHECK_FIELDS(c);
  SWW_REQUIRE(a && g_out && (data || tg_out), "null device pointer");
  BkBufs b;
  bk_small_layout(c, c->ar, b);
  CHECK_ARENA(c);
  return gmaps(c, b, (const double*)a, data, (double*)g_out, (double*)tg_out);
}

static int assemble(sww_ctx* c, BkBufs& b, const double2* p0, int s0, const double2* p1, int s1, const double2* p2,
                    int s2, int lm, double coef) {
  SWW_PROF(c, "bk.assemble");
  // the sm_100a inverse chain only reads the pruning box; cuFFT reads (and destroys) the whole array
  if (c->fft_backend != 1) SWW_CUDA(cudaMemsetAsync(b.X, 0, sizeof(double2) * c->g.Nc, c->stream));
  assemble_kernel<<<SMS * 2, 64, 0, c->stream>>>(c->g, p0, s0, p1, s1, p2, s2, lm, coef, b.X);
  return launch_ok(c);
}

// out (+)= invN * wmap * [Y_lm(r^)] * IFT[X]; box_shell bounds the shells present in X
static int inverse_weighted(sww_ctx* c, BkBufs& b, int box_shell, const double* wmap, int lm, double scale, int first,
                            double* out) {
  if (c->fft_backend == 1) {
    c->n_fft++;
    return sww_sm100_inverse_box(c, b.X, -1, box_shell, wmap, lm, scale, first ? 0 : 1, out);
  }
  SWW_TRY(sww_c2r(c, b.X, b.u));
  acc_real_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, b.u, wmap, lm, scale, first, out);
  return launch_ok(c);
}

static int assemble(sww_ctx* c, BkBufs& b, const double2* p0, int s0, const double2* p1, int s1, const double2* p2,
                    int s2, int lm, double coef);

// include_pix = True: the MAS operators do not cancel for l > 0 because Y_lm(r^) multiplies OUTSIDE M^-1
// (bk/compute_bk_randoms.py:304-306, :344-346):
//     phi = 1/v [ M^-1(n psi_1) + 4pi/(2l+1) sum_m Y_lm(r^) M^-1(n psi_m) ],   C^-1 phi = v M D_w M phi.
// Every term costs one pruned inverse plus one full transform pair; rarely-used branch, not tuned.
static int build_phi_pix(sww_ctx* c, BkBufs& b, const PairTable& P, const double2* spectra, int alpha, bool apply_cinv,
                         double* out) {
  const int nt = c->n_tri;
  const int l_i = alpha / nt, t = alpha % nt;
  const int a = c->tris[3 * t], bb = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
  const size_t be = box_elems(c);
  const double invN = 1.0 / (double)c->g.N;
  auto spec = [&](int i, int j, int l) { return spectra + (size_t)P.get(i, j, l) * be; };
  // term: X assembled -> u = n IFT[X] -> M^-1 u -> out (+)= scale * Y_lm(r^) * (M^-1 u)
  auto term = [&](int lm, int first) -> int {
    if (c->fft_backend == 1) {
      SWW_TRY(sww_sm100_inverse_box(c, b.X, -1, -1, c->nbar, -1, invN, 0, b.u));
      c->n_fft++;
    } else {
      SWW_TRY(sww_c2r(c, b.X, b.tmp));
      SWW_TRY(k_mul(c, b.tmp, c->nbar, invN, b.u));
    }
    SWW_TRY(sww_mas_filter_internal(c, b.u, -1, b.X, b.u));
    acc_real_kernel<<<SMS * 8, 128, 0, c->stream>>>(c->g, b.u, nullptr, lm, 1.0 / c->v_cell, first, out);
    return launch_ok(c);
  };
  if (l_i == 0) {
    SWW_TRY(assemble(c, b, spec(a, bb, 0), cc, spec(a, cc, 0), bb, spec(bb, cc, 0), a, -1, 1.0));
    SWW_TRY(term(-1, 1));
  } else {
    SWW_TRY(assemble(c, b, spec(a, cc, l_i), bb, spec(bb, cc, l_i), a, nullptr, 0, -1, 1.0));
    SWW_TRY(term(-1, 1));
    int lm0 = 0;
    for (int l = 0; l < l_i; ++l) lm0 += 4 * l + 1;
    for (int m = 0; m < 4 * l_i + 1; ++m) {
      SWW_TRY(assemble(c, b, spec(a, bb, 0), cc, nullptr, 0, nullptr, 0, lm0 + m, lcoef(l_i)));
      SWW_TRY(term(lm0 + m, 0));
    }
  }
  if (apply_cinv) {   // v M D_w M phi
    SWW_TRY(sww_mas_filter_internal(c, out, +1, b.X, out));
    SWW_TRY(k_cinv_fkp(c, out, c->nbar_w, c->v_cell, c->shot, c->P_fkp, out, nullptr));
    SWW_TRY(sww_mas_filter_internal(c, out, +1, b.X, out));
  }
  return 0;
}

// one symmetrised phi difference map (kind = compact spectra PS or PST), weighted by wmap
static int build_phi(sww_ctx* c, BkBufs& b, const PairTable& P, const double2* spectra, int alpha, const double* wmap,
                     double* out) {
  const int nt = c->n_tri;
  const int l_i = alpha / nt, t = alpha % nt;
  const int a = c->tris[3 * t], bb = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
  const size_t be = box_elems(c);
  const double invN = 1.0 / (double)c->g.N;
  auto spec = [&](int i, int j, int l) { return spectra + (size_t)P.get(i, j, l) * be; };
  if (c->fft_backend == 1) {
    // fused: the first inverse pass gathers straight from the compact product spectra (no assembled
    // spectrum), and all (2l+1)+1 inverse transforms of the map are summed in the z pass (one write).
    GatherArgs ga[SWW_MAX_ZBATCH] = {};
    int lms[SWW_MAX_ZBATCH];
    int nbat = 0;
    if (l_i == 0) {
      ga[0].p[0] = spec(a, bb, 0); ga[0].s[0] = cc;
      ga[0].p[1] = spec(a, cc, 0); ga[0].s[1] = bb;
      ga[0].p[2] = spec(bb, cc, 0); ga[0].s[2] = a;
      ga[0].n_terms = 3; ga[0].lm = -1; ga[0].coef = 1.0;
      lms[0] = -1;
      nbat = 1;
    } else {
      ga[0].p[0] = spec(a, cc, l_i); ga[0].s[0] = bb;
      ga[0].p[1] = spec(bb, cc, l_i); ga[0].s[1] = a;
      ga[0].n_terms = 2; ga[0].lm = -1; ga[0].coef = 1.0;
      lms[0] = -1;
      nbat = 1;
      int lm0 = 0;
      for (int l = 0; l < l_i; ++l) lm0 += 4 * l + 1;
      for (int m = 0; m < 4 * l_i + 1; ++m, ++nbat) {
        ga[nbat].p[0] = spec(a, bb, 0); ga[nbat].s[0] = cc;
        ga[nbat].n_terms = 1; ga[nbat].lm = lm0 + m; ga[nbat].coef = lcoef(l_i);
        lms[nbat] = lm0 + m;
      }
    }
    c->n_fft += nbat;
    return sww_sm100_inverse_multi(c, nbat, ga, cc, lms, wmap, invN, 0, b.Qstack, out);
  }
  if (l_i == 0) {
    SWW_TRY(assemble(c, b, spec(a, bb, 0), cc, spec(a, cc, 0), bb, spec(bb, cc, 0), a, -1, 1.0));
    return inverse_weighted(c, b, cc, wmap, -1, invN, 1, out);
  }
  SWW_TRY(assemble(c, b, spec(a, cc, l_i), bb, spec(bb, cc, l_i), a, nullptr, 0, -1, 1.0));
  SWW_TRY(inverse_weighted(c, b, bb, wmap, -1, invN, 1, out));
  int lm0 = 0;
  for (int l = 0; l < l_i; ++l) lm0 += 4 * l + 1;
  for (int m = 0; m < 4 * l_i + 1; ++m) {
    SWW_TRY(assemble(c, b, spec(a, bb, 0), cc, nullptr, 0, nullptr, 0, lm0 + m, lcoef(l_i)));
    SWW_TRY(inverse_weighted(c, b, cc, wmap, lm0 + m, invN, 0, out));
  }
  return 0;
}

__global__ void bk_finalize_kernel(double* F, const double* __restrict__ delta, int n) {
  // F (upper triangle holds sum_x Dt_a CD_b) -> /24 * 4/(Delta_a Delta_b), mirrored to the lower triangle
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    int i = idx / n, j = idx - i * n;
    if (i <= j) {
      double v = F[i * n + j] / 12.0 / 2.0 * (4.0 / (delta[i] * delta[j]));
      F[i * n + j] = v;
      F[j * n + i] = v;
    }
  }
}

extern "C" int sww_bk_fisher_pair(sww_ctx* c, sww_devptr aA, sww_devptr aB, sww_devptr delta, sww_devptr gA,
                                  sww_devptr tgA, sww_devptr gB, sww_devptr tgB, sww_devptr fish_out) {
  CHECK_READY(c);
  CHECK_FIELDS(c);
  CHECK_TRIS(c);
  SWW_REQUIRE(aA && aB && delta && gA && tgA && gB && tgB && fish_out, "null device pointer");
  PairTable P = build_pairs(c);
  const int n_slots = (int)P.i.size();
  BkBufs b;
  bk_fisher_layout(c, c->ar, b, n_slots);
  CHECK_ARENA(c);
  const size_t N = c->g.N;
  const int n_k = c->g.n_k, nb = c->n_tri * c->n_l;
  const size_t be = box_elems(c);
  double *GA = (double*)gA, *TGA = (double*)tgA, *GB = (double*)gB, *TGB = (double*)tgB;
  double* F = (double*)fish_out;

  SWW_TRY(gmaps(c, b, (const double*)aA, 0, GA, TGA));
  SWW_TRY(gmaps(c, b, (const double*)aB, 0, GB, TGB));
  bk_weights_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, c->nbar, c->nbar_w, c->shot, c->P_fkp, c->v_cell, b.Wt, b.Wc);
  SWW_TRY(launch_ok(c));

  // product spectra of the A-B difference, kept inside the pruning box only
  for (int kind = 0; kind < 2; ++kind) {
    const double* XA = kind ? TGA : GA;
    const double* XB = kind ? TGB : GB;
    double2* dst = kind ? b.PST : b.PS;
    for (int s = 0; s < n_slots; ++s) {
      size_t o1 = ((size_t)0 * n_k + P.i[s]) * N, o2 = ((size_t)P.l[s] * n_k + P.j[s]) * N;
      {
        SWW_PROF(c, "bk.prod_diff");
        prod_diff_kernel<<<SMS * 8, 256, 0, c->stream>>>(c->g.N, XA + o1, XA + o2, XB + o1, XB + o2, b.tmp);
        SWW_TRY(launch_ok(c));
      }
      if (c->fft_backend == 1) {
        F3Args A = {};
        A.dst = b.X;
        SWW_TRY(sww_sm100_forward_box(c, b.tmp, nullptr, -1, 0, A));
        c->n_fft++;
      } else {
        SWW_TRY(sww_r2c(c, b.tmp, b.X));
      }
      extract_box_kernel<<<SMS * 2, 64, 0, c->stream>>>(c->g, b.X, dst + (size_t)s * be);
      SWW_TRY(launch_ok(c));
    }
  }
  // tilde-phi difference maps, all resident
  for (int al = 0; al < nb; ++al) {
    if (c->include_pix)
      SWW_TRY(build_phi_pix(c, b, P, b.PST, al, false, b.Dt + (size_t)al * N));
    else
      SWW_TRY(build_phi(c, b, P, b.PST, al, b.Wt, b.Dt + (size_t)al * N));
  }
  // C^-1 phi difference maps in column blocks; Gram against the resident tilde maps (rows alpha <= beta only)
  SWW_CUDA(cudaMemsetAsync(F, 0, sizeof(double) * (size_t)nb * nb, c->stream));
  for (int b0 = 0; b0 < nb; b0 += b.blk) {
    int nbk = (b0 + b.blk <= nb) ? b.blk : nb - b0;
    for (int j = 0; j < nbk; ++j) {
      if (c->include_pix)
        SWW_TRY(build_phi_pix(c, b, P, b.PS, b0 + j, true, b.CD + (size_t)j * N));
      else
        SWW_TRY(build_phi(c, b, P, b.PS, b0 + j, b.Wc, b.CD + (size_t)j * N));
    }
    SWW_TRY(sww_gram_upper(c, b.Dt, b.CD, b0 + nbk, nbk, (long long)N, b0 + nbk, F, nb, b0));
  }
  bk_finalize_kernel<<<sww_grid_for((long long)nb * nb, 256), 256, 0, c->stream>>>(F, (const double*)delta, nb);
  return launch_ok(c);
}

// Delta_abc (bk/compute_bk_randoms.py:427-473): data independent
extern "C" int sww_bk_delta(sww_ctx* c, sww_devptr delta_out) {
  CHECK_READY(c);
  CHECK_TRIS(c);
  SWW_REQUIRE(delta_out, "null device pointer");
  BkBufs b;
  bk_small_layout(c, c->ar, b);
  CHECK_ARENA(c);
  const int nt = c->n_tri, n_k = c->g.n_k, nb = nt * c->n_l;
  const size_t N = c->g.N;
  const double invN = 1.0 / (double)c->g.N;
  std::vector<double> D(nb, 0.0);
  for (int t = 0; t < nt; ++t) {
    int a = c->tris[3 * t], bb = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
    D[t] = (a == bb && bb == cc) ? 6.0 : ((a == bb || a == cc || bb == cc) ? 2.0 : 1.0);
  }
  if (c->n_l == 1) {
    SWW_CUDA(cudaMemcpyAsync((void*)delta_out, D.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, c->stream));
    SWW_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  // bins[a] = IFT[Theta_a]: a unit spectrum (compact box layout) restricted to shell a
  double* bins = b.Dt;
  double* Dl = b.Dt + (size_t)n_k * N;
  fill_complex_kernel<<<SMS, 256, 0, c->stream>>>(b.PS, (long long)box_elems(c), 1.0, 0.0);
  SWW_TRY(launch_ok(c));
  for (int a = 0; a < n_k; ++a) {
    SWW_TRY(assemble(c, b, b.PS, a, nullptr, 0, nullptr, 0, -1, 1.0));
    SWW_TRY(inverse_weighted(c, b, a, nullptr, -1, invN, 1, bins + (size_t)a * N));
  }
  // sums needed: for every (l>0, triangle with b==c): N0 and N_l
  // results accumulate into a device vector S: [2 * nb]  (N0 at 2j, Nl at 2j+1)
  double* S = b.nCa;  // reuse a real buffer as scratch
  SWW_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * 2 * (size_t)nb, c->stream));
  for (int l_i = 1; l_i < c->n_l; ++l_i) {
    int lm0 = 0;
    for (int l = 0; l < l_i; ++l) lm0 += 4 * l + 1;
    const int nm = 4 * l_i + 1;
    for (int cc = 0; cc < n_k; ++cc) {
      bool needed = false;
      for (int t = 0; t < nt; ++t)
        if (c->tris[3 * t + 1] == cc && c->tris[3 * t + 2] == cc) needed = true;
      if (!needed) continue;
      // Dl_c[m] = IFT[Theta_c Y_lm(k^)]
      for (int m = 0; m < nm; ++m) {
        SWW_TRY(assemble(c, b, b.PS, cc, nullptr, 0, nullptr, 0, lm0 + m, 1.0));
        SWW_TRY(inverse_weighted(c, b, cc, nullptr, -1, invN, 1, Dl + (size_t)m * N));
      }
      TermBatcher tb(c, S);
      for (int t = 0; t < nt; ++t) {
        int a = c->tris[3 * t], bb = c->tris[3 * t + 1], c2 = c->tris[3 * t + 2];
        if (!(bb == cc && c2 == cc)) continue;
        int j = t + l_i * nt;
        const double* ba = bins + (size_t)a * N;
        const double* bc = bins + (size_t)cc * N;
        SWW_TRY(tb.add(ba, bc, bc, 2 * j, 1.0));
        for (int m = 0; m < nm; ++m) SWW_TRY(tb.add(ba, Dl + (size_t)m * N, Dl + (size_t)m * N, 2 * j + 1, lcoef(l_i)));
      }
      SWW_TRY(tb.flush());
    }
  }
  std::vector<double> hS(2 * (size_t)nb);
  SWW_CUDA(cudaMemcpyAsync(hS.data(), S, sizeof(double) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
  SWW_CUDA(cudaStreamSynchronize(c->stream));
  for (int l_i = 1; l_i < c->n_l; ++l_i)
    for (int t = 0; t < nt; ++t) {
      int a = c->tris[3 * t], bb = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
      int j = t + l_i * nt;
      if (a != bb && bb != cc)
        D[j] = 1.0;
      else if (a == bb && bb != cc)
        D[j] = 2.0;
      else if (a != bb && bb == cc)
        D[j] = 1.0 + hS[2 * j + 1] / hS[2 * j];
      else
        D[j] = 2.0 * (1.0 + 2.0 * hS[2 * j + 1] / hS[2 * j]);
    }
  SWW_CUDA(cudaMemcpyAsync((void*)delta_out, D.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, c->stream));
  SWW_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// q[t][l] = T[a, b, (l, c)] of the triangle (a, b, c)
__global__ void q3_gather_kernel(const int* __restrict__ tris, int n_tri, int n_l, int n_k, const double* __restrict__ T,
                                 double* __restrict__ q) {
  const int nz = n_l * n_k;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_tri * n_l; idx += gridDim.x * blockDim.x) {
    const int t = idx / n_l, l = idx - t * n_l;
    const int a = tris[3 * t], b = tris[3 * t + 1], c = tris[3 * t + 2];
    q[idx] = T[((long long)a * n_k + b) * nz + l * n_k + c];
  }
}

// 3-field term (bk/compute_bk_data.py:350-367): q[t][l] = sum_x g_a^0 g_b^0 g_c^l
extern "C" int sww_bk_q3(sww_ctx* c, sww_devptr g_, sww_devptr q_out) {
  CHECK_READY(c);
  CHECK_TRIS(c);
  SWW_REQUIRE(g_ && q_out, "null device pointer");
  const double* g = (const double*)g_;
  const size_t N = c->g.N;
  const int n_k = c->g.n_k;
  double* q = (double*)q_out;
  if (n_k >= 4 && n_k <= 32 && N % 2 == 0) {
    // ONE trilinear Gram on the FP64 tensor cores, T[i, j, (l, c)] = sum_x g_i^0 g_j^0 g_c^l: every map is streamed once per operand
    // (2x the compulsory traffic) instead of three maps per triangle term (n_tri n_l terms: ~35x); the triangles pick their entries
    const int nz = c->n_l * n_k;
    const size_t tsz = (size_t)n_k * n_k * nz;
    if (c->q1_scratch_elems < 3 * tsz) {
      if (c->d_q1_scratch) cudaFree(c->d_q1_scratch);
      SWW_CUDA(cudaMalloc(&c->d_q1_scratch, 3 * tsz * sizeof(double)));
      c->q1_scratch_elems = 3 * tsz;
    }
    double* T = c->d_q1_scratch;
    SWW_TRY(sww_gram3(c, g, g, g, n_k, n_k, nz, (long long)N, T));
    q3_gather_kernel<<<sww_grid_for((long long)c->n_tri * c->n_l, 128), 128, 0, c->stream>>>(c->d_tris, c->n_tri, c->n_l, n_k, T, q);
    return launch_ok(c);
  }
  SWW_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * (size_t)c->n_tri * c->n_l, c->stream));
  TermBatcher tb(c, q);
  for (int t = 0; t < c->n_tri; ++t) {
    int a = c->tris[3 * t], b = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
    for (int l = 0; l < c->n_l; ++l)
      SWW_TRY(tb.add(g + (size_t)a * N, g + (size_t)b * N, g + ((size_t)l * n_k + cc) * N, t * c->n_l + l, 1.0));
  }
  return tb.flush();
}

// q[t][l] += w * (T1[a,b,lc] + T1[b,a,lc] + T2[a,b,lc] + T2[b,a,lc] + T3[a,b,lc] + T3[b,a,lc])
__global__ void q1_assemble_kernel(const int* __restrict__ tris, int n_tri, int n_l, int n_k, const double* __restrict__ T,
                                   double w, double* __restrict__ q) {
  const int nz = n_l * n_k;
  const long long tsz = (long long)n_k * n_k * nz;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_tri * n_l; idx += gridDim.x * blockDim.x) {
    int t = idx / n_l, l = idx - t * n_l;
    int a = tris[3 * t], b = tris[3 * t + 1], c = tris[3 * t + 2];
    int col = l * n_k + c;
    long long ab = ((long long)a * n_k + b) * nz + col, ba = ((long long)b * n_k + a) * nz + col;
    // the reference adds the two triple sums in this order (bk/compute_bk_data.py:402-403)
    double s1 = T[ab] + T[tsz + ba] + T[2 * tsz + ab];          // g_a gm_b tg_c + g_b gm_c tg_a + g_c gm_a tg_b
    double s2 = T[tsz + ab] + T[ba] + T[2 * tsz + ba];          // g_a tg_b gm_c + g_b tg_c gm_a + g_c tg_a gm_b
    q[idx] += w * s1;
    q[idx] += w * s2;
  }
}

// 1-field term of one MC simulation (bk/compute_bk_data.py:380-403)
extern "C" int sww_bk_q1_accumulate(sww_ctx* c, sww_devptr g_, sww_devptr gm_, sww_devptr tgm_, double inv_two_nmc,
                                    sww_devptr q_inout) {
  CHECK_READY(c);
  CHECK_TRIS(c);
  SWW_REQUIRE(g_ && gm_ && tgm_ && q_inout, "null device pointer");
  const double *g = (const double*)g_, *gm = (const double*)gm_, *tg = (const double*)tgm_;
  const size_t N = c->g.N;
  const int n_k = c->g.n_k;
  if (n_k >= 4 && n_k <= 32 && N % 2 == 0) {
    // three trilinear Grams on the FP64 tensor cores instead of 6 n_tri n_l full-map triple products:
    //   T1[i,j,(l,c)] = sum g_i gm_j tg_c^l ;  T2 = sum g_i tg_j gm_c^l ;  T3 = sum gm_i tg_j g_c^l
    const int nz = c->n_l * n_k;
    const size_t tsz = (size_t)n_k * n_k * nz;
    if (c->q1_scratch_elems < 3 * tsz) {
      if (c->d_q1_scratch) cudaFree(c->d_q1_scratch);
      SWW_CUDA(cudaMalloc(&c->d_q1_scratch, 3 * tsz * sizeof(double)));
      c->q1_scratch_elems = 3 * tsz;
    }
    double* T = c->d_q1_scratch;
    SWW_TRY(sww_gram3(c, g, gm, tg, n_k, n_k, nz, (long long)N, T));
    SWW_TRY(sww_gram3(c, g, tg, gm, n_k, n_k, nz, (long long)N, T + tsz));
    SWW_TRY(sww_gram3(c, gm, tg, g, n_k, n_k, nz, (long long)N, T + 2 * tsz));
    q1_assemble_kernel<<<sww_grid_for((long long)c->n_tri * c->n_l, 128), 128, 0, c->stream>>>(c->d_tris, c->n_tri, c->n_l, n_k, T,
                                                                                         -inv_two_nmc, (double*)q_inout);
    return launch_ok(c);
  }
  TermBatcher tb(c, (double*)q_inout);
  auto M = [&](const double* base, int l, int i) { return base + ((size_t)l * n_k + i) * N; };
  const double w = -inv_two_nmc;
  for (int t = 0; t < c->n_tri; ++t) {
    int a = c->tris[3 * t], b = c->tris[3 * t + 1], cc = c->tris[3 * t + 2];
    for (int l = 0; l < c->n_l; ++l) {
      int o = t * c->n_l + l;
      SWW_TRY(tb.add(M(g, 0, a), M(gm, 0, b), M(tg, l, cc), o, w));
      SWW_TRY(tb.add(M(g, 0, b), M(gm, l, cc), M(tg, 0, a), o, w));
      SWW_TRY(tb.add(M(g, l, cc), M(gm, 0, a), M(tg, 0, b), o, w));
      SWW_TRY(tb.add(M(g, 0, a), M(tg, 0, b), M(gm, l, cc), o, w));
      SWW_TRY(tb.add(M(g, 0, b), M(tg, l, cc), M(gm, 0, a), o, w));
      SWW_TRY(tb.add(M(g, l, cc), M(tg, 0, a), M(gm, 0, b), o, w));
    }
  }
  return tb.flush();
}
