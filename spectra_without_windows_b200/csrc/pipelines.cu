// P_l(k) pipelines and the drop-in operator entry points of libsww.
//
// Formulation (checked against the reference scripts to 1e-12 in tests):
//   F^C_l(k) = 4pi/(2l+1) sum_m Y_lm(k^) FT[n C^-1 a Y_lm(r^)](k),   G_l(k) likewise from A^-1 a
//   u_alpha  = IFT[Theta_a F^C_l],   U_alpha = W u_alpha,  W = n / (n_w (n_w P + s)) on n_w > 0
//   F_ab     = 1/(2 N v_cell) sum_{k in shell b} conj(FT[n U_alpha](k)) G_l'(k)          (Parseval)
//   qbar_a   = 1/(2 N)        sum_{k in shell a} F^C_l(k) conj(FT[W N A^-1 a](k))
// so one realisation costs 2M + 1 + 2 n_b transforms and no derivative map is ever stored
// (the reference materialises 2 n_b maps or spills them to disk: pk/compute_pk_randoms.py:220-274).
#include <cmath>

#include <algorithm>
#include <cstdlib>
#include "sww_internal.cuh"

static inline double lcoef(int l_i) { return 4.0 * M_PI / (4.0 * l_i + 1.0); }

// sum_m Y_lm(k^) FT[f Y_lm(r^)] for every multipole; tmp: real scratch [N]; spec: complex scratch [Nc]
int sww_ylm_project_internal(sww_ctx* c, const double* f, int mas_power, bool with_coef, bool pruned, double* tmp,
                             double2* spec, double2* const* out, const sww_ctx* out_mesh = nullptr) {
  int lm = 0;
  if (c->fft_backend == 1 && pruned) {
    // fused: Y_lm(r^) on the load of the z pass, Y_lm(k^) accumulation in the x-pass epilogue
    for (int l_i = 0; l_i < c->n_l; ++l_i) {
      int nm = 4 * l_i + 1;
      for (int m = 0; m < nm; ++m, ++lm) {
        F3Args A = {};
        A.dst = out[l_i];
        A.lm = lm;
        A.first = (m == 0);
        A.mas_power = mas_power;
        A.coef = (m == nm - 1) ? (with_coef ? lcoef(l_i) : 1.0) : 0.0;
        if (out_mesh) {   // the box modes go straight into spectra laid out for the coarse row mesh
          A.dnx = out_mesh->g.nx;
          A.dny = out_mesh->g.ny;
          A.dnzh = out_mesh->g.nzh;
        }
        SWW_TRY(sww_sm100_forward_box(c, f, nullptr, lm, 1, A));
        c->n_fft++;
      }
    }
    return 0;
  }
  for (int l_i = 0; l_i < c->n_l; ++l_i) {
    int nm = 4 * l_i + 1;
    for (int m = 0; m < nm; ++m, ++lm) {
      SWW_TRY(k_mul_ylm_r(c, f, lm, tmp));
      SWW_TRY(sww_r2c(c, tmp, spec));
      double coef = (m == nm - 1) ? (with_coef ? lcoef(l_i) : 1.0) : 0.0;
      SWW_TRY(k_acc_ylm_k(c, spec, lm, coef, mas_power, m == 0, pruned ? 1 : 0, out[l_i]));
    }
  }
  return 0;
}

struct PkFisherBufs {
  double *nCa, *nAa, *WS, *tmp, *row;
  double2 *FC[SWW_MAX_L], *G[SWW_MAX_L], *X;
  // shell-ordered rows (sm_100a backend): Us[shell_rows][S], Gs[n_l][S], partial sums of the segmented products
  double2 *Us = nullptr, *Gs = nullptr;
  double* sp = nullptr;
  long long S = 0;
};

int sww_rowmesh_prepare(sww_ctx* c);   // rowmesh.cu

// the context the Fisher rows run on: the coarse row mesh when there is one (rowmesh.cu), else the context itself
static inline sww_ctx* row_ctx(sww_ctx* c) { return (c->fft_backend == 1 && c->rowc) ? c->rowc : c; }

static void pk_fisher_layout(sww_ctx* c, Arena& ar, PkFisherBufs& b) {
  size_t N = c->g.N, Nc = c->g.Nc;
  sww_ctx* rc = row_ctx(c);
  ar.reset(c->persist_bytes);
  // F^C_l and G_l are only ever read inside the k_max box: with a row mesh they are stored in its (smaller) layout
  for (int l = 0; l < c->n_l; ++l) b.FC[l] = ar.alloc<double2>((size_t)rc->g.Nc);
  for (int l = 0; l < c->n_l; ++l) b.G[l] = ar.alloc<double2>((size_t)rc->g.Nc);
  // the fused sm_100a chains never materialise the shell-filtered spectrum or the real-space map
  const bool fused_only = c->fft_backend == 1 && c->tight_ws;
  b.X = fused_only ? nullptr : ar.alloc<double2>(Nc);
  b.tmp = fused_only ? nullptr : ar.alloc<double>(N);
  b.nCa = ar.alloc<double>(N);
  b.nAa = ar.alloc<double>(N);
  b.WS = ar.alloc<double>(N);
  b.row = ar.alloc<double>((size_t)c->n_l * c->g.n_k);
  const char* no_store = getenv("SWW_NO_SHELL_STORE");
  if (c->fft_backend == 1 && !(no_store && no_store[0] == '1')) {
    const long long S = sww_sm100_shell_cells(rc);
    if (S > 0 && c->shell_rows < 0) {
      // rows kept in shell order between two segmented-product launches: all n_b of them if they fit the budget
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      double budget = (c->tight_ws ? 0.10 : 0.20) * (double)total_b;
      if (const char* e = getenv("SWW_SHELL_STORE_GB")) budget = atof(e) * 1e9;
      long long rows = (long long)(budget / ((double)S * sizeof(double2)));
      const int n_b = c->n_l * c->g.n_k;
      if (rows > n_b) rows = n_b;
      c->shell_rows = (int)(rows / c->n_l) * c->n_l;
    }
    if (S > 0 && c->shell_rows >= c->n_l) {
      b.S = S;
      b.Us = ar.alloc<double2>((size_t)c->shell_rows * S);
      b.Gs = ar.alloc<double2>((size_t)c->n_l * S);
      b.sp = ar.alloc<double>((size_t)sww_sm100_shell_chunks(rc) * c->shell_rows * 3);
    }
  }
}

struct PkDataBufs {
  double *f, *tmp;
  double2 *FD[SWW_MAX_L], *R, *spec;
};

static void pk_data_layout(sww_ctx* c, Arena& ar, PkDataBufs& b) {
  size_t N = c->g.N, Nc = c->g.Nc;
  ar.reset(c->persist_bytes);
  for (int l = 0; l < c->n_l; ++l) b.FD[l] = ar.alloc<double2>(Nc);
  b.R = ar.alloc<double2>(Nc);
  b.spec = ar.alloc<double2>(Nc);
  b.f = ar.alloc<double>(N);
  b.tmp = ar.alloc<double>(N);
}

struct OpsBufs {
  double *f, *tmp;
  double2 *spec, *F[SWW_MAX_L];
};

static void ops_layout(sww_ctx* c, Arena& ar, OpsBufs& b) {
  size_t N = c->g.N, Nc = c->g.Nc;
  ar.reset(c->persist_bytes);
  b.f = ar.alloc<double>(N);
  b.tmp = ar.alloc<double>(N);
  b.spec = ar.alloc<double2>(Nc);
  for (int l = 0; l < c->n_l; ++l) b.F[l] = ar.alloc<double2>(Nc);
}

size_t sww_bk_pipeline_bytes(sww_ctx* c, int what);  // bk.cu
size_t sww_ml_pipeline_bytes(sww_ctx* c);           // ml.cu

size_t sww_pipeline_bytes(sww_ctx* c, int what) {
  Arena dry;
  dry.dry = true;
  size_t save = c->persist_bytes;
  c->persist_bytes = 0;
  size_t r = 0;
  if (what == SWW_WS_PK_FISHER) {
    PkFisherBufs b;
    pk_fisher_layout(c, dry, b);
    r = dry.peak;
  } else if (what == SWW_WS_PK_DATA) {
    PkDataBufs b;
    pk_data_layout(c, dry, b);
    r = dry.peak;
  } else if (what == SWW_WS_OPS) {
    OpsBufs b;
    ops_layout(c, dry, b);
    r = dry.peak;
  } else if (what == SWW_WS_PK_ML) {
    r = sww_ml_pipeline_bytes(c);
  } else {
    r = sww_bk_pipeline_bytes(c, what);
  }
  c->persist_bytes = save;
  return r;
}

#define CHECK_READY(c)                                                       \
  SWW_REQUIRE(c, "null context");                                            \
  SWW_REQUIRE((c)->ar.base, "no workspace bound (sww_bind_workspace)");

#define CHECK_FIELDS(c) SWW_REQUIRE(!(c)->fields_dirty && (c)->nbar, "fields not set (sww_set_fields)")

#define CHECK_ARENA(c) \
  SWW_REQUIRE((c)->ar.ok(), "workspace too small: pipeline needs %zu bytes, %zu bound", (c)->ar.off, (c)->ar.cap)

// ------------------------------------------------------------------------------------------
extern "C" int sww_pk_fisher_realisation(sww_ctx* c, sww_devptr a_, sww_devptr fish_out, sww_devptr qbar_out) {
  CHECK_READY(c);
  CHECK_FIELDS(c);
  SWW_REQUIRE(a_ && fish_out && qbar_out, "null device pointer");
  if (c->fft_backend == 1 && c->rowc && !c->rowc->row_ready) SWW_TRY(sww_rowmesh_prepare(c));
  PkFisherBufs b;
  pk_fisher_layout(c, c->ar, b);
  CHECK_ARENA(c);
  sww_ctx* rc = row_ctx(c);
  const sww_ctx* om = rc != c ? rc : nullptr;
  if (om) SWW_REQUIRE(rc->row_ready, "row mesh: the filtered weight map could not be built (workspace too small)");
  const double* a = (const double*)a_;
  double* F = (double*)fish_out;
  double* qbar = (double*)qbar_out;
  const int n_k = c->g.n_k, n_l = c->n_l, n_b = n_k * n_l;
  const double N = (double)c->g.N;

  if (c->include_pix) {
    // a_c = M a ; ainv = M^-1 (a / n_A): two extra full transforms pairs, everything else is unchanged
    SWW_REQUIRE(b.X && b.tmp, "include_pix needs the full workspace (unset SWW_TIGHT_WS)");
    SWW_TRY(sww_mas_filter_internal(c, a, +1, b.X, b.nCa));
    SWW_TRY(k_div(c, a, c->nbar_A, b.nAa));
    SWW_TRY(sww_mas_filter_internal(c, b.nAa, -1, b.X, b.tmp));
    SWW_TRY(k_prep_pk(c, a, b.nCa, b.nAa, b.WS, b.nCa, b.tmp));
  } else {
    SWW_TRY(k_prep_pk(c, a, b.nCa, b.nAa, b.WS));
  }
  SWW_TRY(sww_ylm_project_internal(c, b.nCa, 0, true, true, b.tmp, b.X, b.FC, om));
  SWW_TRY(sww_ylm_project_internal(c, b.nAa, 0, true, true, b.tmp, b.X, b.G, om));
  // q-bar: one transform of W N A^-1 a, then one binning pass against F^C_l
  if (c->fft_backend == 1) {
    F3Args A = {};
    for (int q = 0; q < n_l; ++q) A.G[q] = b.FC[q];
    A.n_g = n_l;
    A.scale = 0.5 / N;
    A.out_row = qbar;
    if (om) {
      A.dnx = om->g.nx;
      A.dny = om->g.ny;
      A.dnzh = om->g.nzh;
    }
    SWW_TRY(sww_sm100_forward_box(c, b.WS, nullptr, -1, 2, A));
    c->n_fft++;
    // Fisher rows: I1 -> I2 -> Z(c2r * nW * r2c) -> F2 -> F3(bin); nothing but the row leaves the chip
    // the n_l rows of one k bin share the shell box and the weight map: one batched chain per bin
    // the rows run on rc (the coarse row mesh, or c itself): IFT normalised by the full mesh, and
    // FT_M[w~ u] = (N_M / N) FT[nW u] inside the box, so the bin sums carry N / N_M
    const double row_scale = 0.5 / ((double)rc->g.N * c->v_cell);
    const unsigned long long own0 = rc->n_own;
    int rr = 0;
    if (b.Us) {
      // rows go to HBM in shell order; every block of shells ends with one segmented-product launch
      rr = sww_sm100_shell_sort_g(rc, (const double2* const*)b.G, n_l, b.Gs);
      const int spb = c->shell_rows / n_l;
      for (int ka0 = 0; ka0 < n_k && rr == 0; ka0 += spb) {
        const int nsh = std::min(spb, n_k - ka0);
        for (int k = 0; k < nsh && rr == 0; ++k) {
          rr = sww_sm100_pk_rows_store(rc, n_l, (const double2* const*)b.FC, ka0 + k, rc->d_nW, 1.0 / N,
                                       b.Us + (size_t)k * n_l * b.S);
          c->n_fft += 2 * n_l;
        }
        if (rr == 0) rr = sww_sm100_shell_dot(rc, b.Us, nsh * n_l, b.Gs, n_l, row_scale, b.sp, F, n_l, ka0);
      }
    } else {
      for (int ka = 0; ka < n_k && rr == 0; ++ka) {
        double* rows[SWW_MAX_L];
        for (int l_i = 0; l_i < n_l; ++l_i) rows[l_i] = F + (size_t)(l_i * n_k + ka) * n_b;
        rr = sww_sm100_pk_rows(rc, n_l, (const double2* const*)b.FC, ka, rc->d_nW, 1.0 / N, (const double2* const*)b.G, n_l,
                               row_scale, rows);
        c->n_fft += 2 * n_l;
      }
    }
    if (om) c->n_own += rc->n_own - own0;
    SWW_TRY(rr);
    SWW_TRY(k_symmetrise(c, F, n_b));
    return 0;
  }
  SWW_TRY(sww_r2c(c, b.WS, b.X));
  SWW_TRY(k_shell_bin(c, b.X, (const double2* const*)b.FC, n_l, 0.5 / N, qbar));
  // Fisher rows
  for (int alpha = 0; alpha < n_b; ++alpha) {
    int l_i = alpha / n_k, ka = alpha % n_k;
    SWW_TRY(k_shell_fill(c, b.FC[l_i], ka, 1.0, b.X));
    SWW_TRY(sww_c2r(c, b.X, b.tmp));
    SWW_TRY(k_weight(c, b.tmp, c->d_nW, 1.0 / N));
    SWW_TRY(sww_r2c(c, b.tmp, b.X));
    SWW_TRY(k_shell_bin(c, b.X, (const double2* const*)b.G, n_l, 0.5 / (N * c->v_cell), F + (size_t)alpha * n_b));
  }
  SWW_TRY(k_symmetrise(c, F, n_b));
  return 0;
}

extern "C" int sww_pk_data_q(sww_ctx* c, sww_devptr diff_, sww_devptr q_out) {
  CHECK_READY(c);
  CHECK_FIELDS(c);
  SWW_REQUIRE(diff_ && q_out, "null device pointer");
  PkDataBufs b;
  pk_data_layout(c, c->ar, b);
  CHECK_ARENA(c);
  const double N = (double)c->g.N;
  // f = n C^-1 d   (include_pix: f = n v D_w M d, and no MAS^2 factor: src/covariances_pk.py:386-389)
  const double* d_in = (const double*)diff_;
  if (c->include_pix) {
    SWW_TRY(sww_mas_filter_internal(c, d_in, +1, b.spec, b.tmp));
    d_in = b.tmp;
  }
  SWW_TRY(k_cinv_fkp(c, d_in, c->nbar_w, c->v_cell, c->shot, c->P_fkp, b.f, c->nbar));
  SWW_TRY(sww_ylm_project_internal(c, b.f, c->include_pix ? 0 : 2, true, true, b.tmp, b.spec, b.FD));
  if (c->fft_backend == 1) {
    F3Args A = {};
    for (int q = 0; q < c->n_l; ++q) A.G[q] = b.FD[q];
    A.n_g = c->n_l;
    A.scale = 0.5 / (N * c->v_cell);
    A.out_row = (double*)q_out;
    c->n_fft++;
    return sww_sm100_forward_box(c, b.f, nullptr, -1, 2, A);
  }
  SWW_TRY(sww_r2c(c, b.f, b.R));
  SWW_TRY(k_shell_bin(c, b.R, (const double2* const*)b.FD, c->n_l, 0.5 / (N * c->v_cell), (double*)q_out));
  return 0;
}

// ------------------------------------------------------------------------------------------
// drop-in operators
// ------------------------------------------------------------------------------------------
extern "C" int sww_fft_c2c(sww_ctx* c, sww_devptr in, sww_devptr out, int inverse) {
  SWW_REQUIRE(c && in && out, "null argument");
  return sww_c2c(c, (const double2*)in, (double2*)out, inverse);
}

extern "C" int sww_fft_r2c(sww_ctx* c, sww_devptr in, sww_devptr out) {
  CHECK_READY(c);
  return sww_r2c(c, (const double*)in, (double2*)out);
}

extern "C" int sww_fft_c2r(sww_ctx* c, sww_devptr in, sww_devptr out) {
  CHECK_READY(c);
  SWW_TRY(sww_c2r(c, (double2*)in, (double*)out));
  return k_mul(c, (const double*)out, nullptr, 1.0 / (double)c->g.N, (double*)out);
}

extern "C" int sww_apply_cinv_fkp(sww_ctx* c, sww_devptr x, sww_devptr nbar_w, double v_cell, double shot_fac,
                                  double P_fkp, sww_devptr out) {
  SWW_REQUIRE(c && x && nbar_w && out, "null argument");
  return k_cinv_fkp(c, (const double*)x, (const double*)nbar_w, v_cell, shot_fac, P_fkp, (double*)out, nullptr);
}

extern "C" int sww_apply_n(sww_ctx* c, sww_devptr x, sww_devptr nbar, double v_cell, double shot_fac, sww_devptr out) {
  SWW_REQUIRE(c && x && nbar && out, "null argument");
  return k_apply_n(c, (const double*)x, (const double*)nbar, v_cell, shot_fac, (double*)out);
}

extern "C" int sww_ylm_project(sww_ctx* c, sww_devptr x, int mas_power, sww_devptr out) {
  CHECK_READY(c);
  SWW_REQUIRE(x && out, "null argument");
  SWW_REQUIRE(mas_power >= 0 && mas_power <= 2, "mas_power must be 0, 1 or 2");
  OpsBufs b;
  ops_layout(c, c->ar, b);
  CHECK_ARENA(c);
  double2* outs[SWW_MAX_L];
  for (int l = 0; l < c->n_l; ++l) outs[l] = (double2*)out + (size_t)l * c->g.Nc;
  return sww_ylm_project_internal(c, (const double*)x, mas_power, false, false, b.tmp, b.spec, outs);
}

extern "C" int sww_apply_c_alpha_single(sww_ctx* c, sww_devptr x, sww_devptr nbar, double v_cell, int l_i, int a,
                                        int data, sww_devptr out) {
  CHECK_READY(c);
  SWW_REQUIRE(x && nbar && out, "null argument");
  SWW_REQUIRE(l_i >= 0 && l_i < c->n_l && a >= 0 && a < c->g.n_k, "multipole / bin index out of range");
  OpsBufs b;
  ops_layout(c, c->ar, b);
  CHECK_ARENA(c);
  SWW_TRY(k_mul(c, (const double*)x, (const double*)nbar, 1.0, b.f));
  // only the requested multipole
  int lm0 = 0;
  for (int l = 0; l < l_i; ++l) lm0 += 4 * l + 1;
  int nm = 4 * l_i + 1;
  for (int m = 0; m < nm; ++m) {
    SWW_TRY(k_mul_ylm_r(c, b.f, lm0 + m, b.tmp));
    SWW_TRY(sww_r2c(c, b.tmp, b.spec));
    double coef = (m == nm - 1) ? lcoef(l_i) : 0.0;
    SWW_TRY(k_acc_ylm_k(c, b.spec, lm0 + m, coef, data ? 2 : 0, m == 0, 1, b.F[0]));
  }
  if (c->fft_backend == 1) {
    c->n_fft++;
    return sww_sm100_inverse_box(c, b.F[0], a, -1, (const double*)nbar, -1, 1.0 / ((double)c->g.N * v_cell), 0, (double*)out);
  }
  SWW_TRY(k_shell_fill(c, b.F[0], a, 1.0, b.spec));
  SWW_TRY(sww_c2r(c, b.spec, b.tmp));
  return k_mul(c, b.tmp, (const double*)nbar, 1.0 / ((double)c->g.N * v_cell), (double*)out);
}

extern "C" int sww_paint(sww_ctx* c, sww_devptr pos, sww_devptr w, uint64_t n, const double box_center[3], int scheme,
                         double scale, sww_devptr out) {
  SWW_REQUIRE(c && out && box_center, "null argument");
  SWW_REQUIRE(scheme == 2 || scheme == 3, "scheme must be 2 (CIC) or 3 (TSC)");
  SWW_REQUIRE(n == 0 || (pos && w), "null catalogue pointer");
  return k_paint(c, (const double*)pos, (const double*)w, n, box_center, scheme, scale, (double*)out);
}

extern "C" int sww_set_include_pix(sww_ctx* c, int include_pix) {
  SWW_REQUIRE(c, "null context");
  c->include_pix = include_pix != 0;
  return 0;
}

extern "C" int sww_mas_filter(sww_ctx* c, sww_devptr x, int power, sww_devptr out) {
  CHECK_READY(c);
  SWW_REQUIRE(x && out && (power == 1 || power == -1), "mas_filter: bad argument");
  OpsBufs b;
  ops_layout(c, c->ar, b);
  CHECK_ARENA(c);
  return sww_mas_filter_internal(c, (const double*)x, power, b.spec, (double*)out);
}

// Mean over the Monte-Carlo realisations of the summed F and q-bar (pk/compute_pk_data.py:227-247, the fold of the per-seed
// files) and a final symmetrisation, in place on the device -- what remains of SURVEY's K9 once the sum over ranks is done.
// The sum over ranks itself is ONE all-reduce of n_b^2 + n_b doubles issued by the host through torch.distributed (NCCL):
// the communicator belongs to the host framework, so libsww does not link NCCL (Boundary).
__global__ void mc_finalize_kernel(double* __restrict__ F, double* __restrict__ q, int n, double inv) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    if (i <= j) {
      const double v = 0.5 * (F[i * n + j] + F[j * n + i]) * inv;
      F[i * n + j] = v;
      F[j * n + i] = v;
    }
    if (q && idx < n) q[idx] *= inv;
  }
}

extern "C" int sww_mc_finalize(sww_ctx* c, sww_devptr fish_sum, sww_devptr qbar_sum, int32_t n_b, int32_t n_realisations) {
  SWW_REQUIRE(c && fish_sum && n_b > 0 && n_realisations > 0, "mc_finalize: bad argument");
  mc_finalize_kernel<<<sww_grid_for((long long)n_b * n_b, 256), 256, 0, c->stream>>>((double*)fish_sum, (double*)qbar_sum, n_b,
                                                                                    1.0 / (double)n_realisations);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}
