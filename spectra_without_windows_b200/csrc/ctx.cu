// Context life-cycle, workspace layout and the small C-ABI entry points of libsww.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "sww_internal.cuh"

static thread_local char g_err[1024] = "";

void sww_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sww_last_error(void) { return g_err; }
extern "C" int sww_abi_version(void) { return SWW_ABI_VERSION; }

static int max_abs_index_below(const double* kax, int n, double kmax) {
  // largest |signed index| whose |k| can fall below kmax (modes at or above kmax are in no bin)
  int best = 0;
  for (int i = 0; i < n; ++i) {
    int s = (i >= n / 2.0) ? i - n : i;
    if (std::fabs(kax[i]) < kmax) best = std::max(best, std::abs(s));
  }
  return best;
}

int sww_rowmesh_create(sww_ctx* c, const sww_geometry* geo);   // rowmesh.cu
int sww_rowmesh_prepare(sww_ctx* c);                            // rowmesh.cu

static int ctx_fill(sww_ctx* c, const sww_geometry* geo, bool child) {
  c->device = geo->device;
  c->is_child = child;
  {
    const char* e = getenv("SWW_TIGHT_WS");
    c->tight_ws = e && e[0] == '1';
  }
  Geo& g = c->g;
  g.nx = geo->grid[0];
  g.ny = geo->grid[1];
  g.nz = geo->grid[2];
  g.nzh = g.nz / 2 + 1;
  g.N = (long long)g.nx * g.ny * g.nz;
  g.Nc = (long long)g.nx * g.ny * g.nzh;
  g.n_k = geo->n_k;
  for (int d = 0; d < 3; ++d) c->box[d] = geo->box[d];
  c->v_cell = 1.0 * (geo->box[0] * geo->box[1] * geo->box[2]) / (1.0 * ((double)g.nx * g.ny * g.nz));
  c->lmax = geo->lmax;
  c->n_l = geo->lmax / 2 + 1;
  c->n_lm = 0;
  for (int l = 0; l <= geo->lmax; l += 2) c->n_lm += 2 * l + 1;
  c->h_klo.assign(geo->k_lo, geo->k_lo + geo->n_k);
  c->h_khi.assign(geo->k_hi, geo->k_hi + geo->n_k);
  double kmax = *std::max_element(c->h_khi.begin(), c->h_khi.end());
  // Hermitian (R2C) evaluation requires every binned mode to lie strictly below Nyquist on every axis
  for (int d = 0; d < 3; ++d) {
    double kny = M_PI * geo->grid[d] / geo->box[d];
    SWW_REQUIRE(kmax < kny, "k_max=%.6g is not below the Nyquist frequency %.6g of axis %d", kmax, kny, d);
  }
  g.bx = max_abs_index_below(geo->kax[0], g.nx, kmax);
  g.by = max_abs_index_below(geo->kax[1], g.ny, kmax);
  g.bz = max_abs_index_below(geo->kax[2], g.nz, kmax);
  g.Kx = std::min(2 * g.bx + 1, g.nx);
  g.Ky = std::min(2 * g.by + 1, g.ny);
  g.Kz = std::min(g.bz + 1, g.nzh);
  // tables: rax(3) kax(3) mas(3) klo khi
  size_t nt = 3 * (size_t)(g.nx + g.ny + g.nz) + 2 * (size_t)g.n_k;
  std::vector<double> h(nt);
  size_t o = 0;
  double* dptr[11];
  const double* src[9] = {geo->rax[0], geo->rax[1], geo->rax[2], geo->kax[0], geo->kax[1],
                          geo->kax[2], geo->mas[0], geo->mas[1], geo->mas[2]};
  int len[9] = {g.nx, g.ny, g.nz, g.nx, g.ny, g.nz, g.nx, g.ny, g.nz};
  SWW_CUDA(cudaMalloc(&c->d_tables, nt * sizeof(double)));
  for (int i = 0; i < 9; ++i) {
    memcpy(&h[o], src[i], len[i] * sizeof(double));
    dptr[i] = c->d_tables + o;
    o += len[i];
  }
  memcpy(&h[o], geo->k_lo, g.n_k * sizeof(double));
  c->d_klo = c->d_tables + o;
  o += g.n_k;
  memcpy(&h[o], geo->k_hi, g.n_k * sizeof(double));
  c->d_khi = c->d_tables + o;
  o += g.n_k;
  SWW_CUDA(cudaMemcpy(c->d_tables, h.data(), nt * sizeof(double), cudaMemcpyHostToDevice));
  for (int d = 0; d < 3; ++d) {
    g.rax[d] = dptr[d];
    g.kax[d] = dptr[3 + d];
    g.mas[d] = dptr[6 + d];
  }
  SWW_CUDA(cudaMalloc(&c->d_binmap, (size_t)g.Nc));
  g.binmap = c->d_binmap;
  c->partial_elems = (size_t)SMS * 4 * SWW_MAX_L * SWW_MAX_BINS;
  SWW_CUDA(cudaMalloc(&c->d_partial, c->partial_elems * sizeof(double)));
  SWW_CUDA(cudaMalloc(&c->d_counter, 16 * sizeof(unsigned int)));
  SWW_CUDA(cudaMemset(c->d_counter, 0, 16 * sizeof(unsigned int)));
  SWW_TRY(k_build_binmap(c));
  if (child) {
    // the row mesh only ever runs the hand-written passes: no cuFFT plans
    c->fft_work_bytes = sww_fft_sm100_work_bytes(c);
    c->fft_backend = 1;
  } else {
    SWW_TRY(sww_fft_init(c));
    SWW_TRY(sww_rowmesh_create(c, geo));
  }
  SWW_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int sww_ctx_create_internal(const sww_geometry* geo, sww_ctx** out, bool child) {
  SWW_REQUIRE(geo && out, "null argument");
  SWW_REQUIRE(geo->lmax >= 0 && geo->lmax % 2 == 0 && geo->lmax <= 2 * (SWW_MAX_L - 1),
              "l-max must be even and <= %d", 2 * (SWW_MAX_L - 1));
  SWW_REQUIRE(geo->n_k >= 1 && geo->n_k <= SWW_MAX_BINS, "n_k must be in [1,%d]", SWW_MAX_BINS);
  for (int d = 0; d < 3; ++d) SWW_REQUIRE(geo->grid[d] >= 4, "grid too small");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  SWW_REQUIRE(e == cudaSuccess && ndev > 0, "no CUDA device: libsww has no CPU path (%s)", cudaGetErrorString(e));
  SWW_CUDA(cudaSetDevice(geo->device));
  sww_ctx* c = new sww_ctx();
  int r = ctx_fill(c, geo, child);
  if (r != 0) {
    sww_ctx_destroy(c);   // every failure path releases the tables, plans and the child context
    return r;
  }
  *out = c;
  return 0;
}

extern "C" int sww_ctx_create(const sww_geometry* geo, sww_ctx** out) { return sww_ctx_create_internal(geo, out, false); }

extern "C" int sww_ctx_destroy(sww_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->rowc) sww_ctx_destroy(c->rowc);
  c->rowc = nullptr;
  sww_fft_destroy(c);
  sww_rng_destroy(c);
  cudaFree(c->d_tables);
  cudaFree(c->d_binmap);
  cudaFree(c->d_partial);
  cudaFree(c->d_counter);
  if (c->d_tris) cudaFree(c->d_tris);
  if (c->d_q1_scratch) cudaFree(c->d_q1_scratch);
  if (c->d_paint_scratch) cudaFree(c->d_paint_scratch);
  if (c->d_gram_partial) cudaFree(c->d_gram_partial);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  delete c;
  return 0;
}

extern "C" int sww_ctx_set_stream(sww_ctx* c, sww_stream s) {
  SWW_REQUIRE(c, "null context");
  c->stream = (cudaStream_t)s;
  if (c->fft_work || c->fft_work_bytes == 0) SWW_TRY(sww_fft_bind_work(c));
  if (c->rowc) SWW_TRY(sww_ctx_set_stream(c->rowc, s));
  return 0;
}

extern "C" int sww_launch_counts(sww_ctx* c, uint64_t* own, uint64_t* fft) {
  SWW_REQUIRE(c, "null context");
  if (own) *own = c->n_own;
  if (fft) *fft = c->n_fft;
  return 0;
}

extern "C" int sww_set_fft_backend(sww_ctx* c, int backend) {
  SWW_REQUIRE(c, "null context");
  SWW_REQUIRE(backend == 0 || backend == 1, "unknown FFT backend %d", backend);
  SWW_REQUIRE(!(c->tight_ws && c->ar.base && backend != c->fft_backend), "SWW_TIGHT_WS=1: the FFT backend is fixed once the workspace is bound");
  if (backend == 1) SWW_REQUIRE(sww_fft_sm100_supported(c), "the sm_100a FFT passes need power-of-two axes (nx, ny in [64,1024], nz in [128,1024]) or axes of at most 640 points");
  c->fft_backend = backend;
  return 0;
}

extern "C" int sww_row_mesh(sww_ctx* c, int32_t grid[3]) {
  SWW_REQUIRE(c && grid, "null argument");
  const bool on = c->rowc && c->fft_backend == 1;
  grid[0] = on ? c->rowc->g.nx : 0;
  grid[1] = on ? c->rowc->g.ny : 0;
  grid[2] = on ? c->rowc->g.nz : 0;
  return 0;
}

static size_t persist_own(sww_ctx* c) {
  size_t nW = ((size_t)c->g.N * sizeof(double) + 255) & ~size_t(255);
  size_t fw = (c->fft_work_bytes + 255) & ~size_t(255);
  return nW + fw;
}

// [nW map][FFT work area] of this context, then the same of the coarse row mesh
static size_t persist_layout(sww_ctx* c) { return persist_own(c) + (c->rowc ? persist_own(c->rowc) : 0); }

size_t sww_pipeline_bytes(sww_ctx* c, int what);  // pipelines.cu

extern "C" int sww_workspace_bytes(sww_ctx* c, int what, uint64_t* bytes) {
  SWW_REQUIRE(c && bytes, "null argument");
  *bytes = persist_layout(c) + sww_pipeline_bytes(c, what) + 4096;
  return 0;
}

extern "C" int sww_bind_workspace(sww_ctx* c, sww_devptr ws, uint64_t bytes) {
  SWW_REQUIRE(c, "null context");
  SWW_REQUIRE(ws % 256 == 0, "workspace must be 256-byte aligned");
  size_t p = persist_layout(c);
  SWW_REQUIRE(bytes >= p, "workspace too small: %llu < %zu", (unsigned long long)bytes, p);
  c->ar.base = (char*)ws;
  c->ar.cap = bytes;
  c->persist_bytes = p;
  c->d_nW = (double*)ws;
  c->fft_work = (char*)ws + (((size_t)c->g.N * sizeof(double) + 255) & ~size_t(255));
  c->fft_work_cap = c->fft_work_bytes;
  c->fields_dirty = true;
  SWW_TRY(sww_fft_bind_work(c));
  if (c->rowc) {
    sww_ctx* r = c->rowc;
    char* rb = (char*)ws + persist_own(c);
    r->ar.base = rb;
    r->ar.cap = persist_own(r);
    r->persist_bytes = r->ar.cap;
    r->d_nW = (double*)rb;
    r->fft_work = rb + (((size_t)r->g.N * sizeof(double) + 255) & ~size_t(255));
    r->fft_work_cap = r->fft_work_bytes;
    r->row_ready = false;
  }
  return 0;
}

extern "C" int sww_set_fields(sww_ctx* c, sww_devptr nbar, sww_devptr nbar_weight, sww_devptr nbar_A, double shot_fac,
                              double P_fkp) {
  SWW_REQUIRE(c, "null context");
  SWW_REQUIRE(c->ar.base, "bind a workspace before setting the fields");
  c->nbar = (const double*)nbar;
  c->nbar_w = (const double*)nbar_weight;
  c->nbar_A = (const double*)nbar_A;
  c->shot = shot_fac;
  c->P_fkp = P_fkp;
  SWW_TRY(k_prep_static(c));
  c->fields_dirty = false;
  if (c->rowc) SWW_TRY(sww_rowmesh_prepare(c));
  return 0;
}

extern "C" int sww_get_binmap(sww_ctx* c, sww_devptr out) {
  SWW_REQUIRE(c && out, "null argument");
  SWW_CUDA(cudaMemcpyAsync((void*)out, c->d_binmap, (size_t)c->g.Nc, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

extern "C" int sww_bk_set_triangles(sww_ctx* c, const int32_t* tris, int32_t n_tri) {
  SWW_REQUIRE(c && tris && n_tri > 0, "bad triangle list");
  for (int i = 0; i < 3 * n_tri; ++i) SWW_REQUIRE(tris[i] >= 0 && tris[i] < c->g.n_k, "triangle bin out of range");
  c->tris.assign(tris, tris + 3 * (size_t)n_tri);
  c->n_tri = n_tri;
  if (c->d_tris) cudaFree(c->d_tris);
  SWW_CUDA(cudaMalloc(&c->d_tris, sizeof(int) * 3 * (size_t)n_tri));
  SWW_CUDA(cudaMemcpy(c->d_tris, tris, sizeof(int) * 3 * (size_t)n_tri, cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------------
// profiling
// ------------------------------------------------------------------------------------------
ProfScope::ProfScope(sww_ctx* c_, const char* name) : c(c_->parent ? c_->parent : c_), st(c_->stream), slot(-1) {
  if (!c->prof_on) return;
  int cat = -1;
  for (size_t i = 0; i < c->prof_names.size(); ++i)
    if (c->prof_names[i] == name) cat = (int)i;
  if (cat < 0) {
    cat = (int)c->prof_names.size();
    c->prof_names.push_back(name);
  }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  slot = (int)c->prof_cat.size();
  c->prof_cat.push_back(cat);
  c->prof_ev.push_back(a);
  c->prof_ev.push_back(b);
  cudaEventRecord(a, st);
}

ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(c->prof_ev[2 * slot + 1], st);
}

static void prof_clear(sww_ctx* c) {
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  c->prof_ev.clear();
  c->prof_cat.clear();
  c->prof_names.clear();
}

extern "C" int sww_profile_enable(sww_ctx* c, int enable) {
  SWW_REQUIRE(c, "null context");
  prof_clear(c);
  c->prof_on = enable != 0;
  return 0;
}

extern "C" int sww_profile_report(sww_ctx* c, char* buf, uint64_t n) {
  SWW_REQUIRE(c && buf && n > 0, "null argument");
  SWW_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<double> ms(c->prof_names.size(), 0.0);
  std::vector<long long> cnt(c->prof_names.size(), 0);
  for (size_t i = 0; i < c->prof_cat.size(); ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]);
    ms[c->prof_cat[i]] += t;
    cnt[c->prof_cat[i]]++;
  }
  std::string out;
  char line[256];
  for (size_t i = 0; i < ms.size(); ++i) {
    snprintf(line, sizeof(line), "%-28s %8lld %12.3f\n", c->prof_names[i].c_str(), cnt[i], ms[i]);
    out += line;
  }
  snprintf(buf, n, "%s", out.c_str());
  return 0;
}
