// Internal declarations shared by the libsww translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sww.h"

#define SWW_NO_BIN 255
#define SWW_MAX_L 3     // multipole slots (l = 0,2,4)
#define SWW_MAX_BINS 254

void sww_set_error(const char* fmt, ...);

#define SWW_CUDA(x)                                                                             \
  do {                                                                                          \
    cudaError_t e__ = (x);                                                                      \
    if (e__ != cudaSuccess) {                                                                   \
      sww_set_error("%s:%d CUDA error %s in %s", __FILE__, __LINE__, cudaGetErrorString(e__), #x); \
      return -2;                                                                                \
    }                                                                                           \
  } while (0)

#define SWW_CUFFT(x)                                                                   \
  do {                                                                                 \
    cufftResult r__ = (x);                                                             \
    if (r__ != CUFFT_SUCCESS) {                                                        \
      sww_set_error("%s:%d cuFFT error %d in %s", __FILE__, __LINE__, (int)r__, #x);  \
      return -3;                                                                       \
    }                                                                                  \
  } while (0)

#define SWW_TRY(x)          \
  do {                      \
    int r__ = (x);          \
    if (r__ != 0) return r__; \
  } while (0)

#define SWW_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      sww_set_error(__VA_ARGS__); \
      return -1;                  \
    }                             \
  } while (0)

// Geometry block passed by value to kernels.
struct Geo {
  int nx, ny, nz, nzh;
  long long N, Nc;
  const double* rax[3];
  const double* kax[3];
  const double* mas[3];
  const uint8_t* binmap;
  int n_k;
  // pruning box: signed indices |ix|<=bx, |iy|<=by, 0<=iz<=bz carry every binned mode
  int bx, by, bz, Kx, Ky, Kz;
};

// Bump allocator over the caller-provided workspace.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void reset(size_t start = 0) { off = start; }
  template <typename T>
  T* alloc(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (off > peak) peak = off;
    if (dry) return nullptr;
    return (T*)(base + o);
  }
  bool ok() const { return dry || off <= cap; }
};

struct sww_ctx {
  int device = 0;
  cudaStream_t stream = 0;
  Geo g;
  double box[3];
  double v_cell;
  int lmax, n_l, n_lm;
  std::vector<double> h_klo, h_khi;
  // ctx-owned small tables
  double* d_tables = nullptr;
  double *d_klo = nullptr, *d_khi = nullptr;
  uint8_t* d_binmap = nullptr;
  // borrowed fields
  const double *nbar = nullptr, *nbar_w = nullptr, *nbar_A = nullptr;
  double shot = 0, P_fkp = 1e4;
  bool fields_dirty = true;
  // workspace
  Arena ar;
  size_t persist_bytes = 0;   // [nW map][cufft work area] at the front of the workspace
  double* d_nW = nullptr;     // n * W  with W = n/(n_w (n_w P + s)) [n_w>0]
  void* fft_work = nullptr;
  size_t fft_work_bytes = 0;   // bytes reserved for FFT scratch (cuFFT work area or the sm_100a pass buffers)
  size_t fft_work_cap = 0;
  void* sm100 = nullptr;       // Sm100State (fft_sm100.cu)
  void* rng = nullptr;         // RngScratch (rng.cu)
  const double* ylm_maps = nullptr;   // optional cache [n_lm-1][N] of Y_lm(r^), lm >= 1
  // cuFFT
  cufftHandle plan_d2z = 0, plan_z2d = 0, plan_z2z = 0;
  bool have_z2z = false;
  int fft_backend = 0;
  bool include_pix = false;
  const double* pk_fid = nullptr;   // borrowed: fiducial P_l(|k|) on the full mesh, double[n_l][nx][ny][nz] (ML weights)
  bool tight_ws = false;   // SWW_TIGHT_WS=1: size the workspace for the selected FFT backend only
  // bk
  std::vector<int> tris;
  int* d_tris = nullptr;
  int n_tri = 0;
  // accounting
  unsigned long long n_own = 0, n_fft = 0;
  // scratch for block-level partial sums
  double* d_partial = nullptr;
  size_t partial_elems = 0;
  unsigned int* d_counter = nullptr;
  int shell_rows = -1;   // rows of the shell-ordered store (-1: not decided yet)
  double* d_gram_partial = nullptr;   // split-K partial tiles of the DMMA Grams (gram.cu)
  size_t gram_partial_bytes = 0;
  double* d_q1_scratch = nullptr;
  size_t q1_scratch_elems = 0;
  void* d_paint_scratch = nullptr;    // tile bins of the painter (paint.cu)
  size_t paint_scratch_bytes = 0;
  // coarse "row mesh" (fft_sm100.cu, "row mesh"): a child context on a smaller power-of-two grid with the same box and
  // bins, on which the band-limited Fisher rows u_alpha -> FT[nW u_alpha] are evaluated exactly (no aliasing into the
  // binned modes) against a low-pass-filtered copy of the weight map
  sww_ctx* rowc = nullptr;
  sww_ctx* parent = nullptr;   // child: events of the library's launches are recorded in the parent's profile
  bool is_child = false;
  bool row_ready = false;     // child: the filtered weight map is current
  // profiling
  bool prof_on = false;
  std::vector<std::string> prof_names;
  std::vector<int> prof_cat;
  std::vector<cudaEvent_t> prof_ev;   // pairs
};

// ---- profiling (ctx.cu): events around the library's own launches
struct ProfScope {
  sww_ctx* c;
  cudaStream_t st;
  int slot;
  ProfScope(sww_ctx* c, const char* name);
  ~ProfScope();
};
#define SWW_PROF(c, name) ProfScope prof_scope__(c, name)

// SM count of the current device (148 on B200), queried once: grids are sized in multiples of it
static inline int sww_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      sms = 148;
  }
  return sms;
}
#define SMS sww_sm_count()

static inline int sww_grid_for(long long n, int block, int max_blocks = 0) {
  if (max_blocks <= 0) max_blocks = SMS * 16;
  long long b = (n + block - 1) / block;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- fft.cu
int sww_fft_init(sww_ctx* c);
void sww_fft_destroy(sww_ctx* c);
int sww_fft_bind_work(sww_ctx* c);
int sww_r2c(sww_ctx* c, const double* in, double2* out);
int sww_c2r(sww_ctx* c, double2* in, double* out);  // unnormalised, destroys in
int sww_c2c(sww_ctx* c, const double2* in, double2* out, int inverse);
int sww_fft_sm100_supported(const sww_ctx* c);
int sww_fft_sm100_r2c(sww_ctx* c, const double* in, double2* out);
int sww_fft_sm100_c2r(sww_ctx* c, double2* in, double* out);
int sww_fft_sm100_c2c_supported(const sww_ctx* c);
int sww_fft_sm100_c2c(sww_ctx* c, const double2* in, double2* out, int inverse);   // unnormalised
size_t sww_fft_sm100_work_bytes(const sww_ctx* c);
void sww_fft_sm100_destroy(sww_ctx* c);
void sww_rng_destroy(sww_ctx* c);   // rng.cu
struct F3Args {
  double2* dst;
  int lm, first, mas_power;
  double coef;
  const double2* G[3];
  int n_g;
  double scale;
  double* partial;
  unsigned int* counter;
  double* out_row;
  // MODE 2 batch: nrows planes (P + r * row_pstride) binned against the same G maps in one launch; consecutive
  // CTAs take consecutive rows of the same columns, so the G reads of rows 1.. hit in L2
  int nrows;
  long long row_pstride;
  double* out_rows[SWW_MAX_L];
  // MODE 3: store the in-shell modes in shell order (perm[box cell] = position or -1)
  const int* perm;
  double2* sorted_dst;
  // MODE 0/1 dst and MODE 2 G maps may live on another (coarser) grid with the same box: [dnx][dny][dnzh]  (0: this grid)
  int dnx, dny, dnzh;
  // > 0: (k_o, k_z tile) columns with k_o^2 + k_z0^2 >= klim2 hold no binned mode and are skipped (the first forward pass
  // does not store them either)
  double klim2;
};
// gather description for the first inverse pass: sum of up to three compact box spectra, each kept only
// inside its shell, times coef * Y_lm(k^)  (n_terms == 0: plain spectrum + optional shell filter)
struct GatherArgs {
  const double2* p[3];
  int s[3];
  int n_terms;
  int lm;
  double coef;
  const double2* src;   // n_terms == 0: spectrum [nx][ny][nzh] ...
  int shell;            // ... kept only inside this shell (< 0: whole box)
};
#define SWW_MAX_ZBATCH 10
struct GatherBatch {
  GatherArgs ga[SWW_MAX_ZBATCH];
  int nb;   // 0: plain spectrum mode
};
int sww_sm100_inverse_multi(sww_ctx* c, int nb, const GatherArgs* gas, int box_shell, const int* lms_r,
                            const double* wmap, double scale, int accumulate, double2* qstack, double* real_out);
int sww_sm100_forward_box(sww_ctx* c, const double* real_in, const double* wmap, int lm_r, int mode, F3Args A);
int sww_sm100_inverse_box(sww_ctx* c, const double2* src, int filter_shell, int box_shell, const double* wmap, int lm_r,
                          double scale, int accumulate, double* real_out);
int sww_sm100_pk_row(sww_ctx* c, const double2* F, int shell, const double* wmap, double inv_scale,
                     const double2* const* G, int n_g, double scale, double* out_row);
// the same for nrows spectra sharing one shell (the multipoles of one k bin): the weight map is read once
// shell-ordered Fisher rows (see "shell-ordered spectra" in fft_sm100.cu)
long long sww_sm100_shell_cells(sww_ctx* c);          // S: modes of the output box that fall in a k bin (0: unsupported)
int sww_sm100_shell_chunks(sww_ctx* c);
int sww_sm100_shell_sort_g(sww_ctx* c, const double2* const* G, int n_g, double2* Gs);
int sww_sm100_pk_rows_store(sww_ctx* c, int nrows, const double2* const* F, int shell, const double* wmap, double inv_scale,
                            double2* Us_rows);
int sww_sm100_shell_dot(sww_ctx* c, const double2* Us, int R, const double2* Gs, int n_g, double scale, double* partial,
                        double* F, int n_l, int ka0);
int sww_sm100_pk_rows(sww_ctx* c, int nrows, const double2* const* F, int shell, const double* wmap, double inv_scale,
                      const double2* const* G, int n_g, double scale, double* const* out_rows);

// ---- kernels.cu (launch wrappers)
int k_build_binmap(sww_ctx* c);
int k_prep_static(sww_ctx* c);  // nW map
int k_prep_pk(sww_ctx* c, const double* a, double* nCa, double* nAa, double* WS, const double* a_c = nullptr,
              const double* ainv = nullptr);
int k_div(sww_ctx* c, const double* a, const double* b, double* out);
int k_mas_mul(sww_ctx* c, double2* spec, int power, double scale);
int sww_mas_filter_internal(sww_ctx* c, const double* x, int power, double2* spec, double* out);
int k_cinv_fkp(sww_ctx* c, const double* x, const double* nw, double v_cell, double shot, double P, double* out,
               const double* post_mul);
int k_apply_n(sww_ctx* c, const double* x, const double* nbar, double v_cell, double shot, double* out);
int k_mul(sww_ctx* c, const double* a, const double* b, double s, double* out);
int k_mul_ylm_r(sww_ctx* c, const double* f, int lm, double* out);
int k_acc_ylm_k(sww_ctx* c, const double2* spec, int lm, double coef, int mas_power, int first, int pruned,
                double2* out);
int k_shell_fill(sww_ctx* c, const double2* F, int a, double coef, double2* X);
int k_weight(sww_ctx* c, double* u, const double* w, double s);
int k_shell_bin(sww_ctx* c, const double2* T, const double2* const* G, int n_g, double scale, double* out_row);
int k_symmetrise(sww_ctx* c, double* F, int n);
int k_scale_c(sww_ctx* c, double2* x, double s, long long n);
int k_paint(sww_ctx* c, const double* pos, const double* w, unsigned long long n, const double bc[3], int scheme,
            double scale, double* out);
