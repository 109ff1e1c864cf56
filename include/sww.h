/* sww.h -- C ABI of libsww.so: the B200 (sm_100a) implementation of the Monte-Carlo Fisher /
 * quadratic + cubic estimator hot path of Spectra-Without-Windows.
 *
 * The reference has no FFI layer: its "operator API" is the set of Python functions the four
 * driver scripts import by name (pk/compute_pk_randoms.py:12-13, pk/compute_pk_data.py:12-13,
 * bk/compute_bk_randoms.py:12, bk/compute_bk_data.py:13).  Each entry point below states the
 * reference function / script lines it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on error; sww_last_error() returns a
 *     thread-local message (the Python layer raises the bare `Exception` the scripts use);
 *   - device pointers are passed as 64-bit integers (tensor.data_ptr()); the caller owns all
 *     large buffers (one workspace, bound once) and the library never frees caller memory;
 *   - all work is asynchronous on the context's stream (sww_ctx_set_stream);
 *   - real maps are double[nx][ny][nz] (C order); spectra are the R2C half-spectrum
 *     double2[nx][ny][nz/2+1]; full complex maps (ft/ift mirrors only) are double2[nx][ny][nz];
 *   - one sww_ctx per process and GPU; a context is not thread-safe.
 */
#ifndef SWW_H
#define SWW_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWW_ABI_VERSION 1

typedef struct sww_ctx sww_ctx;
typedef uint64_t sww_devptr; /* device address */
typedef uint64_t sww_stream; /* cudaStream_t   */

/* Geometry and binning, built on the host with the reference's own expressions so that bin
 * membership and the MAS window are bit-identical:
 *   rax  : src/opt_utilities.py:428-430 (fft-ordered mesh coordinate + BoxCenter + half a cell)
 *   kax  : src/opt_utilities.py:420-422
 *   mas  : src/opt_utilities.py:448-463 (1-D factors, float32 pi/N prefactor)
 *   k_lo/k_hi : src/opt_utilities.py:485-487 (np.arange edges)                              */
typedef struct {
  int32_t grid[3];
  double box[3];
  int32_t lmax; /* even, <= 4 */
  int32_t n_k;
  const double* k_lo;
  const double* k_hi;
  const double* rax[3];
  const double* kax[3];
  const double* mas[3];
  int32_t device;
} sww_geometry;

/* workspace selectors for sww_workspace_bytes */
enum { SWW_WS_OPS = 0, SWW_WS_PK_FISHER = 1, SWW_WS_PK_DATA = 2, SWW_WS_BK_FISHER = 3, SWW_WS_BK_DATA = 4, SWW_WS_PK_ML = 5 };

const char* sww_last_error(void);
int sww_abi_version(void);

int sww_ctx_create(const sww_geometry* g, sww_ctx** out);
int sww_ctx_destroy(sww_ctx* c);
int sww_ctx_set_stream(sww_ctx* c, sww_stream s);
/* bytes of caller-owned workspace needed for a pipeline (includes the cuFFT work area) */
int sww_workspace_bytes(sww_ctx* c, int what, uint64_t* bytes);
int sww_bind_workspace(sww_ctx* c, sww_devptr ws, uint64_t bytes);
/* number of kernel launches (own kernels, cuFFT calls) issued by this context so far */
int sww_launch_counts(sww_ctx* c, uint64_t* own_kernels, uint64_t* fft_calls);
/* CUDA-event profiling of the library's own launches, aggregated per kernel family.
 * enable != 0 starts a fresh recording; the report is a text table (name, launches, ms). */
int sww_profile_enable(sww_ctx* c, int enable);
int sww_profile_report(sww_ctx* c, char* buf, uint64_t buf_bytes);
/* choose the FFT backend: 0 = cuFFT, 1 = hand-written sm_100a passes (power-of-two axes, or any axis of <= 640 points) */
int sww_set_fft_backend(sww_ctx* c, int backend);
/* Coarse row mesh of the P_l(k) Fisher rows (the n_b evaluations of applyC_alpha_single,
 * src/covariances_pk.py:404-466, that pk/compute_pk_randoms.py:220-275 loops over): the band-limited products
 * are evaluated exactly on a grid chosen per axis -- the mesh's own length, or the smallest power of two
 * M_i >= 2 (2 b_i + 1), b_i = largest binned mode index, which may be smaller OR larger than the mesh (an axis length with
 * a large prime factor: the weight map's n-periodic spectrum turns the cyclic convolution into a linear one).
 * grid[3] receives M (zeros when the rows run on the full mesh; SWW_ROW_MESH=0 disables it). */
int sww_row_mesh(sww_ctx* c, int32_t grid[3]);

/* n(r) maps: nbar (load_nbar(...,alpha_ran) or painted randoms), nbar_weight, nbar_A
 * (pk/compute_pk_randoms.py:135-136,160,173-179).  Borrowed device pointers, double[N]. */
int sww_set_fields(sww_ctx* c, sww_devptr nbar, sww_devptr nbar_weight, sww_devptr nbar_A,
                   double shot_fac, double P_fkp);

/* Optional cache of the real-space harmonics Y_lm(r^) for l > 0: `maps` is caller-owned device memory of
 * sww_ylm_cache_bytes() bytes which the library fills; 0 disables the cache (harmonics are then
 * re-evaluated in registers).  Trades (n_lm - 1) maps of HBM for the FP64 work of the z passes. */
int sww_ylm_cache_bytes(sww_ctx* c, uint64_t* bytes);
int sww_set_ylm_cache(sww_ctx* c, sww_devptr maps);

/* include_pix (the .param boolean, paramfiles/boss_nC.param:64): forward-model the pixel window.  In the
 * Fisher matrix, q-bar and q_alpha the MAS operators cancel between C^-1 and C_,alpha, so only the
 * preparation of the maps changes: a -> M a in the C^-1 branch, A^-1 a -> M^-1 A^-1 a in the other, and the
 * data-only MAS factors are dropped (src/covariances_pk.py:155-163,371-372,386-389,396-397). */
int sww_set_include_pix(sww_ctx* c, int include_pix);
/* out = IFT[ MAS^power * FT[x] ]  (power = +1 / -1), the building block of the include_pix mirrors */
int sww_mas_filter(sww_ctx* c, sww_devptr x, int power, sww_devptr out);

/* ---- drop-in operators (mirrors of src/opt_utilities.py and src/covariances_pk.py) ---- */
/* ft / ift  (src/opt_utilities.py:502-555) in complex128; inverse normalised by 1/N. */
int sww_fft_c2c(sww_ctx* c, sww_devptr in, sww_devptr out, int inverse);
/* real-to-half-spectrum and back (unnormalised forward, 1/N inverse). c2r destroys `in`. */
int sww_fft_r2c(sww_ctx* c, sww_devptr in, sww_devptr out);
int sww_fft_c2r(sww_ctx* c, sww_devptr in, sww_devptr out);
/* applyCinv_fkp, include_pix=False (src/covariances_pk.py:153-164): out = x v_cell/(n(nP+s)) on n>0 */
int sww_apply_cinv_fkp(sww_ctx* c, sww_devptr x, sww_devptr nbar_w, double v_cell, double shot_fac,
                       double P_fkp, sww_devptr out);
/* applyN, include_pix=False (src/covariances_pk.py:124) */
int sww_apply_n(sww_ctx* c, sww_devptr x, sww_devptr nbar, double v_cell, double shot_fac, sww_devptr out);
/* sum_m Y_lm(k^) FT[x Y_lm(r^)] * MAS^mas_power for every even l <= lmax
 * (src/covariances_pk.py:378-389; NOT yet multiplied by 4pi/(2l+1)).  out: double2[n_l][Nc]. */
int sww_ylm_project(sww_ctx* c, sww_devptr x, int mas_power, sww_devptr out);
/* applyC_alpha_single, include_pix=False (src/covariances_pk.py:404-466): one derivative map */
int sww_apply_c_alpha_single(sww_ctx* c, sww_devptr x, sww_devptr nbar, double v_cell, int l_i, int a,
                             int data, sww_devptr out);
/* device bin-index map (uint8[nx][ny][nz/2+1], 255 = outside every bin) -- compute_filters */
int sww_get_binmap(sww_ctx* c, sww_devptr out);

/* ---- painting (src/opt_utilities.py:219-267, pmesh TSC / CIC without compensation) ---- */
/* out[cell] += scale * sum_p w_p W(x_p - cell); pos double[n][3], w double[n] (device).
 * scheme: 2 = CIC, 3 = TSC. */
int sww_paint(sww_ctx* c, sww_devptr pos, sww_devptr w, uint64_t n, const double box_center[3],
              int scheme, double scale, sww_devptr out);

/* ---- Gaussian random field with numpy's legacy stream (pk/compute_pk_randoms.py:139-140) ---- */
/* out[i] = sigma[i] * g_i, i < n, where g is the sequence np.random.seed(seed); np.random.normal(size=n)
 * (MT19937 + polar Box-Muller with the cached second variate).  Uniforms and accept/reject decisions
 * are bit-exact; the Gaussian values agree with numpy's to <= 2 ulp.  Synchronous on `stream` only: the
 * generator uses one SM, so on a side stream it overlaps with the pipelines of the context stream. */
int sww_grf_legacy(sww_ctx* c, uint32_t seed, sww_devptr sigma, sww_devptr out, uint64_t n, sww_stream stream /* 0: context stream */);
/* The same for up to 8 seeds at once (outs[s] = field of seeds[s], all scaled by the same sigma): the MT19937
 * recurrences are sequential but independent, one CTA each, so a batch costs the time of one field.
 * With stream != 0 the call only enqueues (no host synchronisation); sww_grf_check() afterwards reports the
 * (probability ~1e-15) case of a field left short of accepted polar pairs. */
int sww_grf_legacy_batch(sww_ctx* c, const uint32_t* seeds, int nseeds, sww_devptr sigma, const sww_devptr* outs, uint64_t n,
                         sww_stream stream);
int sww_grf_check(sww_ctx* c);

/* ---- P_l(k) pipelines ---- */
/* One Gaussian realisation `a` (double[N], device) -> q-bar_alpha [n_b] and the symmetrised
 * F_ab [n_b*n_b] (device, double).  Replaces pk/compute_pk_randoms.py:206-281. */
int sww_pk_fisher_realisation(sww_ctx* c, sww_devptr a, sww_devptr fish_out, sww_devptr qbar_out);
/* q_alpha [n_b] of a painted (n_g - alpha n_r)/V_cell map.  Replaces pk/compute_pk_data.py:191-206. */
int sww_pk_data_q(sww_ctx* c, sww_devptr diff, sww_devptr q_out);
/* After the sum over realisations and ranks (one host-side all-reduce of n_b^2 + n_b doubles, torch.distributed / NCCL):
 * the mean and the symmetrisation of pk/compute_pk_data.py:227-247 / pk/compute_pk_randoms.py:281, in place.
 * qbar_sum may be 0 (bispectrum Fisher: pairs instead of realisations, no bias vector). */
int sww_mc_finalize(sww_ctx* c, sww_devptr fish_sum, sww_devptr qbar_sum, int32_t n_b, int32_t n_realisations);

/* ---- maximum-likelihood weights (`weights = ML`; workspace selector SWW_WS_PK_ML) ----
 * These operators work on FULL COMPLEX maps double2[nx][ny][nz], like the reference's complex128 arrays: on the
 * Nyquist planes Y_lm(k^) P_l(k) is not Hermitian, and the imaginary parts feed back through the bilinear
 * conjugate-gradient products, so a half-spectrum formulation would not reproduce the reference to 1e-9.
 * Fiducial multipoles P_l(|k|) interpolated on the mesh (pk/compute_pk_randoms.py:188-193): borrowed device
 * pointer, double[n_l][nx][ny][nz]. */
int sww_set_fiducial_pk(sww_ctx* c, sww_devptr pk_maps);
/* out = Op x on complex maps (x != out); include_pix as set by sww_set_include_pix.
 *   op 0: applyS (src/covariances_pk.py:45-95)   op 1: applyN (:97-124)
 *   op 2: applyC = S + N (:10-43)                op 3: applyCinv_fkp / applyCinv_approx / preconditioner (:126-220) */
int sww_ml_apply(sww_ctx* c, int op, sww_devptr x, sww_devptr out);
/* applyCinv (src/covariances_pk.py:222-327): preconditioned conjugate gradients for C x = pix, FKP inverse as
 * first guess and preconditioner, stall test every 10 iterations.  rel_tol < 0 means rel_tol = None (the
 * absolute test sqrt(new_sum) < abs_tol applies).  stop_reason: 1 tolerance met, 2 stalled, 0 max_it reached. */
int sww_ml_apply_cinv(sww_ctx* c, sww_devptr pix, int max_it, double abs_tol, double rel_tol, sww_devptr out,
                      int32_t* iterations, int32_t* stop_reason);
/* q_alpha = 1/2 sum_x z C_,alpha z of an already C^-1-weighted REAL map z (pk/compute_pk_data.py:198-206);
 * for the complex C^-1 d of the ML weights the caller forms q[Re z] - q[Im z] (the products are bilinear). */
int sww_pk_q_from_cinv(sww_ctx* c, sww_devptr z, sww_devptr q_out);
/* One realisation with ML weights -> q-bar_alpha and symmetrised F_ab: pk/compute_pk_randoms.py:205-281 along
 * its non-low_mem branch (n_b + 1 conjugate-gradient solves with rel_tol, max_it as at :211,:240).  The reference
 * script cannot run this configuration itself (undefined name at :261, double `del` at :278). */
int sww_pk_fisher_realisation_ml(sww_ctx* c, sww_devptr a, int max_it, double rel_tol, sww_devptr fish_out,
                                 sww_devptr qbar_out, int32_t* total_iterations);

/* ---- B_l(k1,k2,k3) pipelines ---- */
/* triangle list (a<=b<=c) set by the host from test_bin (bk/compute_bk_randoms.py:186-211) */
int sww_bk_set_triangles(sww_ctx* c, const int32_t* tris /*[n_tri][3]*/, int32_t n_tri);
/* g_i^l and tilde-g_i^l maps of one field (bk/compute_bk_randoms.py:222-267).  data!=0 applies the
 * extra MAS factor of bk/compute_bk_data.py:270 and skips tilde-g.  Outputs double[n_l][n_k][N]. */
int sww_bk_gmaps(sww_ctx* c, sww_devptr a, int data, sww_devptr g_out, sww_devptr tg_out);
/* Fisher contribution of a pair of realisations (bk/compute_bk_randoms.py:269-517) given Delta_abc
 * [n_b] (device).  fish_out double[n_b*n_b].  g/tilde-g maps of both fields are left in
 * gA,tgA,gB,tgB (double[n_l][n_k][N] each, device) for the 1-field term of the data script. */
int sww_bk_fisher_pair(sww_ctx* c, sww_devptr aA, sww_devptr aB, sww_devptr delta, sww_devptr gA,
                       sww_devptr tgA, sww_devptr gB, sww_devptr tgB, sww_devptr fish_out);
/* Delta_abc degeneracy factors (bk/compute_bk_randoms.py:427-473), data independent. */
int sww_bk_delta(sww_ctx* c, sww_devptr delta_out);
/* 3-field term sum_x g_a^0 g_b^0 g_c^l for every triangle and l (bk/compute_bk_data.py:350-367);
 * q_out double[n_tri][n_l]. */
int sww_bk_q3(sww_ctx* c, sww_devptr g, sww_devptr q_out);
/* 1-field term of one MC simulation (bk/compute_bk_data.py:380-403): q_out[t][l] -= 0.5/N_mc * (...) */
int sww_bk_q1_accumulate(sww_ctx* c, sww_devptr g, sww_devptr g_mc, sww_devptr tg_mc, double inv_two_nmc,
                         sww_devptr q_inout);
/* C[i][j] = sum_x A[i][x] B[j][x]  (FP64 tensor-core Gram, the dense contraction of
 * bk/compute_bk_randoms.py:489-497);  accumulate!=0 adds into C.  A: [m][N], B: [n][N], C: [m][n]. */
int sww_gram_f64(sww_ctx* c, sww_devptr A, sww_devptr B, int32_t m, int32_t n, uint64_t N, int accumulate,
                 sww_devptr C);

#ifdef __cplusplus
}
#endif
#endif /* SWW_H */
