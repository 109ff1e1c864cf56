// Hand-written sm_100a 3-D FFT passes (FP64, power-of-two meshes) with pruning and fused
// prologue / epilogue work.  Replaces the cuFFT plans + separate element-wise kernels on the hot
// loops (fft_backend = 1).
//
// A 3-D real<->half-spectrum transform is five passes over compact intermediates that only hold the
// pruning box along the axes already (or still) in Fourier space:
//
//   inverse:  I1  x-axis  gather box modes from a spectrum [nx][ny][nzh] (optional shell filter Theta_a)
//                         -> P [nx][Ky][Kzp]
//             I2  y-axis  P -> Q [nx][ny][Kzp]
//             Z   z-axis  complex-to-real of length nz through a half-length complex FFT
//   forward:  Z   z-axis  real-to-complex, keeping only iz < Kz          -> R [nx][ny][Kzp]
//             F2  y-axis  R -> P [nx][Ky][Kzp]   (only box rows are stored)
//             F3  x-axis  P -> epilogue: store into a spectrum, accumulate Y_lm(k^)-weighted, or bin
//                         conj(T) G_l' straight into the k-shell histogram (no spectrum is written)
//
// The fused P_l Fisher rows of one k bin are I1 -> I2 -> Z(c2r * nW * r2c) -> F2 -> F3(shell-ordered store), batched over
// the multipoles: the real-space map u_alpha and its weighted transform never touch HBM, and the k-shell reduction of
// all rows is one launch of segmented dot products at the end (shell_dot_kernel).  Kernels of this file:
//   pass_i1 / pass_i2 / pass_f2 / pass_f3   column passes (x or y axis), tiles of 8 z-columns in shared memory
//   pass_zw_kernel                          z pass, one warp per row, generic in nz (inverse, forward or fused)
//   pass_z16_fused_kernel                   the fused z pass for nz = 512: 256 = 16 x 16, two exchanges per row
//   shell_sort_g / shell_dot / shell_dot_reduce   shell-ordered G maps and the segmented products
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "fft_core.cuh"
#include "fft_any.cuh"
#include "sww_internal.cuh"
#include "ylm_gen.cuh"


static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// an axis runs on the templated power-of-two kernels, or -- any other length up to ANY_MAX_N -- on the run-time radix plans of
// fft_any.cuh / fft_any_passes.cuh (the reference's production meshes: 258 x 496 x 274, 172 x 330 x 182, 167 x 307 x 175)
static inline bool col_pow2(int n) { return is_pow2(n) && n >= 64 && n <= 1024; }
static inline bool z_pow2(int n) { return is_pow2(n) && n >= 128 && n <= 1024; }
static inline bool any_len(int n) { return n >= 2 && n <= ANY_MAX_N; }

int sww_fft_sm100_supported(const sww_ctx* c) {
  const Geo& g = c->g;
  if (const char* e = getenv("SWW_NO_ANY_LENGTH"))   // A/B switch: power-of-two meshes only (the other meshes take cuFFT)
    if (e[0] == '1') return col_pow2(g.nx) && col_pow2(g.ny) && z_pow2(g.nz);
  return (col_pow2(g.nx) || any_len(g.nx)) && (col_pow2(g.ny) || any_len(g.ny)) && (z_pow2(g.nz) || any_len(g.nz));
}

// ------------------------------------------------------------------------------------------
// Tile FFT with register-resident first and last stages.
//   * stage 0 takes its inputs from a caller-supplied functor (global memory, a staging buffer or
//     registers left by a previous transform) -- the tile is never filled by a separate pass;
//   * the last stage hands its outputs to a functor (global stores, epilogues, or the registers of
//     the next transform) -- the tile is never drained by a separate pass;
//   * only the exchanges between stages go through shared memory (2 passes per exchange).
// REV selects the reversed radix order, so that "inverse, then forward" chains share the register
// layout at the seam: outputs of the last inverse stage sit at positions bi + r*(N/R) -- exactly what
// stage 0 of a forward transform whose first radix is R consumes.
// ------------------------------------------------------------------------------------------
template <int N, bool REV, int S>
struct PlanRadix {
  static constexpr int value = REV ? FftPlan<N>::R[FftPlan<N>::NS - 1 - S] : FftPlan<N>::R[S];
};

template <int N, int DIR, bool REV, int S, int P, int LDT = FFT_LDT>
struct MidStages {   // stages 1 .. NS-2, fully in shared memory
  static __device__ __forceinline__ void run(double2* tile, const double2* tw, int tid) {
    if constexpr (S < FftPlan<N>::NS - 1) {
      constexpr int R = PlanRadix<N, REV, S>::value;
      constexpr int NT = N / 2;
      constexpr int IPT = 16 / R;
      double2 u[IPT][R];
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        int w = tid + s * NT;
        stage_load<R, DIR, LDT>(tile, N, P, w >> 3, w & 7, tw, u[s]);
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        int w = tid + s * NT;
        stage_store<R, LDT>(tile, N, P, w >> 3, w & 7, u[s]);
      }
      __syncthreads();
      MidStages<N, DIR, REV, S + 1, P * R, LDT>::run(tile, tw, tid);
    }
  }
};

template <int N, bool REV, int S>
struct PlanProd {   // product of the first S radices
  static constexpr int value = PlanProd<N, REV, S - 1>::value * PlanRadix<N, REV, S - 1>::value;
};
template <int N, bool REV>
struct PlanProd<N, REV, 0> {
  static constexpr int value = 1;
};

// Stage 0: u[s][r] = ld(pos = bi + r*T, c), butterfly, scatter into the tile.  The caller must have
// synchronised since the tile was last read.
template <int N, int DIR, bool REV, typename LD>
__device__ __forceinline__ void fft_first(double2* tile, int tid, LD ld) {
  constexpr int R = PlanRadix<N, REV, 0>::value;
  constexpr int NT = N / 2, IPT = 16 / R, T = N / R;
  double2 u[IPT][R];
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    int w = tid + s * NT, bi = w >> 3, c = w & 7;
#pragma unroll
    for (int r = 0; r < R; ++r) u[s][r] = ld(bi + r * T, c);
  }
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    int w = tid + s * NT;
    dftR<R, DIR>(u[s]);
    stage_store<R>(tile, N, 1, w >> 3, w & 7, u[s]);
  }
  __syncthreads();
}

// Stage 0 fed from registers laid out as the outputs of a preceding last stage of the same radix.
template <int N, int DIR, bool REV, int R>
__device__ __forceinline__ void fft_first_regs(double2* tile, int tid, double2 (*u)[R]) {
  static_assert(R == PlanRadix<N, REV, 0>::value, "radix mismatch at the seam");
  constexpr int NT = N / 2, IPT = 16 / R;
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    int w = tid + s * NT;
    dftR<R, DIR>(u[s]);
    stage_store<R>(tile, N, 1, w >> 3, w & 7, u[s]);
  }
  __syncthreads();
}

// Last stage: outputs u[s][r] sit at position bi + r*P of column c (P = N/R).
template <int N, int DIR, bool REV, int R, int LDT = FFT_LDT>
__device__ __forceinline__ void fft_last(const double2* tile, const double2* tw, int tid, double2 (*u)[R]) {
  static_assert(R == PlanRadix<N, REV, FftPlan<N>::NS - 1>::value, "radix mismatch");
  constexpr int NT = N / 2, IPT = 16 / R;
  constexpr int P = N / R;
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    int w = tid + s * NT;
    stage_load<R, DIR, LDT>(tile, N, P, w >> 3, w & 7, tw, u[s]);
  }
}

// ------------------------------------------------------------------------------------------
// Bulk asynchronous copies (the TMA engine, non-tensor form) completing on an mbarrier: one thread issues the copy of a
// whole contiguous tile, the CTA waits on the barrier's phase.  cp.async.bulk / mbarrier are sm_90+ PTX; the SASS shows
// UBLKCP and SYNCS.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// resident CTAs per SM the column-pass launchers size their persistent grids for (occ_for: shared memory allows 4 / 2 / 1 at
// N <= 256 / 512 / 1024): the register allocation is held to it -- without the bound the N = 256 instantiations grew from 102 to
// 192-220 registers when the sphere predicates were added, two CTAs fitted where the grid expected four, and the bispectrum
// passes ran 30 % longer (bisected on C4, profiles/r2y)
#define COL_MINB(N) ((N) <= 256 ? 4 : ((N) == 512 ? 2 : 1))
#define FFT_LAST_RADIX(N, REV) (PlanRadix<N, REV, FftPlan<N>::NS - 1>::value)
#define FFT_MID(N, DIR, REV, tile, tw, tid) \
  MidStages<N, DIR, REV, 1, PlanRadix<N, REV, 0>::value>::run(tile, tw, tid)

struct BoxAxis {
  int n, K, b;  // axis length, number of kept indices, last non-negative kept index
  __host__ __device__ int idx(int j) const { return j <= b ? j : n - K + j; }
  __host__ __device__ bool has(int i) const { return i <= b || i >= n - (K - b - 1); }
  __host__ __device__ int pos(int i) const { return i <= b ? i : i - (n - K); }
};

// ------------------------------------------------------------------------------------------
// Column passes are generic in the axis they transform: AX = 0 (x) or 1 (y).  `a` runs along the
// transformed axis, `o` along the other one; row(a, o) = ix*ny + iy.  The pass order is chosen on the
// host so that the more strongly pruned axis is transformed first on the way forward and last on the
// way back, which minimises the compact intermediates:
//     inverse:  first pass (gather)  spectrum -> P[Ko][nA][Kzp],   second pass  P -> Q[nx][ny][Kzp]
//     forward:  first pass           R -> P[K_B][nC][Kzp],          last pass    P -> epilogue
// ------------------------------------------------------------------------------------------
// 1 / x for a positive, normal x (r^2 of a mesh cell) to 1-2 ulp: hardware seed (20 bits) + two Newton steps -- five instructions
// against the ~25 of the IEEE division sequence with its special-case branches (ncu on the forward z pass: FSEL + BRA + BSSY/BSYNC
// of those sequences were 17 % of the instructions).  The harmonics are compared with the reference at 1e-12.
__device__ __forceinline__ double sww_rcp_pos(double x) {
  double q;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
  double e = fma(-x, q, 1.0);
  q = fma(q, e, q);
  e = fma(-x, q, 1.0);
  return fma(q, e, q);
}

// 1 / (1e-12 + sqrt(r2))  (pk/compute_pk_randoms.py:166) through one reciprocal square root:
// 1/(e + r) = rs (1 - e rs + O((e rs)^2)), rs = 1/r; e rs <= 1e-12/|r| so the dropped term is < 1e-24.
__device__ __forceinline__ double inv_norm_eps(double r2) {
  if (r2 <= 0.0) return 1e12;
  const double rs = rsqrt(r2);
  return rs - 1e-12 * rs * rs;
}

// Sphere pruning of the column passes: within the z tile starting at k_z0 only the indices |i| <= lim along an axis with
// fundamental kF can hold a mode with |k|^2 < klim2 (one index of margin against rounding).  All four passes use this one
// function, so a column skipped by the pass that would write it is skipped by the pass that would read it.
__device__ __forceinline__ int sphere_lim(double klim2, double kz0, double kF, int n) {
  if (klim2 <= 0.0) return n;
  const double r2 = klim2 - kz0 * kz0;
  if (r2 <= 0.0) return -1;
  return (int)(sqrt(r2) / kF) + 1;
}
__device__ __forceinline__ int sidx_abs(int i, int n) { return 2 * i <= n ? i : n - i; }

template <int AX>
__device__ __forceinline__ long long row_of(int a, int o, int ny) {
  return AX == 0 ? (long long)a * ny + o : (long long)o * ny + a;
}

// First inverse pass.  Source: spectrum [nx][ny][nzh] restricted to the box; optional shell filter.
// item = (jo, z-tile).  Rows outside the box are never read.
template <int N, int AX>
__global__ void __launch_bounds__(N / 2, COL_MINB(N))
pass_i1_kernel(Geo g, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* __restrict__ src, int shell,
               const double2* __restrict__ twg, double2* __restrict__ Pbase, GatherBatch gb, BoxAxis cbx, BoxAxis cby,
               long long pstride, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  double2* tw = tile + N * FFT_LDT;
  const int tid = threadIdx.x;
  constexpr int NT = N / 2;
  constexpr int RL = FFT_LAST_RADIX(N, false);
  for (int i = tid; i < N; i += NT) tw[i] = twg[i];
  const int ztiles = Kzp / FFT_TZ;
  const long long per = (long long)bO.K * ztiles;
  const long long items = per * (gb.nb > 0 ? gb.nb : 1);
  for (long long it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const int bat = (int)(it0 / per);
    const long long it = it0 - (long long)bat * per;
    const GatherArgs& ga = gb.ga[bat];
    double2* __restrict__ P = Pbase + (long long)bat * pstride;
    const int jo = (int)(it / ztiles), zt = (int)(it - (long long)jo * ztiles);
    const int io = bO.idx(jo), iz0 = zt * FFT_TZ;
    if (klim2 > 0.0) {
      // every mode of this (k_o, k_z tile) column lies outside the sphere |k| < k_lim: the column is identically zero
      // and the next pass never reads it (same predicate there)
      const int n_o = AX == 0 ? g.ny : g.nx;
      if (sidx_abs(io, n_o) > sphere_lim(klim2, g.kax[2][iz0], g.kax[AX == 0 ? 1 : 0][1], n_o)) continue;
    }
    __syncthreads();
    fft_first<N, -1, false>(tile, tid, [&](int ia, int cc) {
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
      return v;
    });
    FFT_MID(N, -1, false, tile, tw, tid);
    double2 u[16 / RL][RL];
    fft_last<N, -1, false, RL>(tile, tw, tid, u);
#pragma unroll
    for (int s = 0; s < 16 / RL; ++s) {
      int w = tid + s * NT, bi = w >> 3, cc = w & 7;
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        int ia = bi + r * (N / RL);
        P[((long long)jo * N + ia) * Kzp + iz0 + cc] = u[s][r];
      }
    }
  }
}

// Second inverse pass (axis AX, box bA): P[ja][f][iz] -> Q[row][iz]; item = (f, z-tile), f < nF
template <int N, int AX>
__global__ void __launch_bounds__(N / 2, COL_MINB(N))
pass_i2_kernel(int nF, int ny, BoxAxis bA, int Kzp, const double2* __restrict__ Pbase, const double2* __restrict__ twg,
               double2* __restrict__ Qbase, int nb, long long pstride, long long qstride, int q_tiled,
               const double* __restrict__ kax_a, const double* __restrict__ kax_z, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  double2* tw = tile + N * FFT_LDT;
  const int tid = threadIdx.x;
  constexpr int NT = N / 2;
  constexpr int RL = FFT_LAST_RADIX(N, false);
  for (int i = tid; i < N; i += NT) tw[i] = twg[i];
  const int ztiles = Kzp / FFT_TZ;
  const long long per = (long long)nF * ztiles;
  const long long items = per * nb;
  for (long long it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const int bat = (int)(it0 / per);
    const long long it = it0 - (long long)bat * per;
    const double2* __restrict__ P = Pbase + (long long)bat * pstride;
    double2* __restrict__ Q = Qbase + (long long)bat * qstride;
    const int f = (int)(it / ztiles), zt = (int)(it - (long long)f * ztiles);
    const int iz0 = zt * FFT_TZ;
    __syncthreads();
    const int lim = klim2 > 0.0 ? sphere_lim(klim2, kax_z[iz0], kax_a[1], N) : N;   // uniform branch: no table loads without pruning
    fft_first<N, -1, false>(tile, tid, [&](int ia, int cc) {
      double2 v = make_double2(0.0, 0.0);
      // columns outside the sphere |k| < k_lim were skipped by the first pass: they are zero, not stored
      if (bA.has(ia) && sidx_abs(ia, N) <= lim) v = P[((long long)bA.pos(ia) * nF + f) * Kzp + iz0 + cc];
      return v;
    });
    FFT_MID(N, -1, false, tile, tw, tid);
    double2 u[16 / RL][RL];
    fft_last<N, -1, false, RL>(tile, tw, tid, u);
#pragma unroll
    for (int s = 0; s < 16 / RL; ++s) {
      int w = tid + s * NT, bi = w >> 3, cc = w & 7;
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        int ia = bi + r * (N / RL);
        if (q_tiled)   // tile-major Q: [f][z-tile][ia][8], contiguous per item
          Q[(((long long)f * ztiles + zt) * N + ia) * FFT_TZ + cc] = u[s][r];
        else
          Q[row_of<AX>(ia, f, ny) * Kzp + iz0 + cc] = u[s][r];
      }
    }
  }
}

// First forward column pass (axis AX, box bA): R[row][iz] -> P[ja][c][iz] (box rows only); item = (c, z-tile), c < nC
template <int N, int AX>
__global__ void __launch_bounds__(N / 2, COL_MINB(N))
pass_f2_kernel(int nC, int ny, BoxAxis bA, int Kzp, const double2* __restrict__ Rbase, const double2* __restrict__ twg,
               double2* __restrict__ Pbase, int nb, long long rstride, long long pstride, int r_tiled,
               const double* __restrict__ kax_a, const double* __restrict__ kax_z, double klim2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  double2* tw = tile + N * FFT_LDT;
  const int tid = threadIdx.x;
  constexpr int NT = N / 2;
  constexpr int RL = FFT_LAST_RADIX(N, false);
  for (int i = tid; i < N; i += NT) tw[i] = twg[i];
  const int ztiles = Kzp / FFT_TZ;
  const long long per = (long long)nC * ztiles;
  const long long items = per * nb;
  for (long long it0 = blockIdx.x; it0 < items; it0 += gridDim.x) {
    const int bat = (int)(it0 / per);
    const long long it = it0 - (long long)bat * per;
    const double2* __restrict__ R = Rbase + (long long)bat * rstride;
    double2* __restrict__ P = Pbase + (long long)bat * pstride;
    const int ic = (int)(it / ztiles), zt = (int)(it - (long long)ic * ztiles);
    const int iz0 = zt * FFT_TZ;
    const int lim = klim2 > 0.0 ? sphere_lim(klim2, kax_z[iz0], kax_a[1], N) : N;   // uniform branch: no table loads without pruning
    __syncthreads();
    if (r_tiled) {   // tile-major R: this item's tile is one contiguous block [ia][8]
      const double2* __restrict__ Rt = R + ((long long)ic * ztiles + zt) * N * FFT_TZ;
      fft_first<N, 1, false>(tile, tid, [&](int ia, int cc) { return Rt[ia * FFT_TZ + cc]; });
    } else {
      fft_first<N, 1, false>(tile, tid, [&](int ia, int cc) { return R[row_of<AX>(ia, ic, ny) * Kzp + iz0 + cc]; });
    }
    FFT_MID(N, 1, false, tile, tw, tid);
    double2 u[16 / RL][RL];
    fft_last<N, 1, false, RL>(tile, tw, tid, u);
#pragma unroll
    for (int s = 0; s < 16 / RL; ++s) {
      int w = tid + s * NT, bi = w >> 3, cc = w & 7;
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        int ia = bi + r * (N / RL);
        // columns outside the sphere |k| < k_lim hold no binned mode: the last pass skips them, nothing is stored
        if (bA.has(ia) && sidx_abs(ia, N) <= lim) P[((long long)bA.pos(ia) * nC + ic) * Kzp + iz0 + cc] = u[s][r];
      }
    }
  }
}

// The same pass on tile-major R planes with double-buffered bulk asynchronous copies: a tile [N][8] is ONE contiguous block,
// thread 0 issues cp.async.bulk for the NEXT item into the other buffer before the CTA transforms the current one, and the
// CTA waits on the buffer's mbarrier.  Stage 0 reads the tile from shared memory (dense pitch 8: every quarter-warp access is
// a full 128-byte row, conflict-free) instead of global memory.  One CTA per SM (two 64 KB buffers at N = 512).
template <int N, int AX>
__global__ void __launch_bounds__(N / 2, 1)
pass_f2_bulk_kernel(int nC, BoxAxis bA, int Kzp, const double2* __restrict__ Rbase, const double2* __restrict__ twg,
                    double2* __restrict__ Pbase, int nb, long long rstride, long long pstride) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int LDT = FFT_TZ;                    // dense tiles
  constexpr int TE = N * FFT_TZ;                 // elements per tile
  constexpr unsigned BYTES = TE * sizeof(double2);
  double2* tiles = reinterpret_cast<double2*>(smem_raw);
  double2* tw = tiles + 2 * TE;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tw + N);
  const int tid = threadIdx.x;
  constexpr int NT = N / 2;
  constexpr int R0 = PlanRadix<N, false, 0>::value;
  constexpr int RL = FFT_LAST_RADIX(N, false);
  for (int i = tid; i < N; i += NT) tw[i] = twg[i];
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int ztiles = Kzp / FFT_TZ;
  const long long per = (long long)nC * ztiles;
  const long long items = per * nb;
  auto src_of = [&](long long it0) -> const double2* {
    const int bat = (int)(it0 / per);
    const long long it = it0 - (long long)bat * per;     // = ic * ztiles + zt
    return Rbase + (long long)bat * rstride + it * TE;
  };
  unsigned phase0 = 0, phase1 = 0;
  int cur = 0;
  long long it0 = blockIdx.x;
  if (it0 < items && tid == 0) {
    mbar_expect_tx(bar, BYTES);
    bulk_g2s(tiles, src_of(it0), BYTES, bar);
  }
  for (; it0 < items; it0 += gridDim.x, cur ^= 1) {
    const long long nxt = it0 + gridDim.x;
    if (nxt < items && tid == 0) {
      fence_proxy_async();        // the other buffer was last touched by generic-proxy loads / stores (before the barrier below)
      mbar_expect_tx(bar + (cur ^ 1), BYTES);
      bulk_g2s(tiles + (cur ^ 1) * TE, src_of(nxt), BYTES, bar + (cur ^ 1));
    }
    if (cur == 0) {
      mbar_wait(bar, phase0);
      phase0 ^= 1;
    } else {
      mbar_wait(bar + 1, phase1);
      phase1 ^= 1;
    }
    double2* tile = tiles + cur * TE;
    const int bat = (int)(it0 / per);
    const long long it = it0 - (long long)bat * per;
    double2* __restrict__ P = Pbase + (long long)bat * pstride;
    const int ic = (int)(it / ztiles), zt = (int)(it - (long long)ic * ztiles);
    const int iz0 = zt * FFT_TZ;
    {   // stage 0 from the tile: all loads, barrier, butterflies + scatter
      constexpr int IPT = 16 / R0, T = N / R0;
      double2 u0[IPT][R0];
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        const int w = tid + s * NT, bi = w >> 3, c = w & 7;
#pragma unroll
        for (int r = 0; r < R0; ++r) u0[s][r] = tile[(bi + r * T) * LDT + c];
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        const int w = tid + s * NT;
        dftR<R0, 1>(u0[s]);
        stage_store<R0, LDT>(tile, N, 1, w >> 3, w & 7, u0[s]);
      }
      __syncthreads();
    }
    MidStages<N, 1, false, 1, R0, LDT>::run(tile, tw, tid);
    double2 u[16 / RL][RL];
    fft_last<N, 1, false, RL, LDT>(tile, tw, tid, u);
#pragma unroll
    for (int s = 0; s < 16 / RL; ++s) {
      int w = tid + s * NT, bi = w >> 3, cc = w & 7;
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        int ia = bi + r * (N / RL);
        if (bA.has(ia)) P[((long long)bA.pos(ia) * nC + ic) * Kzp + iz0 + cc] = u[s][r];
      }
    }
    __syncthreads();   // every thread is done with this buffer before it is refilled two items later
  }
}

// ------------------------------------------------------------------------------------------
// F3: x-axis forward + epilogue straight from the registers of the last stage.  item = (jy, z-tile)
//   MODE 0: dst[cell] = T                         (box cells of a spectrum [nx][ny][nzh])
//   MODE 1: dst[cell] = (first ? 0 : dst[cell]) + Y_lm(k^) T ; on the last m: *= MAS^p * coef
//   MODE 2: shell binning  hist[q*n_k + bin] += w_herm Re(conj(T) G_q[cell])   -> out row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_f(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cell of mode (ix, iy, iz) of this grid in a spectrum laid out for another grid [dnx][dny][dnzh] of the same box (the
// coarse row mesh): negative frequencies keep their distance from the end of the axis.  dnx == 0: this grid.
__device__ __forceinline__ long long remap_cell(const Geo& g, const F3Args& A, int ix, int iy, int iz) {
  if (A.dnx == 0) return ((long long)ix * g.ny + iy) * g.nzh + iz;
  const int jx = 2 * ix < g.nx ? ix : ix - g.nx + A.dnx;
  const int jy = 2 * iy < g.ny ? iy : iy - g.ny + A.dny;
  return ((long long)jx * A.dny + jy) * A.dnzh + iz;
}

template <int N, int AX, int MODE>
__global__ void __launch_bounds__(N / 2, (N >= 512 ? (N >= 1024 ? 1 : 2) : 3))
pass_f3_kernel(Geo g, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* __restrict__ P,
               const double2* __restrict__ twg, F3Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  double2* tw = tile + N * FFT_LDT;
  double* hist = reinterpret_cast<double*>(tw + N);  // MODE 2: [nwarps][n_g*n_k]
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  constexpr int NT = N / 2;
  constexpr int NW = NT / 32;
  constexpr int RL = FFT_LAST_RADIX(N, false);
  const int warp = tid >> 5, lane = tid & 31;
  const int nb = A.n_g * g.n_k;
  for (int i = tid; i < N; i += NT) tw[i] = twg[i];
  if (MODE == 2)
    for (int i = tid; i < NW * nb; i += NT) hist[i] = 0.0;
  const int ztiles = Kzp / FFT_TZ;
  const long long items = (long long)bO.K * ztiles;
  const bool nz_even = (g.nz % 2) == 0;
  // MODE 2 row batch: block = blk * nrows + row
  const int nrows = (MODE == 2 && A.nrows > 1) ? A.nrows : 1;
  const int row = (int)(blockIdx.x % (unsigned)nrows), blk = (int)(blockIdx.x / (unsigned)nrows);
  const int nblk = (int)(gridDim.x / (unsigned)nrows);
  if (MODE == 2 && nrows > 1) P += (long long)row * A.row_pstride;
  for (long long it = blk; it < items; it += nblk) {
    const int jo = (int)(it / ztiles), zt = (int)(it - (long long)jo * ztiles);
    const int io = bO.idx(jo), iz0 = zt * FFT_TZ;
    if (A.klim2 > 0.0) {   // no binned mode in this (k_o, k_z tile) column: nothing to store or bin (uniform per CTA)
      const int n_o = AX == 0 ? g.ny : g.nx;
      if (sidx_abs(io, n_o) > sphere_lim(A.klim2, g.kax[2][iz0], g.kax[AX == 0 ? 1 : 0][1], n_o)) continue;
    }
    __syncthreads();
    fft_first<N, 1, false>(tile, tid,
                           [&](int ia, int cc) { return P[((long long)jo * N + ia) * Kzp + iz0 + cc]; });
    FFT_MID(N, 1, false, tile, tw, tid);
    double2 u[16 / RL][RL];
    fft_last<N, 1, false, RL>(tile, tw, tid, u);
    // epilogue straight from registers (uniform trip count: the warp collectives of MODE 2 are safe)
    if (MODE == 3) {
      // shell-ordered store: all permutation loads first (the compiler cannot hoist them over the stores)
      int pq[16 / RL][RL];
#pragma unroll
      for (int s = 0; s < 16 / RL; ++s) {
        const int w_ = tid + s * NT, bi = w_ >> 3, cc = w_ & 7, iz = iz0 + cc;
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          const int ia = bi + r * (N / RL);
          const bool inb = bA.has(ia) && iz < Kz;
          pq[s][r] = inb ? __ldg(&A.perm[((long long)jo * bA.K + bA.pos(ia)) * Kzp + iz]) : -1;
        }
      }
#pragma unroll
      for (int s = 0; s < 16 / RL; ++s)
#pragma unroll
        for (int r = 0; r < RL; ++r)
          if (pq[s][r] >= 0) A.sorted_dst[pq[s][r]] = u[s][r];
    } else if (MODE != 2) {
#pragma unroll
      for (int s = 0; s < 16 / RL; ++s) {
        const int w_ = tid + s * NT, bi = w_ >> 3, cc = w_ & 7, iz = iz0 + cc;
        // MODE 1 accumulates into dst: all previous values of the batch are loaded before the first store (the stores to dst would
        // otherwise order every load behind them: one memory latency per element)
        long long cells[RL];
        double2 old[RL];
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          const int ia = bi + r * (N / RL);
          const bool inb = bA.has(ia) && iz < Kz;
          const int ix = AX == 0 ? ia : io, iy = AX == 0 ? io : ia;
          cells[r] = inb ? remap_cell(g, A, ix, iy, iz) : -1;
          old[r] = make_double2(0.0, 0.0);
          if (MODE == 1 && inb && !A.first) old[r] = A.dst[cells[r]];
        }
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          const int ia = bi + r * (N / RL);
          const bool inb = cells[r] >= 0;
          const int ix = AX == 0 ? ia : io, iy = AX == 0 ? io : ia;
          const long long cell = cells[r];
          const double2 t = u[s][r];
          if (MODE == 0) {
            if (inb) A.dst[cell] = t;
          } else if (inb) {
            double kx = g.kax[0][ix], ky = g.kax[1][iy], kz = g.kax[2][iz];
            double inv = inv_norm_eps(kx * kx + ky * ky + kz * kz);
            double yk = sww_ylm_dyn(A.lm, kx * inv, ky * inv, kz * inv);
            double2 o = old[r];
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
            A.dst[cell] = o;
          }
        }
      }
    } else {
      // shell binning.  The spectrum goes back to the tile (natural order) so that only box rows are
      // visited and the global loads (bin index, then G_l') can be batched four elements deep.
      __syncthreads();   // all threads are done reading the tile in the last stage
#pragma unroll
      for (int s = 0; s < 16 / RL; ++s) {
        const int w_ = tid + s * NT, bi = w_ >> 3, cc = w_ & 7;
#pragma unroll
        for (int r = 0; r < RL; ++r) tile[(bi + r * (N / RL)) * FFT_LDT + cc] = u[s][r];
      }
      __syncthreads();
      // |k| is even in the transformed axis, so the cells (+j, iz) and (-j, iz) share a bin: each lane takes
      // such a pair.  Lane = (16 consecutive j) x (2 consecutive iz); along j the bin index only changes in
      // contiguous runs, so a segmented shuffle reduction over the 16 same-parity lanes (4 steps) leaves
      // every run's sum in its head lane, and heads of one parity hold distinct bins: two conflict-free
      // rounds of adds into the warp-private histogram.  Fixed order -> deterministic.
      constexpr int GRP = 2;
      const bool full = (bA.K == N);
      const int npos = full ? N / 2 + 1 : bA.b + 1;
      const int nsteps = (npos + 15) >> 4;
      const int nitems = nsteps * 4;                     // (step, iz pair)
      const int par = lane & 1, sub = lane >> 1;
      for (int q0 = warp; q0 < nitems; q0 += GRP * NW) { // uniform per warp (collectives below)
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
          mir[j] = inb && ja > 0 && (!full || ja < N / 2);
          cellp[j] = row_of<AX>(jq[j], io, g.ny) * g.nzh + iz;
          {
            const int jm = mir[j] ? N - jq[j] : 0;
            gcp[j] = AX == 0 ? remap_cell(g, A, jq[j], io, iz) : remap_cell(g, A, io, jq[j], iz);
            gcm[j] = AX == 0 ? remap_cell(g, A, jm, io, iz) : remap_cell(g, A, io, jm, iz);
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
            const double2 t = tile[jq[j] * FFT_LDT + ccq[j]];
            double2 tm = make_double2(0.0, 0.0);
            if (mir[j]) tm = tile[(N - jq[j]) * FFT_LDT + ccq[j]];
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
    for (int j = tid; j < nb; j += NT) {
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
      for (int j = tid; j < nb; j += NT) {
        double sm = 0.0;
        for (int bl = 0; bl < nblk; ++bl) sm += partial[(long long)bl * nb + j];
        out_row[j] = sm * A.scale;
      }
      if (tid == 0) *counter = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Z pass.  8 rows (ix, iy0..iy0+7) per tile, half-length H = nz/2 complex FFTs.
//   INV: complex rows Q[row][0..Kz_in) (pitch Kzp_in)  -> real row x[0..nz)
//   FWD: real row -> complex rows R[row][0..Kz_out) (pitch Kzp_out)
//   INV && FWD: x * wmap * scale goes from the registers of the last inverse stage straight into the
//               first forward stage (reversed radix plan): the real-space map never leaves registers
//   INV only : real_out[row][z] = x * scale [* wmap] [* Y_lm(r^)]  (optionally accumulated)
//   FWD only : x = real_in[row][z] [* wmap] [* Y_lm(r^)]
// ------------------------------------------------------------------------------------------
struct ZArgs {
  const double2* Q;
  int Kz_in, Kzp_in;
  double2* R;
  int Kz_out, Kzp_out;
  const double* real_in;
  double* real_out;
  const double* wmap;
  double scale;
  int lm;          // >= 0: multiply the real row by Y_lm(r^)
  int accumulate;  // INV only: real_out += ...
  // INV only: nb > 1 sums nb inverse transforms (planes Q + b*q_stride), each times Y_{lms[b]}(r^) (or 1)
  int nb;
  int lms[SWW_MAX_ZBATCH];
  long long q_stride;
  long long r_stride;    // INV && FWD: nb > 1 transforms nb planes (Q + b*q_stride -> R + b*r_stride) per weight load
  const double* ymaps;   // optional cache [lm-1][N] of Y_lm(r^)
  long long N;
  // Tile-major intermediates (fused row chain, z16p kernel only): instead of rows [nx*ny][Kzp] the complex planes are laid
  // out as the 8-column tiles of the adjacent column pass, [f][z-tile][ia][8] with ia along that pass's transform axis
  // (q_tile / r_tile = 1: x, 2: y, 0: row-major).  The column pass then moves whole tiles as ONE contiguous block
  // (64 KB at N = 512) and this FP64-bound kernel absorbs the 128-byte pieces.
  int q_tile, r_tile;
};

// ------------------------------------------------------------------------------------------
// Z pass, warp-per-row variant.  Each warp owns RW = max(1, 256/H) rows and a private padded
// shared-memory buffer; stages are separated by __syncwarp only, so the warps of an SM are fully
// independent and hide each other's global-memory latency (the CTA-cooperative variant above stalls
// all its warps at every barrier).  Radix plans end (inverse) / start (forward) with radix 8, so the
// seam between the inverse and the forward transform is register-resident:
//   H = 64: 8,8   H = 128: 4,4,8   H = 256: 4,8,8   H = 512: 8,8,8      (forward: reversed)
// Element index i of a row lives at i + (i >> 3) (one pad slot per 8): conflict-free 128-bit accesses
// for every stage pattern.
// ------------------------------------------------------------------------------------------
template <int H> struct WPlan;
template <> struct WPlan<64>  { static constexpr int NS = 2; static constexpr int R[3] = {8, 8, 1}; };
template <> struct WPlan<128> { static constexpr int NS = 3; static constexpr int R[3] = {4, 4, 8}; };
template <> struct WPlan<256> { static constexpr int NS = 3; static constexpr int R[3] = {4, 8, 8}; };
template <> struct WPlan<512> { static constexpr int NS = 3; static constexpr int R[3] = {8, 8, 8}; };

__device__ __forceinline__ int zpad(int i) { return i + (i >> 3); }

template <int H, bool REV, int S>
struct WRadix {
  static constexpr int value = REV ? WPlan<H>::R[WPlan<H>::NS - 1 - S] : WPlan<H>::R[S];
};
template <int H, bool REV, int S>
struct WProd {
  static constexpr int value = WProd<H, REV, S - 1>::value * WRadix<H, REV, S - 1>::value;
};
template <int H, bool REV>
struct WProd<H, REV, 0> {
  static constexpr int value = 1;
};

// geometry of the warp's work: RW rows, butterfly index vb -> (row rho, bi)
template <int H>
struct WGeo {
  static constexpr int RW = (H >= 256) ? 1 : 256 / H;
  static constexpr int ROWLEN = H + H / 8 + 2;      // padded row (+ room for the Nyquist element at index H)
  static constexpr int ELEMS = RW * H / 32;         // complex values per lane (8, or 16 for H = 512)
};

// Per-stage twiddle tables: tab[rl*P + k] = W_H^(k * 2^rl * H/(P R)), rl = 0,1,2 (w^1, w^2, w^4), k < P.
// Consecutive lanes read consecutive k: conflict-free (the single W_H table would be read with
// strides 1, 2, 4 ... times H/(P R): up to 4-way conflicts, 24 % of all shared wavefronts under ncu).
template <int H, bool REV, int S>
struct WTabOff {   // offset of stage S's table (stage 0 has none)
  static constexpr int value = WTabOff<H, REV, S - 1>::value + (S - 1 >= 1 ? 3 * WProd<H, REV, S - 1>::value : 0);
};
template <int H, bool REV>
struct WTabOff<H, REV, 0> {
  static constexpr int value = 0;
};
template <int H, bool REV>
struct WTabSize {
  static constexpr int value = WTabOff<H, REV, WPlan<H>::NS>::value;
};

template <int H, bool REV, int S>
__device__ __forceinline__ void w_fill_tab(double2* tabs, const double2* __restrict__ twh_g, int tid, int nthreads) {
  if constexpr (S < WPlan<H>::NS) {
    if constexpr (S >= 1) {
      constexpr int R = WRadix<H, REV, S>::value;
      constexpr int P = WProd<H, REV, S>::value;
      double2* tab = tabs + WTabOff<H, REV, S>::value;
      for (int e = tid; e < 3 * P; e += nthreads) {
        const int rl = e / P, k = e - rl * P, r = 1 << rl;
        tab[e] = (r < R) ? twh_g[k * r * (H / (P * R))] : make_double2(1.0, 0.0);
      }
    }
    w_fill_tab<H, REV, S + 1>(tabs, twh_g, tid, nthreads);
  }
}

template <int R, int DIR>
__device__ __forceinline__ void w_twiddle(const double2* tab, int P, int k, double2* u) {
  double2 w[R];
#pragma unroll
  for (int r = 1, rl = 0; r < R; r <<= 1, ++rl) {
    w[r] = tab[rl * P + k];
    if (DIR < 0) w[r].y = -w[r].y;
  }
#pragma unroll
  for (int r = 3; r < R; ++r) {
    if ((r & (r - 1)) != 0) {
      int hb = r >= 4 ? 4 : 2;
      w[r] = c_mul(w[hb], w[r - hb]);
    }
  }
#pragma unroll
  for (int r = 1; r < R; ++r) u[r] = c_mul(u[r], w[r]);
}

// middle stage S (load, twiddle, butterfly, store) entirely inside the warp
template <int H, int DIR, bool REV, int S>
__device__ __forceinline__ void w_mid_stage(double2* buf, const double2* tw, int lane) {
  constexpr int R = WRadix<H, REV, S>::value;
  constexpr int P = WProd<H, REV, S>::value;
  constexpr int BPR = H / R;                         // butterflies per row
  constexpr int IPT = WGeo<H>::RW * BPR / 32;
  double2 u[IPT][R];
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
    const double2* row = buf + rho * WGeo<H>::ROWLEN;
#pragma unroll
    for (int r = 0; r < R; ++r) u[s][r] = row[zpad(bi + r * BPR)];
    w_twiddle<R, DIR>(tw + WTabOff<H, REV, S>::value, P, bi & (P - 1), u[s]);
    dftR<R, DIR>(u[s]);
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < IPT; ++s) {
    const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
    double2* row = buf + rho * WGeo<H>::ROWLEN;
    const int k = bi & (P - 1), j = (bi - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) row[zpad(j + r * P)] = u[s][r];
  }
  __syncwarp();
}

template <int H, int DIR, bool REV>
__device__ __forceinline__ void w_mid_stages(double2* buf, const double2* tw, int lane) {
  if constexpr (WPlan<H>::NS == 3) w_mid_stage<H, DIR, REV, 1>(buf, tw, lane);
}

#ifndef ZW_INV_ONE_CTA
#define ZW_INV_ONE_CTA 0   // 1: inverse-only variants at 255 registers / 8 warps per SM -- measured slower (the kernel is latency bound)
#endif
template <int H, bool INV, bool FWD, bool YLM>
__global__ void __launch_bounds__(256, ((H >= 512 && INV) || (INV && !FWD && ZW_INV_ONE_CTA)) ? 1 : 2)
pass_zw_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = 8;
  constexpr int RW = WGeo<H>::RW, ROWLEN = WGeo<H>::ROWLEN;
  constexpr int SE = 8;                       // seam radix
  constexpr int BPS = H / SE;                 // butterflies per row at the seam; outputs at bi + r*BPS
  constexpr int IPS = RW * BPS / 32;          // seam butterflies per lane
  constexpr int TI = WTabSize<H, false>::value, TF = WTabSize<H, true>::value;
  double2* twn = reinterpret_cast<double2*>(smem_raw);   // [H+1]  exp(-2 pi i j / nz)
  double2* tabi = twn + (H + 8);                         // per-stage twiddles of the inverse plan
  double2* tabf = tabi + TI;                             // ... of the (reversed) forward plan
  double2* bufs = tabf + TF;                             // [WARPS][RW*ROWLEN (+8)]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double2* buf = bufs + warp * (RW * ROWLEN + 8);
  // inverse variants: the complex input rows of the NEXT plane stream into a second per-warp buffer (cp.async)
  // while the current plane is transformed, so no row waits on a DRAM round trip
  constexpr bool PREF = INV;
  double2* inb = PREF ? bufs + WARPS * (RW * ROWLEN + 8) + warp * (RW * ROWLEN + 8) : buf;
  for (int i = tid; i <= H; i += 256) twn[i] = twn_g[i];
  w_fill_tab<H, false, 0>(tabi, twh_g, tid, 256);
  w_fill_tab<H, true, 0>(tabf, twh_g, tid, 256);
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / RW;
  if (PREF) {
    const long long g0 = (long long)blockIdx.x * WARPS + warp;
    if (g0 < ngroups) {
      for (int rho = 0; rho < RW; ++rho) {
        const double2* src = A.Q + (g0 * RW + rho) * A.Kzp_in;
        double2* row = inb + rho * ROWLEN;
        for (int k = lane; k < A.Kz_in; k += 32) {
          unsigned sa = (unsigned)__cvta_generic_to_shared(row + zpad(k));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    }
  }
  for (long long grp = (long long)blockIdx.x * WARPS + warp; grp < ngroups; grp += (long long)gridDim.x * WARPS) {
    const long long row0 = grp * RW;
    const int nfb_all = (INV && FWD && A.nb > 1) ? A.nb : 1;   // planes transformed one after the other (fused variant)
    const int npl = (INV && A.nb > 1) ? A.nb : 1;                // planes of one row group in prefetch order
    int cur_fb = 0;
    double2 u[IPS][SE];
    // early prefetch of the weights only where the registers allow it: not with 16 values per lane
    // (H = 512) and not in the batched inverse-only variant (which carries an accumulator)
    constexpr bool EARLY_W = (IPS == 1) && (FWD || !INV);
    double2 wv[EARLY_W ? IPS : 1][SE];
    // weights of the real-space step: issued first, consumed after the inverse transform
    if (EARLY_W && A.wmap) {
#pragma unroll
      for (int s = 0; s < IPS; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
#pragma unroll
        for (int r = 0; r < SE; ++r)
          wv[s][r] = *reinterpret_cast<const double2*>(A.wmap + (row0 + rho) * g.nz + 2 * (bi + r * BPS));
      }
    }
    // inverse transform of the rows of one plane -> registers u (outputs at m = bi + r*BPS)
    auto prefetch_rows = [&](const double2* __restrict__ Qbase, long long r0) {
      for (int rho = 0; rho < RW; ++rho) {
        const double2* src = Qbase + (r0 + rho) * A.Kzp_in;
        double2* row = inb + rho * ROWLEN;
        for (int k = lane; k < A.Kz_in; k += 32) {
          unsigned sa = (unsigned)__cvta_generic_to_shared(row + zpad(k));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    auto do_inverse = [&](const double2* __restrict__ Qbase) {
      // 1. complex input rows -> buffer (natural positions, padded)
      __syncwarp();
      if (PREF) {
        asm volatile("cp.async.wait_group 0;\n" ::);
      } else
      for (int rho = 0; rho < RW; ++rho) {
        const double2* src = Qbase + (row0 + rho) * A.Kzp_in;
        double2* row = buf + rho * ROWLEN;
        for (int k0 = lane; k0 < A.Kz_in; k0 += 128) {
          double2 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k0 + 32 * q < A.Kz_in) v[q] = src[k0 + 32 * q];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k0 + 32 * q < A.Kz_in) row[zpad(k0 + 32 * q)] = v[q];
        }
      }
      __syncwarp();
      // 2. stage 0 of the inverse transform, inputs Y[k] built on the fly
      {
        constexpr int R = WRadix<H, false, 0>::value;
        constexpr int BPR = H / R;
        constexpr int IPT = RW * BPR / 32;
        double2 y[IPT][R];
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
          const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
          const double2* row = inb + rho * ROWLEN;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int k = bi + r * BPR, hk = H - k;
            y[s][r] = make_double2(0.0, 0.0);
            if (k < A.Kz_in || hk < A.Kz_in) {   // both halves empty for most k of a low shell
              double2 zk = (k < A.Kz_in) ? row[zpad(k)] : make_double2(0.0, 0.0);
              double2 zh = (hk < A.Kz_in) ? row[zpad(hk)] : make_double2(0.0, 0.0);
              if (k == 0) {
                zk.y = 0.0;
                zh.y = 0.0;
              }
              double2 a = make_double2(zk.x + zh.x, zk.y - zh.y);
              double2 d = make_double2(zk.x - zh.x, zk.y + zh.y);
              double2 w = twn[k];
              w.y = -w.y;
              double2 b = c_mul(d, w);
              y[s][r] = make_double2(a.x - b.y, a.y + b.x);
            }
          }
          dftR<R, -1>(y[s]);
        }
        __syncwarp();
        if (PREF) {
          // the staging buffer is free again: start the copy of the next plane of this warp
          const bool wrap = cur_fb + 1 >= npl;
          const long long ngrp = wrap ? grp + (long long)gridDim.x * WARPS : grp;
          if (ngrp < ngroups) prefetch_rows(A.Q + (wrap ? 0LL : (long long)(cur_fb + 1) * A.q_stride), ngrp * RW);
        }
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
          const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
          double2* row = buf + rho * ROWLEN;
#pragma unroll
          for (int r = 0; r < R; ++r) row[zpad(bi * R + r)] = y[s][r];
        }
        __syncwarp();
      }
      w_mid_stages<H, -1, false>(buf, tabi, lane);
      // last inverse stage (radix 8) -> registers, outputs at m = bi + r*BPS
#pragma unroll
      for (int s = 0; s < IPS; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
        const double2* row = buf + rho * ROWLEN;
#pragma unroll
        for (int r = 0; r < SE; ++r) u[s][r] = row[zpad(bi + r * BPS)];
        w_twiddle<SE, -1>(tabi + WTabOff<H, false, WPlan<H>::NS - 1>::value, BPS, bi, u[s]);
        dftR<SE, -1>(u[s]);
      }
    };
    // fused variant: nb > 1 transforms nb planes per weight load (the multipoles of one k bin)
    const int nfb = nfb_all;
    for (int fb = 0; fb < nfb; ++fb) {
    cur_fb = fb;
    if (INV) {
      if (!FWD && A.nb > 1) {
        // sum of nb inverse transforms, each times Y_lm(r^) (or 1): the real-space map is written once
        double2 acc[IPS][SE];
#pragma unroll
        for (int s = 0; s < IPS; ++s)
#pragma unroll
          for (int r = 0; r < SE; ++r) acc[s][r] = make_double2(0.0, 0.0);
        for (int bq = 0; bq < A.nb; ++bq) {
          cur_fb = bq;
          do_inverse(A.Q + (long long)bq * A.q_stride);
          const int lmb = A.lms[bq];
          if (lmb >= 1 && A.ymaps) {
            const double* ym = A.ymaps + (long long)(lmb - 1) * A.N;
#pragma unroll
            for (int s = 0; s < IPS; ++s) {
              const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
              double2 yv[SE];
#pragma unroll
              for (int r = 0; r < SE; ++r)
                yv[r] = *reinterpret_cast<const double2*>(ym + (row0 + rho) * g.nz + 2 * (bi + r * BPS));
#pragma unroll
              for (int r = 0; r < SE; ++r) {
                acc[s][r].x += yv[r].x * u[s][r].x;
                acc[s][r].y += yv[r].y * u[s][r].y;
              }
            }
            continue;
          }
#pragma unroll
          for (int s = 0; s < IPS; ++s) {
            const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
            // Y_lm(r^) along the row: degree-l Horner polynomial in z with row coefficients, times r^-l (ylm_gen.cuh)
            double Ac[3], xy2 = 0.0;
            bool odd = false;
            const int deg = lmb >= 0 ? SWW_YLM_ZDEG(lmb) : 0;
            if (lmb >= 0) {
              const long long row = row0 + rho;
              const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
              const double X = g.rax[0][ix], Y = g.rax[1][iy];
              xy2 = X * X + Y * Y;
              sww_ylm_zpack(lmb, X, Y, Ac, &odd);
            }
#pragma unroll
            for (int r = 0; r < SE; ++r) {
              double f0 = 1.0, f1 = 1.0;
              if (lmb >= 0) {
                const int m = bi + r * BPS;
                const double Z0 = g.rax[2][2 * m], Z1 = g.rax[2][2 * m + 1];
                const double w0 = Z0 * Z0, w1 = Z1 * Z1;
                f0 = sww_ylm_zpacked(deg, Ac, odd, Z0, w0, deg ? sww_rcp_pos(xy2 + w0) : 1.0);
                f1 = sww_ylm_zpacked(deg, Ac, odd, Z1, w1, deg ? sww_rcp_pos(xy2 + w1) : 1.0);
              }
              acc[s][r].x += f0 * u[s][r].x;
              acc[s][r].y += f1 * u[s][r].y;
            }
          }
        }
#pragma unroll
        for (int s = 0; s < IPS; ++s)
#pragma unroll
          for (int r = 0; r < SE; ++r) u[s][r] = acc[s][r];
      } else {
        do_inverse(A.Q + (long long)fb * A.q_stride);
      }
    } else {
#pragma unroll
      for (int s = 0; s < IPS; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
#pragma unroll
        for (int r = 0; r < SE; ++r)
          u[s][r] = *reinterpret_cast<const double2*>(A.real_in + (row0 + rho) * g.nz + 2 * (bi + r * BPS));
      }
    }
    // real-space step on registers
#pragma unroll
    for (int s = 0; s < IPS; ++s) {
      const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
      const int ws = EARLY_W ? s : 0;
      if (!EARLY_W && A.wmap) {
#pragma unroll
        for (int r = 0; r < SE; ++r)
          wv[0][r] = *reinterpret_cast<const double2*>(A.wmap + (row0 + rho) * g.nz + 2 * (bi + r * BPS));
      }
      // on-the-fly Y_lm(r^): row coefficients of the degree-l polynomial in z (ylm_gen.cuh), hoisted out of the cell loop
      const bool ylm_fly = YLM && !(INV && !FWD && A.nb > 1) && !(A.lm >= 1 && A.ymaps);
      double Ac[3], xy2 = 0.0;
      bool odd = false;
      const int deg = ylm_fly ? SWW_YLM_ZDEG(A.lm) : 0;
      if (ylm_fly) {
        const long long row = row0 + rho;
        const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
        const double X = g.rax[0][ix], Y = g.rax[1][iy];
        xy2 = X * X + Y * Y;
        sww_ylm_zpack(A.lm, X, Y, Ac, &odd);
      }
#pragma unroll
      for (int r = 0; r < SE; ++r) {
        double f0 = A.scale, f1 = A.scale;
        if (A.wmap) {
          f0 *= wv[ws][r].x;
          f1 *= wv[ws][r].y;
        }
        if (YLM && !(INV && !FWD && A.nb > 1) && A.lm >= 1 && A.ymaps) {
          const double2 yv = *reinterpret_cast<const double2*>(A.ymaps + (long long)(A.lm - 1) * A.N + (row0 + rho) * g.nz +
                                                               2 * (bi + r * BPS));
          f0 *= yv.x;
          f1 *= yv.y;
        } else if (ylm_fly) {
          const int m = bi + r * BPS;
          const double Z0 = g.rax[2][2 * m], Z1 = g.rax[2][2 * m + 1];
          const double w0 = Z0 * Z0, w1 = Z1 * Z1;
          f0 *= sww_ylm_zpacked(deg, Ac, odd, Z0, w0, deg ? sww_rcp_pos(xy2 + w0) : 1.0);
          f1 *= sww_ylm_zpacked(deg, Ac, odd, Z1, w1, deg ? sww_rcp_pos(xy2 + w1) : 1.0);
        }
        u[s][r].x *= f0;
        u[s][r].y *= f1;
      }
    }
    if (INV && !FWD) {
      if (A.accumulate) {
#pragma unroll
        for (int s = 0; s < IPS; ++s) {
          const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
#pragma unroll
          for (int r = 0; r < SE; ++r)
            wv[0][r] = *reinterpret_cast<const double2*>(A.real_out + (row0 + rho) * g.nz + 2 * (bi + r * BPS));
#pragma unroll
          for (int r = 0; r < SE; ++r) {
            u[s][r].x += wv[0][r].x;
            u[s][r].y += wv[0][r].y;
          }
        }
      }
#pragma unroll
      for (int s = 0; s < IPS; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
#pragma unroll
        for (int r = 0; r < SE; ++r)
          *reinterpret_cast<double2*>(A.real_out + (row0 + rho) * g.nz + 2 * (bi + r * BPS)) = u[s][r];
      }
    }
    if (FWD) {
      // forward stage 0 (radix 8, reversed plan) straight from the registers
      __syncwarp();
#pragma unroll
      for (int s = 0; s < IPS; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPS, bi = vb - rho * BPS;
        double2* row = buf + rho * ROWLEN;
        dftR<SE, 1>(u[s]);
#pragma unroll
        for (int r = 0; r < SE; ++r) row[zpad(bi * SE + r)] = u[s][r];
      }
      __syncwarp();
      w_mid_stages<H, 1, true>(buf, tabf, lane);
      // last forward stage -> natural order in the buffer
      {
        constexpr int R = WRadix<H, true, WPlan<H>::NS - 1>::value;
        constexpr int BPR = H / R;
        constexpr int IPT = RW * BPR / 32;
        double2 v[IPT][R];
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
          const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
          const double2* row = buf + rho * ROWLEN;
#pragma unroll
          for (int r = 0; r < R; ++r) v[s][r] = row[zpad(bi + r * BPR)];
          w_twiddle<R, 1>(tabf + WTabOff<H, true, WPlan<H>::NS - 1>::value, BPR, bi, v[s]);
          dftR<R, 1>(v[s]);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
          const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
          double2* row = buf + rho * ROWLEN;
#pragma unroll
          for (int r = 0; r < R; ++r) row[zpad(bi + r * BPR)] = v[s][r];
        }
        __syncwarp();
      }
      // T[k] = 1/2 (V[k] + conj V[H-k]) - i/2 e^{-2 pi i k/nz} (V[k] - conj V[H-k]),  V[H] = V[0]
      for (int rho = 0; rho < RW; ++rho) {
        const double2* row = buf + rho * ROWLEN;
        double2* dst = A.R + (long long)fb * A.r_stride + (row0 + rho) * A.Kzp_out;
        for (int k = lane; k < A.Kz_out; k += 32) {
          int k1 = (k == H) ? 0 : k, k2 = (k == 0 || k == H) ? 0 : H - k;
          double2 v1 = row[zpad(k1)], v2 = row[zpad(k2)];
          double2 ev = make_double2(0.5 * (v1.x + v2.x), 0.5 * (v1.y - v2.y));
          double2 dv = make_double2(0.5 * (v1.x - v2.x), 0.5 * (v1.y + v2.y));
          double2 od = c_mul(make_double2(dv.y, -dv.x), twn[k]);
          dst[k] = make_double2(ev.x + od.x, ev.y + od.y);
        }
      }
    }
    }   // fb
  }
}

// ------------------------------------------------------------------------------------------
// Complex z pass (C2C transforms: the ft / ift mirrors and the maximum-likelihood operators, which keep full complex
// maps): in-place FFT of length H = nz along the contiguous axis, one warp per row (two or four rows for short rows), the
// stages of the warp-per-row machinery above with no real-data pre / post-processing.
// ------------------------------------------------------------------------------------------
template <int H, int DIR>
__global__ void __launch_bounds__(256, 2)
pass_zc_kernel(long long nrows, double2* __restrict__ data, const double2* __restrict__ tw_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = 8;
  constexpr int RW = WGeo<H>::RW, ROWLEN = WGeo<H>::ROWLEN;
  constexpr int T = WTabSize<H, false>::value;
  double2* tabs = reinterpret_cast<double2*>(smem_raw);
  double2* bufs = tabs + T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double2* buf = bufs + warp * (RW * ROWLEN + 8);
  w_fill_tab<H, false, 0>(tabs, tw_g, tid, 256);
  __syncthreads();
  const long long ngroups = nrows / RW;
  for (long long grp = (long long)blockIdx.x * WARPS + warp; grp < ngroups; grp += (long long)gridDim.x * WARPS) {
    double2* rows = data + grp * RW * H;
    __syncwarp();
#pragma unroll
    for (int rho = 0; rho < RW; ++rho)
      for (int i = lane; i < H; i += 32) buf[rho * ROWLEN + zpad(i)] = rows[rho * H + i];
    __syncwarp();
    {   // stage 0: no twiddles
      constexpr int R = WRadix<H, false, 0>::value;
      constexpr int BPR = H / R;
      constexpr int IPT = RW * BPR / 32;
      double2 u[IPT][R];
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
        const double2* row = buf + rho * ROWLEN;
#pragma unroll
        for (int r = 0; r < R; ++r) u[s][r] = row[zpad(bi + r * BPR)];
        dftR<R, DIR>(u[s]);
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < IPT; ++s) {
        const int vb = lane + 32 * s, rho = vb / BPR, bi = vb - rho * BPR;
        double2* row = buf + rho * ROWLEN;
#pragma unroll
        for (int r = 0; r < R; ++r) row[zpad(bi * R + r)] = u[s][r];
      }
      __syncwarp();
    }
    w_mid_stage<H, DIR, false, 1>(buf, tabs, lane);
    if constexpr (WPlan<H>::NS == 3) w_mid_stage<H, DIR, false, 2>(buf, tabs, lane);
#pragma unroll
    for (int rho = 0; rho < RW; ++rho)
      for (int i = lane; i < H; i += 32) rows[rho * H + i] = buf[rho * ROWLEN + zpad(i)];
  }
}

// ------------------------------------------------------------------------------------------
// Fused z pass for nz = 512 (H = 256 = 16 x 16): c2r * wmap * r2c with TWO exchanges through shared
// memory per row instead of the eleven of the generic warp-per-row kernel above, which is bound by the
// 128 B/clk shared-memory crossbar, not by the FP64 pipe.
//   * a warp owns two rows; lane = (row rho, butterfly b < 16) holds 16 complex values of its row;
//   * both half-length transforms are two radix-16 stages: stage 0 straight from registers (inverse: the
//     Hermitian pre-processing of the staged input row; forward: the weighted outputs of the inverse),
//     one exchange (store at 16 b + r, load at b + 16 r, pitch 17 -> conflict-free), stage 1 with the
//     inter-stage twiddles W_256^(b r) built from four lane-constant registers (w, w^2, w^4, w^8);
//   * the r2c post-processing pairs V[k] with V[H-k], which lives in lane 16 - b at register 15 - r:
//     one shuffle per output, no pass through shared memory; only k < Kz_out <= 128 is produced (r < 8);
//   * input rows of the next plane stream into a per-warp staging buffer with cp.async.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

template <int DIR>
__device__ __forceinline__ void twiddle16(double2* u, double2 w1, double2 w2, double2 w4, double2 w8) {
  if (DIR < 0) {
    w1.y = -w1.y;
    w2.y = -w2.y;
    w4.y = -w4.y;
    w8.y = -w8.y;
  }
#pragma unroll
  for (int j = 8; j < 16; ++j) u[j] = c_mul(u[j], w8);
  u[1] = c_mul(u[1], w1);
  u[9] = c_mul(u[9], w1);
  u[2] = c_mul(u[2], w2);
  u[10] = c_mul(u[10], w2);
  {
    const double2 w3 = c_mul(w1, w2);
    u[3] = c_mul(u[3], w3);
    u[11] = c_mul(u[11], w3);
    const double2 w7 = c_mul(w3, w4);
    u[7] = c_mul(u[7], w7);
    u[15] = c_mul(u[15], w7);
  }
  u[4] = c_mul(u[4], w4);
  u[12] = c_mul(u[12], w4);
  {
    const double2 w5 = c_mul(w1, w4);
    u[5] = c_mul(u[5], w5);
    u[13] = c_mul(u[13], w5);
  }
  {
    const double2 w6 = c_mul(w2, w4);
    u[6] = c_mul(u[6], w6);
    u[14] = c_mul(u[14], w6);
  }
}

// the same from a shared-memory table tab[r * 16 + b] = W_256^(b r): 15 conflict-free loads instead of 11 complex products
template <int DIR>
__device__ __forceinline__ void twiddle16_tab(double2* u, const double2* __restrict__ tab, int b) {
#pragma unroll
  for (int r = 1; r < 16; ++r) {
    double2 w = tab[r * 16 + b];
    if (DIR < 0) w.y = -w.y;
    u[r] = c_mul(u[r], w);
  }
}

#define Z16_H 256
#define Z16_ROW 272   // 256 + one pad slot per 16

template <int WARPS, int MINB, bool WREG>
__global__ void __launch_bounds__(WARPS * 32, MINB)
pass_z16_fused_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g, int in_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = Z16_H, ROW = Z16_ROW;
  double2* twn = reinterpret_cast<double2*>(smem_raw);   // [H + 8]  exp(-2 pi i j / nz), j <= H
  double2* tab = twn + (H + 8);                          // [16][16] inter-stage twiddles W_256^(b r) at [r * 16 + b]
  double2* bufs = tab + 256;                             // [WARPS][2 * ROW] exchange buffers
  double2* inbs = bufs + WARPS * 2 * ROW;                // [WARPS][2 * in_len] staged input rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = lane >> 4, b = lane & 15;
  double2* buf = bufs + warp * 2 * ROW + rho * ROW;
  double2* inw = inbs + warp * 2 * in_len;
  const double2* inb = inw + rho * in_len;
  for (int i = tid; i <= H; i += WARPS * 32) twn[i] = twn_g[i];
  for (int i = tid; i < 256; i += WARPS * 32) tab[i] = twh_g[((i >> 4) * (i & 15)) & (H - 1)];
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / 2;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int nfb = A.nb > 1 ? A.nb : 1;
  const int Kz_in = A.Kz_in, Kz_out = A.Kz_out;
  const double half = 0.5 * A.scale;   // 1/2 of the r2c post-processing and the caller's scale, folded into the weights
  auto prefetch_rows = [&](const double2* __restrict__ Qbase, long long r0) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const double2* src = Qbase + (r0 + rr) * A.Kzp_in;
      double2* row = inw + rr * in_len;
      for (int k = lane; k < Kz_in; k += 32) {
        unsigned sa = (unsigned)__cvta_generic_to_shared(row + k);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  long long grp = (long long)blockIdx.x * WARPS + warp;
  if (grp < ngroups) prefetch_rows(A.Q, grp * 2);
  for (; grp < ngroups; grp += gstep) {
    const long long row = grp * 2 + rho;
    const double* __restrict__ wrow = A.wmap + row * g.nz + 2 * b;
    // the weights of this row pair are needed after the inverse transform of every plane: pull them towards L2 now
    // WREG: the weights of the row pair sit in registers for all planes of the group (loaded here, consumed after
    // the first inverse transform); otherwise they are pulled towards L2 now and loaded at the point of use
    double2 wreg[WREG ? 16 : 1];
    if (WREG) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 wv = *reinterpret_cast<const double2*>(wrow + 32 * r);
        wreg[WREG ? r : 0] = make_double2(wv.x * half, wv.y * half);
      }
    } else {
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A.wmap + grp * 2 * g.nz + lane * 16));
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A.wmap + grp * 2 * g.nz + 512 + lane * 16));
    }
    for (int fb = 0; fb < nfb; ++fb) {
      double2 u[16];
      // ---- inverse: Hermitian pre-processing of the staged row -> stage 0 (radix 16, no twiddles)
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u[r] = make_double2(0.0, 0.0);
        if (16 * r < Kz_in || 241 - 16 * r < Kz_in) {   // warp-uniform: some k or H-k of this block is kept
          const int k = b + 16 * r, hk = H - k;
          if (k < Kz_in || hk < Kz_in) {
            double2 zk = (k < Kz_in) ? inb[k] : make_double2(0.0, 0.0);
            double2 zh = (hk < Kz_in) ? inb[hk] : make_double2(0.0, 0.0);
            if (k == 0) {
              zk.y = 0.0;
              zh.y = 0.0;
            }
            const double2 a = make_double2(zk.x + zh.x, zk.y - zh.y);
            const double2 d = make_double2(zk.x - zh.x, zk.y + zh.y);
            double2 w = twn[k];
            w.y = -w.y;
            const double2 bb = c_mul(d, w);
            u[r] = make_double2(a.x - bb.y, a.y + bb.x);
          }
        }
      }
      __syncwarp();
      {
        // the staging buffer is free again: start the copy of the next plane of this warp
        const bool wrap = fb + 1 >= nfb;
        const long long ngrp = wrap ? grp + gstep : grp;
        if (ngrp < ngroups) prefetch_rows(A.Q + (wrap ? 0LL : (long long)(fb + 1) * A.q_stride), ngrp * 2);
      }
      dft16<-1>(u);
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      twiddle16_tab<-1>(u, tab, b);
      dft16<-1>(u);
      // ---- real-space step: x[2m], x[2m+1] at m = b + 16 r, times the weight map
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        double2 wv = WREG ? wreg[WREG ? r : 0] : *reinterpret_cast<const double2*>(wrow + 32 * r);
        if (!WREG) wv = make_double2(wv.x * half, wv.y * half);
        u[r].x *= wv.x;
        u[r].y *= wv.y;
      }
      // ---- forward: stage 0 from the same registers, one exchange, stage 1
      dft16<1>(u);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      twiddle16_tab<1>(u, tab, b);
      dft16<1>(u);
      // ---- r2c post-processing: T[k] = 1/2 (V[k] + conj V[H-k]) - i/2 e^{-2 pi i k/nz} (V[k] - conj V[H-k])
      double2* __restrict__ dst = A.R + (long long)fb * A.r_stride + row * A.Kzp_out;
      const int src_lane = (rho << 4) | ((16 - b) & 15);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double2 v2;
        v2.x = __shfl_sync(0xffffffffu, u[15 - r].x, src_lane);
        v2.y = __shfl_sync(0xffffffffu, u[15 - r].y, src_lane);
        if (b == 0) v2 = u[(16 - r) & 15];
        const int k = b + 16 * r;
        if (16 * r < Kz_out && k < Kz_out) {
          const double2 v1 = u[r];
          const double2 ev = make_double2(v1.x + v2.x, v1.y - v2.y);
          const double2 dv = make_double2(v1.x - v2.x, v1.y + v2.y);
          const double2 od = c_mul(make_double2(dv.y, -dv.x), twn[k]);
          dst[k] = make_double2(ev.x + od.x, ev.y + od.y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// pass_z16p_fused_kernel: the kernel above with table-driven, branch-free pre- and post-processing (round 2).
// ncu on the kernel above (profiles/r1e) put 52 % of its stall samples in the Hermitian pre-processing and the r2c
// post-processing -- 16 + 8 basic blocks of "3 shared loads -> a 10-deep chain of dependent FP64 operations", one after
// the other -- for a fifth of its FP64 work.  Here:
//   * Kz_in <= 128 = H/2 means that of Z[k] and Z[H-k] at most ONE is kept, so
//         Y[k] = Z[k] (1 + i conj w_k)          (k < 128)       Y[k] = conj(Z[H-k]) (1 - i conj w_k)      (k > 128)
//     is ONE complex product with a table value: preA[k] = (1 - s_k, c_k), preB[j] = (1 + s_j, c_j), w = c - i s.
//     NBI = number of 16-blocks that can be non-zero at either end (template): the blocks in between are compile-time
//     zeros, all loads of the phase are issued together, and the 2 NBI products are independent;
//   * T[k] = V[k] postA[k] + conj(V[H-k]) postB[k] with postA = h (1 - s, -c), postB = h (1 + s, c), h = scale / 2:
//     all shuffles first, then 8 independent 4-deep FMA chains (8 FP64 operations per output instead of 10), and the
//     weights stay unscaled in registers (their load is only waited for after the first inverse transform).
// ------------------------------------------------------------------------------------------
template <int NBI, bool WREG, int MINB>
__global__ void __launch_bounds__(128, MINB)
pass_z16p_fused_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g, int in_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = Z16_H, ROW = Z16_ROW, WARPS = 4;
  double2* preA = reinterpret_cast<double2*>(smem_raw);   // [128]
  double2* preB = preA + 128;                            // [136] (index 128 is never kept: Kz_in <= 128)
  double2* postA = preB + 136;                           // [128]
  double2* postB = postA + 128;                          // [128]
  double2* tab = postB + 128;                            // [16][16] inter-stage twiddles W_256^(b r) at [r * 16 + b]
  double2* bufs = tab + 256;                             // [WARPS][2 * ROW] exchange buffers
  double2* inbs = bufs + WARPS * 2 * ROW;                // [WARPS][2 * in_len] staged input rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = lane >> 4, b = lane & 15;
  double2* buf = bufs + warp * 2 * ROW + rho * ROW;
  double2* inw = inbs + warp * 2 * in_len;
  const double2* inb = inw + rho * in_len;
  const double half = 0.5 * A.scale;
  for (int i = tid; i <= 128; i += WARPS * 32) {
    const double2 w = twn_g[i];                          // (c, -s)
    preB[i] = make_double2(1.0 - w.y, w.x);
    if (i < 128) {
      preA[i] = make_double2(1.0 + w.y, w.x);
      postA[i] = make_double2(half * (1.0 + w.y), -half * w.x);
      postB[i] = make_double2(half * (1.0 - w.y), half * w.x);
    }
  }
  for (int i = tid; i < 256; i += WARPS * 32) tab[i] = twh_g[((i >> 4) * (i & 15)) & (H - 1)];
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / 2;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int nfb = A.nb > 1 ? A.nb : 1;
  const int Kz_in = A.Kz_in, Kz_out = A.Kz_out;
  const int in_last = in_len - 1;
  // element (row, k) of a tile-major plane: ((f * ztiles + k / 8) * n_ax + ia) * 8 + k % 8
  auto tile_base = [&](long long row_, int axis, int Kzp, long long& zstep) -> long long {
    const int ix = (int)(row_ / g.ny), iy = (int)(row_ - (long long)ix * g.ny);
    const int ia = axis == 0 ? ix : iy, f = axis == 0 ? iy : ix, n_ax = axis == 0 ? g.nx : g.ny;
    zstep = (long long)n_ax * 8;
    return ((long long)f * (Kzp >> 3) * n_ax + ia) * 8;
  };
  auto prefetch_rows = [&](const double2* __restrict__ Qbase, long long r0) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      double2* row = inw + rr * in_len;
      if (A.q_tile > 0) {
        long long zstep;
        const double2* src = Qbase + tile_base(r0 + rr, A.q_tile - 1, A.Kzp_in, zstep);
        for (int k = lane; k < Kz_in; k += 32) {
          unsigned sa = (unsigned)__cvta_generic_to_shared(row + k);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + (k >> 3) * zstep + (k & 7)));
        }
      } else {
        const double2* src = Qbase + (r0 + rr) * A.Kzp_in;
        for (int k = lane; k < Kz_in; k += 32) {
          unsigned sa = (unsigned)__cvta_generic_to_shared(row + k);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  long long grp = (long long)blockIdx.x * WARPS + warp;
  if (grp < ngroups) prefetch_rows(A.Q, grp * 2);
  for (; grp < ngroups; grp += gstep) {
    const long long row = grp * 2 + rho;
    const double* __restrict__ wrow = A.wmap + row * g.nz + 2 * b;
    long long r_zstep = 0;
    const long long r_off = A.r_tile > 0 ? tile_base(row, A.r_tile - 1, A.Kzp_out, r_zstep) : row * A.Kzp_out;
    // WREG: the weights of the row pair stay in registers for all planes (168 registers, 12 warps / SM); otherwise they are
    // re-read per plane (L2 hits after the first) right before the inverse transform that hides their latency
    double2 wreg[16];
    if (WREG) {
#pragma unroll
      for (int r = 0; r < 16; ++r) wreg[r] = *reinterpret_cast<const double2*>(wrow + 32 * r);
    }
    for (int fb = 0; fb < nfb; ++fb) {
      double2 u[16];
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncwarp();
      // ---- inverse: Hermitian pre-processing, one table product per kept element, all loads up front
#pragma unroll
      for (int r = NBI; r < 16 - NBI; ++r) u[r] = make_double2(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < 2 * NBI; ++q) {
        const int r = q < NBI ? q : 16 - 2 * NBI + q;
        const int k = b + 16 * r;
        const int idx = q < NBI ? k : H - k;            // the kept partner: Z[k] or Z[H - k]
        double2 v = inb[idx < in_last ? idx : in_last];
        if (idx >= Kz_in) v = make_double2(0.0, 0.0);
        if (q == 0 && b == 0) v.y = 0.0;                // Z[0] is real
        if (q < NBI) {                                  // Z[k] (1 + i conj w_k)
          const double2 cw = preA[k];
          u[r] = make_double2(v.x * cw.x - v.y * cw.y, v.x * cw.y + v.y * cw.x);
        } else {                                        // conj(Z[H - k]) (1 - i conj w_k)
          const double2 cw = preB[idx];
          u[r] = make_double2(v.x * cw.x + v.y * cw.y, v.x * cw.y - v.y * cw.x);
        }
      }
      __syncwarp();
      {
        // the staging buffer is free again: start the copy of the next plane of this warp
        const bool wrap = fb + 1 >= nfb;
        const long long ngrp = wrap ? grp + gstep : grp;
        if (ngrp < ngroups) prefetch_rows(A.Q + (wrap ? 0LL : (long long)(fb + 1) * A.q_stride), ngrp * 2);
      }
      dft16<-1>(u);
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      if (!WREG) {
#pragma unroll
        for (int r = 0; r < 16; ++r) wreg[r] = *reinterpret_cast<const double2*>(wrow + 32 * r);
      }
      twiddle16_tab<-1>(u, tab, b);
      dft16<-1>(u);
      // ---- real-space step: x[2m], x[2m+1] at m = b + 16 r, times the weight map (scale / 2 sits in the post tables)
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u[r].x *= wreg[r].x;
        u[r].y *= wreg[r].y;
      }
      // ---- forward: stage 0 from the same registers, one exchange, stage 1
      dft16<1>(u);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      twiddle16_tab<1>(u, tab, b);
      dft16<1>(u);
      // ---- r2c post-processing: T[k] = V[k] postA[k] + conj(V[H-k]) postB[k]; V[H-k] sits in lane 16 - b, register 15 - r
      double2* __restrict__ dst = A.R + (long long)fb * A.r_stride + r_off;
      const bool r_tiled = A.r_tile > 0;
      const int src_lane = (rho << 4) | ((16 - b) & 15);
#pragma unroll
      for (int h4 = 0; h4 < 8; h4 += 4) {   // two batches of four outputs: shuffles first, then independent FMA chains
        double2 v2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v2[q].x = __shfl_sync(0xffffffffu, u[15 - h4 - q].x, src_lane);
          v2[q].y = __shfl_sync(0xffffffffu, u[15 - h4 - q].y, src_lane);
        }
        if (b == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v2[q] = u[(16 - h4 - q) & 15];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = b + 16 * (h4 + q);
          const double2 v1 = u[h4 + q], pa = postA[k], pb = postB[k];
          double2 t;
          t.x = v1.x * pa.x - v1.y * pa.y + (v2[q].x * pb.x + v2[q].y * pb.y);
          t.y = v1.x * pa.y + v1.y * pa.x + (v2[q].x * pb.y - v2[q].y * pb.x);
          if (k < Kz_out) dst[r_tiled ? (k >> 3) * r_zstep + (k & 7) : k] = t;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused z pass for nz = 1024 (H = 512 = 2 x 16 x 16), one row per warp.  EXPERIMENTAL: selected only with SWW_Z32=1
// (written at the end of round 1 without GPU time left to validate it; its index algebra is pinned by the numpy model
// tests/test_z32_model_cpu.py, the nz = 512 kernel above is its validated template).
//   * the half-warps p = lane >> 4 run the even / odd 256-point transforms of the 512-point half-length FFT in the
//     two-exchange layout of pass_z16_fused_kernel (lane b < 16 holds 16 complex values at n = b + 16 r);
//   * inverse: decimation in frequency -- A[n] = Y[n] + Y[n+256] (p = 0), B[n] = (Y[n] - Y[n+256]) e^{+2 pi i n/512}
//     (p = 1), both formed straight from the staged Hermitian row, so x_c[2q + p] comes out of half-warp p;
//   * forward: decimation in time -- E, O from the same registers, joined by ONE xor-16 shuffle exchange:
//     V[j] = E[j] + W^j O[j] (kept by p = 0), V[j + 256] = E[j] - W^j O[j] (kept by p = 1);
//   * r2c post-processing on the p = 0 lanes: V[512 - k] sits in half-warp 1, lane 16 - b, register 15 - r.
// ------------------------------------------------------------------------------------------
#define Z32_H 512
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
pass_z32_fused_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g, int in_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = Z32_H, ROW = Z16_ROW;
  double2* twn = reinterpret_cast<double2*>(smem_raw);   // [H + 8]  exp(-2 pi i j / nz), j <= H
  double2* tab = twn + (H + 8);                          // [16][16] inter-stage twiddles W_256^(b r) at [r * 16 + b]
  double2* bufs = tab + 256;                             // [WARPS][2 * ROW] exchange buffers (one per half-warp)
  double2* inbs = bufs + WARPS * 2 * ROW;                // [WARPS][in_len] staged input row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p = lane >> 4, b = lane & 15;
  double2* buf = bufs + warp * 2 * ROW + p * ROW;
  double2* inb = inbs + warp * in_len;
  for (int i = tid; i <= H; i += WARPS * 32) twn[i] = twn_g[i];
  for (int i = tid; i < 256; i += WARPS * 32) tab[i] = twh_g[(2 * (i >> 4) * (i & 15)) & (H - 1)];   // W_512^(2 b r)
  __syncthreads();
  const long long nrows = (long long)g.nx * g.ny;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int nfb = A.nb > 1 ? A.nb : 1;
  const int Kz_in = A.Kz_in, Kz_out = A.Kz_out;
  const double half = 0.5 * A.scale;
  auto prefetch_row = [&](const double2* __restrict__ Qbase, long long r0) {
    const double2* src = Qbase + r0 * A.Kzp_in;
    for (int k = lane; k < Kz_in; k += 32) {
      unsigned sa = (unsigned)__cvta_generic_to_shared(inb + k);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
    }
  };
  // Y[k] of the half-length sequence from the staged row (zero where neither k nor H - k is kept)
  auto hermY = [&](int k) -> double2 {
    const int hk = H - k;
    if (!(k < Kz_in || hk < Kz_in)) return make_double2(0.0, 0.0);
    double2 zk = (k < Kz_in) ? inb[k] : make_double2(0.0, 0.0);
    double2 zh = (hk < Kz_in) ? inb[hk] : make_double2(0.0, 0.0);
    if (k == 0) {
      zk.y = 0.0;
      zh.y = 0.0;
    }
    const double2 a = make_double2(zk.x + zh.x, zk.y - zh.y);
    const double2 d = make_double2(zk.x - zh.x, zk.y + zh.y);
    double2 w = twn[k];
    w.y = -w.y;
    const double2 bb = c_mul(d, w);
    return make_double2(a.x - bb.y, a.y + bb.x);
  };
  long long row = (long long)blockIdx.x * WARPS + warp;
  if (row < nrows) prefetch_row(A.Q, row);
  asm volatile("cp.async.commit_group;\n" ::);
  for (; row < nrows; row += gstep) {
    const double* __restrict__ wrow = A.wmap + row * g.nz + 4 * b + 2 * p;
    double2 wreg[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const double2 wv = *reinterpret_cast<const double2*>(wrow + 64 * r);
      wreg[r] = make_double2(wv.x * half, wv.y * half);
    }
    for (int fb = 0; fb < nfb; ++fb) {
      double2 u[16];
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncwarp();
      // ---- inverse: pre-processing + radix-2 decimation in frequency, straight into stage 0 of the 256-point transforms
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u[r] = make_double2(0.0, 0.0);
        // warp-uniform: some n, 512 - n, n + 256 or 256 - n of this block of 16 is kept
        if (16 * r < Kz_in || 497 - 16 * r < Kz_in || 16 * r + 256 < Kz_in || 241 - 16 * r < Kz_in) {
          const int n = b + 16 * r;
          const double2 y0 = hermY(n), y1 = hermY(n + 256);
          if (p == 0) {
            u[r] = make_double2(y0.x + y1.x, y0.y + y1.y);
          } else {
            double2 w = twn[2 * n];      // e^{-2 pi i n/512}; conjugate for the inverse
            w.y = -w.y;
            u[r] = c_mul(make_double2(y0.x - y1.x, y0.y - y1.y), w);
          }
        }
      }
      __syncwarp();
      {
        const bool wrap = fb + 1 >= nfb;
        const long long nrow = wrap ? row + gstep : row;
        if (nrow < nrows) prefetch_row(A.Q + (wrap ? 0LL : (long long)(fb + 1) * A.q_stride), nrow);
        asm volatile("cp.async.commit_group;\n" ::);
      }
      dft16<-1>(u);
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      twiddle16_tab<-1>(u, tab, b);
      dft16<-1>(u);
      // ---- real-space step: u[r] = x_c[2 q + p] = (x[4q + 2p], x[4q + 2p + 1]) at q = b + 16 r
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u[r].x *= wreg[r].x;
        u[r].y *= wreg[r].y;
      }
      // ---- forward: E (p = 0) and O (p = 1) from the same registers
      dft16<1>(u);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
      twiddle16_tab<1>(u, tab, b);
      dft16<1>(u);
      // ---- radix-2 join: V[j] = E[j] + W_512^j O[j] (p = 0),  V[j + 256] = E[j] - W_512^j O[j] (p = 1),  j = b + 16 r
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        double2 o;
        o.x = __shfl_xor_sync(0xffffffffu, u[r].x, 16);
        o.y = __shfl_xor_sync(0xffffffffu, u[r].y, 16);
        const double2 e = p == 0 ? u[r] : o;           // E[j]
        const double2 od = p == 0 ? o : u[r];          // O[j]
        const double2 t = c_mul(od, twn[2 * (b + 16 * r)]);
        u[r] = p == 0 ? make_double2(e.x + t.x, e.y + t.y) : make_double2(e.x - t.x, e.y - t.y);
      }
      // ---- r2c post-processing on the p = 0 lanes: T[k] = (V[k] + conj V[H-k]) - i e^{-2 pi i k/nz} (V[k] - conj V[H-k])
      //      (the 1/2 is folded into the weights);  V[512 - k] = V[(256 - k) + 256]: half-warp 1, lane 16 - b, register
      //      15 - r; for b = 0 lane 16 holds it in register 16 - r, so that lane offers that register instead
      double2* __restrict__ dst = A.R + (long long)fb * A.r_stride + row * A.Kzp_out;
      const int src_lane = 16 | ((16 - b) & 15);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if (16 * r < Kz_out) {                          // warp-uniform
          const double2 offer = (lane == 16) ? u[(16 - r) & 15] : u[15 - r];
          double2 v2;
          v2.x = __shfl_sync(0xffffffffu, offer.x, src_lane);
          v2.y = __shfl_sync(0xffffffffu, offer.y, src_lane);
          const int k = b + 16 * r;
          if (k == 0) v2 = u[0];                        // V[512] = V[0]
          if (p == 0 && k < Kz_out) {
            const double2 v1 = u[r];
            const double2 ev = make_double2(v1.x + v2.x, v1.y - v2.y);
            const double2 dv = make_double2(v1.x - v2.x, v1.y + v2.y);
            const double2 od = c_mul(make_double2(dv.y, -dv.x), twn[k]);
            dst[k] = make_double2(ev.x + od.x, ev.y + od.y);
          }
        }
      }
    }
  }
}

#include "fft_any_passes.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Sm100State {
  AnyPlan plan_x, plan_y, plan_z;   // run-time radix plans of the axes that are not powers of two
  bool any_x = false, any_y = false, any_z = false;
  double2 *tw_x = nullptr, *tw_y = nullptr, *tw_h = nullptr, *tw_n = nullptr;
  double2* tw_c = nullptr;      // exp(-2 pi i j / nz), j < nz: complex z pass (built on first use)
  double2 *P = nullptr, *Q = nullptr, *R = nullptr;
  size_t buf_elems = 0;
  std::vector<int> shell_bx, shell_by, shell_bz;  // per-bin inverse boxes
  // shell order of the output box (built on first use)
  int shell_state = 0;          // 0: not built, 1: ready, -1: unsupported
  long long S = 0;              // in-shell modes
  int* d_perm = nullptr;        // [Ko][Ka][Kzp] -> position in shell order, or -1
  int4* d_chunks = nullptr;     // (start, len, bin, 0)
  int* d_bin_chunk = nullptr;   // [n_k + 1] chunk range of every bin
  int n_chunks = 0;
};

static std::vector<double2> make_tw(int n, int count) {
  std::vector<double2> t(count);
  for (int j = 0; j < count; ++j) {
    long double a = -2.0L * M_PIl * (long double)j / (long double)n;
    t[j] = make_double2((double)cosl(a), (double)sinl(a));
  }
  return t;
}

static int up8(int k) { return (k + 7) & ~7; }

// elements of each of the three pass buffers: one full plane, or n_l pruned planes (batched Fisher rows)
static size_t pass_buf_elems(const sww_ctx* c) {
  const Geo& g = c->g;
  size_t full = (size_t)g.nx * g.ny * up8(g.nzh);
  size_t batch = (size_t)c->n_l * g.nx * g.ny * up8(g.Kz);
  return full > batch ? full : batch;
}

size_t sww_fft_sm100_work_bytes(const sww_ctx* c) {
  if (!sww_fft_sm100_supported(c)) return 0;
  return 3 * pass_buf_elems(c) * sizeof(double2) + 1024;
}

static int max_abs_idx(const sww_ctx* c, int axis, double kmax) {
  int n = axis == 0 ? c->g.nx : (axis == 1 ? c->g.ny : c->g.nz);
  double kF = 2.0 * M_PI / c->box[axis];
  int best = 0;
  for (int i = 0; i <= n / 2; ++i)
    if (i * kF < kmax) best = i;
  return best;
}

int sww_fft_sm100_init(sww_ctx* c) {
  if (c->sm100) return 0;
  SWW_REQUIRE(sww_fft_sm100_supported(c), "the sm_100a FFT passes need power-of-two axes (nx, ny in [64,1024], nz in [128,1024]) or axes of at most %d points", ANY_MAX_N);
  Sm100State* s = new Sm100State();
  const Geo& g = c->g;
  const int H = g.nz / 2;
  s->any_x = !col_pow2(g.nx);
  s->any_y = !col_pow2(g.ny);
  s->any_z = !z_pow2(g.nz);
  s->plan_x = any_make_plan(g.nx);
  s->plan_y = any_make_plan(g.ny);
  s->plan_z = any_make_plan(g.nz);
  auto up = [&](const std::vector<double2>& h, double2** d) -> int {
    SWW_CUDA(cudaMalloc(d, h.size() * sizeof(double2)));
    SWW_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(double2), cudaMemcpyHostToDevice));
    return 0;
  };
  SWW_TRY(up(make_tw(g.nx, g.nx), &s->tw_x));
  SWW_TRY(up(make_tw(g.ny, g.ny), &s->tw_y));
  SWW_TRY(up(make_tw(H, H), &s->tw_h));
  SWW_TRY(up(make_tw(g.nz, H + 1), &s->tw_n));
  if (s->any_z) SWW_TRY(up(make_tw(g.nz, g.nz), &s->tw_c));
  for (int a = 0; a < g.n_k; ++a) {
    s->shell_bx.push_back(max_abs_idx(c, 0, c->h_khi[a]));
    s->shell_by.push_back(max_abs_idx(c, 1, c->h_khi[a]));
    s->shell_bz.push_back(max_abs_idx(c, 2, c->h_khi[a]));
  }
  c->sm100 = s;
  return 0;
}

void sww_fft_sm100_destroy(sww_ctx* c) {
  Sm100State* s = (Sm100State*)c->sm100;
  if (!s) return;
  cudaFree(s->tw_x);
  cudaFree(s->tw_y);
  cudaFree(s->tw_h);
  cudaFree(s->tw_n);
  if (s->tw_c) cudaFree(s->tw_c);
  if (s->d_perm) cudaFree(s->d_perm);
  if (s->d_chunks) cudaFree(s->d_chunks);
  if (s->d_bin_chunk) cudaFree(s->d_bin_chunk);
  delete s;
  c->sm100 = nullptr;
}

static int bind_bufs(sww_ctx* c, Sm100State* s) {
  SWW_REQUIRE(c->fft_work, "no workspace bound (sww_bind_workspace)");
  size_t e = pass_buf_elems(c);
  SWW_REQUIRE(c->fft_work_cap >= 3 * e * sizeof(double2),
              "workspace was sized before the sm_100a FFT backend was selected; select it before binding");
  s->P = (double2*)c->fft_work;
  s->Q = s->P + e;
  s->R = s->Q + e;
  s->buf_elems = e;
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) SWW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// SWW_GRID_OVERSUB (default 1): the passes are persistent kernels with one static share of the work per resident CTA, so a
// CTA slot taken by a co-running kernel (the single-CTA MT19937 recurrences of the device RNG, NCCL, another context)
// delays a full share to the end of the launch.  A factor G > 1 launches G times more, G times shorter CTAs, so that a
// missing slot costs 1/G of a launch instead.  Not yet evaluated on a GPU (added after the round's budget was spent).
static long long grid_oversub() {
  static long long v = 0;
  if (v == 0) {
    const char* e = getenv("SWW_GRID_OVERSUB");
    v = e ? atoll(e) : 1;
    if (v < 1 || v > 64) v = 1;
  }
  return v;
}

static inline int grid_for_items(long long items, int per_sm) {
  long long g = (long long)SMS * per_sm * grid_oversub();
  return (int)(items < g ? items : g);
}

#define DISPATCH_N(n, MACRO)                                         \
  switch (n) {                                                       \
    case 64: MACRO(64); break;                                       \
    case 128: MACRO(128); break;                                     \
    case 256: MACRO(256); break;                                     \
    case 512: MACRO(512); break;                                     \
    case 1024: MACRO(1024); break;                                   \
    default: SWW_REQUIRE(false, "unsupported FFT length %d", n);     \
  }

static size_t col_smem(int N, int hist_doubles) {
  return (size_t)N * FFT_LDT * 16 + (size_t)N * 16 + (size_t)hist_doubles * 8;
}

static int occ_for(size_t smem) {
  int o = (int)((227 * 1024) / (smem + 1024));
  if (o < 1) o = 1;
  if (o > 4) o = 4;
  return o;
}
// resident CTAs per SM of a column-pass kernel as the device will actually schedule it (registers included), at most what the
// shared memory allows (occ_for): the persistent grids are sized from this, so a register count that drifts past a threshold
// costs occupancy but never a second, partial wave
static std::mutex g_occ_mutex;
static std::map<std::tuple<const void*, int, size_t>, int> g_occ_cache;
template <typename K>
static int occ_of(K kernel, int threads, size_t smem) {
  const auto key = std::make_tuple((const void*)kernel, threads, smem);
  {
    std::lock_guard<std::mutex> lk(g_occ_mutex);
    auto it = g_occ_cache.find(key);
    if (it != g_occ_cache.end()) return it->second;
  }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
  const int cap = occ_for(smem);
  nb = nb < cap ? nb : cap;
  std::lock_guard<std::mutex> lk(g_occ_mutex);
  g_occ_cache[key] = nb;
  return nb;
}

// ---- column-pass launchers, generic in the transformed axis -------------------------------------
static BoxAxis box_axis_of(int n, int b) {
  int K = 2 * b + 1;
  if (K >= n) return BoxAxis{n, n, n - 1};
  return BoxAxis{n, K, b};
}

struct AxisPlan {
  BoxAxis bx, by;   // kept indices along x and y
  bool x_first_fwd; // forward: transform x before y (x is the more strongly pruned axis)
};

static AxisPlan make_plan(const Geo& g, BoxAxis bx, BoxAxis by) {
  AxisPlan p;
  p.bx = bx;
  p.by = by;
  p.x_first_fwd = (long long)bx.K * g.ny <= (long long)by.K * g.nx;
  return p;
}

static const double2* tw_of(Sm100State* s, int ax) { return ax == 0 ? s->tw_x : s->tw_y; }

template <int N, int AX>
static int launch_i1(sww_ctx* c, Sm100State* s, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* src, int shell,
                     const GatherBatch& gb, long long pstride, double klim2) {
  size_t smem = col_smem(N, 0);
  long long items = (long long)bO.K * (Kzp / FFT_TZ) * (gb.nb > 0 ? gb.nb : 1);
  SWW_TRY(set_smem(pass_i1_kernel<N, AX>, smem));
  int grid = grid_for_items(items, occ_of(pass_i1_kernel<N, AX>, N / 2, smem));
  pass_i1_kernel<N, AX><<<grid, N / 2, smem, c->stream>>>(c->g, bA, bO, Kz, Kzp, src, shell, tw_of(s, AX), s->P, gb,
                                                         box_axis_of(c->g.nx, c->g.bx), box_axis_of(c->g.ny, c->g.by),
                                                         pstride, klim2);
  return 0;
}
template <int N, int AX>
static int launch_i2(sww_ctx* c, Sm100State* s, int nF, BoxAxis bA, int Kzp, double2* Qdst, int nb, long long pstride,
                     long long qstride, int q_tiled, double klim2) {
  size_t smem = col_smem(N, 0);
  long long items = (long long)nF * (Kzp / FFT_TZ) * nb;
  SWW_TRY(set_smem(pass_i2_kernel<N, AX>, smem));
  int grid = grid_for_items(items, occ_of(pass_i2_kernel<N, AX>, N / 2, smem));
  pass_i2_kernel<N, AX><<<grid, N / 2, smem, c->stream>>>(nF, c->g.ny, bA, Kzp, s->P, tw_of(s, AX), Qdst, nb, pstride, qstride, q_tiled,
                                                         c->g.kax[AX], c->g.kax[2], klim2);
  return 0;
}
template <int N, int AX>
static int launch_f2(sww_ctx* c, Sm100State* s, int nC, BoxAxis bA, int Kzp, int nb, long long rstride, long long pstride,
                     int r_tiled, double klim2) {
  size_t smem = col_smem(N, 0);
  long long items = (long long)nC * (Kzp / FFT_TZ) * nb;
  if constexpr (N <= 512) {
    const char* eb = getenv("SWW_F2_BULK");
    if (r_tiled && eb && eb[0] == '1' && klim2 <= 0.0) {   // opt-in (measured slower than the two-CTA kernel: see sww_sm100_pk_rows_store)
      // double-buffered bulk asynchronous tile loads (TMA engine), one CTA per SM
      const size_t bsm = (size_t)2 * N * FFT_TZ * 16 + (size_t)N * 16 + 64;
      const int per_sm = std::max(1, std::min(2, (int)((227 * 1024) / (bsm + 1024))));
      int bgrid = grid_for_items(items, per_sm);
      SWW_TRY(set_smem(pass_f2_bulk_kernel<N, AX>, bsm));
      pass_f2_bulk_kernel<N, AX><<<bgrid, N / 2, bsm, c->stream>>>(nC, bA, Kzp, s->R, tw_of(s, AX), s->P, nb, rstride, pstride);
      return 0;
    }
  }
  SWW_TRY(set_smem(pass_f2_kernel<N, AX>, smem));
  int grid = grid_for_items(items, occ_of(pass_f2_kernel<N, AX>, N / 2, smem));
  pass_f2_kernel<N, AX><<<grid, N / 2, smem, c->stream>>>(nC, c->g.ny, bA, Kzp, s->R, tw_of(s, AX), s->P, nb, rstride, pstride, r_tiled,
                                                         c->g.kax[AX], c->g.kax[2], klim2);
  return 0;
}
template <int N, int AX>
static int launch_f3(sww_ctx* c, Sm100State* s, int mode, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, F3Args A,
                     const double2* Psrc) {
  int hist = (mode == 2) ? (N / 2 / 32) * A.n_g * c->g.n_k : 0;
  size_t smem = col_smem(N, hist);
  long long items = (long long)bO.K * (Kzp / FFT_TZ);
  const int nrows = (mode == 2 && A.nrows > 1) ? A.nrows : 1;
  int occ = occ_for(smem);
  if (mode == 0) { SWW_TRY(set_smem(pass_f3_kernel<N, AX, 0>, smem)); occ = occ_of(pass_f3_kernel<N, AX, 0>, N / 2, smem); }
  else if (mode == 1) { SWW_TRY(set_smem(pass_f3_kernel<N, AX, 1>, smem)); occ = occ_of(pass_f3_kernel<N, AX, 1>, N / 2, smem); }
  else if (mode == 3) { SWW_TRY(set_smem(pass_f3_kernel<N, AX, 3>, smem)); occ = occ_of(pass_f3_kernel<N, AX, 3>, N / 2, smem); }
  else { SWW_TRY(set_smem(pass_f3_kernel<N, AX, 2>, smem)); occ = occ_of(pass_f3_kernel<N, AX, 2>, N / 2, smem); }
  int grid = grid_for_items(items * nrows, occ);
  if (nrows > 1) grid = std::max(1, grid / nrows) * nrows;
  if (mode == 2) {
    SWW_REQUIRE(nrows <= 16, "F3 bin: too many rows in a batch");
    SWW_REQUIRE((size_t)grid * A.n_g * c->g.n_k <= c->partial_elems, "F3 bin: partial buffer too small");
    A.partial = c->d_partial;
    A.counter = c->d_counter;
  }
  if (mode == 0) {
    SWW_TRY(set_smem(pass_f3_kernel<N, AX, 0>, smem));
    pass_f3_kernel<N, AX, 0><<<grid, N / 2, smem, c->stream>>>(c->g, bA, bO, Kz, Kzp, Psrc, tw_of(s, AX), A);
  } else if (mode == 1) {
    SWW_TRY(set_smem(pass_f3_kernel<N, AX, 1>, smem));
    pass_f3_kernel<N, AX, 1><<<grid, N / 2, smem, c->stream>>>(c->g, bA, bO, Kz, Kzp, Psrc, tw_of(s, AX), A);
  } else if (mode == 3) {
    SWW_TRY(set_smem(pass_f3_kernel<N, AX, 3>, smem));
    pass_f3_kernel<N, AX, 3><<<grid, N / 2, smem, c->stream>>>(c->g, bA, bO, Kz, Kzp, Psrc, tw_of(s, AX), A);
  } else {
    SWW_TRY(set_smem(pass_f3_kernel<N, AX, 2>, smem));
    pass_f3_kernel<N, AX, 2><<<grid, N / 2, smem, c->stream>>>(c->g, bA, bO, Kz, Kzp, Psrc, tw_of(s, AX), A);
  }
  return 0;
}

// ---- launchers of the any-length column passes (fft_any_passes.cuh) ----
static const AnyPlan& plan_of(Sm100State* s, int ax) { return ax == 0 ? s->plan_x : s->plan_y; }
static bool any_ax(Sm100State* s, int ax) { return ax == 0 ? s->any_x : s->any_y; }
// resident CTAs per SM of one of the any-length kernels (registers, shared memory, threads), from the occupancy API
template <typename K>
static int any_occ(K kernel, size_t smem, int threads) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
  if (nb > 8) nb = 8;
  return nb;
}
template <int AX>
static int launch_i1_any(sww_ctx* c, Sm100State* s, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, const double2* src, int shell,
                         const GatherBatch& gb, long long pstride, double klim2) {
  const AnyPlan& pl = plan_of(s, AX);
  const size_t smem = ANY_SMEM_TILES(pl.n);
  const int threads = any_threads_for(pl);
  long long items = (long long)bO.K * ((Kzp + FFT_TZ - 1) / FFT_TZ) * (gb.nb > 0 ? gb.nb : 1);
  SWW_REQUIRE(items < (1LL << 31), "any-length pass: too many work items");
  SWW_TRY(set_smem(gpass_i1_kernel<AX>, smem));
  int grid = grid_for_items(items, any_occ(gpass_i1_kernel<AX>, smem, threads));
  gpass_i1_kernel<AX><<<grid, threads, smem, c->stream>>>(c->g, pl, bA, bO, Kz, Kzp, src, shell, tw_of(s, AX), s->P, gb,
                                                          box_axis_of(c->g.nx, c->g.bx), box_axis_of(c->g.ny, c->g.by), pstride,
                                                          klim2);
  return 0;
}
template <int AX>
static int launch_i2_any(sww_ctx* c, Sm100State* s, int nF, BoxAxis bA, int Kzp, double2* Qdst, int nb, long long pstride,
                         long long qstride, int q_tiled, double klim2) {
  SWW_REQUIRE(!q_tiled, "tile-major planes need power-of-two axes");
  const AnyPlan& pl = plan_of(s, AX);
  const size_t smem = ANY_SMEM_TILES(pl.n);
  const int threads = any_threads_for(pl);
  long long items = (long long)nF * ((Kzp + FFT_TZ - 1) / FFT_TZ) * nb;
  SWW_REQUIRE(items < (1LL << 31), "any-length pass: too many work items");
  SWW_TRY(set_smem(gpass_i2_kernel<AX>, smem));
  int grid = grid_for_items(items, any_occ(gpass_i2_kernel<AX>, smem, threads));
  gpass_i2_kernel<AX><<<grid, threads, smem, c->stream>>>(pl, nF, c->g.ny, bA, Kzp, s->P, tw_of(s, AX), Qdst, nb, pstride, qstride,
                                                          c->g.kax[AX], c->g.kax[2], klim2);
  return 0;
}
template <int AX>
static int launch_f2_any(sww_ctx* c, Sm100State* s, int nC, BoxAxis bA, int Kzp, int nb, long long rstride, long long pstride,
                         int r_tiled, double klim2) {
  SWW_REQUIRE(!r_tiled, "tile-major planes need power-of-two axes");
  const AnyPlan& pl = plan_of(s, AX);
  const size_t smem = ANY_SMEM_TILES(pl.n);
  const int threads = any_threads_for(pl);
  long long items = (long long)nC * ((Kzp + FFT_TZ - 1) / FFT_TZ) * nb;
  SWW_REQUIRE(items < (1LL << 31), "any-length pass: too many work items");
  SWW_TRY(set_smem(gpass_f2_kernel<AX>, smem));
  int grid = grid_for_items(items, any_occ(gpass_f2_kernel<AX>, smem, threads));
  gpass_f2_kernel<AX><<<grid, threads, smem, c->stream>>>(pl, nC, c->g.ny, bA, Kzp, s->R, tw_of(s, AX), s->P, nb, rstride, pstride,
                                                          c->g.kax[AX], c->g.kax[2], klim2);
  return 0;
}
template <int AX>
static int launch_f3_any(sww_ctx* c, Sm100State* s, int mode, BoxAxis bA, BoxAxis bO, int Kz, int Kzp, F3Args A,
                         const double2* Psrc) {
  const AnyPlan& pl = plan_of(s, AX);
  int threads = any_threads_for(pl);
  // MODE 2 keeps one histogram per warp behind the tiles: fewer warps when the two do not fit the SM's shared memory
  while (mode == 2 && threads > 64 && ANY_SMEM_TILES(pl.n) + (size_t)(threads / 32) * A.n_g * c->g.n_k * 8 > 220 * 1024) threads -= 32;
  const int hist = (mode == 2) ? (threads / 32) * A.n_g * c->g.n_k : 0;
  const size_t smem = ANY_SMEM_TILES(pl.n) + (size_t)hist * 8;
  long long items = (long long)bO.K * ((Kzp + FFT_TZ - 1) / FFT_TZ);
  const int nrows = (mode == 2 && A.nrows > 1) ? A.nrows : 1;
  SWW_REQUIRE(items < (1LL << 31), "any-length pass: too many work items");
#define LAUNCH_F3_ANY(MM)                                                                                                  \
  {                                                                                                                        \
    SWW_TRY((set_smem(gpass_f3_kernel<AX, MM>, smem)));                                                                    \
    int grid = grid_for_items(items * nrows, any_occ(gpass_f3_kernel<AX, MM>, smem, threads));                             \
    if (nrows > 1) grid = std::max(1, grid / nrows) * nrows;                                                               \
    if (MM == 2) {                                                                                                         \
      SWW_REQUIRE(nrows <= 16, "F3 bin: too many rows in a batch");                                                        \
      SWW_REQUIRE((size_t)grid * A.n_g * c->g.n_k <= c->partial_elems, "F3 bin: partial buffer too small");               \
      A.partial = c->d_partial;                                                                                            \
      A.counter = c->d_counter;                                                                                            \
    }                                                                                                                      \
    gpass_f3_kernel<AX, MM><<<grid, threads, smem, c->stream>>>(c->g, pl, bA, bO, Kz, Kzp, Psrc, tw_of(s, AX), A);        \
  }
  if (mode == 0) LAUNCH_F3_ANY(0)
  else if (mode == 1) LAUNCH_F3_ANY(1)
  else if (mode == 3) LAUNCH_F3_ANY(3)
  else LAUNCH_F3_ANY(2)
#undef LAUNCH_F3_ANY
  return 0;
}
// any-length axis: the run-time plan kernels; power-of-two axis: the templated ones
#define DISPATCH_AX_ANY(s, n, ax, CALL, CALL_ANY)                     \
  if (any_ax(s, ax)) {                                                \
    if (ax == 0) { SWW_TRY((CALL_ANY(0))); } else { SWW_TRY((CALL_ANY(1))); } \
  } else {                                                            \
    DISPATCH_AX(n, ax, CALL)                                          \
  }

#define DISPATCH_AX(n, ax, CALL)                                                   \
  switch (n) {                                                                     \
    case 64: if (ax == 0) { SWW_TRY((CALL(64, 0))); } else { SWW_TRY((CALL(64, 1))); } break;       \
    case 128: if (ax == 0) { SWW_TRY((CALL(128, 0))); } else { SWW_TRY((CALL(128, 1))); } break;    \
    case 256: if (ax == 0) { SWW_TRY((CALL(256, 0))); } else { SWW_TRY((CALL(256, 1))); } break;    \
    case 512: if (ax == 0) { SWW_TRY((CALL(512, 0))); } else { SWW_TRY((CALL(512, 1))); } break;    \
    case 1024: if (ax == 0) { SWW_TRY((CALL(1024, 0))); } else { SWW_TRY((CALL(1024, 1))); } break; \
    default: SWW_REQUIRE(false, "unsupported FFT length %d", n);                   \
  }

static int axis_len(const Geo& g, int ax) { return ax == 0 ? g.nx : g.ny; }

// inverse column passes: spectrum (box, optional shell filter) -> Q[nx][ny][Kzp]
static int run_inverse_cols(sww_ctx* c, Sm100State* s, const AxisPlan& pl, int Kz, int Kzp, const double2* src, int shell,
                            const GatherArgs* gap = nullptr, int nb = 1, double2* Qdst = nullptr, long long qstride = 0,
                            int q_tiled = 0, double klim2 = 0.0) {
  if (!Qdst) Qdst = s->Q;
  const int a1 = pl.x_first_fwd ? 1 : 0;   // first inverse axis = the weakly pruned one
  const int a2 = 1 - a1;
  const BoxAxis b1 = a1 == 0 ? pl.bx : pl.by, b2 = a2 == 0 ? pl.bx : pl.by;
  const int n1 = axis_len(c->g, a1), n2 = axis_len(c->g, a2);
  // a batch shares the P buffer: plane stride = what one pruned plane of this box needs; the batch is
  // processed in chunks of as many planes as fit
  const long long pstride = (long long)b2.K * n1 * Kzp;
  int chunk = (int)std::min<long long>(nb, (long long)(s->buf_elems / (size_t)pstride));
  SWW_REQUIRE(chunk >= 1, "pass buffer too small for one pruned plane");
  for (int b0 = 0; b0 < nb; b0 += chunk) {
    const int nbc = std::min(chunk, nb - b0);
    GatherBatch gb = {};
    if (gap) {
      gb.nb = nbc;
      for (int b = 0; b < nbc; ++b) gb.ga[b] = gap[b0 + b];
    }
    {
      SWW_PROF(c, a1 == 0 ? "fft.I1 x-inverse (gather)" : "fft.I1 y-inverse (gather)");
#define CALL(NN, AA) launch_i1<NN, AA>(c, s, b1, b2, Kz, Kzp, src, shell, gb, pstride, klim2)
#define CALL_ANY(AA) launch_i1_any<AA>(c, s, b1, b2, Kz, Kzp, src, shell, gb, pstride, klim2)
      DISPATCH_AX_ANY(s, n1, a1, CALL, CALL_ANY)
#undef CALL_ANY
#undef CALL
      c->n_own++;
      SWW_CUDA(cudaGetLastError());
    }
    {
      SWW_PROF(c, a2 == 0 ? "fft.I2 x-inverse" : "fft.I2 y-inverse");
#define CALL(NN, AA) launch_i2<NN, AA>(c, s, n1, b2, Kzp, Qdst + (long long)b0 * qstride, nbc, pstride, qstride, q_tiled, klim2)
#define CALL_ANY(AA) launch_i2_any<AA>(c, s, n1, b2, Kzp, Qdst + (long long)b0 * qstride, nbc, pstride, qstride, q_tiled, klim2)
      DISPATCH_AX_ANY(s, n2, a2, CALL, CALL_ANY)
#undef CALL_ANY
#undef CALL
      c->n_own++;
      SWW_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

// forward column passes: R[nx][ny][Kzp] -> epilogue `mode`
static int run_forward_cols(sww_ctx* c, Sm100State* s, const AxisPlan& pl, int Kz, int Kzp, int mode, F3Args A,
                            int nb = 1, long long rstride = 0, const F3Args* Ab = nullptr, int r_tiled = 0, double klim2 = 0.0) {
  const int a1 = pl.x_first_fwd ? 0 : 1;   // first forward axis = the strongly pruned one
  const int a2 = 1 - a1;
  const BoxAxis b1 = a1 == 0 ? pl.bx : pl.by, b2 = a2 == 0 ? pl.bx : pl.by;
  const int n1 = axis_len(c->g, a1), n2 = axis_len(c->g, a2);
  const long long pstride = (long long)b1.K * n2 * Kzp;
  SWW_REQUIRE((size_t)pstride * nb <= s->buf_elems, "forward batch of %d planes does not fit the pass buffer", nb);
  {
    SWW_PROF(c, a1 == 0 ? "fft.F2 x-forward" : "fft.F2 y-forward");
#define CALL(NN, AA) launch_f2<NN, AA>(c, s, n2, b1, Kzp, nb, rstride, pstride, r_tiled, klim2)
#define CALL_ANY(AA) launch_f2_any<AA>(c, s, n2, b1, Kzp, nb, rstride, pstride, r_tiled, klim2)
    DISPATCH_AX_ANY(s, n1, a1, CALL, CALL_ANY)
#undef CALL_ANY
#undef CALL
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
  }
  if (mode == 2 && nb > 1 && Ab && nb <= SWW_MAX_L) {
    // one launch bins all rows of the batch against the shared G maps
    SWW_PROF(c, "fft.F3 last-forward+bin");
    F3Args Ax = Ab[0];
    Ax.nrows = nb;
    Ax.row_pstride = pstride;
    for (int b = 0; b < nb; ++b) Ax.out_rows[b] = Ab[b].out_row;
    const double2* Psrc = s->P;
#define CALL(NN, AA) launch_f3<NN, AA>(c, s, mode, b2, b1, Kz, Kzp, Ax, Psrc)
#define CALL_ANY(AA) launch_f3_any<AA>(c, s, mode, b2, b1, Kz, Kzp, Ax, Psrc)
    DISPATCH_AX_ANY(s, n2, a2, CALL, CALL_ANY)
#undef CALL_ANY
#undef CALL
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  for (int b = 0; b < nb; ++b) {
    SWW_PROF(c, mode == 2 ? "fft.F3 last-forward+bin"
                          : (mode == 1 ? "fft.F3 last-forward+ylm" : (mode == 3 ? "fft.F3 last-fwd+shell-store" : "fft.F3 last-forward+store")));
    const F3Args& Ax = Ab ? Ab[b] : A;
    const double2* Psrc = s->P + (long long)b * pstride;
#define CALL(NN, AA) launch_f3<NN, AA>(c, s, mode, b2, b1, Kz, Kzp, Ax, Psrc)
#define CALL_ANY(AA) launch_f3_any<AA>(c, s, mode, b2, b1, Kz, Kzp, Ax, Psrc)
    DISPATCH_AX_ANY(s, n2, a2, CALL, CALL_ANY)
#undef CALL_ANY
#undef CALL
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
  }
  return 0;
}

template <int H, bool INV, bool FWD, bool YLM>
static int launch_zw(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 8;
  size_t smem = (size_t)(H + 8 + WTabSize<H, false>::value + WTabSize<H, true>::value) * 16 +
                (size_t)WARPS * (WGeo<H>::RW * WGeo<H>::ROWLEN + 8) * 16 * (INV ? 2 : 1);
  long long ngroups = (long long)c->g.nx * c->g.ny / WGeo<H>::RW;
  long long ctas = (ngroups + WARPS - 1) / WARPS;
  const long long slots = (long long)SMS * 2 * grid_oversub();
  int grid = (int)(ctas < slots ? ctas : slots);
  SWW_TRY(set_smem(pass_zw_kernel<H, INV, FWD, YLM>, smem));
  pass_zw_kernel<H, INV, FWD, YLM><<<grid, 256, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n);
  return 0;
}

// two-exchange fused z pass (nz = 512, outputs below nz/4): see pass_z16_fused_kernel
template <int WARPS, int MINB, bool WREG>
static int launch_z16_v(sww_ctx* c, Sm100State* s, ZArgs A) {
  const int in_len = up8(A.Kz_in);
  const size_t smem = (size_t)(Z16_H + 8 + 256 + WARPS * 2 * Z16_ROW + WARPS * 2 * in_len) * sizeof(double2);
  const long long ngroups = (long long)c->g.nx * c->g.ny / 2;
  const long long ctas = (ngroups + WARPS - 1) / WARPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > MINB) per_sm = MINB;
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  SWW_TRY((set_smem(pass_z16_fused_kernel<WARPS, MINB, WREG>, smem)));
  pass_z16_fused_kernel<WARPS, MINB, WREG><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n, in_len);
  return 0;
}

static int launch_z32(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4, MINB = 3;
  const int in_len = up8(A.Kz_in);
  const size_t smem = (size_t)(Z32_H + 8 + 256 + WARPS * 2 * Z16_ROW + WARPS * in_len) * sizeof(double2);
  const long long nrows = (long long)c->g.nx * c->g.ny;
  const long long ctas = (nrows + WARPS - 1) / WARPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > MINB) per_sm = MINB;
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm);
  SWW_TRY((set_smem(pass_z32_fused_kernel<WARPS, MINB>, smem)));
  pass_z32_fused_kernel<WARPS, MINB><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n, in_len);
  return 0;
}

// ------------------------------------------------------------------------------------------
// pass_z16f_kernel: the FORWARD half of pass_z16p_fused_kernel on its own (nz = 512, outputs below nz/4): the 2M + 1 transforms
// FT[f Y_lm(r^)] at the head of every realisation.  A half-warp owns a row: lane b loads x[2m], x[2m+1] at m = b + 16 r straight
// into the registers of forward stage 0, multiplies by the weight map and by Y_lm(r^) evaluated along the row (Horner in z
// with the row's coefficients, 1 / r^2 by sww_rcp_pos; the z table sits in shared memory), and runs 16 x 16 with ONE exchange and
// the table-driven r2c post-processing.  The three-stage warp-per-row kernel it replaces spent 67 % of the shared-memory
// bandwidth on its two exchanges (ncu, profiles/r3f).
// ------------------------------------------------------------------------------------------
template <bool YLM>
__global__ void __launch_bounds__(128, 4)
pass_z16f_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = Z16_H, ROW = Z16_ROW, WARPS = 4;
  double2* postA = reinterpret_cast<double2*>(smem_raw);   // [128]
  double2* postB = postA + 128;                           // [128]
  double2* tab = postB + 128;                             // [16][16] inter-stage twiddles W_256^(b r) at [r * 16 + b]
  double2* zt = tab + 256;                                // [256] (z[2m], z[2m+1]) of the mesh axis
  double2* bufs = zt + 256;                               // [WARPS][2 * ROW] exchange buffers
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = lane >> 4, b = lane & 15;
  double2* buf = bufs + warp * 2 * ROW + rho * ROW;
  const double half = 0.5 * A.scale;
  for (int i = tid; i < 128; i += WARPS * 32) {
    const double2 w = twn_g[i];                          // (c, -s)
    postA[i] = make_double2(half * (1.0 + w.y), -half * w.x);
    postB[i] = make_double2(half * (1.0 - w.y), half * w.x);
  }
  for (int i = tid; i < 256; i += WARPS * 32) {
    tab[i] = twh_g[((i >> 4) * (i & 15)) & (H - 1)];
    zt[i] = make_double2(g.rax[2][2 * i], g.rax[2][2 * i + 1]);
  }
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / 2;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int Kz_out = A.Kz_out;
  const bool ymap = YLM && A.lm >= 1 && A.ymaps != nullptr;
  for (long long grp = (long long)blockIdx.x * WARPS + warp; grp < ngroups; grp += gstep) {
    const long long row = grp * 2 + rho;
    double2 u[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = *reinterpret_cast<const double2*>(A.real_in + row * g.nz + 2 * b + 32 * r);
    if (A.wmap) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 w = *reinterpret_cast<const double2*>(A.wmap + row * g.nz + 2 * b + 32 * r);
        u[r].x *= w.x;
        u[r].y *= w.y;
      }
    }
    if (ymap) {
      const double* ym = A.ymaps + (long long)(A.lm - 1) * A.N + row * g.nz + 2 * b;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 y = *reinterpret_cast<const double2*>(ym + 32 * r);
        u[r].x *= y.x;
        u[r].y *= y.y;
      }
    } else if (YLM) {
      const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
      const double X = g.rax[0][ix], Y = g.rax[1][iy], xy2 = X * X + Y * Y;
      double Ac[3];
      bool odd = false;
      sww_ylm_zpack(A.lm, X, Y, Ac, &odd);
      const int deg = SWW_YLM_ZDEG(A.lm);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 z = zt[b + 16 * r];
        const double w0 = z.x * z.x, w1 = z.y * z.y;
        u[r].x *= sww_ylm_zpacked(deg, Ac, odd, z.x, w0, deg ? sww_rcp_pos(xy2 + w0) : 1.0);
        u[r].y *= sww_ylm_zpacked(deg, Ac, odd, z.y, w1, deg ? sww_rcp_pos(xy2 + w1) : 1.0);
      }
    }
    // ---- forward: stage 0 from the registers, one exchange, stage 1
    dft16<1>(u);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
    twiddle16_tab<1>(u, tab, b);
    dft16<1>(u);
    // ---- r2c post-processing: T[k] = V[k] postA[k] + conj(V[H-k]) postB[k]; V[H-k] sits in lane 16 - b, register 15 - r
    double2* __restrict__ dst = A.R + row * A.Kzp_out;
    const int src_lane = (rho << 4) | ((16 - b) & 15);
#pragma unroll
    for (int h4 = 0; h4 < 8; h4 += 4) {
      double2 v2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v2[q].x = __shfl_sync(0xffffffffu, u[15 - h4 - q].x, src_lane);
        v2[q].y = __shfl_sync(0xffffffffu, u[15 - h4 - q].y, src_lane);
      }
      if (b == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v2[q] = u[(16 - h4 - q) & 15];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = b + 16 * (h4 + q);
        const double2 v1 = u[h4 + q], pa = postA[k], pb = postB[k];
        double2 t;
        t.x = v1.x * pa.x - v1.y * pa.y + (v2[q].x * pb.x + v2[q].y * pb.y);
        t.y = v1.x * pa.y + v1.y * pa.x + (v2[q].x * pb.y - v2[q].y * pb.x);
        if (k < Kz_out) dst[k] = t;
      }
    }
  }
}

template <bool YLM>
static int launch_z16f(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)(128 + 128 + 256 + 256 + WARPS * 2 * Z16_ROW) * sizeof(double2);
  const long long ngroups = (long long)c->g.nx * c->g.ny / 2;
  const long long ctas = (ngroups + WARPS - 1) / WARPS;
  SWW_TRY((set_smem(pass_z16f_kernel<YLM>, smem)));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pass_z16f_kernel<YLM>, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  pass_z16f_kernel<YLM><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n);
  return 0;
}

// ------------------------------------------------------------------------------------------
// pass_z32f_kernel: the same for nz = 1024 -- the forward half of pass_z32_fused_kernel on its own.  A warp owns a row: half-warp
// p transforms the samples x_c[2q + p] = (x[4q + 2p], x[4q + 2p + 1]) (q = b + 16 r) with the 16 x 16 kernel, the two 256-point
// results are joined by one radix-2 step through an xor-16 shuffle, and the p = 0 lanes post-process and store k < K_z,out.
// ------------------------------------------------------------------------------------------
template <bool YLM>
__global__ void __launch_bounds__(128, 4)
pass_z32f_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = Z32_H, ROW = Z16_ROW, WARPS = 4;
  double2* twn = reinterpret_cast<double2*>(smem_raw);   // [H + 8]  exp(-2 pi i j / nz), j <= H
  double2* tab = twn + (H + 8);                          // [16][16] inter-stage twiddles W_512^(2 b r) at [r * 16 + b]
  double2* zt = tab + 256;                               // [512] (z[2i], z[2i+1]) of the mesh axis
  double2* bufs = zt + 512;                              // [WARPS][2 * ROW] exchange buffers (one per half-warp)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p = lane >> 4, b = lane & 15;
  double2* buf = bufs + warp * 2 * ROW + p * ROW;
  for (int i = tid; i <= H; i += WARPS * 32) twn[i] = twn_g[i];
  for (int i = tid; i < 256; i += WARPS * 32) tab[i] = twh_g[(2 * (i >> 4) * (i & 15)) & (H - 1)];
  for (int i = tid; i < 512; i += WARPS * 32) zt[i] = make_double2(g.rax[2][2 * i], g.rax[2][2 * i + 1]);
  __syncthreads();
  const long long nrows = (long long)g.nx * g.ny;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int Kz_out = A.Kz_out;
  const double half = 0.5 * A.scale;
  const bool ymap = YLM && A.lm >= 1 && A.ymaps != nullptr;
  for (long long row = (long long)blockIdx.x * WARPS + warp; row < nrows; row += gstep) {
    const long long off = row * g.nz + 4 * b + 2 * p;
    double2 u[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = *reinterpret_cast<const double2*>(A.real_in + off + 64 * r);
    if (A.wmap) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 w = *reinterpret_cast<const double2*>(A.wmap + off + 64 * r);
        u[r].x *= w.x;
        u[r].y *= w.y;
      }
    }
    if (ymap) {
      const double* ym = A.ymaps + (long long)(A.lm - 1) * A.N + off;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 y = *reinterpret_cast<const double2*>(ym + 64 * r);
        u[r].x *= y.x * half;
        u[r].y *= y.y * half;
      }
    } else if (YLM) {
      const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
      const double X = g.rax[0][ix], Y = g.rax[1][iy], xy2 = X * X + Y * Y;
      double Ac[3];
      bool odd = false;
      sww_ylm_zpack(A.lm, X, Y, Ac, &odd);
      const int deg = SWW_YLM_ZDEG(A.lm);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 z = zt[2 * (b + 16 * r) + p];
        const double w0 = z.x * z.x, w1 = z.y * z.y;
        u[r].x *= half * sww_ylm_zpacked(deg, Ac, odd, z.x, w0, deg ? sww_rcp_pos(xy2 + w0) : 1.0);
        u[r].y *= half * sww_ylm_zpacked(deg, Ac, odd, z.y, w1, deg ? sww_rcp_pos(xy2 + w1) : 1.0);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u[r].x *= half;
        u[r].y *= half;
      }
    }
    // ---- forward: E (p = 0) and O (p = 1), 16 x 16 with one exchange
    dft16<1>(u);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[17 * b + r] = u[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = buf[b + 17 * r];
    twiddle16_tab<1>(u, tab, b);
    dft16<1>(u);
    // ---- radix-2 join: V[j] = E[j] + W_512^j O[j] (p = 0),  V[j + 256] = E[j] - W_512^j O[j] (p = 1),  j = b + 16 r
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double2 o;
      o.x = __shfl_xor_sync(0xffffffffu, u[r].x, 16);
      o.y = __shfl_xor_sync(0xffffffffu, u[r].y, 16);
      const double2 e = p == 0 ? u[r] : o;
      const double2 od = p == 0 ? o : u[r];
      const double2 t = c_mul(od, twn[2 * (b + 16 * r)]);
      u[r] = p == 0 ? make_double2(e.x + t.x, e.y + t.y) : make_double2(e.x - t.x, e.y - t.y);
    }
    // ---- r2c post-processing on the p = 0 lanes (see pass_z32_fused_kernel)
    double2* __restrict__ dst = A.R + row * A.Kzp_out;
    const int src_lane = 16 | ((16 - b) & 15);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (16 * r < Kz_out) {                          // warp-uniform
        const double2 offer = (lane == 16) ? u[(16 - r) & 15] : u[15 - r];
        double2 v2;
        v2.x = __shfl_sync(0xffffffffu, offer.x, src_lane);
        v2.y = __shfl_sync(0xffffffffu, offer.y, src_lane);
        const int k = b + 16 * r;
        if (k == 0) v2 = u[0];                        // V[512] = V[0]
        if (p == 0 && k < Kz_out) {
          const double2 v1 = u[r];
          const double2 ev = make_double2(v1.x + v2.x, v1.y - v2.y);
          const double2 dv = make_double2(v1.x - v2.x, v1.y + v2.y);
          const double2 od = c_mul(make_double2(dv.y, -dv.x), twn[k]);
          dst[k] = make_double2(ev.x + od.x, ev.y + od.y);
        }
      }
    }
  }
}

template <bool YLM>
static int launch_z32f(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)(Z32_H + 8 + 256 + 512 + WARPS * 2 * Z16_ROW) * sizeof(double2);
  const long long nrows = (long long)c->g.nx * c->g.ny;
  const long long ctas = (nrows + WARPS - 1) / WARPS;
  SWW_TRY((set_smem(pass_z32f_kernel<YLM>, smem)));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pass_z32f_kernel<YLM>, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  pass_z32f_kernel<YLM><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n);
  return 0;
}

// ------------------------------------------------------------------------------------------
// pass_z8f_kernel: forward-only z pass for nz = 256 (the bispectrum meshes), 128 = 16 x 8 with ONE exchange.  A quarter-warp owns a
// row: lane b (0..7) loads x_c[m] = (x[2m], x[2m+1]) at m = b + 8 r into 16 registers, runs the 16-point stage over r, multiplies
// by W_128^(b q), exchanges so that lane c holds q = c and q = c + 8 for all eight b, and finishes with two 8-point transforms:
// V[q + 16 s].  The few outputs below nz/8 are post-processed from a natural-order copy in shared memory.
// ------------------------------------------------------------------------------------------
#define Z8_ROW 152   // 16 * 9 exchange slots (pitch 9: conflict-free transposed reads), reused for V[0..128]
template <bool YLM>
__global__ void __launch_bounds__(128, 4)
pass_z8f_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = 128, WARPS = 4;
  double2* postA = reinterpret_cast<double2*>(smem_raw);   // [64]
  double2* postB = postA + 64;                            // [64]
  double2* tab = postB + 64;                              // [16][8]  W_128^(b q) at [q * 8 + b]
  double2* zt = tab + 128;                                // [128] (z[2m], z[2m+1])
  double2* bufs = zt + 128;                               // [WARPS][4 rows][Z8_ROW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = lane >> 3, b = lane & 7;
  double2* buf = bufs + (warp * 4 + rho) * Z8_ROW;
  const double half = 0.5 * A.scale;
  for (int i = tid; i < 64; i += WARPS * 32) {
    const double2 w = twn_g[i];                          // (c, -s) of 2 pi i / nz
    postA[i] = make_double2(half * (1.0 + w.y), -half * w.x);
    postB[i] = make_double2(half * (1.0 - w.y), half * w.x);
  }
  for (int i = tid; i < 128; i += WARPS * 32) {
    tab[i] = twh_g[((i >> 3) * (i & 7)) & (H - 1)];
    zt[i] = make_double2(g.rax[2][2 * i], g.rax[2][2 * i + 1]);
  }
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / 4;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int Kz_out = A.Kz_out;
  const bool ymap = YLM && A.lm >= 1 && A.ymaps != nullptr;
  for (long long grp = (long long)blockIdx.x * WARPS + warp; grp < ngroups; grp += gstep) {
    const long long row = grp * 4 + rho;
    const long long off = row * g.nz + 2 * b;
    double2 u[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = *reinterpret_cast<const double2*>(A.real_in + off + 16 * r);
    if (A.wmap) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 w = *reinterpret_cast<const double2*>(A.wmap + off + 16 * r);
        u[r].x *= w.x;
        u[r].y *= w.y;
      }
    }
    if (ymap) {
      const double* ym = A.ymaps + (long long)(A.lm - 1) * A.N + off;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 y = *reinterpret_cast<const double2*>(ym + 16 * r);
        u[r].x *= y.x;
        u[r].y *= y.y;
      }
    } else if (YLM) {
      const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
      const double X = g.rax[0][ix], Y = g.rax[1][iy], xy2 = X * X + Y * Y;
      double Ac[3];
      bool odd = false;
      sww_ylm_zpack(A.lm, X, Y, Ac, &odd);
      const int deg = SWW_YLM_ZDEG(A.lm);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double2 z = zt[b + 8 * r];
        const double w0 = z.x * z.x, w1 = z.y * z.y;
        u[r].x *= sww_ylm_zpacked(deg, Ac, odd, z.x, w0, deg ? sww_rcp_pos(xy2 + w0) : 1.0);
        u[r].y *= sww_ylm_zpacked(deg, Ac, odd, z.y, w1, deg ? sww_rcp_pos(xy2 + w1) : 1.0);
      }
    }
    // ---- stage 0: 16-point transforms over r; twiddle W_128^(b q); exchange
    dft16<1>(u);
#pragma unroll
    for (int q = 1; q < 16; ++q) u[q] = c_mul(u[q], tab[q * 8 + b]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; ++q) buf[q * 9 + b] = u[q];
    __syncwarp();
    double2 v0[8], v1[8];
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      v0[bb] = buf[b * 9 + bb];            // q = b
      v1[bb] = buf[(b + 8) * 9 + bb];      // q = b + 8
    }
    // ---- stage 1: 8-point transforms over the old lane index -> V[q + 16 s]
    dft8<1>(v0);
    dft8<1>(v1);
    __syncwarp();
#pragma unroll
    for (int sft = 0; sft < 8; ++sft) {
      buf[b + 16 * sft] = v0[sft];
      buf[b + 8 + 16 * sft] = v1[sft];
    }
    __syncwarp();
    // ---- r2c post-processing: T[k] = V[k] postA[k] + conj(V[H-k]) postB[k], V[H] = V[0]
    double2* __restrict__ dst = A.R + row * A.Kzp_out;
    for (int k = b; k < Kz_out; k += 8) {
      const double2 va = buf[k], vb = buf[k == 0 ? 0 : H - k], pa = postA[k], pb = postB[k];
      double2 t;
      t.x = va.x * pa.x - va.y * pa.y + (vb.x * pb.x + vb.y * pb.y);
      t.y = va.x * pa.y + va.y * pa.x + (vb.x * pb.y - vb.y * pb.x);
      dst[k] = t;
    }
    __syncwarp();
  }
}

template <bool YLM>
static int launch_z8f(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)(64 + 64 + 128 + 128 + WARPS * 4 * Z8_ROW) * sizeof(double2);
  const long long ngroups = (long long)c->g.nx * c->g.ny / 4;
  const long long ctas = (ngroups + WARPS - 1) / WARPS;
  SWW_TRY((set_smem(pass_z8f_kernel<YLM>, smem)));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pass_z8f_kernel<YLM>, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  pass_z8f_kernel<YLM><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n);
  return 0;
}

// ------------------------------------------------------------------------------------------
// pass_z8i_kernel: inverse-only z pass for nz = 256 with K_z,in <= 64 -- the batched inverse of the bispectrum phi maps
// (sum over planes of Y_lm(r^) IFT[...], times weights, one write).  Geometry of pass_z8f_kernel run backwards: lane c of a
// quarter-warp gathers Y[q + 16 s] for q = c, c + 8 (table-driven Hermitian pre-processing straight from the complex rows: at
// most one of Z[k], Z[H-k] is kept), two 8-point transforms over s, twiddle, ONE exchange, a 16-point transform over q:
// x_c[b + 8 r] in 16 registers.  The harmonic planes (one degree l) are summed as row polynomials in z -- r^l Y_lm = P_lm(x, y; z),
// three FMAs per cell and plane -- and multiplied by r^-l once per cell, then the plain planes are added: no cached harmonic maps
// (3/4 of the old pass's traffic), no division per plane.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3)
pass_z8i_kernel(Geo g, ZArgs A, const double2* __restrict__ twh_g, const double2* __restrict__ twn_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = 128, WARPS = 4;
  double2* preA = reinterpret_cast<double2*>(smem_raw);   // [64]
  double2* preB = preA + 64;                             // [72]
  double2* tab = preB + 72;                              // [16][8]  W_128^(b q) at [q * 8 + b]
  double2* zt = tab + 128;                               // [128] (z[2m], z[2m+1])
  double2* bufs = zt + 128;                              // [WARPS][4 rows][Z8_ROW]
  double2* stgs = bufs + WARPS * 4 * Z8_ROW;             // [WARPS][4 rows][64] staged complex input rows (cp.async)
  __shared__ int ord[SWW_MAX_ZBATCH];                    // planes in processing order: harmonic planes first
  __shared__ int n_harm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = lane >> 3, b = lane & 7;
  double2* buf = bufs + (warp * 4 + rho) * Z8_ROW;
  double2* stg = stgs + (warp * 4 + rho) * 64;
  for (int i = tid; i <= 64; i += WARPS * 32) {
    const double2 w = twn_g[i];                          // (c, -s)
    preB[i] = make_double2(1.0 - w.y, w.x);
    if (i < 64) preA[i] = make_double2(1.0 + w.y, w.x);
  }
  if (tid == 0) {
    const int np_ = A.nb > 1 ? A.nb : 1;
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int bq = 0; bq < np_; ++bq) {
        const int lmb = A.nb > 1 ? A.lms[bq] : A.lm;
        if ((lmb >= 0) == (pass == 0)) ord[n++] = bq;
        if (pass == 0 && bq == np_ - 1) n_harm = n;
      }
  }
  for (int i = tid; i < 128; i += WARPS * 32) {
    tab[i] = twh_g[((i >> 3) * (i & 7)) & (H - 1)];
    zt[i] = make_double2(g.rax[2][2 * i], g.rax[2][2 * i + 1]);
  }
  __syncthreads();
  const long long ngroups = (long long)g.nx * g.ny / 4;
  const long long gstep = (long long)gridDim.x * WARPS;
  const int Kz_in = A.Kz_in;
  const int npl = A.nb > 1 ? A.nb : 1;
  // the complex row of plane `pl` for this quarter-warp's row -> staging buffer (asynchronous; 8 lanes x 16 bytes per step)
  auto prefetch = [&](int pl, long long r_) {
    const double2* src = A.Q + (long long)pl * A.q_stride + r_ * A.Kzp_in;
    for (int k = b; k < Kz_in; k += 8) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(stg + k);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + k));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  long long grp = (long long)blockIdx.x * WARPS + warp;
  if (grp < ngroups) prefetch(ord[0], grp * 4 + rho);
  for (; grp < ngroups; grp += gstep) {
    const long long row = grp * 4 + rho;
    const int ix = (int)(row / g.ny), iy = (int)(row - (long long)ix * g.ny);
    const double X = g.rax[0][ix], Y = g.rax[1][iy], xy2 = X * X + Y * Y;
    double2 acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = make_double2(0.0, 0.0);
    int deg_s = 0;
    for (int i = 0; i < npl; ++i) {
      const int bq = ord[i];
      const bool harm = i < n_harm;
      const int lmb = A.nb > 1 ? A.lms[bq] : A.lm;
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncwarp();
      // ---- gather Y[q + 16 s], q = b + 8 j, from the staged row: one table product per kept element
      double2 y[2][8];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int sft = 0; sft < 8; ++sft) {
          const int k = b + 8 * j + 16 * sft;
          double2 o = make_double2(0.0, 0.0);
          if (k < Kz_in) {
            double2 v = stg[k];
            if (k == 0) v.y = 0.0;
            const double2 cw = preA[k];
            o = make_double2(v.x * cw.x - v.y * cw.y, v.x * cw.y + v.y * cw.x);
          } else if (H - k < Kz_in) {
            const double2 v = stg[H - k], cw = preB[H - k];
            o = make_double2(v.x * cw.x + v.y * cw.y, v.x * cw.y - v.y * cw.x);
          }
          y[j][sft] = o;
        }
      __syncwarp();
      {
        // the staging buffer is free again: start the copy of the next plane (or of the next row group's first plane)
        const bool wrap = i + 1 >= npl;
        const long long ngrp = wrap ? grp + gstep : grp;
        if (ngrp < ngroups) prefetch(ord[wrap ? 0 : i + 1], ngrp * 4 + rho);
      }
      dft8<-1>(y[0]);
      dft8<-1>(y[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
          double2 w = tab[(b + 8 * j) * 8 + bb];
          w.y = -w.y;
          buf[(b + 8 * j) * 9 + bb] = c_mul(y[j][bb], w);
        }
      __syncwarp();
      double2 u[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) u[q] = buf[q * 9 + b];
      __syncwarp();
      dft16<-1>(u);
      // ---- u[r] = (x[2m], x[2m+1]) at m = b + 8 r
      if (harm) {
        double Ac[3];
        bool odd = false;
        sww_ylm_zpack(lmb, X, Y, Ac, &odd);
        deg_s = SWW_YLM_ZDEG(lmb);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double2 z = zt[b + 8 * r];
          const double w0 = z.x * z.x, w1 = z.y * z.y;
          double p0 = (Ac[2] * w0 + Ac[1]) * w0 + Ac[0], p1 = (Ac[2] * w1 + Ac[1]) * w1 + Ac[0];
          if (odd) {
            p0 *= z.x;
            p1 *= z.y;
          }
          acc[r].x += p0 * u[r].x;
          acc[r].y += p1 * u[r].y;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          acc[r].x += u[r].x;
          acc[r].y += u[r].y;
        }
      }
      if (i == n_harm - 1 && deg_s > 0) {
        // r^-l once per cell, after the last harmonic plane
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double2 z = zt[b + 8 * r];
          double q0 = sww_rcp_pos(xy2 + z.x * z.x), q1 = sww_rcp_pos(xy2 + z.y * z.y);
          if (deg_s == 4) {
            q0 *= q0;
            q1 *= q1;
          }
          acc[r].x *= q0;
          acc[r].y *= q1;
        }
      }
    }
    // ---- weights, scale, one write
    double* __restrict__ out = A.real_out + row * g.nz + 2 * b;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double f0 = A.scale, f1 = A.scale;
      if (A.wmap) {
        const double2 w = *reinterpret_cast<const double2*>(A.wmap + row * g.nz + 2 * b + 16 * r);
        f0 *= w.x;
        f1 *= w.y;
      }
      double2 o = make_double2(acc[r].x * f0, acc[r].y * f1);
      if (A.accumulate) {
        const double2 old = *reinterpret_cast<const double2*>(out + 16 * r);
        o.x += old.x;
        o.y += old.y;
      }
      *reinterpret_cast<double2*>(out + 16 * r) = o;
    }
  }
}

static int launch_z8i(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)(64 + 72 + 128 + 128 + WARPS * 4 * Z8_ROW + WARPS * 4 * 64) * sizeof(double2);
  const long long ngroups = (long long)c->g.nx * c->g.ny / 4;
  const long long ctas = (ngroups + WARPS - 1) / WARPS;
  SWW_TRY((set_smem(pass_z8i_kernel, smem)));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pass_z8i_kernel, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  pass_z8i_kernel<<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n);
  return 0;
}

// the harmonic planes of a batched inverse share one polynomial degree (the planes of one phi map do): pass_z8i_kernel's sum
static bool one_harmonic_degree(const ZArgs& A) {
  if (A.nb <= 1) return true;
  int deg = -1;
  for (int b = 0; b < A.nb; ++b)
    if (A.lms[b] >= 0) {
      const int d = SWW_YLM_ZDEG(A.lms[b]);
      if (deg >= 0 && d != deg) return false;
      deg = d;
    }
  return true;
}

template <int NBI, bool WREG, int MINB>
static int launch_z16p_v(sww_ctx* c, Sm100State* s, ZArgs A) {
  constexpr int WARPS = 4;
  const int in_len = up8(A.Kz_in);
  const size_t smem = (size_t)(128 + 136 + 128 + 128 + 256 + WARPS * 2 * Z16_ROW + WARPS * 2 * in_len) * sizeof(double2);
  const long long ngroups = (long long)c->g.nx * c->g.ny / 2;
  const long long ctas = (ngroups + WARPS - 1) / WARPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > MINB) per_sm = MINB;
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(ctas, (long long)SMS * per_sm * grid_oversub());
  SWW_TRY((set_smem(pass_z16p_fused_kernel<NBI, WREG, MINB>, smem)));
  pass_z16p_fused_kernel<NBI, WREG, MINB><<<grid, WARPS * 32, smem, c->stream>>>(c->g, A, s->tw_h, s->tw_n, in_len);
  return 0;
}

template <bool WREG, int MINB>
static int launch_z16p(sww_ctx* c, Sm100State* s, ZArgs A) {
  if (A.Kz_in <= 32) return launch_z16p_v<2, WREG, MINB>(c, s, A);
  if (A.Kz_in <= 64) return launch_z16p_v<4, WREG, MINB>(c, s, A);
  if (A.Kz_in <= 96) return launch_z16p_v<6, WREG, MINB>(c, s, A);
  return launch_z16p_v<8, WREG, MINB>(c, s, A);
}

static int launch_z16(sww_ctx* c, Sm100State* s, ZArgs A) {
  const char* e = getenv("SWW_Z16_VARIANT");
  const int v = e ? atoi(e) : 0;
  if (A.Kz_in <= 128) {
    // table-driven pre / post-processing; the staged row keeps at most one of Z[k], Z[H-k]
    // measured per C2 realisation (82 launches): 8 warps / SM at 255 registers without spills 88.0 ms, 12 warps / SM at 168
    // registers (24-56 B of spills) 90.6 ms, 16 warps / SM with the weights re-read per plane 106.9 ms -- the kernel
    // wants instruction-level parallelism inside the warp more than it wants warps
    if (v == 0) return launch_z16p<true, 2>(c, s, A);
    if (v == 6) return launch_z16p<false, 4>(c, s, A);
    if (v == 7) return launch_z16p<true, 3>(c, s, A);
  }
  switch (v) {
    case 1: return launch_z16_v<8, 2, false>(c, s, A);   // 16 warps/SM, weights loaded at the point of use (exposed latency)
    case 2: return launch_z16_v<8, 1, true>(c, s, A);    // 8 warps/SM, 224 registers
    case 3: return launch_z16_v<4, 4, false>(c, s, A);   // 16 warps/SM in 4-warp CTAs
    case 4: return launch_z16_v<4, 2, true>(c, s, A);    // 8 warps/SM in 4-warp CTAs, weights in registers
    // 12 warps/SM (3 CTAs of 4 warps, 168 registers, no spills): the weights of a row pair stay in registers for all
    // planes of the group -- measured fastest (100.5 ms per C2 realisation vs 101.2 / 111.3 / 113.4 for cases 1 / 2 / 4)
    default: return launch_z16_v<4, 3, true>(c, s, A);
  }
}

// A/B switches (tests, profiling) are looked up per launch so that a test can flip them; only the value "1" enables one
static inline bool env_on(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1' && e[1] == 0;
}

// the fused nz = 512 kernel with table-driven pre / post-processing (pass_z16p_fused_kernel) takes this launch
static bool z16p_takes(const sww_ctx* c, const ZArgs& A) {
  const char* e = getenv("SWW_Z16_VARIANT");
  const int v = e ? atoi(e) : 0;
  return c->g.nz / 2 == Z16_H && A.lm < 0 && A.wmap && A.Kz_out <= 128 && A.Kz_in <= 128 && !env_on("SWW_NO_Z16") &&
         (c->g.nx * (long long)c->g.ny) % 2 == 0 && (v == 0 || v == 6 || v == 7);
}

static int run_z(sww_ctx* c, Sm100State* s, bool inv, bool fwd, ZArgs A) {
  SWW_PROF(c, inv && fwd ? "fft.Z c2r*w*r2c fused" : (inv ? "fft.Z c2r" : "fft.Z r2c"));
  A.ymaps = c->ylm_maps;
  A.N = c->g.N;
  const int H = c->g.nz / 2;
  const bool ylm = A.lm >= 0;
  if (s->any_z) {
    SWW_REQUIRE(!(A.q_tile || A.r_tile), "tile-major planes need power-of-two axes");
    const AnyPlan& pl = s->plan_z;
    const size_t smem = ANY_SMEM_TILES(pl.n);
    const int threads = any_threads_for(pl);
    const long long tiles = ((long long)c->g.nx * c->g.ny + 15) / 16;
#define LAUNCH_Z_ANY(II, FF)                                                                              \
  {                                                                                                       \
    SWW_TRY((set_smem(gpass_z_kernel<II, FF>, smem)));                                                    \
    const int grid = grid_for_items(tiles, any_occ(gpass_z_kernel<II, FF>, smem, threads));               \
    gpass_z_kernel<II, FF><<<grid, threads, smem, c->stream>>>(c->g, pl, A, s->tw_c);                     \
  }
    if (inv && fwd) {
      SWW_REQUIRE(!ylm, "fused z pass has no Y_lm variant");
      LAUNCH_Z_ANY(true, true)
    } else if (inv) {
      LAUNCH_Z_ANY(true, false)
    } else {
      LAUNCH_Z_ANY(false, true)
    }
#undef LAUNCH_Z_ANY
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  const bool no_z16 = env_on("SWW_NO_Z16");
  if (inv && fwd && !ylm && H == Z32_H && A.wmap && A.Kz_out <= 256 && A.Kz_in <= Z32_H + 1 && env_on("SWW_Z32")) {
    // experimental (not yet validated on a GPU): opt-in only
    SWW_TRY(launch_z32(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  SWW_REQUIRE(!(A.q_tile || A.r_tile) || (inv && fwd && z16p_takes(c, A)), "tile-major planes need the fused nz = 512 kernel");
  if (inv && fwd && !ylm && H == Z16_H && A.wmap && A.Kz_out <= 128 && A.Kz_in <= Z16_H + 1 && !no_z16 &&
      (c->g.nx * (long long)c->g.ny) % 2 == 0) {
    SWW_TRY(launch_z16(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  if (!inv && fwd && H == Z16_H && A.Kz_out <= 128 && (c->g.nx * (long long)c->g.ny) % 2 == 0 && !no_z16 && !env_on("SWW_NO_Z16F")) {
    // forward-only transforms at nz = 512: the 16 x 16 kernel with one exchange (pass_z16f_kernel)
    if (ylm) SWW_TRY(launch_z16f<true>(c, s, A));
    else SWW_TRY(launch_z16f<false>(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  if (inv && !fwd && H == 128 && A.Kz_in <= 64 && (c->g.nx * (long long)c->g.ny) % 4 == 0 && !no_z16 && env_on("SWW_Z8I") &&
      one_harmonic_degree(A)) {
    // inverse-only transforms at nz = 256 (bispectrum phi maps): 8 x 16 with one exchange, polynomial harmonics (pass_z8i_kernel).
    // OPT-IN (SWW_Z8I=1): validated against the warp-per-row kernel, but measured slower on the C4 pair -- 442 ms against 340 ms
    // (558 ms before the cp.async staging of the next plane's rows): 16 values per lane cost 168 registers, 12 warps per SM, and
    // the pass is bound by its per-plane latencies, not by the cached-harmonics traffic this kernel removes (profiles/r3s)
    SWW_TRY(launch_z8i(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  if (!inv && fwd && H == 128 && A.Kz_out <= 64 && (c->g.nx * (long long)c->g.ny) % 4 == 0 && !no_z16 && !env_on("SWW_NO_Z16F")) {
    // forward-only transforms at nz = 256: 16 x 8 with one exchange (pass_z8f_kernel)
    if (ylm) SWW_TRY(launch_z8f<true>(c, s, A));
    else SWW_TRY(launch_z8f<false>(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  if (!inv && fwd && H == Z32_H && A.Kz_out <= 256 && !no_z16 && !env_on("SWW_NO_Z16F")) {
    // forward-only transforms at nz = 1024: two 16 x 16 halves per row and a radix-2 join (pass_z32f_kernel)
    if (ylm) SWW_TRY(launch_z32f<true>(c, s, A));
    else SWW_TRY(launch_z32f<false>(c, s, A));
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
#define M(HH)                                                      \
  if (inv && fwd) {                                                \
    SWW_REQUIRE(!ylm, "fused z pass has no Y_lm variant");        \
    SWW_TRY((launch_zw<HH, true, true, false>(c, s, A)));          \
  } else if (inv) {                                                \
    if (ylm) SWW_TRY((launch_zw<HH, true, false, true>(c, s, A))); \
    else SWW_TRY((launch_zw<HH, true, false, false>(c, s, A)));    \
  } else {                                                         \
    if (ylm) SWW_TRY((launch_zw<HH, false, true, true>(c, s, A))); \
    else SWW_TRY((launch_zw<HH, false, true, false>(c, s, A)));    \
  }
  switch (H) {
    case 64: M(64); break;
    case 128: M(128); break;
    case 256: M(256); break;
    case 512: M(512); break;
    default: SWW_REQUIRE(false, "unsupported z length %d", 2 * H);
  }
#undef M
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

static BoxAxis full_axis(int n) { return BoxAxis{n, n, n - 1}; }
static BoxAxis box_axis(int n, int b) {
  int K = 2 * b + 1;
  if (K >= n) return full_axis(n);
  return BoxAxis{n, K, b};
}

static int get_state(sww_ctx* c, Sm100State** out) {
  SWW_TRY(sww_fft_sm100_init(c));
  Sm100State* s = (Sm100State*)c->sm100;
  SWW_TRY(bind_bufs(c, s));
  *out = s;
  return 0;
}

// ---- generic transforms (full box): used by every pipeline when fft_backend == 1 ----
int sww_fft_sm100_r2c(sww_ctx* c, const double* in, double2* out) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  const Geo& g = c->g;
  AxisPlan pl = make_plan(g, full_axis(g.nx), full_axis(g.ny));
  int Kz = g.nzh, Kzp = up8(Kz);
  ZArgs Z = {};
  Z.R = s->R;
  Z.Kz_out = Kz;
  Z.Kzp_out = Kzp;
  Z.real_in = in;
  Z.scale = 1.0;
  Z.lm = -1;
  SWW_TRY(run_z(c, s, false, true, Z));
  F3Args A = {};
  A.dst = out;
  return run_forward_cols(c, s, pl, Kz, Kzp, 0, A);
}

int sww_fft_sm100_c2r(sww_ctx* c, double2* in, double* out) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  const Geo& g = c->g;
  AxisPlan pl = make_plan(g, full_axis(g.nx), full_axis(g.ny));
  int Kz = g.nzh, Kzp = up8(Kz);
  SWW_TRY(run_inverse_cols(c, s, pl, Kz, Kzp, in, -1));
  ZArgs Z = {};
  Z.Q = s->Q;
  Z.Kz_in = Kz;
  Z.Kzp_in = Kzp;
  Z.real_out = out;
  Z.scale = 1.0;
  Z.lm = -1;
  return run_z(c, s, true, false, Z);
}

// ---- complex 3-D transforms on the same passes: z (complex, in place) + the two column passes over the full box ----
int sww_fft_sm100_c2c_supported(const sww_ctx* c) {
  const int nz = c->g.nz;
  if (!sww_fft_sm100_supported(c)) return 0;
  if (nz == 128 || nz == 256 || nz == 512) return 1;
  // any-length z axis: the planes have pitch nz; a ragged last 8-column tile is only handled by the any-length column passes
  // (power-of-two x / y axes of at most ANY_MAX_N points are sent through them for such a mesh)
  return !z_pow2(nz) && any_len(nz) && (nz % 8 == 0 || (any_len(c->g.nx) && any_len(c->g.ny)));
}

template <int DIR>
static int run_zc(sww_ctx* c, Sm100State* s, double2* data) {
  SWW_PROF(c, DIR > 0 ? "fft.Zc forward (complex)" : "fft.Zc inverse (complex)");
  const int H = c->g.nz;
  const long long nrows = (long long)c->g.nx * c->g.ny;
  if (s->any_z) {
    const AnyPlan& pl = s->plan_z;
    const size_t smem = ANY_SMEM_TILES(pl.n);
    const int threads = any_threads_for(pl);
    SWW_TRY((set_smem(gpass_zc_kernel<DIR>, smem)));
    const int grid = grid_for_items((nrows + 7) / 8, any_occ(gpass_zc_kernel<DIR>, smem, threads));
    gpass_zc_kernel<DIR><<<grid, threads, smem, c->stream>>>(pl, nrows, data, s->tw_c);
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
#define M(HH)                                                                                                     \
  {                                                                                                               \
    const size_t smem = (size_t)(WTabSize<HH, false>::value + 8 * (WGeo<HH>::RW * WGeo<HH>::ROWLEN + 8)) * 16;   \
    const long long ctas = (nrows / WGeo<HH>::RW + 7) / 8;                                                        \
    const int grid = (int)std::min<long long>(ctas, (long long)SMS * 2);                                          \
    SWW_TRY((set_smem(pass_zc_kernel<HH, DIR>, smem)));                                                           \
    pass_zc_kernel<HH, DIR><<<grid, 256, smem, c->stream>>>(nrows, data, s->tw_c);                               \
  }
  switch (H) {
    case 64: M(64); break;
    case 128: M(128); break;
    case 256: M(256); break;
    case 512: M(512); break;
    default: SWW_REQUIRE(false, "complex z pass: unsupported length %d", H);
  }
#undef M
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// out = DFT(in) (forward, e^{-i..}) or the unnormalised inverse; in == out allowed.  The pass buffers of the real
// transforms (one contiguous region) serve as the single intermediate of N complex values.
int sww_fft_sm100_c2c(sww_ctx* c, const double2* in, double2* out, int inverse) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(sww_fft_sm100_c2c_supported(c), "complex transforms on the sm_100a passes: unsupported mesh (power-of-two nz must be 128, 256 or 512)");
  Geo& g = c->g;
  if (!s->tw_c) {
    std::vector<double2> t = make_tw(g.nz, g.nz);
    SWW_CUDA(cudaMalloc(&s->tw_c, t.size() * sizeof(double2)));
    SWW_CUDA(cudaMemcpy(s->tw_c, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice));
  }
  SWW_REQUIRE(3 * s->buf_elems >= (size_t)g.N, "pass buffers too small for a complex transform");
  // the column passes index a spectrum as [nx][ny][nzh]: for complex data every z index is a column
  const int save_nzh = g.nzh;
  const size_t save_elems = s->buf_elems;
  double2* save_R = s->R;
  g.nzh = g.nz;
  s->buf_elems = 3 * save_elems;
  const bool save_ax = s->any_x, save_ay = s->any_y;
  if (s->any_z && g.nz % 8 != 0) s->any_x = s->any_y = true;   // ragged last z tile: any-length column passes on every axis
  AxisPlan pl = make_plan(g, full_axis(g.nx), full_axis(g.ny));
  const int Kz = g.nz, Kzp = g.nz;
  int rc = 0;
  if (!inverse) {
    if (in != out) rc = cudaMemcpyAsync(out, in, sizeof(double2) * (size_t)g.N, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess ? 0 : -2;
    if (rc == 0) rc = run_zc<1>(c, s, out);
    s->R = out;
    F3Args A = {};
    A.dst = out;
    if (rc == 0) rc = run_forward_cols(c, s, pl, Kz, Kzp, 0, A);
  } else {
    rc = run_inverse_cols(c, s, pl, Kz, Kzp, in, -1, nullptr, 1, out, 0);
    if (rc == 0) rc = run_zc<-1>(c, s, out);
  }
  g.nzh = save_nzh;
  s->buf_elems = save_elems;
  s->R = save_R;
  s->any_x = save_ax;
  s->any_y = save_ay;
  return rc;
}

// ---- pruned / fused chains ----
// forward transform of (real_in * wmap * Y_lm(r^)) restricted to the k_max box, F3 epilogue `mode`
int sww_sm100_forward_box(sww_ctx* c, const double* real_in, const double* wmap, int lm_r, int mode, F3Args A) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  const Geo& g = c->g;
  AxisPlan pl = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  int Kz = g.Kz, Kzp = up8(Kz);
  ZArgs Z = {};
  Z.R = s->R;
  Z.Kz_out = Kz;
  Z.Kzp_out = Kzp;
  Z.real_in = real_in;
  Z.wmap = wmap;
  Z.scale = 1.0;
  Z.lm = lm_r;
  SWW_TRY(run_z(c, s, false, true, Z));
  return run_forward_cols(c, s, pl, Kz, Kzp, mode, A);
}

// inverse transform of the box part of `src` (optionally only shell `shell`) into a real map:
// real_out (+)= scale * wmap * Y_lm(r^) * IFT[...]
int sww_sm100_inverse_box(sww_ctx* c, const double2* src, int shell, int box_shell, const double* wmap, int lm_r,
                          double scale, int accumulate, double* real_out) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  const Geo& g = c->g;
  int b_x = g.bx, b_y = g.by, b_z = g.bz;
  if (shell >= 0 && box_shell < 0) box_shell = shell;
  if (box_shell >= 0) {
    b_x = s->shell_bx[box_shell];
    b_y = s->shell_by[box_shell];
    b_z = s->shell_bz[box_shell];
  }
  AxisPlan pl = make_plan(g, box_axis(g.nx, b_x), box_axis(g.ny, b_y));
  int Kz = b_z + 1 < g.nzh ? b_z + 1 : g.nzh, Kzp = up8(Kz);
  SWW_TRY(run_inverse_cols(c, s, pl, Kz, Kzp, src, shell));
  ZArgs Z = {};
  Z.Q = s->Q;
  Z.Kz_in = Kz;
  Z.Kzp_in = Kzp;
  Z.real_out = real_out;
  Z.wmap = wmap;
  Z.scale = scale;
  Z.lm = lm_r;
  Z.accumulate = accumulate;
  return run_z(c, s, true, false, Z);
}

// fused P_l Fisher row: out_row[q*n_k + b] = scale * sum_{k in b} w Re[conj(FT[wmap * IFT[Theta_a F]]) G_q]
int sww_sm100_pk_row(sww_ctx* c, const double2* F, int shell, const double* wmap, double inv_scale,
                     const double2* const* G, int n_g, double scale, double* out_row) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  const Geo& g = c->g;
  AxisPlan pli = make_plan(g, box_axis(g.nx, s->shell_bx[shell]), box_axis(g.ny, s->shell_by[shell]));
  int Kzi = s->shell_bz[shell] + 1, Kzpi = up8(Kzi);
  AxisPlan plo = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  int Kzo = g.Kz, Kzpo = up8(Kzo);
  SWW_TRY(run_inverse_cols(c, s, pli, Kzi, Kzpi, F, shell));
  ZArgs Z = {};
  Z.Q = s->Q;
  Z.Kz_in = Kzi;
  Z.Kzp_in = Kzpi;
  Z.R = s->R;
  Z.Kz_out = Kzo;
  Z.Kzp_out = Kzpo;
  Z.wmap = wmap;
  Z.scale = inv_scale;
  Z.lm = -1;
  SWW_TRY(run_z(c, s, true, true, Z));
  F3Args A = {};
  for (int q = 0; q < n_g; ++q) A.G[q] = G[q];
  A.n_g = n_g;
  A.scale = scale;
  A.out_row = out_row;
  return run_forward_cols(c, s, plo, Kzo, Kzpo, 2, A);
}

// sum over nb gathered spectra of  wmap * scale * Y_{lms_r[b]}(r^) * IFT[ gather_b ]  -> real_out, one write.
// Every gather is restricted to the box of shell `box_shell`; the nb intermediate planes live in qstack.
int sww_sm100_inverse_multi(sww_ctx* c, int nb, const GatherArgs* gas, int box_shell, const int* lms_r,
                            const double* wmap, double scale, int accumulate, double2* qstack, double* real_out) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(nb >= 1 && nb <= SWW_MAX_ZBATCH, "inverse_multi: batch of %d planes", nb);
  const Geo& g = c->g;
  int b_x = g.bx, b_y = g.by, b_z = g.bz;
  if (box_shell >= 0) {
    b_x = s->shell_bx[box_shell];
    b_y = s->shell_by[box_shell];
    b_z = s->shell_bz[box_shell];
  }
  AxisPlan pl = make_plan(g, box_axis(g.nx, b_x), box_axis(g.ny, b_y));
  int Kz = b_z + 1 < g.nzh ? b_z + 1 : g.nzh, Kzp = up8(Kz);
  const long long stride = (long long)g.nx * g.ny * Kzp;
  ZArgs Z = {};
  bool any_ylm = false;
  SWW_TRY(run_inverse_cols(c, s, pl, Kz, Kzp, nullptr, -1, gas, nb, qstack, stride));
  for (int b = 0; b < nb; ++b) {
    Z.lms[b] = lms_r[b];
    any_ylm = any_ylm || lms_r[b] >= 0;
  }
  Z.Q = qstack;
  Z.Kz_in = Kz;
  Z.Kzp_in = Kzp;
  Z.real_out = real_out;
  Z.wmap = wmap;
  Z.scale = scale;
  Z.accumulate = accumulate;
  Z.nb = nb;
  Z.q_stride = stride;
  Z.lm = (nb == 1) ? lms_r[0] : (any_ylm ? 0 : -1);   // selects the Y_lm-capable kernel variant
  return run_z(c, s, true, false, Z);
}

size_t sww_sm100_qstack_bytes(const sww_ctx* c, int nb) {
  return (size_t)nb * c->g.nx * c->g.ny * up8(c->g.Kz) * sizeof(double2);
}

// ------------------------------------------------------------------------------------------
// Shell-ordered spectra.  The Fisher row F[alpha][(l', b)] = scale * sum_{k in b} w Re[conj(U_alpha) G_l'] needs a
// k-shell reduction of every transformed row.  Instead of binning inside the last forward pass, that pass
// scatters the modes that fall in a bin into SHELL ORDER -- sorted by (bin, outer index, inner index, iz), a
// static permutation of the output box -- so that bin b of every row is one contiguous segment, and the
// reduction becomes a set of segmented dot products with perfectly coalesced reads (shell_dot_kernel), against
// G maps sorted once per realisation with the half-spectrum weight folded in.
// ------------------------------------------------------------------------------------------
#define SHELL_CH 4096      // modes per chunk (a chunk never straddles a bin)
#define SHELL_RT 8         // rows per CTA

static int build_shell_order(sww_ctx* c, Sm100State* s) {
  if (s->shell_state != 0) return 0;
  const Geo& g = c->g;
  AxisPlan pl = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  const int a1 = pl.x_first_fwd ? 0 : 1, a2 = 1 - a1;
  const BoxAxis bO = a1 == 0 ? pl.bx : pl.by, bA = a2 == 0 ? pl.bx : pl.by;
  const int Kz = g.Kz, Kzp = up8(Kz);
  std::vector<uint8_t> bm((size_t)g.nx * g.ny * g.nzh);
  SWW_CUDA(cudaMemcpy(bm.data(), g.binmap, bm.size(), cudaMemcpyDeviceToHost));
  std::vector<long long> cnt(g.n_k + 1, 0);
  auto cell_of = [&](int jo, int ja, int iz) -> size_t {
    int io = bO.idx(jo), ia = bA.idx(ja);
    int ix = a2 == 0 ? ia : io, iy = a2 == 0 ? io : ia;
    return ((size_t)ix * g.ny + iy) * g.nzh + iz;
  };
  for (int jo = 0; jo < bO.K; ++jo)
    for (int ja = 0; ja < bA.K; ++ja)
      for (int iz = 0; iz < Kz; ++iz) {
        int b = bm[cell_of(jo, ja, iz)];
        if (b != SWW_NO_BIN) cnt[b]++;
      }
  std::vector<long long> off(g.n_k + 1, 0);
  for (int b = 0; b < g.n_k; ++b) off[b + 1] = off[b] + cnt[b];
  const long long S = off[g.n_k];
  if (S <= 0 || S >= (1LL << 31)) {
    s->shell_state = -1;
    return 0;
  }
  std::vector<int> perm((size_t)bO.K * bA.K * Kzp, -1);
  std::vector<long long> cur(off.begin(), off.end() - 1);
  for (int jo = 0; jo < bO.K; ++jo)
    for (int ja = 0; ja < bA.K; ++ja)
      for (int iz = 0; iz < Kz; ++iz) {
        int b = bm[cell_of(jo, ja, iz)];
        if (b != SWW_NO_BIN) perm[((size_t)jo * bA.K + ja) * Kzp + iz] = (int)cur[b]++;
      }
  std::vector<int4> chunks;
  std::vector<int> bin_chunk(g.n_k + 1, 0);
  for (int b = 0; b < g.n_k; ++b) {
    bin_chunk[b] = (int)chunks.size();
    for (long long st = off[b]; st < off[b + 1]; st += SHELL_CH) {
      long long len = std::min<long long>(SHELL_CH, off[b + 1] - st);
      chunks.push_back(make_int4((int)st, (int)len, b, 0));
    }
  }
  bin_chunk[g.n_k] = (int)chunks.size();
  SWW_CUDA(cudaMalloc(&s->d_perm, perm.size() * sizeof(int)));
  SWW_CUDA(cudaMemcpy(s->d_perm, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));
  SWW_CUDA(cudaMalloc(&s->d_chunks, std::max<size_t>(1, chunks.size()) * sizeof(int4)));
  SWW_CUDA(cudaMemcpy(s->d_chunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  SWW_CUDA(cudaMalloc(&s->d_bin_chunk, bin_chunk.size() * sizeof(int)));
  SWW_CUDA(cudaMemcpy(s->d_bin_chunk, bin_chunk.data(), bin_chunk.size() * sizeof(int), cudaMemcpyHostToDevice));
  s->S = S;
  s->n_chunks = (int)chunks.size();
  s->shell_state = 1;
  return 0;
}

long long sww_sm100_shell_cells(sww_ctx* c) {
  if (!sww_fft_sm100_supported(c) || sww_fft_sm100_init(c) != 0) return 0;
  Sm100State* s = (Sm100State*)c->sm100;
  if (build_shell_order(c, s) != 0 || s->shell_state != 1) return 0;
  return s->S;
}
int sww_sm100_shell_chunks(sww_ctx* c) {
  Sm100State* s = (Sm100State*)c->sm100;
  return s && s->shell_state == 1 ? s->n_chunks : 0;
}

// Gs[q][perm] = w(iz) * G_q[cell]
__global__ void shell_sort_g_kernel(Geo g, BoxAxis bA, BoxAxis bO, int ax, int Kz, int Kzp, const int* __restrict__ perm,
                                    long long S, F3Args A, double2* __restrict__ Gs) {
  const long long total = (long long)bO.K * bA.K * Kzp;
  const bool nz_even = (g.nz % 2) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = perm[i];
    if (p < 0) continue;
    const int iz = (int)(i % Kzp);
    const long long r = i / Kzp;
    const int ja = (int)(r % bA.K), jo = (int)(r / bA.K);
    const int ia = bA.idx(ja), io = bO.idx(jo);
    const long long cell = (ax == 0 ? (long long)ia * g.ny + io : (long long)io * g.ny + ia) * g.nzh + iz;
    const double w = (iz == 0 || (nz_even && iz == g.nz / 2)) ? 1.0 : 2.0;
    for (int q = 0; q < A.n_g; ++q) {
      double2 v = A.G[q][cell];
      Gs[(long long)q * S + p] = make_double2(w * v.x, w * v.y);
    }
  }
}

int sww_sm100_shell_sort_g(sww_ctx* c, const double2* const* G, int n_g, double2* Gs) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(s->shell_state == 1, "shell order not available");
  SWW_PROF(c, "shell.sort G");
  const Geo& g = c->g;
  AxisPlan pl = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  const int a1 = pl.x_first_fwd ? 0 : 1, a2 = 1 - a1;
  const BoxAxis bO = a1 == 0 ? pl.bx : pl.by, bA = a2 == 0 ? pl.bx : pl.by;
  F3Args A = {};
  for (int q = 0; q < n_g; ++q) A.G[q] = G[q];
  A.n_g = n_g;
  shell_sort_g_kernel<<<SMS * 8, 256, 0, c->stream>>>(g, bA, bO, a2, g.Kz, up8(g.Kz), s->d_perm, s->S, A, Gs);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// partial[(chunk * R + r) * 3 + q] = sum_{i in chunk} Re[conj(Us[r][i]) Gs[q][i]]
// block = chunk * ntile + tile (tile fastest: the CTAs sharing a chunk's G segment run together -> L2 hits)
__global__ void __launch_bounds__(256)
shell_dot_kernel(const double2* __restrict__ Us, int R, long long S, const double2* __restrict__ Gs, int n_g,
                 const int4* __restrict__ chunks, int ntile, double* __restrict__ partial) {
  __shared__ double red[8][SHELL_RT * 3];
  const int chunk = blockIdx.x / ntile, tile = blockIdx.x - chunk * ntile;
  const int4 ch = chunks[chunk];
  const int r0 = tile * SHELL_RT;
  const int nr = min(SHELL_RT, R - r0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[SHELL_RT][3];
#pragma unroll
  for (int r = 0; r < SHELL_RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.0;
  const double2* __restrict__ Up = Us + (long long)r0 * S + ch.x;
  const double2* __restrict__ Gp = Gs + ch.x;
  for (int i = tid; i < ch.y; i += 256) {
    double2 gq[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) gq[q] = q < n_g ? Gp[(long long)q * S + i] : make_double2(0.0, 0.0);
    double2 u[SHELL_RT];
#pragma unroll
    for (int r = 0; r < SHELL_RT; ++r) u[r] = r < nr ? Up[(long long)r * S + i] : make_double2(0.0, 0.0);
#pragma unroll
    for (int r = 0; r < SHELL_RT; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) acc[r][q] += u[r].x * gq[q].x + u[r].y * gq[q].y;
  }
#pragma unroll
  for (int r = 0; r < SHELL_RT; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double v = acc[r][q];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) red[warp][r * 3 + q] = v;
    }
  __syncthreads();
  if (tid < SHELL_RT * 3) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][tid];
    const int r = tid / 3, q = tid - 3 * r;
    if (r < nr) partial[((long long)chunk * R + r0 + r) * 3 + q] = v;
  }
}

// F row of (local row r): shell ka0 + r / n_l, multipole r % n_l;  F[row][q * n_k + b] = scale * sum of the bin's chunks
__global__ void shell_dot_reduce_kernel(const double* __restrict__ partial, int R, int n_g, int n_k, int n_l, int ka0,
                                        const int* __restrict__ bin_chunk, double scale, double* __restrict__ F) {
  const int total = R * n_g * n_k;
  const int n_b = n_k * n_l;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx % n_k, q = (idx / n_k) % n_g, r = idx / (n_k * n_g);
    double v = 0.0;
    for (int ch = bin_chunk[b]; ch < bin_chunk[b + 1]; ++ch) v += partial[((long long)ch * R + r) * 3 + q];
    const int ka = ka0 + r / n_l, l_i = r % n_l;
    F[(long long)(l_i * n_k + ka) * n_b + q * n_k + b] = v * scale;
  }
}

int sww_sm100_shell_dot(sww_ctx* c, const double2* Us, int R, const double2* Gs, int n_g, double scale, double* partial,
                        double* F, int n_l, int ka0) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(s->shell_state == 1 && R >= 1 && n_g >= 1 && n_g <= 3, "shell_dot: bad arguments");
  const int ntile = (R + SHELL_RT - 1) / SHELL_RT;
  {
    SWW_PROF(c, "shell.dot rows x G");
    shell_dot_kernel<<<(unsigned)((long long)s->n_chunks * ntile), 256, 0, c->stream>>>(Us, R, s->S, Gs, n_g, s->d_chunks, ntile,
                                                                                       partial);
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
  }
  {
    SWW_PROF(c, "shell.dot reduce");
    shell_dot_reduce_kernel<<<sww_grid_for((long long)R * n_g * c->g.n_k, 128), 128, 0, c->stream>>>(
        partial, R, n_g, c->g.n_k, n_l, ka0, s->d_bin_chunk, scale, F);
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
  }
  return 0;
}

// the n_l rows of shell `shell`: I1 -> I2 -> Z(c2r * wmap * r2c) -> F2 -> F3(store in shell order) into Us_rows[b * S]
int sww_sm100_pk_rows_store(sww_ctx* c, int nrows, const double2* const* F, int shell, const double* wmap, double inv_scale,
                            double2* Us_rows) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(nrows >= 1 && nrows <= SWW_MAX_L && s->shell_state == 1, "pk_rows_store: %d rows", nrows);
  const Geo& g = c->g;
  AxisPlan pli = make_plan(g, box_axis(g.nx, s->shell_bx[shell]), box_axis(g.ny, s->shell_by[shell]));
  int Kzi = s->shell_bz[shell] + 1, Kzpi = up8(Kzi);
  AxisPlan plo = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  int Kzo = g.Kz, Kzpo = up8(Kzo);
  const long long qstride = (long long)g.nx * g.ny * Kzpi, rstride = (long long)g.nx * g.ny * Kzpo;
  SWW_REQUIRE((size_t)rstride * nrows <= s->buf_elems, "pk_rows: pass buffers too small for %d planes", nrows);
  GatherArgs ga[SWW_MAX_L] = {};
  for (int b = 0; b < nrows; ++b) {
    ga[b].src = F[b];
    ga[b].shell = shell;
  }
  ZArgs Z = {};
  Z.Q = s->Q;
  Z.Kz_in = Kzi;
  Z.Kzp_in = Kzpi;
  Z.R = s->R;
  Z.Kz_out = Kzo;
  Z.Kzp_out = Kzpo;
  Z.wmap = wmap;
  Z.scale = inv_scale;
  Z.lm = -1;
  Z.nb = nrows;
  Z.q_stride = qstride;
  Z.r_stride = rstride;
  // tile-major Q and R when the fused nz = 512 kernel runs the z pass: the column passes next to it (second inverse,
  // first forward) then move each 8-column tile as one contiguous block
  const char* etl = getenv("SWW_TILE_MAJOR");
  // Measured on C2 / C5 (profiles/r2j): R tile-major F2 42.3 -> 40.7 ms but the z pass 92.2 -> 93.6 ms; Q tile-major I2
  // 24.5 -> 23.9 ms, z pass +5 ms; the bulk-copy F2 kernel (one CTA per SM) 44.3 ms.  F2 already moves 5.3 TB/s (82 % of
  // the measured HBM peak) in row-major form, so the layout is OFF by default: SWW_TILE_MAJOR = q | r | b(oth) opts in.
  if (z16p_takes(c, Z) && etl && (etl[0] == 'q' || etl[0] == 'r' || etl[0] == 'b')) {
    if (etl[0] != 'r') Z.q_tile = (pli.x_first_fwd ? 0 : 1) + 1;      // axis of the second inverse pass
    if (etl[0] != 'q') Z.r_tile = (plo.x_first_fwd ? 0 : 1) + 1;      // axis of the first forward pass
  }
  // columns of the (k_o, k_z) plane outside the sphere of the shell (inverse) / of k_max (forward) are skipped by the
  // column passes (a margin of 1e-12 keeps modes that numpy's sqrt comparison might still bin); SWW_NO_SPHERE_SKIP=1: off
  const bool sphere = !env_on("SWW_NO_SPHERE_SKIP");
  const double khi_a = c->h_khi[shell], kmax_all = *std::max_element(c->h_khi.begin(), c->h_khi.end());
  const double klim_i = sphere ? khi_a * khi_a * (1.0 + 1e-12) : 0.0, klim_o = sphere ? kmax_all * kmax_all * (1.0 + 1e-12) : 0.0;
  SWW_TRY(run_inverse_cols(c, s, pli, Kzi, Kzpi, nullptr, -1, ga, nrows, s->Q, qstride, Z.q_tile, klim_i));
  SWW_TRY(run_z(c, s, true, true, Z));
  F3Args Ab[SWW_MAX_L] = {};
  for (int b = 0; b < nrows; ++b) {
    Ab[b].perm = s->d_perm;
    Ab[b].sorted_dst = Us_rows + (long long)b * s->S;
    Ab[b].klim2 = klim_o;
  }
  return run_forward_cols(c, s, plo, Kzo, Kzpo, 3, Ab[0], nrows, rstride, Ab, Z.r_tile, klim_o);
}

int sww_sm100_pk_rows(sww_ctx* c, int nrows, const double2* const* F, int shell, const double* wmap, double inv_scale,
                      const double2* const* G, int n_g, double scale, double* const* out_rows) {
  Sm100State* s;
  SWW_TRY(get_state(c, &s));
  SWW_REQUIRE(nrows >= 1 && nrows <= SWW_MAX_L, "pk_rows: %d rows", nrows);
  const Geo& g = c->g;
  AxisPlan pli = make_plan(g, box_axis(g.nx, s->shell_bx[shell]), box_axis(g.ny, s->shell_by[shell]));
  int Kzi = s->shell_bz[shell] + 1, Kzpi = up8(Kzi);
  AxisPlan plo = make_plan(g, box_axis(g.nx, g.bx), box_axis(g.ny, g.by));
  int Kzo = g.Kz, Kzpo = up8(Kzo);
  const long long qstride = (long long)g.nx * g.ny * Kzpi, rstride = (long long)g.nx * g.ny * Kzpo;
  SWW_REQUIRE((size_t)rstride * nrows <= s->buf_elems, "pk_rows: pass buffers too small for %d planes", nrows);
  GatherArgs ga[SWW_MAX_L] = {};
  for (int b = 0; b < nrows; ++b) {
    ga[b].src = F[b];
    ga[b].shell = shell;
  }
  ZArgs Z = {};
  Z.Q = s->Q;
  Z.Kz_in = Kzi;
  Z.Kzp_in = Kzpi;
  Z.R = s->R;
  Z.Kz_out = Kzo;
  Z.Kzp_out = Kzpo;
  Z.wmap = wmap;
  Z.scale = inv_scale;
  Z.lm = -1;
  Z.nb = nrows;
  Z.q_stride = qstride;
  Z.r_stride = rstride;
  // tile-major Q and R when the fused nz = 512 kernel runs the z pass: the column passes next to it (second inverse,
  // first forward) then move each 8-column tile as one contiguous block
  const char* etl = getenv("SWW_TILE_MAJOR");
  // Measured on C2 / C5 (profiles/r2j): R tile-major F2 42.3 -> 40.7 ms but the z pass 92.2 -> 93.6 ms; Q tile-major I2
  // 24.5 -> 23.9 ms, z pass +5 ms; the bulk-copy F2 kernel (one CTA per SM) 44.3 ms.  F2 already moves 5.3 TB/s (82 % of
  // the measured HBM peak) in row-major form, so the layout is OFF by default: SWW_TILE_MAJOR = q | r | b(oth) opts in.
  if (z16p_takes(c, Z) && etl && (etl[0] == 'q' || etl[0] == 'r' || etl[0] == 'b')) {
    if (etl[0] != 'r') Z.q_tile = (pli.x_first_fwd ? 0 : 1) + 1;      // axis of the second inverse pass
    if (etl[0] != 'q') Z.r_tile = (plo.x_first_fwd ? 0 : 1) + 1;      // axis of the first forward pass
  }
  // columns of the (k_o, k_z) plane outside the sphere of the shell (inverse) / of k_max (forward) are skipped by the
  // column passes (a margin of 1e-12 keeps modes that numpy's sqrt comparison might still bin); SWW_NO_SPHERE_SKIP=1: off
  const bool sphere = !env_on("SWW_NO_SPHERE_SKIP");
  const double khi_a = c->h_khi[shell], kmax_all = *std::max_element(c->h_khi.begin(), c->h_khi.end());
  const double klim_i = sphere ? khi_a * khi_a * (1.0 + 1e-12) : 0.0, klim_o = sphere ? kmax_all * kmax_all * (1.0 + 1e-12) : 0.0;
  SWW_TRY(run_inverse_cols(c, s, pli, Kzi, Kzpi, nullptr, -1, ga, nrows, s->Q, qstride, Z.q_tile, klim_i));
  SWW_TRY(run_z(c, s, true, true, Z));
  F3Args Ab[SWW_MAX_L] = {};
  for (int b = 0; b < nrows; ++b) {
    for (int q = 0; q < n_g; ++q) Ab[b].G[q] = G[q];
    Ab[b].n_g = n_g;
    Ab[b].scale = scale;
    Ab[b].out_row = out_rows[b];
    Ab[b].klim2 = klim_o;
  }
  return run_forward_cols(c, s, plo, Kzo, Kzpo, 2, Ab[0], nrows, rstride, Ab, Z.r_tile, klim_o);
}
