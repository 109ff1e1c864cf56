// Any-length tile FFT (FP64) for the meshes that are not powers of two -- the reference's production grids
// (paramfiles/boss_nC.param:23,33: 258 x 496 x 274 and 172 x 330 x 182; 167 x 307 x 175 for the Nseries bispectra).
//
// Same tile as fft_core.cuh: double2[n][FFT_LDT], 8 independent columns, but the transform is an out-of-place
// Stockham autosort between TWO tiles with a run-time radix plan n = R0 R1 ... (largest radix first).  Every stage
// is a DENSE radix-R DFT: output q of sub-transform j is  sum_r in[j + r n/R] * W_n^(r e),  e = k n/(Ns R) + q n/R,
// k = j mod Ns  -- stage twiddle and DFT_R matrix element are both powers of W_n, so one table lookup per term and
// the exponent advances by a fixed step (one add and one conditional subtract).  A thread forms a block of
// 4 outputs x 4 columns in registers, so every shared-memory input is used 4 times and every table value 4 times:
// 16 complex FMAs per 8 shared loads.  Stage 0 has Ns = 1, its exponents do not depend on j: the lanes of a warp read
// the same table entries (broadcast), which is why the large prime factor (43, 137, 307 ...) goes first.
//
// Cost per point = 4 * sum(R_s) DFMA (258 = 43*3*2: 192;  496 = 31*4*4: 156;  274 = 137*2: 556), against the
// 64 DFMA/clk/SM of the FP64 pipe -- these passes are FP64-bound, not HBM-bound; large prime factors are the
// price of running the reference's meshes without padding (padding would change every k bin).
//
// Compiles for the host as well (tests/test_fft_any_cpu.py builds tools/test_fft_any.cpp against this header).
#pragma once
#include <math.h>

#include "fft_core.cuh"

#define ANY_MAX_STAGES 12
#define ANY_MAX_N 640      // two tiles + the table must fit 227 KB of shared memory: n * (2 * 9 + 1) * 16 B

struct AnyPlan {
  int n, ns;
  int R[ANY_MAX_STAGES];
};

// prime factors, pairs of twos merged into fours, largest radix first
static inline AnyPlan any_make_plan(int n) {
  AnyPlan p;
  p.n = n;
  p.ns = 0;
  int f[32], nf = 0, m = n;
  for (int d = 2; d * d <= m; ++d)
    while (m % d == 0) {
      f[nf++] = d;
      m /= d;
    }
  if (m > 1) f[nf++] = m;
  int twos = 0;
  for (int i = 0; i < nf; ++i) twos += f[i] == 2;
  int r[32], nr = 0;
  for (int i = 0; i < nf; ++i)
    if (f[i] != 2) r[nr++] = f[i];
  for (; twos >= 2; twos -= 2) r[nr++] = 4;
  if (twos) r[nr++] = 2;
  for (int i = 0; i < nr; ++i)   // descending
    for (int j = i + 1; j < nr; ++j)
      if (r[j] > r[i]) {
        int t = r[i];
        r[i] = r[j];
        r[j] = t;
      }
  for (int i = 0; i < nr && i < ANY_MAX_STAGES; ++i) p.R[p.ns++] = r[i];
  for (int i = p.ns; i < ANY_MAX_STAGES; ++i) p.R[i] = 1;
  return p;
}

// dense-radix work of one stage per point, in complex multiply-adds (the row-mesh cost model, rowmesh.cu)
static inline double any_radix_cost(int R) { return (double)R; }

// work items of stage 0 (4 outputs x 4 columns each): the launchers size the CTA from it
static inline int any_stage_items(int n, int R) { return ((R + 3) / 4) * (n / R) * 2; }

// threads per CTA (a multiple of 32, 64 ... 256) that wastes the fewest slots of the most expensive stage
static inline int any_threads_for(const AnyPlan& p) {
  const int items = any_stage_items(p.n, p.R[0]);
  int best = 256;
  double best_eff = 0.0;
  for (int t = 256; t >= 64; t -= 32) {
    const int rounds = (items + t - 1) / t;
    const double eff = (double)items / ((double)rounds * t);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = t;
    }
  }
  return best;
}

#define ANY_QB 4
#define ANY_CB 4

// one dense radix-R Stockham stage: in -> out (both [n][FFT_LDT]); tw[j] = exp(-2 pi i j / n)
template <int DIR>
FFT_HD void any_stage(const double2* in, double2* out, const double2* tw, int n, int R, int Ns, int tid, int nthr) {
  const int m = n / R, nqb = (R + ANY_QB - 1) / ANY_QB;
  const int total = nqb * m * (FFT_TZ / ANY_CB);
  const int tstep = n / (Ns * R);
  for (int it = tid; it < total; it += nthr) {
    const int cb = it & 1, t = it >> 1;
    const int qb = t / m, j = t - qb * m;
    const int k = j % Ns, q0 = qb * ANY_QB;
    int step[ANY_QB], idx[ANY_QB];
    double2 acc[ANY_QB][ANY_CB];
#pragma unroll
    for (int i = 0; i < ANY_QB; ++i) {
      step[i] = (k * tstep + (q0 + i) * m) % n;
      idx[i] = 0;
#pragma unroll
      for (int c = 0; c < ANY_CB; ++c) acc[i][c] = double2{0.0, 0.0};
    }
    const double2* xp = in + j * FFT_LDT + cb * ANY_CB;
    for (int r = 0; r < R; ++r) {
      double2 x[ANY_CB];
#pragma unroll
      for (int c = 0; c < ANY_CB; ++c) x[c] = xp[c];
      xp += m * FFT_LDT;
#pragma unroll
      for (int i = 0; i < ANY_QB; ++i) {
        double2 w = tw[idx[i]];
        if (DIR < 0) w.y = -w.y;
        idx[i] += step[i];
        if (idx[i] >= n) idx[i] -= n;
#pragma unroll
        for (int c = 0; c < ANY_CB; ++c) {
          acc[i][c].x = fma(x[c].x, w.x, acc[i][c].x);   // four FMAs per complex multiply-add (the expression form
          acc[i][c].x = fma(-x[c].y, w.y, acc[i][c].x);  // compiles to a multiply, an FMA and an add)
          acc[i][c].y = fma(x[c].x, w.y, acc[i][c].y);
          acc[i][c].y = fma(x[c].y, w.x, acc[i][c].y);
        }
      }
    }
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int i = 0; i < ANY_QB; ++i)
      if (q0 + i < R) {
        double2* op = out + (j0 + (q0 + i) * Ns) * FFT_LDT + cb * ANY_CB;
#pragma unroll
        for (int c = 0; c < ANY_CB; ++c) op[c] = acc[i][c];
      }
  }
}

#ifdef __CUDACC__
// Transform of the tile in b0 (filled and synchronised by the caller); returns the tile that holds the result (natural
// order), synchronised.  The other tile is scratch.
template <int DIR>
__device__ __forceinline__ double2* any_fft_tile(double2* b0, double2* b1, const double2* tw, const AnyPlan& pl, int tid, int nthr) {
  double2 *in = b0, *out = b1;
  int Ns = 1;
  for (int s = 0; s < pl.ns; ++s) {
    const int R = pl.R[s];
    any_stage<DIR>(in, out, tw, pl.n, R, Ns, tid, nthr);
    __syncthreads();
    Ns *= R;
    double2* t = in;
    in = out;
    out = t;
  }
  return in;
}
#endif
