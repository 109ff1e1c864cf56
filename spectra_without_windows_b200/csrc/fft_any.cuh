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
// Odd radices use the symmetric form of the DFT matrix (any_stage_sym): with s_r = y_r + y_{R-r}, d_r = y_r - y_{R-r},
//   X_q, X_{R-q} = y_0 + sum_{r <= (R-1)/2} s_r cos(2 pi r q / R)  -/+  i sum_r d_r sin(2 pi r q / R),
// real table values times complex data: R^2 FMAs per transform instead of 4 R^2.
//
// Cost per point = 4 * sum(R_s) DFMA for the plain form (258 = 43*3*2: 192;  496 = 31*4*4: 156;  274 = 137*2: 556), against the
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
static inline double any_radix_cost(int R) { return (R & 1) ? 0.5 * R : 2.0; }

// Register tiles of a work item.  ncu on the first version (4 x 4 plain, 4 x 2 symmetric; profiles/r2q) showed the passes
// latency-bound at 8-10 warps per SM -- shared memory holds one or two CTAs and a tile only has n * 8 / 16 items -- with the
// FP64 pipe at 18-36 %: smaller tiles (more items, more warps) win over fewer shared-memory loads per FMA.
#define ANY_QB 4     // plain form: outputs per item
#define ANY_CB 2     //             columns per item
#define ANY_SQB 2    // symmetric form: output pairs (q, R - q) per item
#define ANY_SCB 2    //                 columns per item
#define ANY_MAX_THREADS 512

// work items of a stage
static inline int any_stage_items(int n, int R) {
  if (R & 1) return (((R - 1) / 2 + ANY_SQB - 1) / ANY_SQB) * (n / R) * (FFT_TZ / ANY_SCB);
  if (R == 2 || R == 4) return (n / R) * (FFT_TZ / 2);
  return ((R + ANY_QB - 1) / ANY_QB) * (n / R) * (FFT_TZ / ANY_CB);
}

// threads per CTA (a multiple of 32, 64 ... 512): all items of the most expensive stage (stage 0) in one round if they fit,
// else the count that wastes the fewest slots
static inline int any_threads_for(const AnyPlan& p) {
  const int items = any_stage_items(p.n, p.R[0]);
  if (items <= ANY_MAX_THREADS) return items < 64 ? 64 : (items + 31) / 32 * 32;
  int best = ANY_MAX_THREADS;
  double best_eff = 0.0;
  for (int t = ANY_MAX_THREADS; t >= 256; t -= 32) {
    const int rounds = (items + t - 1) / t;
    const double eff = (double)items / ((double)rounds * t);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = t;
    }
  }
  return best;
}


// one dense radix-R Stockham stage: in -> out (both [n][FFT_LDT]); tw[j] = exp(-2 pi i j / n)
template <int DIR>
FFT_HD void any_stage(const double2* in, double2* out, const double2* tw, int n, int R, int Ns, int tid, int nthr) {
  const int m = n / R, nqb = (R + ANY_QB - 1) / ANY_QB;
  const int total = nqb * m * (FFT_TZ / ANY_CB);
  const int tstep = n / (Ns * R);
  for (int it = tid; it < total; it += nthr) {
    const int cb = it % (FFT_TZ / ANY_CB), t = it / (FFT_TZ / ANY_CB);
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

// Odd radix R, symmetric form.  item = (block of 4 output pairs q / R-q, sub-transform j, 2 columns).
template <int DIR, bool TWID>
FFT_HD void any_stage_sym(const double2* in, double2* out, const double2* tw, int n, int R, int Ns, int tid, int nthr) {
  constexpr int NCB = FFT_TZ / ANY_SCB;
  const int m = n / R, h = (R - 1) / 2, nqb = (h + ANY_SQB - 1) / ANY_SQB;
  const int total = nqb * m * NCB;
  const int tstep = n / (Ns * R);
  for (int it = tid; it < total; it += nthr) {
    const int cb = it % NCB, t = it / NCB;
    const int qb = t / m, j = t - qb * m;
    const int k = j % Ns, q0 = 1 + qb * ANY_SQB;
    const int kt = k * tstep;            // stage twiddle of input r: W_n^(r kt)
    int step[ANY_SQB], idx[ANY_SQB];
    double2 A[ANY_SQB][ANY_SCB], B[ANY_SQB][ANY_SCB], S0[ANY_SCB];
#pragma unroll
    for (int i = 0; i < ANY_SQB; ++i) {
      step[i] = ((q0 + i) * m) % n;
      idx[i] = step[i];
#pragma unroll
      for (int c = 0; c < ANY_SCB; ++c) A[i][c] = B[i][c] = double2{0.0, 0.0};
    }
#pragma unroll
    for (int c = 0; c < ANY_SCB; ++c) S0[c] = double2{0.0, 0.0};
    const double2* xlo = in + (j + m) * FFT_LDT + cb * ANY_SCB;
    const double2* xhi = in + (j + (R - 1) * m) * FFT_LDT + cb * ANY_SCB;
    int ilo = kt, ihi = (int)(((long long)(R - 1) * kt) % n);
    for (int r = 1; r <= h; ++r) {
      double2 s[ANY_SCB], d[ANY_SCB];
      double2 tl = double2{1.0, 0.0}, th = double2{1.0, 0.0};
      if (TWID) {
        tl = tw[ilo];
        th = tw[ihi];
        if (DIR < 0) {
          tl.y = -tl.y;
          th.y = -th.y;
        }
        ilo += kt;
        if (ilo >= n) ilo -= n;
        ihi -= kt;
        if (ihi < 0) ihi += n;
      }
#pragma unroll
      for (int c = 0; c < ANY_SCB; ++c) {
        double2 a = xlo[c], b = xhi[c];
        if (TWID) {
          a = c_mul(a, tl);
          b = c_mul(b, th);
        }
        s[c] = c_add(a, b);
        d[c] = c_sub(a, b);
      }
      xlo += m * FFT_LDT;
      xhi -= m * FFT_LDT;
      if (qb == 0) {
#pragma unroll
        for (int c = 0; c < ANY_SCB; ++c) S0[c] = c_add(S0[c], s[c]);
      }
#pragma unroll
      for (int i = 0; i < ANY_SQB; ++i) {
        const double2 w = tw[idx[i]];    // (cos, -sin) of 2 pi r q / R
        idx[i] += step[i];
        if (idx[i] >= n) idx[i] -= n;
#pragma unroll
        for (int c = 0; c < ANY_SCB; ++c) {
          A[i][c].x = fma(s[c].x, w.x, A[i][c].x);
          A[i][c].y = fma(s[c].y, w.x, A[i][c].y);
          B[i][c].x = fma(d[c].x, w.y, B[i][c].x);
          B[i][c].y = fma(d[c].y, w.y, B[i][c].y);
        }
      }
    }
    // B = -sum d sin:  forward X_q = y0 + A + iB, X_{R-q} = y0 + A - iB;  inverse: the other way round
    const int j0 = (j - k) * R + k;
    const double2* y0p = in + j * FFT_LDT + cb * ANY_SCB;
#pragma unroll
    for (int i = 0; i < ANY_SQB; ++i)
      if (q0 + i <= h) {
        double2* op = out + (j0 + (q0 + i) * Ns) * FFT_LDT + cb * ANY_SCB;
        double2* om = out + (j0 + (R - q0 - i) * Ns) * FFT_LDT + cb * ANY_SCB;
#pragma unroll
        for (int c = 0; c < ANY_SCB; ++c) {
          const double2 y0 = y0p[c];
          const double2 e = double2{y0.x + A[i][c].x, y0.y + A[i][c].y};
          const double2 f = DIR > 0 ? double2{-B[i][c].y, B[i][c].x} : double2{B[i][c].y, -B[i][c].x};
          op[c] = c_add(e, f);
          om[c] = c_sub(e, f);
        }
      }
    if (qb == 0) {
      double2* o0 = out + j0 * FFT_LDT + cb * ANY_SCB;
#pragma unroll
      for (int c = 0; c < ANY_SCB; ++c) o0[c] = c_add(y0p[c], S0[c]);
    }
  }
}

// Radix 2 and 4 as butterflies (no multiplications inside the DFT): item = (sub-transform j, 2 columns)
template <int DIR, int R>
FFT_HD void any_stage_bfly(const double2* in, double2* out, const double2* tw, int n, int Ns, int tid, int nthr) {
  constexpr int CB = 2, NCB = FFT_TZ / CB;
  const int m = n / R;
  const int total = m * NCB;
  const int tstep = n / (Ns * R);
  for (int it = tid; it < total; it += nthr) {
    const int cb = it % NCB, j = it / NCB;
    const int k = j % Ns, kt = k * tstep;
    double2 w[R];
#pragma unroll
    for (int r = 1; r < R; ++r) {
      w[r] = tw[r * kt];               // r kt < n
      if (DIR < 0) w[r].y = -w[r].y;
    }
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      double2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] = in[(j + r * m) * FFT_LDT + cb * CB + c];
      if (Ns > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) u[r] = c_mul(u[r], w[r]);
      }
      dftR<R, DIR>(u);
#pragma unroll
      for (int r = 0; r < R; ++r) out[(j0 + r * Ns) * FFT_LDT + cb * CB + c] = u[r];
    }
  }
}

// one stage: the symmetric form for odd radices, butterflies for 2 and 4 (the plain dense form serves any other even radix)
template <int DIR>
FFT_HD void any_stage_auto(const double2* in, double2* out, const double2* tw, int n, int R, int Ns, int tid, int nthr) {
  if (R & 1) {
    if (Ns == 1)
      any_stage_sym<DIR, false>(in, out, tw, n, R, Ns, tid, nthr);
    else
      any_stage_sym<DIR, true>(in, out, tw, n, R, Ns, tid, nthr);
  } else if (R == 4) {
    any_stage_bfly<DIR, 4>(in, out, tw, n, Ns, tid, nthr);
  } else if (R == 2) {
    any_stage_bfly<DIR, 2>(in, out, tw, n, Ns, tid, nthr);
  } else {
    any_stage<DIR>(in, out, tw, n, R, Ns, tid, nthr);
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
    any_stage_auto<DIR>(in, out, tw, pl.n, R, Ns, tid, nthr);
    __syncthreads();
    Ns *= R;
    double2* t = in;
    in = out;
    out = t;
  }
  return in;
}
#endif
