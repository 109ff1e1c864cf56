// Shared-memory FFT building blocks of the hand-written sm_100a passes (FP64).
//
// A tile is double2[N][LDT]: 8 independent 1-D sequences ("columns") of length N interleaved
// with a row pitch of LDT = 9 elements.  The 8 lanes of a quarter-warp always work on the 8
// columns of one row, so every 128-bit shared-memory access of the butterflies covers 8 distinct
// 16-byte bank groups (conflict-free), and the odd pitch makes the transposing loads/stores of the
// z pass conflict-free as well.
//
// The transform is an in-place Stockham autosort (natural order in, natural order out) with
// radix-8/4 stages: every thread loads all its inputs of a stage into registers, the CTA
// synchronises, every thread stores.  16 complex values per thread per stage (NT = N/2 threads).
//
// The same functions compile for the host (tools/test_fft_core.cpp emulates the CTA serially).
#pragma once
#include <vector_types.h>

#ifndef __CUDACC__
#define FFT_HD inline
#else
#define FFT_HD __host__ __device__ __forceinline__
#endif

#define FFT_TZ 8   // columns per tile
#define FFT_LDT 9  // row pitch of a tile (elements)

FFT_HD double2 c_add(double2 a, double2 b) { return double2{a.x + b.x, a.y + b.y}; }
FFT_HD double2 c_sub(double2 a, double2 b) { return double2{a.x - b.x, a.y - b.y}; }
FFT_HD double2 c_mul(double2 a, double2 b) { return double2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
FFT_HD double2 c_conj(double2 a) { return double2{a.x, -a.y}; }
// multiply by -i (forward) or +i (inverse)
template <int DIR>
FFT_HD double2 c_rot(double2 a) {
  return DIR > 0 ? double2{a.y, -a.x} : double2{-a.y, a.x};
}

// DIR = +1: forward (e^{-i...}), DIR = -1: inverse (e^{+i...}), both unnormalised.  Natural-order output.
template <int DIR>
FFT_HD void dft2(double2& a, double2& b) {
  double2 t = a;
  a = c_add(t, b);
  b = c_sub(t, b);
}

template <int DIR>
FFT_HD void dft4(double2* u) {
  // stage 1
  double2 a0 = c_add(u[0], u[2]), a1 = c_sub(u[0], u[2]);
  double2 a2 = c_add(u[1], u[3]), a3 = c_rot<DIR>(c_sub(u[1], u[3]));
  u[0] = c_add(a0, a2);
  u[2] = c_sub(a0, a2);
  u[1] = c_add(a1, a3);
  u[3] = c_sub(a1, a3);
}

template <int DIR>
FFT_HD void dft8(double2* u) {
  const double h = 0.70710678118654752440;
  // even / odd 4-point transforms
  double2 e[4] = {u[0], u[2], u[4], u[6]};
  double2 o[4] = {u[1], u[3], u[5], u[7]};
  dft4<DIR>(e);
  dft4<DIR>(o);
  // twiddles w8^k, k = 1,2,3  (forward: e^{-i pi k/4})
  double2 t1 = DIR > 0 ? double2{h * (o[1].x + o[1].y), h * (o[1].y - o[1].x)}
                       : double2{h * (o[1].x - o[1].y), h * (o[1].y + o[1].x)};
  double2 t2 = c_rot<DIR>(o[2]);
  double2 t3 = DIR > 0 ? double2{h * (-o[3].x + o[3].y), h * (-o[3].y - o[3].x)}
                       : double2{h * (-o[3].x - o[3].y), h * (-o[3].y + o[3].x)};
  u[0] = c_add(e[0], o[0]);
  u[4] = c_sub(e[0], o[0]);
  u[1] = c_add(e[1], t1);
  u[5] = c_sub(e[1], t1);
  u[2] = c_add(e[2], t2);
  u[6] = c_sub(e[2], t2);
  u[3] = c_add(e[3], t3);
  u[7] = c_sub(e[3], t3);
}

// 16-point DFT as 4 x 4 (decimation in time): A[j][q] = DFT4 over s of u[j + 4 s]; twiddle w16^(j q);
// X[q + 4 t] = DFT4 over j.
template <int DIR>
FFT_HD void dft16(double2* u) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
  double2 a[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double2 t[4] = {u[j], u[j + 4], u[j + 8], u[j + 12]};
    dft4<DIR>(t);
#pragma unroll
    for (int q = 0; q < 4; ++q) a[j][q] = t[q];
  }
  // w16^n = (cos(pi n/8), -DIR sin(pi n/8)), n = j*q in {1,2,3,4,6,9}
  const double cs[10] = {1.0, c1, h, s1, 0.0, -s1, -h, -c1, -1.0, -c1};
  const double sn[10] = {0.0, s1, h, c1, 1.0, c1, h, s1, 0.0, -s1};
#pragma unroll
  for (int j = 1; j < 4; ++j)
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const int n = j * q;
      if (n == 4) {   // w16^4 = -i (forward) / +i (inverse): a rotation, no arithmetic
        a[j][q] = c_rot<DIR>(a[j][q]);
      } else {
        double2 w = double2{cs[n], DIR > 0 ? -sn[n] : sn[n]};
        a[j][q] = c_mul(a[j][q], w);
      }
    }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double2 t[4] = {a[0][q], a[1][q], a[2][q], a[3][q]};
    dft4<DIR>(t);
#pragma unroll
    for (int s = 0; s < 4; ++s) u[q + 4 * s] = t[s];
  }
}

template <int R, int DIR>
FFT_HD void dftR(double2* u) {
  if (R == 16)
    dft16<DIR>(u);
  else if (R == 8)
    dft8<DIR>(u);
  else if (R == 4)
    dft4<DIR>(u);
  else
    dft2<DIR>(u[0], u[1]);
}

// Twiddles w^r, r < R, of one butterfly: w^1, w^2, w^4 (, w^8) come from the table, the others are
// products of two table values (<= 2 roundings), which costs less than the shared-memory loads.
template <int R, int DIR>
FFT_HD void stage_twiddle(int N, int p, int k, const double2* tw, double2* u) {
  if (p > 1) {
    const int step = k * (N / (p * R));
    double2 w[R];
#pragma unroll
    for (int r = 1; r < R; r <<= 1) {
      w[r] = tw[step * r];
      if (DIR < 0) w[r].y = -w[r].y;
    }
#pragma unroll
    for (int r = 3; r < R; ++r) {
      if ((r & (r - 1)) != 0) {  // not a power of two: highest set bit times the rest
        int hb = r >= 8 ? 8 : (r >= 4 ? 4 : 2);
        w[r] = c_mul(w[hb], w[r - hb]);
      }
    }
#pragma unroll
    for (int r = 1; r < R; ++r) u[r] = c_mul(u[r], w[r]);
  }
}

// One Stockham stage, radix R, for work item (butterfly bi, column c) of a tile [N][FFT_LDT].
//   p  : product of the radices of the previous stages (1 for the first stage)
//   tw : table W_N[j] = exp(-2 pi i j / N), j < N
template <int R, int DIR, int LDT = FFT_LDT>
FFT_HD void stage_load(const double2* tile, int N, int p, int bi, int c, const double2* tw, double2* u) {
  const int T = N / R;
  const int k = bi & (p - 1);
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = tile[(bi + r * T) * LDT + c];
  stage_twiddle<R, DIR>(N, p, k, tw, u);
  dftR<R, DIR>(u);
}

template <int R, int LDT = FFT_LDT>
FFT_HD void stage_store(double2* tile, int N, int p, int bi, int c, const double2* u) {
  (void)N;
  const int k = bi & (p - 1);
  const int j = (bi - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) tile[(j + r * p) * LDT + c] = u[r];
}

// Radix plan: N = R0*R1[*R2].
// The plan is palindromic-friendly: the reversed plan is used for the transform that follows a fused
// real-space step, so that the last stage of one and the first stage of the next share registers.
template <int N> struct FftPlan;
template <> struct FftPlan<64>   { static constexpr int NS = 2; static constexpr int R[4] = {8, 8, 1, 1}; };
template <> struct FftPlan<128>  { static constexpr int NS = 2; static constexpr int R[4] = {8, 16, 1, 1}; };
template <> struct FftPlan<256>  { static constexpr int NS = 2; static constexpr int R[4] = {16, 16, 1, 1}; };
template <> struct FftPlan<512>  { static constexpr int NS = 3; static constexpr int R[4] = {8, 8, 8, 1}; };
template <> struct FftPlan<1024> { static constexpr int NS = 3; static constexpr int R[4] = {4, 16, 16, 1}; };
