// Host emulation of the tile FFT (csrc/fft_core.cuh): checks every plan against a naive DFT.
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include tools/test_fft_core.cpp -o /tmp/test_fft_core && /tmp/test_fft_core
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../spectra_without_windows_b200/csrc/fft_core.cuh"

template <int N, int R, int DIR>
void run_stage(std::vector<double2>& tile, int p, const std::vector<double2>& tw) {
  const int NT = N / 2;
  const int items = (N / R) * FFT_TZ;
  const int ipt = items / NT;
  std::vector<double2> regs((size_t)NT * ipt * R);
  for (int tid = 0; tid < NT; ++tid)
    for (int s = 0; s < ipt; ++s) {
      int w = tid + s * NT;
      stage_load<R, DIR>(tile.data(), N, p, w >> 3, w & 7, tw.data(), &regs[((size_t)tid * ipt + s) * R]);
    }
  for (int tid = 0; tid < NT; ++tid)
    for (int s = 0; s < ipt; ++s) {
      int w = tid + s * NT;
      stage_store<R>(tile.data(), N, p, w >> 3, w & 7, &regs[((size_t)tid * ipt + s) * R]);
    }
}

template <int N, int DIR>
double test() {
  std::vector<double2> tw(N);
  for (int j = 0; j < N; ++j) tw[j] = double2{cos(-2.0 * M_PI * j / N), sin(-2.0 * M_PI * j / N)};
  std::vector<double2> tile((size_t)N * FFT_LDT), ref((size_t)N * FFT_TZ), in((size_t)N * FFT_TZ);
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < FFT_TZ; ++c) {
      double2 v{drand48() - 0.5, drand48() - 0.5};
      tile[i * FFT_LDT + c] = v;
      in[i * FFT_TZ + c] = v;
    }
  for (int c = 0; c < FFT_TZ; ++c)
    for (int k = 0; k < N; ++k) {
      double sr = 0, si = 0;
      for (int n = 0; n < N; ++n) {
        double ang = -DIR * 2.0 * M_PI * (double)((long long)k * n % N) / N;
        double cr = cos(ang), ci = sin(ang);
        double2 v = in[n * FFT_TZ + c];
        sr += v.x * cr - v.y * ci;
        si += v.x * ci + v.y * cr;
      }
      ref[k * FFT_TZ + c] = double2{sr, si};
    }
  using P = FftPlan<N>;
  int p = 1;
  for (int s = 0; s < P::NS; ++s) {
    int R = P::R[s];
    if (R == 16) run_stage<N, 16, DIR>(tile, p, tw);
    else if (R == 8) run_stage<N, 8, DIR>(tile, p, tw);
    else if (R == 4) run_stage<N, 4, DIR>(tile, p, tw);
    else run_stage<N, 2, DIR>(tile, p, tw);
    p *= R;
  }
  double err = 0, mx = 0;
  for (int k = 0; k < N; ++k)
    for (int c = 0; c < FFT_TZ; ++c) {
      double2 a = tile[k * FFT_LDT + c], b = ref[k * FFT_TZ + c];
      err = fmax(err, fmax(fabs(a.x - b.x), fabs(a.y - b.y)));
      mx = fmax(mx, fmax(fabs(b.x), fabs(b.y)));
    }
  return err / mx;
}

int main() {
  double worst = 0;
#define T(N) { double a = test<N, 1>(), b = test<N, -1>(); printf("N=%4d fwd %.2e inv %.2e\n", N, a, b); worst = fmax(worst, fmax(a, b)); }
  T(64) T(128) T(256) T(512) T(1024)
  if (worst > 1e-13) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
