// Host emulation of the any-length tile FFT (csrc/fft_any.cuh): every radix plan of the reference's production mesh sizes
// (and a sweep of small / prime / composite lengths) against a naive long-double DFT.
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include tools/test_fft_any.cpp -o /tmp/test_fft_any && /tmp/test_fft_any
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../spectra_without_windows_b200/csrc/fft_any.cuh"

template <int DIR>
double test(int n, int nthr) {
  AnyPlan pl = any_make_plan(n);
  int prod = 1;
  for (int s = 0; s < pl.ns; ++s) prod *= pl.R[s];
  if (prod != n) return 1.0;
  std::vector<double2> tw(n);
  for (int j = 0; j < n; ++j) {
    long double a = -2.0L * M_PIl * j / n;
    tw[j] = double2{(double)cosl(a), (double)sinl(a)};
  }
  std::vector<double2> b0((size_t)n * FFT_LDT), b1((size_t)n * FFT_LDT, double2{1e300, 1e300}), in((size_t)n * FFT_TZ);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < FFT_TZ; ++c) {
      double2 v{drand48() - 0.5, drand48() - 0.5};
      b0[i * FFT_LDT + c] = v;
      in[i * FFT_TZ + c] = v;
    }
  double2 *a = b0.data(), *b = b1.data();
  int Ns = 1;
  for (int s = 0; s < pl.ns; ++s) {
    for (int tid = 0; tid < nthr; ++tid) any_stage_auto<DIR>(a, b, tw.data(), n, pl.R[s], Ns, tid, nthr);
    Ns *= pl.R[s];
    std::swap(a, b);
  }
  double err = 0, mx = 0;
  for (int c = 0; c < FFT_TZ; c += 3)
    for (int k = 0; k < n; ++k) {
      long double sr = 0, si = 0;
      for (int j = 0; j < n; ++j) {
        long double ang = -DIR * 2.0L * M_PIl * (long double)((long long)k * j % n) / n;
        long double cr = cosl(ang), ci = sinl(ang);
        double2 v = in[j * FFT_TZ + c];
        sr += v.x * cr - v.y * ci;
        si += v.x * ci + v.y * cr;
      }
      double2 g = a[k * FFT_LDT + c];
      err = fmax(err, fmax(fabs(g.x - (double)sr), fabs(g.y - (double)si)));
      mx = fmax(mx, fmax(fabs((double)sr), fabs((double)si)));
    }
  return err / mx;
}

int main() {
  const int sizes[] = {2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 18, 20, 24, 30, 32, 49, 60, 64, 97, 100, 121, 128, 167, 172, 175, 182,
                       258, 274, 307, 330, 496, 512, 640};
  double worst = 0;
  for (int n : sizes) {
    AnyPlan pl = any_make_plan(n);
    double a = test<1>(n, 1), b = test<-1>(n, any_threads_for(pl));
    printf("n=%4d plan", n);
    for (int s = 0; s < pl.ns; ++s) printf(" %d", pl.R[s]);
    printf("  threads %d  fwd %.2e inv %.2e\n", any_threads_for(pl), a, b);
    worst = fmax(worst, fmax(a, b));
  }
  if (worst > 2e-14) { printf("FAIL %.3e\n", worst); return 1; }
  printf("OK\n");
  return 0;
}
