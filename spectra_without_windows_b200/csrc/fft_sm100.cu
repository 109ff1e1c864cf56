// Hand-written sm_100a 3-D FFT passes (power-of-two meshes).  Filled in after the cuFFT path.
#include "sww_internal.cuh"

int sww_fft_sm100_supported(const sww_ctx* c) { return 0; }
int sww_fft_sm100_r2c(sww_ctx* c, const double* in, double2* out) {
  sww_set_error("sm_100a FFT backend not available for this mesh");
  return -1;
}
int sww_fft_sm100_c2r(sww_ctx* c, double2* in, double* out) {
  sww_set_error("sm_100a FFT backend not available for this mesh");
  return -1;
}
