// 3-D FFT engine of libsww.  Backend 0: cuFFT plans whose work area lives in the caller's
// workspace.  Backend 1 (power-of-two meshes): hand-written sm_100a passes, see fft_sm100.cu.
#include "sww_internal.cuh"


static int make_plan(cufftHandle* plan, const sww_ctx* c, cufftType type, size_t* ws) {
  long long n[3] = {c->g.nx, c->g.ny, c->g.nz};
  SWW_CUFFT(cufftCreate(plan));
  SWW_CUFFT(cufftSetAutoAllocation(*plan, 0));
  SWW_CUFFT(cufftMakePlanMany64(*plan, 3, n, nullptr, 1, 0, nullptr, 1, 0, type, 1, ws));
  return 0;
}

int sww_fft_init(sww_ctx* c) {
  size_t w1 = 0, w2 = 0;
  SWW_TRY(make_plan(&c->plan_d2z, c, CUFFT_D2Z, &w1));
  SWW_TRY(make_plan(&c->plan_z2d, c, CUFFT_Z2D, &w2));
  c->fft_work_bytes = w1 > w2 ? w1 : w2;
  size_t w3 = sww_fft_sm100_work_bytes(c);
  if (w3 > c->fft_work_bytes) c->fft_work_bytes = w3;
  return 0;
}

static int ensure_z2z(sww_ctx* c) {
  // ft/ift mirror only (not on the hot path): this plan lets cuFFT allocate its own work area.
  if (c->have_z2z) return 0;
  long long n[3] = {c->g.nx, c->g.ny, c->g.nz};
  size_t w = 0;
  SWW_CUFFT(cufftCreate(&c->plan_z2z));
  SWW_CUFFT(cufftMakePlanMany64(c->plan_z2z, 3, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2Z, 1, &w));
  SWW_CUFFT(cufftSetStream(c->plan_z2z, c->stream));
  c->have_z2z = true;
  return 0;
}

void sww_fft_destroy(sww_ctx* c) {
  if (c->plan_d2z) cufftDestroy(c->plan_d2z);
  if (c->plan_z2d) cufftDestroy(c->plan_z2d);
  if (c->have_z2z) cufftDestroy(c->plan_z2z);
  sww_fft_sm100_destroy(c);
}

int sww_fft_bind_work(sww_ctx* c) {
  if (!c->plan_d2z) return 0;   // row-mesh child: hand-written passes only
  SWW_CUFFT(cufftSetWorkArea(c->plan_d2z, c->fft_work));
  SWW_CUFFT(cufftSetWorkArea(c->plan_z2d, c->fft_work));
  SWW_CUFFT(cufftSetStream(c->plan_d2z, c->stream));
  SWW_CUFFT(cufftSetStream(c->plan_z2d, c->stream));
  if (c->have_z2z) SWW_CUFFT(cufftSetStream(c->plan_z2z, c->stream));
  return 0;
}

int sww_r2c(sww_ctx* c, const double* in, double2* out) {
  SWW_REQUIRE(c->fft_work || c->fft_work_bytes == 0, "no workspace bound (sww_bind_workspace)");
  c->n_fft++;
  if (c->fft_backend == 1) return sww_fft_sm100_r2c(c, in, out);
  SWW_CUFFT(cufftExecD2Z(c->plan_d2z, (cufftDoubleReal*)in, (cufftDoubleComplex*)out));
  return 0;
}

int sww_c2r(sww_ctx* c, double2* in, double* out) {
  SWW_REQUIRE(c->fft_work || c->fft_work_bytes == 0, "no workspace bound (sww_bind_workspace)");
  c->n_fft++;
  if (c->fft_backend == 1) return sww_fft_sm100_c2r(c, in, out);
  SWW_CUFFT(cufftExecZ2D(c->plan_z2d, (cufftDoubleComplex*)in, (cufftDoubleReal*)out));
  return 0;
}

int sww_c2c(sww_ctx* c, const double2* in, double2* out, int inverse) {
  c->n_fft++;
  if (c->fft_backend == 1 && c->fft_work && sww_fft_sm100_c2c_supported(c)) {
    // hand-written passes: complex z pass + the two column passes over the full box
    SWW_TRY(sww_fft_sm100_c2c(c, in, out, inverse));
    if (inverse) SWW_TRY(k_scale_c(c, out, 1.0 / (double)c->g.N, c->g.N));
    return 0;
  }
  SWW_TRY(ensure_z2z(c));
  SWW_CUFFT(cufftExecZ2Z(c->plan_z2z, (cufftDoubleComplex*)in, (cufftDoubleComplex*)out,
                         inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
  if (inverse) SWW_TRY(k_scale_c(c, out, 1.0 / (double)c->g.N, c->g.N));
  return 0;
}
