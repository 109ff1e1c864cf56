// B_l(k1,k2,k3) pipelines of libsww.
#include "sww_internal.cuh"

size_t sww_bk_pipeline_bytes(sww_ctx* c, int what) { return 0; }

#define NOTYET(name) \
  sww_set_error(name ": not implemented yet"); \
  return -9;

extern "C" int sww_bk_gmaps(sww_ctx* c, sww_devptr a, int data, sww_devptr g_out, sww_devptr tg_out) { NOTYET("sww_bk_gmaps") }
extern "C" int sww_bk_fisher_pair(sww_ctx* c, sww_devptr aA, sww_devptr aB, sww_devptr delta, sww_devptr gA,
                                  sww_devptr tgA, sww_devptr gB, sww_devptr tgB, sww_devptr fish_out) { NOTYET("sww_bk_fisher_pair") }
extern "C" int sww_bk_delta(sww_ctx* c, sww_devptr delta_out) { NOTYET("sww_bk_delta") }
extern "C" int sww_bk_q3(sww_ctx* c, sww_devptr g, sww_devptr q_out) { NOTYET("sww_bk_q3") }
extern "C" int sww_bk_q1_accumulate(sww_ctx* c, sww_devptr g, sww_devptr g_mc, sww_devptr tg_mc, double inv_two_nmc,
                                    sww_devptr q_inout) { NOTYET("sww_bk_q1_accumulate") }
extern "C" int sww_gram_f64(sww_ctx* c, sww_devptr A, sww_devptr B, int32_t m, int32_t n, uint64_t N, int accumulate,
                            sww_devptr C) { NOTYET("sww_gram_f64") }
