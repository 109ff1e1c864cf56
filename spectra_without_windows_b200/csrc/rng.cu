// Device-side reproduction of numpy's legacy Gaussian stream
//     np.random.seed(s); np.random.normal(loc=0, scale=sigma, size=N)        (pk/compute_pk_randoms.py:139-140)
// = MT19937 (init_genrand seeding) -> 53-bit doubles from word pairs -> polar Box-Muller with rejection,
// second variate cached.  The host draws 5.4 s per 512^3 field on one core (SURVEY.md finding 5); this
// removes the host from the Monte-Carlo loop.
//
//   * mt_words_kernel: ONE CTA walks the recurrence (it is inherently sequential in blocks of 624
//     words, but 227 words of a block are independent: three sub-steps per block) and streams the
//     tempered words to global memory;
//   * polar attempts are aligned to groups of 4 words whatever their outcome, so acceptance is
//     evaluated in parallel, ranked with a prefix sum, and accepted pairs are written in stream order
//     [f*x2, f*x1] exactly as legacy_gauss caches them.
// The uniform stream and the accept/reject decisions are bit-exact; the Gaussian values agree with
// glibc's to the last bit or two (CUDA's log/sqrt are IEEE-accurate to <= 1 ulp).
#include "mt_jump_tab.h"
#include "sww_internal.cuh"

#define MT_N 624
#define MT_M 397
#define MT_DEG 19937

__device__ __forceinline__ unsigned mt_mix(unsigned a, unsigned b, unsigned m) {
  unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// state: 624 words in global memory (updated in place); writes nblocks*624 tempered words.
// One CTA per stream of a batch (blockIdx.x): independent seeds advance side by side on separate SMs.
__global__ void __launch_bounds__(256) mt_words_kernel(unsigned* __restrict__ state, unsigned* __restrict__ out, long long nblocks,
                                                       long long word_stride) {
  __shared__ unsigned mt[MT_N];
  const int t = threadIdx.x;
  state += (long long)blockIdx.x * MT_N;
  out += (long long)blockIdx.x * word_stride;
  for (int i = t; i < MT_N; i += 256) mt[i] = state[i];
  __syncthreads();
  for (long long b = 0; b < nblocks; ++b) {
    unsigned v = 0;
    // words 0..226: depend on old words only
    if (t < 227) v = mt_mix(mt[t], mt[t + 1], mt[t + MT_M]);
    __syncthreads();
    if (t < 227) mt[t] = v;
    __syncthreads();
    // words 227..453: mt[k+397-624] = mt[k-227] is new
    if (t < 227) v = mt_mix(mt[t + 227], mt[t + 228], mt[t]);
    __syncthreads();
    if (t < 227) mt[t + 227] = v;
    __syncthreads();
    // words 454..623
    if (t < 170) {
      int k = t + 454;
      v = mt_mix(mt[k], mt[(k + 1) % MT_N], mt[k - 227]);
    }
    __syncthreads();
    if (t < 170) mt[t + 454] = v;
    __syncthreads();
    unsigned* o = out + b * MT_N;
    for (int i = t; i < MT_N; i += 256) o[i] = mt_temper(mt[i]);
  }
  __syncthreads();
  for (int i = t; i < MT_N; i += 256) state[i] = mt[i];
}

// ------------------------------------------------------------------------------------------
// Jump-ahead (Haramoto et al. 2008): ONE legacy stream split over many CTAs.  Sub-stream u of a field starts u * L0 words
// into the stream (L0 = MT_JUMP_SUB_BLOCKS * 624); its 624-word window is g(F) applied to the seeded window, with
// g = x^(u L0) mod phi and F the one-word step.  With Q_k = x^(2^k L0) mod phi from the generated table
// (tools/mt_jump.py -> mt_jump_tab.h), the windows of all sub-streams follow in ceil(log2 U) rounds: round k multiplies
// the windows whose index has bit k set by Q_k.  Because F is linear, q(F) w = XOR over the set bits i of q of F^i w, and
// F^i w is just the window of w's OWN sequence at offset i: the CTA extends the window by 19937 words in shared memory and
// every thread j accumulates h[j] = XOR_i q_i x[i + j] -- no sequential Horner chain.
// ------------------------------------------------------------------------------------------
#define MT_JUMP_X (MT_N + MT_DEG + 3)
__global__ void mt_replicate_kernel(const unsigned* __restrict__ key, unsigned* __restrict__ windows, int U) {
  const unsigned* k = key + (long long)blockIdx.y * MT_N;
  unsigned* w = windows + ((long long)blockIdx.y * U + blockIdx.x) * MT_N;
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) w[i] = k[i];
}

__global__ void __launch_bounds__(640) mt_jump_round_kernel(unsigned* __restrict__ windows, int U, int k,
                                                            const unsigned* __restrict__ polys) {
  const int u = blockIdx.x;
  if (!((u >> k) & 1)) return;
  extern __shared__ unsigned mt_sm[];
  unsigned* x = mt_sm;                 // [MT_JUMP_X] the window and the next 19937 words of its sequence
  unsigned* q = mt_sm + MT_JUMP_X;     // [624] bits of Q_k
  unsigned* w = windows + ((long long)blockIdx.y * U + u) * MT_N;
  const int t = threadIdx.x;
  for (int i = t; i < MT_N; i += 640) {
    x[i] = w[i];
    q[i] = polys[k * MT_N + i];
  }
  __syncthreads();
  // x[n + 624] = mix(x[n], x[n + 1], x[n + 397]): 227 consecutive n only depend on words already there
  for (int base = 0; base < MT_DEG; base += 227) {
    const int n = base + t;
    if (t < 227 && n < MT_DEG) x[n + MT_N] = mt_mix(x[n], x[n + 1], x[n + MT_M]);
    __syncthreads();
  }
  if (t < MT_N) {
    unsigned acc = 0u;
    for (int wi = 0; wi < MT_N; ++wi) {
      const unsigned bits = q[wi];
      if (bits == 0u) continue;        // uniform: q is shared by the CTA
      const unsigned* xp = x + 32 * wi + t;
#pragma unroll
      for (int b = 0; b < 32; ++b) acc ^= xp[b] & (0u - ((bits >> b) & 1u));
    }
    w[t] = acc;
  }
}

// sub-stream u = u0 + blockIdx.x of seed blockIdx.y: MT_JUMP_SUB_BLOCKS blocks of tempered words from its window
__global__ void __launch_bounds__(256) mt_words_sub_kernel(const unsigned* __restrict__ windows, int U, int u0,
                                                           unsigned* __restrict__ out, long long word_stride) {
  __shared__ unsigned mt[MT_N];
  const int t = threadIdx.x;
  const unsigned* w = windows + ((long long)blockIdx.y * U + u0 + blockIdx.x) * MT_N;
  out += (long long)blockIdx.y * word_stride + (long long)blockIdx.x * MT_JUMP_SUB_BLOCKS * MT_N;
  for (int i = t; i < MT_N; i += 256) mt[i] = w[i];
  __syncthreads();
  for (int b = 0; b < MT_JUMP_SUB_BLOCKS; ++b) {
    unsigned v = 0;
    if (t < 227) v = mt_mix(mt[t], mt[t + 1], mt[t + MT_M]);
    __syncthreads();
    if (t < 227) mt[t] = v;
    __syncthreads();
    if (t < 227) v = mt_mix(mt[t + 227], mt[t + 228], mt[t]);
    __syncthreads();
    if (t < 227) mt[t + 227] = v;
    __syncthreads();
    if (t < 170) {
      int k = t + 454;
      v = mt_mix(mt[k], mt[(k + 1) % MT_N], mt[k - 227]);
    }
    __syncthreads();
    if (t < 170) mt[t + 454] = v;
    __syncthreads();
    unsigned* o = out + (long long)b * MT_N;
    for (int i = t; i < MT_N; i += 256) o[i] = mt_temper(mt[i]);
  }
}

__device__ __forceinline__ bool polar_attempt(const unsigned* __restrict__ w, long long i, double& x1, double& x2, double& r2) {
  const uint4 q = *reinterpret_cast<const uint4*>(w + 4 * i);
  const double d1 = ((double)(q.x >> 5) * 67108864.0 + (double)(q.y >> 6)) / 9007199254740992.0;
  const double d2 = ((double)(q.z >> 5) * 67108864.0 + (double)(q.w >> 6)) / 9007199254740992.0;
  x1 = __dsub_rn(__dmul_rn(2.0, d1), 1.0);
  x2 = __dsub_rn(__dmul_rn(2.0, d2), 1.0);
  r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
  return !(r2 >= 1.0 || r2 == 0.0);
}

#define PA_BLOCK 1024
// pass 1: accepted attempts per block of 1024 attempts
__global__ void __launch_bounds__(PA_BLOCK) polar_count_kernel(const unsigned* __restrict__ w, long long nattempts,
                                                                unsigned* __restrict__ block_counts, long long word_stride,
                                                                long long count_stride) {
  __shared__ unsigned wc[PA_BLOCK / 32];
  w += (long long)blockIdx.y * word_stride;
  block_counts += (long long)blockIdx.y * count_stride;
  const long long i = (long long)blockIdx.x * PA_BLOCK + threadIdx.x;
  double x1, x2, r2;
  bool ok = (i < nattempts) && polar_attempt(w, i, x1, x2, r2);
  unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int k = 0; k < PA_BLOCK / 32; ++k) s += wc[k];
    block_counts[blockIdx.x] = s;
  }
}

// pass 2: exclusive scan of the block counts (single CTA), total appended at [nblocks]
// offsets[nblocks] = accepted pairs of this chunk; *running = pairs emitted by earlier chunks (read by the emit
// pass through offsets[nblocks+1], then advanced)
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned* __restrict__ counts, long long* __restrict__ offsets, int nblocks,
                                                            long long* __restrict__ running, long long count_stride,
                                                            long long offset_stride) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  counts += (long long)blockIdx.x * count_stride;
  offsets += (long long)blockIdx.x * offset_stride;
  running += blockIdx.x;
  const int per = (nblocks + 1023) / 1024;
  long long s = 0;
  for (int k = 0; k < per; ++k) {
    int i = t * per + k;
    if (i < nblocks) s += counts[i];
  }
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int k = 0; k < 1024; ++k) {
      long long v = part[k];
      part[k] = run;
      run += v;
    }
    offsets[nblocks] = run;
    offsets[nblocks + 1] = *running;
    *running += run;
  }
  __syncthreads();
  long long run = part[t];
  for (int k = 0; k < per; ++k) {
    int i = t * per + k;
    if (i < nblocks) {
      offsets[i] = run;
      run += counts[i];
    }
  }
}

// pass 3: accepted attempt with stream rank j writes out[2j] = sigma*f*x2, out[2j+1] = sigma*f*x1
#define RNG_MAX_BATCH 8
struct RngOuts {
  double* p[RNG_MAX_BATCH];
};
__global__ void __launch_bounds__(PA_BLOCK) polar_emit_kernel(const unsigned* __restrict__ w, long long nattempts,
                                                               const long long* __restrict__ offsets, int nblocks,
                                                               const double* __restrict__ sigma, RngOuts outs, long long nout,
                                                               long long word_stride, long long offset_stride) {
  __shared__ unsigned wc[PA_BLOCK / 32];
  w += (long long)blockIdx.y * word_stride;
  offsets += (long long)blockIdx.y * offset_stride;
  double* __restrict__ out = outs.p[blockIdx.y];
  const long long i = (long long)blockIdx.x * PA_BLOCK + threadIdx.x;
  double x1 = 0, x2 = 0, r2 = 1;
  bool ok = (i < nattempts) && polar_attempt(w, i, x1, x2, r2);
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wc[warp] = __popc(m);
  __syncthreads();
  unsigned before = 0;
  for (int k = 0; k < warp; ++k) before += wc[k];
  if (!ok) return;
  const long long j = offsets[nblocks + 1] + offsets[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
  const double f = sqrt(-2.0 * log(r2) / r2);
  const long long o = 2 * j;
  if (o < nout) out[o] = 0.0 + sigma[o] * (f * x2);
  if (o + 1 < nout) out[o + 1] = 0.0 + sigma[o + 1] * (f * x1);
}

// asynchronous shortfall guard: sticky device flag, read by sww_grf_check
__global__ void rng_guard_kernel(const long long* __restrict__ running, long long need_pairs, int nseeds, int* __restrict__ flag) {
  if (threadIdx.x < nseeds && running[threadIdx.x] < need_pairs) *flag = 1;
}

struct RngScratch {
  int* flag = nullptr;
  unsigned* state = nullptr;
  unsigned* polys = nullptr;     // device copy of MT_JUMP_POLY
  unsigned* windows = nullptr;   // [batch][U][624] sub-stream windows of the current fields
  long long cap_windows = 0;
  unsigned* words = nullptr;
  unsigned* counts = nullptr;
  long long* offsets = nullptr;
  long long* running = nullptr;
  long long cap_attempts = 0;
  int cap_batch = 0;
  unsigned h_state[8 * MT_N];   // host copy of the seeded states (source of an asynchronous upload)
  long long word_stride = 0, count_stride = 0, offset_stride = 0;
};

static RngScratch& rng_of(sww_ctx* c) {
  if (!c->rng) c->rng = new RngScratch();
  return *(RngScratch*)c->rng;
}

void sww_rng_destroy(sww_ctx* c) {
  RngScratch* r = (RngScratch*)c->rng;
  if (!r) return;
  cudaFree(r->flag);
  cudaFree(r->state);
  cudaFree(r->polys);
  cudaFree(r->windows);
  cudaFree(r->words);
  cudaFree(r->counts);
  cudaFree(r->offsets);
  cudaFree(r->running);
  delete r;
  c->rng = nullptr;
}

static int rng_reserve(RngScratch& g_rng, long long attempts, int batch) {
  if (attempts <= g_rng.cap_attempts && batch <= g_rng.cap_batch) return 0;
  if (g_rng.words) {
    cudaFree(g_rng.words);
    cudaFree(g_rng.counts);
    cudaFree(g_rng.offsets);
  }
  if (!g_rng.state) {
    SWW_CUDA(cudaMalloc(&g_rng.state, RNG_MAX_BATCH * MT_N * sizeof(unsigned)));
    SWW_CUDA(cudaMalloc(&g_rng.running, RNG_MAX_BATCH * sizeof(long long)));
    SWW_CUDA(cudaMalloc(&g_rng.flag, sizeof(int)));
    SWW_CUDA(cudaMemset(g_rng.flag, 0, sizeof(int)));
    SWW_CUDA(cudaMalloc(&g_rng.polys, sizeof(MT_JUMP_POLY)));
    SWW_CUDA(cudaMemcpy(g_rng.polys, MT_JUMP_POLY, sizeof(MT_JUMP_POLY), cudaMemcpyHostToDevice));
  }
  if (batch < g_rng.cap_batch) batch = g_rng.cap_batch;
  long long nb = (attempts + PA_BLOCK - 1) / PA_BLOCK;
  g_rng.word_stride = attempts * 4 + MT_N;
  g_rng.count_stride = nb + 1;
  g_rng.offset_stride = nb + 2;
  SWW_CUDA(cudaMalloc(&g_rng.words, (size_t)g_rng.word_stride * batch * sizeof(unsigned)));
  SWW_CUDA(cudaMalloc(&g_rng.counts, (size_t)g_rng.count_stride * batch * sizeof(unsigned)));
  SWW_CUDA(cudaMalloc(&g_rng.offsets, (size_t)g_rng.offset_stride * batch * sizeof(long long)));
  g_rng.cap_attempts = attempts;
  g_rng.cap_batch = batch;
  return 0;
}

// outs[s][i] = sigma[i] * gauss_i(seed s), i < n, with the numpy legacy streams of `seeds`; the nseeds recurrences
// (one CTA each, inherently sequential) advance side by side, so a batch costs the time of one field
extern "C" int sww_grf_legacy_batch(sww_ctx* c, const uint32_t* seeds, int nseeds, sww_devptr sigma_, const sww_devptr* outs_,
                                    uint64_t n, sww_stream stream_) {
  SWW_REQUIRE(c && seeds && sigma_ && outs_ && n > 0, "null argument");
  SWW_REQUIRE(nseeds >= 1 && nseeds <= RNG_MAX_BATCH, "grf batch of %d seeds (1..%d)", nseeds, RNG_MAX_BATCH);
  // runs on `stream_` (0: the context stream) and returns when the fields are complete on that stream; the
  // generator occupies one SM per seed, so a side stream overlaps it with the Fisher kernels of another field
  struct StreamSwap {
    sww_ctx* c;
    cudaStream_t old;
    StreamSwap(sww_ctx* c_, cudaStream_t s) : c(c_), old(c_->stream) { if (s) c->stream = s; }
    ~StreamSwap() { c->stream = old; }
  } swap_(c, (cudaStream_t)stream_);
  SWW_PROF(c, "rng.legacy gaussian field");
  const double* sigma = (const double*)sigma_;
  RngOuts outs = {};
  for (int b = 0; b < nseeds; ++b) {
    SWW_REQUIRE(outs_[b], "null output pointer");
    outs.p[b] = (double*)outs_[b];
  }
  // chunks of 624*2^14 words = 2.56 M attempts (a whole number of MT blocks)
  const long long blocks_per_chunk = 4LL << 14;                 // x624 words, divisible by 4 words per attempt
  const long long att_per_chunk = blocks_per_chunk * MT_N / 4;
  RngScratch& g_rng = rng_of(c);
  SWW_TRY(rng_reserve(g_rng, att_per_chunk, nseeds));
  // init_genrand seeding (numpy mt19937_seed)
  unsigned* h = g_rng.h_state;
  for (int b = 0; b < nseeds; ++b) {
    unsigned s = seeds[b];
    for (int pos = 0; pos < MT_N; ++pos) {
      h[b * MT_N + pos] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (unsigned)pos + 1u;
    }
  }
  SWW_CUDA(cudaMemcpyAsync(g_rng.state, h, sizeof(unsigned) * MT_N * nseeds, cudaMemcpyHostToDevice, c->stream));   // pageable source: staged before return
  SWW_CUDA(cudaMemsetAsync(g_rng.running, 0, sizeof(long long) * RNG_MAX_BATCH, c->stream));
  const long long need_pairs = ((long long)n + 1) / 2;
  const int nb = (int)((att_per_chunk + PA_BLOCK - 1) / PA_BLOCK);
  // accepted pairs per chunk ~ Binomial(A, pi/4): plan the chunk count 8 sigma on the safe side, no read-back per chunk
  const double p = 0.78539816339744831, sd = sqrt((double)att_per_chunk * p * (1.0 - p));
  long long chunks = (long long)ceil((double)need_pairs / ((double)att_per_chunk * p - 8.0 * sd));
  long long done_pairs = 0;
  const bool async = stream_ != 0;   // side stream: no host synchronisation, the shortfall guard is a device flag (sww_grf_check)
  // jump-ahead: every chunk is generated by `subs` CTAs per seed instead of one (SWW_RNG_JUMP=0: the sequential walker)
  const char* ej = getenv("SWW_RNG_JUMP");
  const int subs = (int)(blocks_per_chunk / MT_JUMP_SUB_BLOCKS);
  const long long spare = async ? 0 : 1;                        // the synchronous path may ask for one more chunk
  const long long U = (chunks + spare) * subs;
  const bool jump = !(ej && ej[0] == '0') && U < (1LL << MT_JUMP_K);
  if (jump) {
    if (g_rng.cap_windows < U * nseeds) {
      if (g_rng.windows) cudaFree(g_rng.windows);
      g_rng.windows = nullptr;
      g_rng.cap_windows = 0;
      SWW_CUDA(cudaMalloc(&g_rng.windows, (size_t)U * RNG_MAX_BATCH * MT_N * sizeof(unsigned)));
      g_rng.cap_windows = U * RNG_MAX_BATCH;
    }
    const size_t jsm = (size_t)(MT_JUMP_X + MT_N) * sizeof(unsigned);
    SWW_CUDA(cudaFuncSetAttribute(mt_jump_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
    mt_replicate_kernel<<<dim3((unsigned)U, nseeds), 256, 0, c->stream>>>(g_rng.state, g_rng.windows, (int)U);
    c->n_own++;
    for (int k = 0; (1LL << k) < U; ++k) {
      mt_jump_round_kernel<<<dim3((unsigned)U, nseeds), 640, jsm, c->stream>>>(g_rng.windows, (int)U, k, g_rng.polys);
      c->n_own++;
    }
    SWW_CUDA(cudaGetLastError());
  }
  long long chunk_no = 0;
  while (done_pairs < need_pairs) {
    for (long long k = 0; k < chunks; ++k, ++chunk_no) {
      if (jump) {
        SWW_REQUIRE((chunk_no + 1) * subs <= U, "device RNG: more chunks needed than sub-stream windows were prepared");
        mt_words_sub_kernel<<<dim3(subs, nseeds), 256, 0, c->stream>>>(g_rng.windows, (int)U, (int)(chunk_no * subs), g_rng.words,
                                                                      g_rng.word_stride);
      } else {
        mt_words_kernel<<<nseeds, 256, 0, c->stream>>>(g_rng.state, g_rng.words, blocks_per_chunk, g_rng.word_stride);
      }
      polar_count_kernel<<<dim3(nb, nseeds), PA_BLOCK, 0, c->stream>>>(g_rng.words, att_per_chunk, g_rng.counts, g_rng.word_stride,
                                                                        g_rng.count_stride);
      scan_counts_kernel<<<nseeds, 1024, 0, c->stream>>>(g_rng.counts, g_rng.offsets, nb, g_rng.running, g_rng.count_stride,
                                                          g_rng.offset_stride);
      polar_emit_kernel<<<dim3(nb, nseeds), PA_BLOCK, 0, c->stream>>>(g_rng.words, att_per_chunk, g_rng.offsets, nb, sigma, outs,
                                                                       (long long)n, g_rng.word_stride, g_rng.offset_stride);
      c->n_own += 4;
    }
    SWW_CUDA(cudaGetLastError());
    if (async) {
      rng_guard_kernel<<<1, 32, 0, c->stream>>>(g_rng.running, need_pairs, nseeds, g_rng.flag);
      c->n_own++;
      break;
    }
    // one read-back at the end guards the (astronomically unlikely) shortfall: the slowest stream decides
    long long run[RNG_MAX_BATCH];
    SWW_CUDA(cudaMemcpyAsync(run, g_rng.running, sizeof(long long) * nseeds, cudaMemcpyDeviceToHost, c->stream));
    SWW_CUDA(cudaStreamSynchronize(c->stream));
    done_pairs = run[0];
    for (int b = 1; b < nseeds; ++b) done_pairs = run[b] < done_pairs ? run[b] : done_pairs;
    chunks = 1;
  }
  return 0;
}

// synchronises the device and reports whether any asynchronous batch fell short of accepted pairs (probability ~1e-15
// per field with the 8-sigma chunk plan; the caller would regenerate on the context stream)
extern "C" int sww_grf_check(sww_ctx* c) {
  SWW_REQUIRE(c, "null context");
  if (!c->rng) return 0;
  RngScratch& g_rng = rng_of(c);
  if (!g_rng.flag) return 0;
  int f = 0;
  SWW_CUDA(cudaDeviceSynchronize());
  SWW_CUDA(cudaMemcpy(&f, g_rng.flag, sizeof(int), cudaMemcpyDeviceToHost));
  SWW_REQUIRE(f == 0, "device RNG: a field was left incomplete (too few accepted polar pairs); regenerate it synchronously");
  return 0;
}

extern "C" int sww_grf_legacy(sww_ctx* c, uint32_t seed, sww_devptr sigma_, sww_devptr out_, uint64_t n, sww_stream stream_) {
  SWW_REQUIRE(out_, "null argument");
  return sww_grf_legacy_batch(c, &seed, 1, sigma_, &out_, n, stream_);
}
