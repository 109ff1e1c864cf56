// FP64 tensor-core Gram contraction  C[i][j] = sum_k A[i][k] B[j][k]  over N mesh cells
// (the dense contraction of bk/compute_bk_randoms.py:489-497; also usable for the P_l Fisher at
// small meshes, pk/compute_pk_randoms.py:270-274).
//
// tcgen05 has no f64 kind, so on sm_100a the FP64 tensor pipe is reached through
// mma.sync.aligned.m8n8k4.f64 (DMMA).  Layout: A is [m][K], B is [n][K] with K (cells) contiguous,
// i.e. exactly the row-major A / column-major B operands of the instruction.  The K dimension is
// huge (1e7..1e9) and m,n are a few hundred, so the kernel is split-K: CTA (mt, nt, ks) reduces a
// 128x64 output tile over its slice of K with a 3-stage cp.async pipeline, partials are written to
// a scratch buffer and reduced in a fixed order (deterministic, no atomics).
#include "sww_internal.cuh"

#define TM 128
#define TN 64
#define KC 32
#define KPAD 4
#define STAGES 3
#define GTHREADS 256

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GramSmem {
  double A[STAGES][TM][KC + KPAD];
  double B[STAGES][TN][KC + KPAD];
};

template <bool VEC16>
__global__ void __launch_bounds__(GTHREADS, 1)
gram_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, int m, int n, long long K,
                 long long chunks_per_split, double* __restrict__ partial, int Mpad, int Npad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GramSmem& S = *reinterpret_cast<GramSmem*>(smem_raw);
  const int mt = blockIdx.x, nt = blockIdx.y, ks = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps, warp tile 32 x 32
  const long long total_chunks = (K + KC - 1) / KC;
  long long c_begin = (long long)ks * chunks_per_split;
  long long c_end = c_begin + chunks_per_split;
  if (c_end > total_chunks) c_end = total_chunks;
  const int row0 = mt * TM, col0 = nt * TN;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int stage, long long chunk) {
    const long long k0 = chunk * KC;
    if (VEC16) {
      // 16-byte pieces: (TM+TN) rows x KC/2 pieces
      for (int p = tid; p < (TM + TN) * (KC / 2); p += GTHREADS) {
        int r = p / (KC / 2), q = p - r * (KC / 2);
        long long k = k0 + 2 * q;
        if (r < TM) {
          int gi = row0 + r;
          bool ok = (gi < m) && (k + 1 < K);
          const double* src = ok ? A + (long long)gi * K + k : A;
          cp_async16(&S.A[stage][r][2 * q], src, ok);   // K even here: no odd tail element
        } else {
          int rr = r - TM, gj = col0 + rr;
          bool ok = (gj < n) && (k + 1 < K);
          const double* src = ok ? B + (long long)gj * K + k : B;
          cp_async16(&S.B[stage][rr][2 * q], src, ok);
        }
      }
    } else {
      for (int p = tid; p < (TM + TN) * KC; p += GTHREADS) {
        int r = p / KC, q = p - r * KC;
        long long k = k0 + q;
        if (r < TM) {
          int gi = row0 + r;
          bool ok = (gi < m) && (k < K);
          cp_async8(&S.A[stage][r][q], ok ? A + (long long)gi * K + k : A, ok);
        } else {
          int rr = r - TM, gj = col0 + rr;
          bool ok = (gj < n) && (k < K);
          cp_async8(&S.B[stage][rr][q], ok ? B + (long long)gj * K + k : B, ok);
        }
      }
    }
  };

  const long long nchunks = c_end > c_begin ? c_end - c_begin : 0;
  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, c_begin + s);
    cp_async_commit();
  }
  for (long long it = 0; it < nchunks; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // prefetch chunk it + STAGES-1 into the stage consumed at iteration it-1
    {
      long long nx = it + STAGES - 1;
      if (nx < nchunks) load_stage((int)(nx % STAGES), c_begin + nx);
      cp_async_commit();
    }
    const int st = (int)(it % STAGES);
    const int ar = wm * 32 + (lane >> 2), br = wn * 32 + (lane >> 2), kk = lane & 3;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = S.A[st][ar + 8 * i][k4 + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = S.B[st][br + 8 * j][k4 + kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  // write the partial tile: partial[ks][Mpad][Npad]
  double* P = partial + (long long)ks * Mpad * Npad;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = row0 + wm * 32 + 8 * i + (lane >> 2);
      int cidx = col0 + wn * 32 + 8 * j + 2 * (lane & 3);
      *reinterpret_cast<double2*>(&P[(long long)r * Npad + cidx]) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

__global__ void gram_reduce_kernel(const double* __restrict__ partial, int ksplit, int Mpad, int Npad, int m, int n,
                                   int accumulate, double* __restrict__ C, int ldc, int col0) {
  long long total = (long long)m * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
    double s = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) s += partial[((long long)ks * Mpad + i) * Npad + j];
    double* dst = &C[(long long)i * ldc + col0 + j];
    *dst = accumulate ? *dst + s : s;
  }
}

static double* g_partial = nullptr;
static size_t g_partial_bytes = 0;

static int gram_run(sww_ctx* c, const double* A, const double* B, int m, int n, long long K, int accumulate, double* C,
                    int ldc, int col0) {
  SWW_REQUIRE(m > 0 && n > 0 && K > 0, "gram: empty operand");
  SWW_PROF(c, "gram.dmma f64");
  int mtiles = (m + TM - 1) / TM, ntiles = (n + TN - 1) / TN;
  int Mpad = mtiles * TM, Npad = ntiles * TN;
  long long total_chunks = (K + KC - 1) / KC;
  int ksplit = (148 * 2 + mtiles * ntiles - 1) / (mtiles * ntiles);
  if (ksplit > total_chunks) ksplit = (int)total_chunks;
  if (ksplit < 1) ksplit = 1;
  long long cps = (total_chunks + ksplit - 1) / ksplit;
  ksplit = (int)((total_chunks + cps - 1) / cps);
  size_t need = (size_t)ksplit * Mpad * Npad * sizeof(double);
  if (need > g_partial_bytes) {
    if (g_partial) cudaFree(g_partial);
    SWW_CUDA(cudaMalloc(&g_partial, need));
    g_partial_bytes = need;
  }
  size_t smem = sizeof(GramSmem);
  bool vec16 = (K % 2 == 0) && (((uintptr_t)A % 16) == 0) && (((uintptr_t)B % 16) == 0);
  dim3 grid(mtiles, ntiles, ksplit);
  if (vec16) {
    SWW_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_dmma_kernel<true><<<grid, GTHREADS, smem, c->stream>>>(A, B, m, n, K, cps, g_partial, Mpad, Npad);
  } else {
    SWW_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_dmma_kernel<false><<<grid, GTHREADS, smem, c->stream>>>(A, B, m, n, K, cps, g_partial, Mpad, Npad);
  }
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  gram_reduce_kernel<<<sww_grid_for((long long)m * n, 256), 256, 0, c->stream>>>(g_partial, ksplit, Mpad, Npad, m, n,
                                                                                accumulate, C, ldc, col0);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

int sww_gram_upper(sww_ctx* c, const double* A, const double* B, int m, int n, long long K, int row_limit, double* C,
                   int ldc, int col0) {
  (void)row_limit;
  return gram_run(c, A, B, m, n, K, 0, C, ldc, col0);
}

extern "C" int sww_gram_f64(sww_ctx* c, sww_devptr A, sww_devptr B, int32_t m, int32_t n, uint64_t N, int accumulate,
                            sww_devptr C) {
  SWW_REQUIRE(c && A && B && C, "null argument");
  return gram_run(c, (const double*)A, (const double*)B, m, n, (long long)N, accumulate, (double*)C, n, 0);
}
