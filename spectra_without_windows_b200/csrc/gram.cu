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

// chunk length of the plain Gram: 16 keeps a 3-stage tile set at 90 KB, so two CTAs (16 warps) share an SM and one
// CTA's DMMAs cover the other's barrier / fragment-load bubbles
#ifndef KCG
#define KCG 32
#endif
#ifndef GRAM_CTAS
#define GRAM_CTAS 2
#endif
#ifndef STAGES_G
#define STAGES_G 2
#endif
struct GramSmem {
  double A[STAGES_G][TM][KCG + KPAD];
  double B[STAGES_G][TN][KCG + KPAD];
};

template <bool VEC16>
__global__ void __launch_bounds__(GTHREADS, GRAM_CTAS)
gram_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, int m, int n, long long K,
                 long long chunks_per_split, double* __restrict__ partial, int Mpad, int Npad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GramSmem& S = *reinterpret_cast<GramSmem*>(smem_raw);
  const int mt = blockIdx.x, nt = blockIdx.y, ks = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps, warp tile 32 x 32
  const long long total_chunks = (K + KCG - 1) / KCG;
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
    const long long k0 = chunk * KCG;
    if (VEC16) {
      // 16-byte pieces: (TM+TN) rows x KC/2 pieces
      for (int p = tid; p < (TM + TN) * (KCG / 2); p += GTHREADS) {
        int r = p / (KCG / 2), q = p - r * (KCG / 2);
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
      for (int p = tid; p < (TM + TN) * KCG; p += GTHREADS) {
        int r = p / KCG, q = p - r * KCG;
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
  for (int s = 0; s < STAGES_G - 1; ++s) {
    if (s < nchunks) load_stage(s, c_begin + s);
    cp_async_commit();
  }
  for (long long it = 0; it < nchunks; ++it) {
    cp_async_wait<STAGES_G - 2>();
    __syncthreads();
    // prefetch chunk it + STAGES-1 into the stage consumed at iteration it-1
    {
      long long nx = it + STAGES_G - 1;
      if (nx < nchunks) load_stage((int)(nx % STAGES_G), c_begin + nx);
      cp_async_commit();
    }
    const int st = (int)(it % STAGES_G);
    const int ar = wm * 32 + (lane >> 2), br = wn * 32 + (lane >> 2), kk = lane & 3;
#pragma unroll
    for (int k4 = 0; k4 < KCG; k4 += 4) {
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

// ------------------------------------------------------------------------------------------
// Trilinear Gram   C[i*nj + j][c] = sum_k X[i][k] Y[j][k] Z[c][k]
// (the 1-field term of bk/compute_bk_data.py:380-403 is six such contractions per MC simulation).
// Same tiling as above, but the A operand is never materialised: the fragment is the product of two
// shared-memory rows, formed in registers right before the DMMA.
// ------------------------------------------------------------------------------------------
#define G3_XROWS 36
#define G3_YROWS 32
#define G3_ZROWS 40
struct Gram3Stage {
  double X[G3_XROWS][KC + KPAD];
  double Y[G3_YROWS][KC + KPAD];
  double Z[G3_ZROWS][KC + KPAD];
};
struct Gram3Smem {
  Gram3Stage s[STAGES];
};

// Tile: 128 (i,j) rows x 8 NJ columns.  Warps 0-3 and 4-7 cover the same four 32-row slabs and split each
// K chunk in halves (the column count is small, so the second warp column of the plain Gram would idle);
// the two halves are summed through shared memory at the end.
#ifndef GRAM3_CTAS
#define GRAM3_CTAS 2
#endif
template <int NJ>
__global__ void __launch_bounds__(GTHREADS, GRAM3_CTAS)
gram3_dmma_kernel(const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z, int ni, int nj,
                  int nz, long long K, long long chunks_per_split, double* __restrict__ partial, int Mpad, int Npad) {
  constexpr int TN3 = 8 * NJ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Gram3Smem& S = *reinterpret_cast<Gram3Smem*>(smem_raw);
  const int mt = blockIdx.x, nt = blockIdx.y, ks = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wk = warp >> 2;
  const long long total_chunks = (K + KC - 1) / KC;
  long long c_begin = (long long)ks * chunks_per_split;
  long long c_end = c_begin + chunks_per_split;
  if (c_end > total_chunks) c_end = total_chunks;
  const int row0 = mt * TM, col0 = nt * TN3;
  const int m = ni * nj;
  const int i_first = row0 / nj;                       // first X row this tile needs
  int i_last = (row0 + TM - 1) / nj;
  if (i_last > ni - 1) i_last = ni - 1;
  const int nxr = i_last - i_first + 1;                // <= G3_XROWS (checked on the host)

  double acc[4][NJ][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // load slots: every thread copies the same (row, 16-byte piece) of each chunk, so the source pointer and the
  // shared-memory offset are computed once; per chunk only the K offset moves.  K is even (checked on the host).
  constexpr int PITCH = KC + KPAD;
  constexpr int STAGE_DOUBLES = (int)(sizeof(Gram3Stage) / sizeof(double));
  constexpr int MAXSLOT = ((G3_XROWS + G3_YROWS + G3_ZROWS) * (KC / 2) + GTHREADS - 1) / GTHREADS;
  double* sm = reinterpret_cast<double*>(smem_raw);
  const int rows = nxr + nj + TN3;
  const double* src[MAXSLOT];
  int dst[MAXSLOT], kq0[MAXSLOT];
#pragma unroll
  for (int sl = 0; sl < MAXSLOT; ++sl) {
    int v = tid + sl * GTHREADS;
    int r = v / (KC / 2), q = 2 * (v - r * (KC / 2));
    src[sl] = nullptr;
    dst[sl] = 0;
    kq0[sl] = q;
    if (r < nxr) {
      src[sl] = X + (long long)(i_first + r) * K + q;
      dst[sl] = (int)(&S.s[0].X[r][q] - sm);
    } else if (r < nxr + nj) {
      src[sl] = Y + (long long)(r - nxr) * K + q;
      dst[sl] = (int)(&S.s[0].Y[r - nxr][q] - sm);
    } else if (r < rows) {
      int rr = r - nxr - nj;
      dst[sl] = (int)(&S.s[0].Z[rr][q] - sm);
      src[sl] = (col0 + rr < nz) ? Z + (long long)(col0 + rr) * K + q : nullptr;
      if (!src[sl]) kq0[sl] = -1;                      // zero-fill row
    } else {
      kq0[sl] = -2;                                    // no work
    }
  }
  auto load_stage = [&](int stage, long long chunk) {
    const long long k0 = chunk * KC;
#pragma unroll
    for (int sl = 0; sl < MAXSLOT; ++sl) {
      if (kq0[sl] == -2) continue;
      bool ok = kq0[sl] >= 0 && k0 + kq0[sl] + 1 < K;
      cp_async16(sm + stage * STAGE_DOUBLES + dst[sl], ok ? (const void*)(src[sl] + k0) : (const void*)X, ok);
    }
  };

  // operand rows of this thread: A rows (lane>>2) + 8 i of the warp slab -> (X row, Y row); rows past m compute
  // garbage into the padded part of the partial tile, which the reduction never reads
  const int kk = lane & 3;
  const double *xp[4], *yp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = row0 + wm * 32 + (lane >> 2) + 8 * i;
    int rr = r < m ? r : row0;
    xp[i] = &S.s[0].X[rr / nj - i_first][wk * (KC / 2) + kk];
    yp[i] = &S.s[0].Y[rr % nj][wk * (KC / 2) + kk];
  }
  const double* zp = &S.s[0].Z[lane >> 2][wk * (KC / 2) + kk];
  const long long nchunks = c_end > c_begin ? c_end - c_begin : 0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, c_begin + s);
    cp_async_commit();
  }
  int st = 0;
  for (long long it = 0; it < nchunks; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      long long nx = it + STAGES - 1;
      int sn = st + STAGES - 1;
      if (sn >= STAGES) sn -= STAGES;
      if (nx < nchunks) load_stage(sn, c_begin + nx);
      cp_async_commit();
    }
    const int so = st * STAGE_DOUBLES;
#pragma unroll
    for (int k4 = 0; k4 < KC / 2; k4 += 4) {
      double a[4], b[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = xp[i][so + k4] * yp[i][so + k4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) b[j] = zp[so + 8 * j * PITCH + k4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (++st == STAGES) st = 0;
  }
  cp_async_wait<0>();
  __syncthreads();
  // sum the two K halves: warps 4-7 park their accumulators in (now idle) shared memory
  double2* park = reinterpret_cast<double2*>(smem_raw);          // [4 warps][4*NJ][32 lanes]
  if (wk == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) park[(wm * 4 * NJ + i * NJ + j) * 32 + lane] = make_double2(acc[i][j][0], acc[i][j][1]);
  }
  __syncthreads();
  if (wk == 0) {
    double* P = partial + (long long)ks * Mpad * Npad;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        double2 o = park[(wm * 4 * NJ + i * NJ + j) * 32 + lane];
        int r = row0 + wm * 32 + 8 * i + (lane >> 2);
        int cidx = col0 + 8 * j + 2 * (lane & 3);
        *reinterpret_cast<double2*>(&P[(long long)r * Npad + cidx]) = make_double2(acc[i][j][0] + o.x, acc[i][j][1] + o.y);
      }
  }
}

// split-K partial tiles: per-context scratch (two contexts may run Grams on two streams at the same time)
static int gram_scratch(sww_ctx* c, size_t need, double** out) {
  if (need > c->gram_partial_bytes) {
    if (c->d_gram_partial) {
      SWW_CUDA(cudaStreamSynchronize(c->stream));   // earlier launches of this context may still read the old buffer
      cudaFree(c->d_gram_partial);
    }
    SWW_CUDA(cudaMalloc(&c->d_gram_partial, need));
    c->gram_partial_bytes = need;
  }
  *out = c->d_gram_partial;
  return 0;
}

static int gram_run(sww_ctx* c, const double* A, const double* B, int m, int n, long long K, int accumulate, double* C,
                    int ldc, int col0) {
  SWW_REQUIRE(m > 0 && n > 0 && K > 0, "gram: empty operand");
  SWW_PROF(c, "gram.dmma f64");
  int mtiles = (m + TM - 1) / TM, ntiles = (n + TN - 1) / TN;
  int Mpad = mtiles * TM, Npad = ntiles * TN;
  long long total_chunks = (K + KCG - 1) / KCG;
  int ksplit = (SMS * 2 * GRAM_CTAS + mtiles * ntiles - 1) / (mtiles * ntiles);   // two waves of resident CTAs
  if (ksplit > total_chunks) ksplit = (int)total_chunks;
  if (ksplit < 1) ksplit = 1;
  long long cps = (total_chunks + ksplit - 1) / ksplit;
  ksplit = (int)((total_chunks + cps - 1) / cps);
  size_t need = (size_t)ksplit * Mpad * Npad * sizeof(double);
  double* g_partial = nullptr;
  SWW_TRY(gram_scratch(c, need, &g_partial));
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

template <int NJ>
static int gram3_launch(sww_ctx* c, const double* X, const double* Y, const double* Z, int ni, int nj, int nz, long long K,
                        double* C) {
  const int m = ni * nj, TN3 = 8 * NJ;
  int mtiles = (m + TM - 1) / TM, ntiles = (nz + TN3 - 1) / TN3;
  int Mpad = mtiles * TM, Npad = ntiles * TN3;
  long long total_chunks = (K + KC - 1) / KC;
  int ksplit = (SMS * 2 * GRAM3_CTAS + mtiles * ntiles - 1) / (mtiles * ntiles);
  if (ksplit > total_chunks) ksplit = (int)total_chunks;
  if (ksplit < 1) ksplit = 1;
  long long cps = (total_chunks + ksplit - 1) / ksplit;
  ksplit = (int)((total_chunks + cps - 1) / cps);
  size_t need = (size_t)ksplit * Mpad * Npad * sizeof(double);
  double* g_partial = nullptr;
  SWW_TRY(gram_scratch(c, need, &g_partial));
  size_t smem = sizeof(Gram3Smem);
  SWW_CUDA(cudaFuncSetAttribute(gram3_dmma_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(mtiles, ntiles, ksplit);
  gram3_dmma_kernel<NJ><<<grid, GTHREADS, smem, c->stream>>>(X, Y, Z, ni, nj, nz, K, cps, g_partial, Mpad, Npad);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  gram_reduce_kernel<<<sww_grid_for((long long)m * nz, 256), 256, 0, c->stream>>>(g_partial, ksplit, Mpad, Npad, m, nz, 0, C, nz, 0);
  c->n_own++;
  SWW_CUDA(cudaGetLastError());
  return 0;
}

// C[(i*nj + j) * nz + c] = sum_k X[i][k] Y[j][k] Z[c][k]
int sww_gram3(sww_ctx* c, const double* X, const double* Y, const double* Z, int ni, int nj, int nz, long long K, double* C) {
  SWW_REQUIRE(ni > 0 && nj >= 4 && nj <= G3_YROWS && nz > 0, "gram3: unsupported operand counts");
  SWW_REQUIRE(TM / nj + 2 <= G3_XROWS, "gram3: too many X rows per tile");
  SWW_REQUIRE(K % 2 == 0 && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)Z % 16 == 0), "gram3: operands must be 16-byte aligned with an even cell count");
  SWW_PROF(c, "gram3.dmma f64");
  // column blocks of 24, 32 or 40: whichever pads nz least
  int best = 3, waste = 1 << 30;
  for (int nj8 = 3; nj8 <= 5; ++nj8) {
    int w = (nz + 8 * nj8 - 1) / (8 * nj8) * 8 * nj8 - nz;
    if (w < waste) { waste = w; best = nj8; }
  }
  if (best == 3) return gram3_launch<3>(c, X, Y, Z, ni, nj, nz, K, C);
  if (best == 4) return gram3_launch<4>(c, X, Y, Z, ni, nj, nz, K, C);
  return gram3_launch<5>(c, X, Y, Z, ni, nj, nz, K, C);
}
