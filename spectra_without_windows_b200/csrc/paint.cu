// Mesh painting (TSC / CIC), tile-binned and bit-deterministic.
//
// Restates FKPCatalog.to_mesh(window='tsc' | 'cic', compensated=False, interlaced=False) as used at
// src/opt_utilities.py:219-267: base cell = rint((x - BoxCenter)/H) (TSC: weights 1/2(1/2-d)^2, 3/4-d^2, 1/2(1/2+d)^2 on
// cells c-1..c+1) or floor (CIC: 1-d, d on c, c+1), periodic wrap.
//
//   1. paint_count   : particles are binned by the 8x8x8-cell TILE of their base cell (counts per tile, max |w|);
//   2. paint_scan    : exclusive scan of the counts;
//   3. paint_scatter : particle indices grouped by tile (order inside a tile is irrelevant, see 4);
//   4. paint_tile    : one CTA per tile deposits its particles into a 10x10x10 shared-memory block (tile + one-cell halo)
//                      of 64-bit FIXED-POINT accumulators -- integer shared-memory atomics are native and, being
//                      associative, give the same bits in any order; the scale 2^e of a tile is chosen from its particle
//                      count and max |w| so that the block cannot overflow, i.e. the resolution is ~2^-61 of the tile's
//                      largest possible cell.  The block is then added to the mesh with plain (non-atomic) FP64
//                      read-modify-writes: tiles are processed in 8 (odd tile counts: up to 27) colour phases whose
//                      10^3 footprints never overlap, so every cell has one writer per phase and a fixed order of phases.
//
// Against the first version (one thread per particle, 27 FP64 RED atomics to the mesh; kept below as paint_kernel for
// meshes with fewer than 16 cells on an axis): same result to ~1e-16 of a cell, but run-to-run bit-identical, and the
// mesh is touched tile by tile instead of at random (a 512^3 mesh does not fit the 126 MB L2).
#include <algorithm>
#include <cmath>

#include "sww_internal.cuh"


struct PaintGeo {
  int n[3];        // mesh
  int T;           // tile edge in cells: 8, or 4 when the tiles are crowded (small mesh, many particles)
  int t[3];        // tiles per axis
  double bc[3], H[3];
  int scheme;      // 3 = TSC, 2 = CIC
};

__device__ __forceinline__ int base_cell(double s, int scheme) { return (int)(scheme == 3 ? rint(s) : floor(s)); }

__device__ __forceinline__ int wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

__device__ __forceinline__ long long tile_of(const PaintGeo& G, const double* __restrict__ pos, unsigned long long p, int* cw) {
  int tq[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double s = (pos[3 * p + d] - G.bc[d]) / G.H[d];
    const int c = wrap(base_cell(s, G.scheme), G.n[d]);
    if (cw) cw[d] = c;
    tq[d] = c / G.T;
  }
  return ((long long)tq[0] * G.t[1] + tq[1]) * G.t[2] + tq[2];
}

__global__ void paint_count_kernel(PaintGeo G, const double* __restrict__ pos, const double* __restrict__ w,
                                   unsigned long long n, int* __restrict__ count, unsigned long long* __restrict__ maxw) {
  double m = 0.0;
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    atomicAdd(&count[tile_of(G, pos, p, nullptr)], 1);
    m = fmax(m, fabs(w[p]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(maxw, (unsigned long long)__double_as_longlong(m));   // monotone for m >= 0
}

// exclusive scan of `count` (nt entries) by one CTA of 1024 threads: chunk sums, scan of the sums, chunk offsets
__global__ void __launch_bounds__(1024) paint_scan_kernel(const int* __restrict__ count, long long nt, long long* __restrict__ offs) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x;
  const long long chunk = (nt + 1023) / 1024, lo = tid * chunk, hi = lo + chunk < nt ? lo + chunk : nt;
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += count[i];
  part[tid] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    long long v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  long long run = part[tid] - s;
  for (long long i = lo; i < hi; ++i) {
    offs[i] = run;
    run += count[i];
  }
  if (tid == 1023) offs[nt] = part[1023];
}

__global__ void paint_scatter_kernel(PaintGeo G, const double* __restrict__ pos, unsigned long long n,
                                     const long long* __restrict__ offs, int* __restrict__ cursor,
                                     unsigned int* __restrict__ order) {
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const long long t = tile_of(G, pos, p, nullptr);
    order[offs[t] + atomicAdd(&cursor[t], 1)] = (unsigned int)p;
  }
}

// one CTA per tile of colour (cx, cy, cz); colours along an axis: tile parity, and 2 for the last tile of an odd count
template <int NS, int PT_T>   // stencil width (3 = TSC, 2 = CIC), tile edge
__global__ void __launch_bounds__(PT_T == 4 ? 256 : 128)
paint_tile_kernel(PaintGeo G, int3 ncol, int3 col, const double* __restrict__ pos, const double* __restrict__ w, double scale,
                  const long long* __restrict__ offs, const unsigned int* __restrict__ order,
                  const unsigned long long* __restrict__ maxw, double* __restrict__ out) {
  constexpr int PT_B = PT_T + 2, PT_CELLS = PT_B * PT_B * PT_B;   // block = tile + one-cell halo
  constexpr int PT_THREADS = PT_T == 4 ? 256 : 128;                // crowded (4^3) tiles get twice the lanes
  __shared__ unsigned long long acc[PT_CELLS];
  const int tid = threadIdx.x;
  // tiles of this colour along each axis
  auto n_of = [](int t, int nc, int c) { return nc == 2 ? (t - c + 1) / 2 : (c == 2 ? 1 : (t - 1 - c + 1) / 2); };
  auto t_of = [](int t, int nc, int c, int j) { return (nc == 3 && c == 2) ? t - 1 : 2 * j + c; };
  const int m0 = n_of(G.t[0], ncol.x, col.x), m1 = n_of(G.t[1], ncol.y, col.y), m2 = n_of(G.t[2], ncol.z, col.z);
  const long long ntile = (long long)m0 * m1 * m2;
  const double wmax = __longlong_as_double((long long)*maxw) * fabs(scale);
  for (long long it = blockIdx.x; it < ntile; it += gridDim.x) {
    const int j2 = (int)(it % m2), j1 = (int)((it / m2) % m1), j0 = (int)(it / ((long long)m2 * m1));
    const int tq[3] = {t_of(G.t[0], ncol.x, col.x, j0), t_of(G.t[1], ncol.y, col.y, j1), t_of(G.t[2], ncol.z, col.z, j2)};
    const long long t = ((long long)tq[0] * G.t[1] + tq[1]) * G.t[2] + tq[2];
    const long long p0 = offs[t], p1 = offs[t + 1];
    if (p1 == p0) continue;                                   // uniform per CTA
    // fixed-point scale of this tile: (p1 - p0) * wmax * 2^e < 2^62
    int ex;
    frexp((double)(p1 - p0) * wmax, &ex);                     // value < 2^ex
    const int e = 61 - ex;
    const double S = ldexp(1.0, e), invS = ldexp(1.0, -e);
    __syncthreads();
    for (int i = tid; i < PT_CELLS; i += PT_THREADS) acc[i] = 0ull;
    __syncthreads();
    const int org[3] = {tq[0] * PT_T, tq[1] * PT_T, tq[2] * PT_T};
    for (long long q = p0 + tid; q < p1; q += PT_THREADS) {
      const unsigned long long p = order[q];
      double wt[3][3];
      int l0[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double s = (pos[3 * p + d] - G.bc[d]) / G.H[d];
        const int cb = base_cell(s, NS);
        const double dd = s - (double)cb;
        l0[d] = wrap(cb, G.n[d]) - org[d] + (NS == 3 ? 0 : 1);   // first stencil cell in block coordinates
        if (NS == 3) {
          wt[d][0] = 0.5 * (0.5 - dd) * (0.5 - dd);
          wt[d][1] = 0.75 - dd * dd;
          wt[d][2] = 0.5 * (0.5 + dd) * (0.5 + dd);
        } else {
          wt[d][0] = 1.0 - dd;
          wt[d][1] = dd;
          wt[d][2] = 0.0;
        }
      }
      const double ww = w[p] * scale;
#pragma unroll
      for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
          const double wab = ww * wt[0][a] * wt[1][b];
#pragma unroll
          for (int cz = 0; cz < NS; ++cz) {
            const long long v = __double2ll_rn(wab * wt[2][cz] * S);
            atomicAdd(&acc[((l0[0] + a) * PT_B + l0[1] + b) * PT_B + l0[2] + cz], (unsigned long long)v);
          }
        }
    }
    __syncthreads();
    // flush: one writer per mesh cell within a colour phase -> plain read-modify-write
    for (int i = tid; i < PT_CELLS; i += PT_THREADS) {
      const long long v = (long long)acc[i];
      if (v == 0) continue;
      const int lz = i % PT_B, ly = (i / PT_B) % PT_B, lx = i / (PT_B * PT_B);
      const int gx = wrap(org[0] + lx - 1, G.n[0]), gy = wrap(org[1] + ly - 1, G.n[1]), gz = wrap(org[2] + lz - 1, G.n[2]);
      double* o = out + ((long long)gx * G.n[1] + gy) * G.n[2] + gz;
      *o += (double)v * invS;
    }
  }
}

// v1 (meshes with fewer than 16 cells on an axis): one thread per particle, FP64 RED atomics straight to L2
__global__ void paint_kernel(Geo g, const double* __restrict__ pos, const double* __restrict__ w, unsigned long long n,
                             double bcx, double bcy, double bcz, double Hx, double Hy, double Hz, int scheme,
                             double scale, double* __restrict__ out) {
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    double s[3] = {(pos[3 * p] - bcx) / Hx, (pos[3 * p + 1] - bcy) / Hy, (pos[3 * p + 2] - bcz) / Hz};
    int c0[3];
    double wt[3][3];
    int ns = scheme;  // 3 = TSC, 2 = CIC
    for (int d = 0; d < 3; ++d) {
      if (scheme == 3) {
        double c = rint(s[d]);
        double dd = s[d] - c;
        c0[d] = (int)c - 1;
        wt[d][0] = 0.5 * (0.5 - dd) * (0.5 - dd);
        wt[d][1] = 0.75 - dd * dd;
        wt[d][2] = 0.5 * (0.5 + dd) * (0.5 + dd);
      } else {
        double c = floor(s[d]);
        double dd = s[d] - c;
        c0[d] = (int)c;
        wt[d][0] = 1.0 - dd;
        wt[d][1] = dd;
        wt[d][2] = 0.0;
      }
    }
    double ww = w[p] * scale;
    int dims[3] = {g.nx, g.ny, g.nz};
    for (int a = 0; a < ns; ++a) {
      int i = ((c0[0] + a) % dims[0] + dims[0]) % dims[0];
      for (int b = 0; b < ns; ++b) {
        int j = ((c0[1] + b) % dims[1] + dims[1]) % dims[1];
        double wab = ww * wt[0][a] * wt[1][b];
        for (int cz = 0; cz < ns; ++cz) {
          int k = ((c0[2] + cz) % dims[2] + dims[2]) % dims[2];
          atomicAdd(&out[((long long)i * g.ny + j) * g.nz + k], wab * wt[2][cz]);
        }
      }
    }
  }
}

static int paint_scratch(sww_ctx* c, size_t bytes) {
  if (c->paint_scratch_bytes >= bytes) return 0;
  if (c->d_paint_scratch) cudaFree(c->d_paint_scratch);
  c->d_paint_scratch = nullptr;
  c->paint_scratch_bytes = 0;
  SWW_CUDA(cudaMalloc(&c->d_paint_scratch, bytes));
  c->paint_scratch_bytes = bytes;
  return 0;
}

int k_paint(sww_ctx* c, const double* pos, const double* w, unsigned long long n, const double bc[3], int scheme,
            double scale, double* out) {
  if (n == 0) return 0;
  const Geo& g = c->g;
  const double Hx = c->box[0] / g.nx, Hy = c->box[1] / g.ny, Hz = c->box[2] / g.nz;
  const char* e = getenv("SWW_PAINT_V1");
  const bool v1 = (e && e[0] == '1') || g.nx < 16 || g.ny < 16 || g.nz < 16 || n >= (1ull << 32);
  if (v1) {
    SWW_PROF(c, "paint.v1 global atomics");
    paint_kernel<<<sww_grid_for((long long)n, 256, SMS * 16), 256, 0, c->stream>>>(g, pos, w, n, bc[0], bc[1], bc[2], Hx, Hy, Hz,
                                                                                 scheme, scale, out);
    c->n_own++;
    SWW_CUDA(cudaGetLastError());
    return 0;
  }
  PaintGeo G;
  G.n[0] = g.nx;
  G.n[1] = g.ny;
  G.n[2] = g.nz;
  const double H[3] = {Hx, Hy, Hz};
  // 8^3 tiles; crowded tiles (> 512 particles on average: 1.5e7 randoms on 128^3) serialise on one CTA each, so those
  // meshes take 4^3 tiles -- eight times the CTAs per colour phase for a 3.4x instead of 1.95x halo overhead
  {
    const long long nt8 = (long long)((g.nx + 7) / 8) * ((g.ny + 7) / 8) * ((g.nz + 7) / 8);
    G.T = (n / (unsigned long long)nt8 > 512ull) ? 4 : 8;
    if (const char* et = getenv("SWW_PAINT_TILE")) {
      if (et[0] == '4') G.T = 4;
      if (et[0] == '8') G.T = 8;
    }
  }
  for (int d = 0; d < 3; ++d) {
    G.t[d] = (G.n[d] + G.T - 1) / G.T;
    G.bc[d] = bc[d];
    G.H[d] = H[d];
  }
  G.scheme = scheme;
  const long long nt = (long long)G.t[0] * G.t[1] * G.t[2];
  // scratch: offs[nt+1] (i64) | maxw (u64) | count[nt] | cursor[nt] (i32) | order[n] (u32)
  const size_t b_offs = (size_t)(nt + 1) * 8, b_max = 8, b_cnt = (size_t)nt * 4, b_ord = (size_t)n * 4;
  SWW_TRY(paint_scratch(c, b_offs + b_max + 2 * b_cnt + b_ord + 64));
  char* base = (char*)c->d_paint_scratch;
  long long* offs = (long long*)base;
  unsigned long long* maxw = (unsigned long long*)(base + b_offs);
  int* count = (int*)(base + b_offs + b_max);
  int* cursor = count + nt;
  unsigned int* order = (unsigned int*)(cursor + nt);
  SWW_CUDA(cudaMemsetAsync(maxw, 0, b_max + 2 * b_cnt, c->stream));
  const int pg = sww_grid_for((long long)n, 256, SMS * 16);
  {
    SWW_PROF(c, "paint.bin (count, scan, scatter)");
    paint_count_kernel<<<pg, 256, 0, c->stream>>>(G, pos, w, n, count, maxw);
    paint_scan_kernel<<<1, 1024, 0, c->stream>>>(count, nt, offs);
    paint_scatter_kernel<<<pg, 256, 0, c->stream>>>(G, pos, n, offs, cursor, order);
    c->n_own += 3;
    SWW_CUDA(cudaGetLastError());
  }
  {
    SWW_PROF(c, "paint.tiles (smem fixed-point)");
    int3 ncol = make_int3(G.t[0] % 2 ? 3 : 2, G.t[1] % 2 ? 3 : 2, G.t[2] % 2 ? 3 : 2);
    for (int cx = 0; cx < ncol.x; ++cx)
      for (int cy = 0; cy < ncol.y; ++cy)
        for (int cz = 0; cz < ncol.z; ++cz) {
          long long tiles = (long long)((G.t[0] + 1) / 2) * ((G.t[1] + 1) / 2) * ((G.t[2] + 1) / 2);
          int grid = (int)std::min<long long>(tiles, (long long)SMS * 16);
#define PT_LAUNCH(NS_, T_) \
  paint_tile_kernel<NS_, T_><<<grid, (T_) == 4 ? 256 : 128, 0, c->stream>>>(G, ncol, make_int3(cx, cy, cz), pos, w, scale, offs, order, maxw, out)
          if (scheme == 3) {
            if (G.T == 8) PT_LAUNCH(3, 8); else PT_LAUNCH(3, 4);
          } else {
            if (G.T == 8) PT_LAUNCH(2, 8); else PT_LAUNCH(2, 4);
          }
#undef PT_LAUNCH
          c->n_own++;
        }
    SWW_CUDA(cudaGetLastError());
  }
  return 0;
}
