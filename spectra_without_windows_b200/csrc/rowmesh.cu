// Coarse "row mesh" of the P_l(k) Fisher rows.
//
// A Fisher row is  T_alpha(k) = FT[ nW * IFT[Theta_a F_l] ](k)  for |k| < k_max only, and its input is confined to
// shell a (|k| < k_max as well).  With b_i the largest mode index of the k_max box along axis i, the product only involves
// the modes |q_i| <= 2 b_i of the weight map nW, and on a grid of M_i >= 2 (2 b_i + 1) points per axis none of the
// products' modes aliases back into |k_i| <= b_i.  So the whole row chain (I1 -> I2 -> Z(c2r * w * r2c) -> F2 -> F3) is run
// on the smallest supported power-of-two grid M with that property -- same box, same k_F, same bins -- against
//     w~ = IFT_M[ band-limit_{|q_i| < M_i/2}( FT_n[nW] ) ] / N_n ,
// computed once per sww_set_fields, and  T_alpha(k) = (N_n / N_M) * FT_M[ w~ * IFT_M[Theta_a F_l] ](k)  to round-off.
// The reference evaluates these products on the full mesh (src/covariances_pk.py:404-466 via
// pk/compute_pk_randoms.py:220-275); nothing but the summation order changes.  BASELINE config 5 (1024^3, k_max = 0.4):
// the rows run on 512 x 1024 x 512, a quarter of the cells.  Disabled with SWW_ROW_MESH=0.
#include <cstdlib>
#include <vector>

#include "sww_internal.cuh"

int sww_ctx_create_internal(const sww_geometry* geo, sww_ctx** out, bool child);   // ctx.cu

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int sww_rowmesh_create(sww_ctx* c, const sww_geometry* geo) {
  if (!sww_fft_sm100_supported(c)) return 0;
  if (const char* e = getenv("SWW_ROW_MESH"))
    if (e[0] == '0') return 0;
  const Geo& g = c->g;
  const int n[3] = {g.nx, g.ny, g.nz}, b[3] = {g.bx, g.by, g.bz}, lo[3] = {64, 64, 128};
  int M[3];
  bool smaller = false;
  for (int d = 0; d < 3; ++d) {
    M[d] = std::max(lo[d], next_pow2(2 * (2 * b[d] + 1)));
    if (M[d] >= n[d]) M[d] = n[d];
    smaller = smaller || M[d] < n[d];
  }
  if (!smaller) return 0;
  // 1-D tables of the coarse grid: k = signed index * k_F with k_F = kax[1] (src/opt_utilities.py:420-424: the same
  // product of an integer and 2 pi / L, so equal indices carry bit-identical k on both grids); r^ and MAS are unused
  std::vector<double> tab[9];
  sww_geometry cg = *geo;
  for (int d = 0; d < 3; ++d) {
    cg.grid[d] = M[d];
    const double kF = geo->kax[d][1];
    tab[d].assign(M[d], 0.0);
    tab[3 + d].resize(M[d]);
    tab[6 + d].assign(M[d], 1.0);
    for (int i = 0; i < M[d]; ++i) {
      const int s = (i >= M[d] / 2.0) ? i - M[d] : i;
      tab[3 + d][i] = (double)s * kF;
    }
    cg.rax[d] = tab[d].data();
    cg.kax[d] = tab[3 + d].data();
    cg.mas[d] = tab[6 + d].data();
  }
  sww_ctx* r = nullptr;
  SWW_TRY(sww_ctx_create_internal(&cg, &r, true));
  if (r->g.bx != g.bx || r->g.by != g.by || r->g.bz != g.bz) {   // cannot happen: same k tables below k_max
    sww_ctx_destroy(r);
    return 0;
  }
  r->stream = c->stream;
  r->parent = c;
  c->rowc = r;
  return 0;
}

// dst[coarse cell] = src[fine cell] for |q_i| < M_i / 2, zero elsewhere (Nyquist planes of the coarse grid included)
__global__ void band_copy_kernel(int nx, int ny, int nzh, int mx, int my, int mzh, const double2* __restrict__ src,
                                 double2* __restrict__ dst) {
  const long long total = (long long)mx * my * mzh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(i % mzh);
    const long long r = i / mzh;
    const int jy = (int)(r % my), jx = (int)(r / my);
    const int sx = 2 * jx < mx ? jx : jx - mx, sy = 2 * jy < my ? jy : jy - my;
    const bool nyq = (mx < nx && 2 * jx == mx) || (my < ny && 2 * jy == my) || (mzh < nzh && iz == mzh - 1);
    double2 v = make_double2(0.0, 0.0);
    if (!nyq) {
      const int ix = sx < 0 ? sx + nx : sx, iy = sy < 0 ? sy + ny : sy;
      v = src[((long long)ix * ny + iy) * nzh + iz];
    }
    dst[i] = v;
  }
}

__global__ void scale_real_kernel(long long n, double s, double* __restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= s;
}

// the filtered weight map of the row mesh from the parent's nW map (called at the end of sww_set_fields)
int sww_rowmesh_prepare(sww_ctx* c) {
  sww_ctx* r = c->rowc;
  if (!r) return 0;
  r->row_ready = false;
  if (c->fft_backend != 1 || !r->ar.base) return 0;
  const Geo &g = c->g, &m = r->g;
  Arena ar = c->ar;
  ar.reset(c->persist_bytes);
  double2* specF = ar.alloc<double2>((size_t)g.Nc);
  double2* specC = ar.alloc<double2>((size_t)m.Nc);
  if (!ar.ok()) return 0;   // workspace of a pipeline that never runs rows: stay on the full mesh
  r->stream = c->stream;
  SWW_TRY(sww_fft_sm100_r2c(c, c->d_nW, specF));
  band_copy_kernel<<<sww_grid_for(m.Nc, 256), 256, 0, c->stream>>>(g.nx, g.ny, g.nzh, m.nx, m.ny, m.nzh, specF, specC);
  SWW_CUDA(cudaGetLastError());
  SWW_TRY(sww_fft_sm100_c2r(r, specC, r->d_nW));
  scale_real_kernel<<<sww_grid_for(m.N, 256), 256, 0, c->stream>>>(m.N, 1.0 / (double)g.N, r->d_nW);
  SWW_CUDA(cudaGetLastError());
  c->n_own += 2;
  r->fields_dirty = false;
  r->row_ready = true;
  return 0;
}
