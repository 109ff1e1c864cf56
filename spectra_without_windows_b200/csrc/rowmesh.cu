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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_any.cuh"
#include "sww_internal.cuh"

int sww_ctx_create_internal(const sww_geometry* geo, sww_ctx** out, bool child);   // ctx.cu

static inline bool is_pow2_i(int n) { return n > 0 && (n & (n - 1)) == 0; }
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Cost model of one Fisher-row chain on a candidate row mesh, in picoseconds per row-mesh (x, y) column, from the measured
// passes (profiles/r2p, r2i): the z pass touches M_z / 2 complex values per row and transform, the column passes K_z = b_z + 1
// planes whatever M_z is.  Tuned power-of-two kernels: ~3 ps per complex value along z, ~8 ps along x / y scaled by the
// kept fraction of the axis (box pruning, at least half); any-length kernels: ~0.7 ps per unit of dense-radix work
// (fft_any.cuh: any_radix_cost).
static double pass_cost(int n, int d, int b) {
  const bool p2 = d == 2 ? (is_pow2_i(n) && n >= 128 && n <= 1024) : (is_pow2_i(n) && n >= 64 && n <= 1024);
  if (p2) return d == 2 ? 3.0 : 8.0 * std::max(0.5, std::min(1.0, (2.0 * b + 1.0) / n));
  AnyPlan pl = any_make_plan(n);
  double sum = 0.0;
  for (int i = 0; i < pl.ns; ++i) sum += any_radix_cost(pl.R[i]);
  return 0.7 * sum;
}
static double chain_cost(const int* T, const int* b) {
  return (double)T[0] * T[1] * (2.0 * pass_cost(T[2], 2, b[2]) * (T[2] / 2.0) + 2.0 * (b[2] + 1.0) * (pass_cost(T[0], 0, b[0]) + pass_cost(T[1], 1, b[1])));
}

int sww_rowmesh_create(sww_ctx* c, const sww_geometry* geo) {
  if (!sww_fft_sm100_supported(c)) return 0;
  if (const char* e = getenv("SWW_ROW_MESH"))
    if (e[0] == '0') return 0;
  const Geo& g = c->g;
  const int n[3] = {g.nx, g.ny, g.nz}, b[3] = {g.bx, g.by, g.bz}, lo[3] = {64, 64, 128};
  // per axis: stay on the mesh's own grid, or move to the smallest power of two that holds the band-limited product
  // without aliasing (M >= 2 (2 b + 1)) -- smaller than the mesh when k_max is well below Nyquist (BASELINE config 5), LARGER
  // than the mesh when the axis length has a large prime factor whose dense transform costs more than the extra cells
  // (paramfiles/boss_nC.param: 258 x 496 x 274 with k_max = 0.41).  The cheapest of the 2^3 combinations under the pass-cost
  // model wins; SWW_ROW_MESH_DIMS=mx,my,mz forces one (each entry the axis length or the power of two).
  int P[3], M[3];
  for (int d = 0; d < 3; ++d) {
    P[d] = std::max(lo[d], next_pow2(2 * (2 * b[d] + 1)));
    if (P[d] > 1024) P[d] = n[d];
  }
  double best = -1.0;
  for (int mask = 0; mask < 8; ++mask) {
    int T[3];
    for (int d = 0; d < 3; ++d) T[d] = (mask >> d) & 1 ? P[d] : n[d];
    const double cost = chain_cost(T, b);
    if (best < 0.0 || cost < best * (1.0 - 1e-9)) {
      best = cost;
      for (int d = 0; d < 3; ++d) M[d] = T[d];
    }
  }
  if (const char* e = getenv("SWW_ROW_MESH_DIMS")) {
    int v[3];
    if (sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3) {
      for (int d = 0; d < 3; ++d) SWW_REQUIRE(v[d] == n[d] || v[d] == P[d], "SWW_ROW_MESH_DIMS: axis %d must be %d or %d", d, n[d], P[d]);
      for (int d = 0; d < 3; ++d) M[d] = v[d];
    }
  }
  if (M[0] == n[0] && M[1] == n[1] && M[2] == n[2]) return 0;
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

// Row-mesh spectrum of the weight map from the mesh's own:  dst(d) = src(d mod n) for the signed row-mesh index d.
// A smaller axis (M < n) keeps |d| < M / 2 (its Nyquist plane is dropped); a LARGER axis (M > n) keeps |d| <= 2 b, the only
// differences k - k' of two box modes, and reads the n-periodic spectrum there -- the cyclic convolution of the mesh,
// wrap-around included, becomes a linear one on the row mesh.  src is a half spectrum: modes whose reduced z index is
// negative come from the Hermitian partner.
__global__ void band_copy_kernel(int nx, int ny, int nz, int mx, int my, int mz, int bx, int by, int bz, const double2* __restrict__ src,
                                 double2* __restrict__ dst) {
  const int nzh = nz / 2 + 1, mzh = mz / 2 + 1;
  const long long total = (long long)mx * my * mzh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dz = (int)(i % mzh);
    const long long r = i / mzh;
    const int jy = (int)(r % my), jx = (int)(r / my);
    const int dx = 2 * jx < mx ? jx : jx - mx, dy = 2 * jy < my ? jy : jy - my;
    bool keep = true;
    keep = keep && (mx < nx ? 2 * jx != mx : (mx == nx || abs(dx) <= 2 * bx));
    keep = keep && (my < ny ? 2 * jy != my : (my == ny || abs(dy) <= 2 * by));
    keep = keep && (mz < nz ? dz != mzh - 1 || (mz % 2) : (mz == nz || dz <= 2 * bz));
    double2 v = make_double2(0.0, 0.0);
    if (keep) {
      // reduce to the mesh's signed range (-n/2, n/2]
      int sx = dx % nx, sy = dy % ny, sz = dz % nz;
      if (2 * sx > nx) sx -= nx;
      if (2 * sx <= -nx) sx += nx;
      if (2 * sy > ny) sy -= ny;
      if (2 * sy <= -ny) sy += ny;
      if (2 * sz > nz) sz -= nz;
      const bool cj = sz < 0;
      if (cj) {
        sx = -sx;
        sy = -sy;
        sz = -sz;
      }
      const int ix = sx < 0 ? sx + nx : sx, iy = sy < 0 ? sy + ny : sy;
      v = src[((long long)ix * ny + iy) * nzh + sz];
      if (cj) v.y = -v.y;
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
  band_copy_kernel<<<sww_grid_for(m.Nc, 256), 256, 0, c->stream>>>(g.nx, g.ny, g.nz, m.nx, m.ny, m.nz, g.bx, g.by, g.bz, specF, specC);
  SWW_CUDA(cudaGetLastError());
  SWW_TRY(sww_fft_sm100_c2r(r, specC, r->d_nW));
  scale_real_kernel<<<sww_grid_for(m.N, 256), 256, 0, c->stream>>>(m.N, 1.0 / (double)g.N, r->d_nW);
  SWW_CUDA(cudaGetLastError());
  c->n_own += 2;
  r->fields_dirty = false;
  r->row_ready = true;
  return 0;
}
