// setup_kernels.cuh -- once-per-hierarchy kernels of MG.setup (multigrid.py:100-344 of the reference) that used to run on the
// host: the Galerkin product A_{l+1} = R_l A_l P_l (multigrid.py:276) straight into the padded block-sparse layout of the
// coarse-level kernels, and the dense inverse of the coarsest operator (multigrid.py:342-344).  Complex128 throughout; every
// output number is summed by ONE thread in a fixed order, so the result does not depend on the launch or on the rank.
#pragma once
#include "common.cuh"

namespace dmlmc {

// the fine operator as a list of scalar entries per row
struct GalOp {
  int kind;                                     // 0: Wilson-Dirac stencil in link form, 1: padded block-sparse rows
  int LX, LT; const Cx<double>* Ut; const Cx<double>* Ux; Cx<double> diag;
  int bs, bpr; const int* col; const Cx<double>* vals;
  __device__ __forceinline__ int entries_per_row() const { return kind == 0 ? 9 : bpr * bs; }
};
// the aggregation: closed form of multigrid.py:203-227 (rows == nullptr) or indexed equal-sized blocks
struct GalTr {
  int aggr, dofi, nvec, m;                      // m = fine rows per coarse block
  const int* rows; const int* cblk; const Cx<double>* pv;
  __device__ __forceinline__ int block_of(int r) const {
    if (cblk != nullptr) return __ldg(cblk + r);
    const int j = r / aggr, q = (r - j * aggr) % dofi;
    return 2 * j + (q >= (dofi >> 1) ? 1 : 0);
  }
  __device__ __forceinline__ int row_of(int I, int i) const {
    if (rows != nullptr) return __ldg(rows + (size_t)I * m + i);
    const int hd = dofi >> 1;
    return (I >> 1) * aggr + (i / hd) * dofi + (i % hd) + (I & 1) * hd;
  }
};

// entry e of row r: column *j and value *a; false for padding.  Stencil rows (lattice.py: row = s V + x LT + t):
//   e = 0: diagonal;  1, 2: forward t (spin 0, 1);  3, 4: backward t;  5, 6: forward x;  7, 8: backward x
__device__ __forceinline__ bool gal_entry(const GalOp& A, int r, int e, int* j, Cx<double>* a) {
  if (A.kind == 1) {
    const int rb = r / A.bs, ri = r - rb * A.bs;
    const int s = e / A.bs, jj = e - s * A.bs;
    const int cb = __ldg(A.col + (size_t)rb * A.bpr + s);
    if (cb < 0) return false;
    *j = cb * A.bs + jj;
    *a = ldc_ro<double>(A.vals, (((size_t)rb * A.bpr + s) * A.bs + ri) * A.bs + jj);
    return true;
  }
  const int V = A.LX * A.LT;
  const int sr = r / V, site = r - sr * V;
  const int x = site / A.LT, t = site - x * A.LT;
  if (e == 0) { *j = r; *a = A.diag; return true; }
  const int d = (e - 1) >> 1, sc = (e - 1) & 1;
  Cx<double> u; int cs;
  if (d == 0)      { cs = x * A.LT + (t + 1 == A.LT ? 0 : t + 1); u = ldc_ro<double>(A.Ut, site); }
  else if (d == 1) { cs = x * A.LT + (t == 0 ? A.LT - 1 : t - 1); u = cconj(ldc_ro<double>(A.Ut, cs)); }
  else if (d == 2) { cs = (x + 1 == A.LX ? 0 : x + 1) * A.LT + t; u = ldc_ro<double>(A.Ux, site); }
  else             { cs = (x == 0 ? A.LX - 1 : x - 1) * A.LT + t; u = cconj(ldc_ro<double>(A.Ux, cs)); }
  *j = sc * V + cs;
  const bool same = sr == sc;
  Cx<double> v;
  if (d == 0)      v = same ? cx<double>(-u.re, -u.im) : u;                                  // -(1 - s1) U
  else if (d == 1) v = cx<double>(-u.re, -u.im);                                             // -(1 + s1) U^*
  else if (d == 2) v = same ? cx<double>(-u.re, -u.im) : (sr == 0 ? cx<double>(u.im, -u.re) : cx<double>(-u.im, u.re));   // -(1 - s2) U
  else             v = same ? cx<double>(-u.re, -u.im) : (sr == 0 ? cx<double>(-u.im, u.re) : cx<double>(u.im, -u.re));   // -(1 + s2) U^*
  *a = v;
  return true;
}

// One CTA per coarse block row I.  Phase 1: the distinct coarse blocks J reached from the fine rows of I, sorted (the slot
// order of the padded layout: increasing block column, as scipy's sorted BSR gives it).  Phase 2: one thread per output number
//   A_c[I][w][u][v] = sum_{r in I} conj(P[r][u]) sum_{entries (r, j) with block(j) = J_w} A[r][j] P[j][v].
// col_out[nbc][cap] (-1 = padding), vals_out[nbc][cap][c][c] (zero where padded: cleared by the caller), *maxw = the largest
// number of slots any row needed (cap + 1 if a row overflowed: the caller retries with a larger cap).
__global__ void __launch_bounds__(256)
galerkin_kernel(GalOp A, GalTr T, int cap, int* __restrict__ col_out, Cx<double>* __restrict__ vals_out, int* __restrict__ maxw) {
  extern __shared__ int gal_list[];             // [cap]
  __shared__ int s_w, s_over;
  const int I = blockIdx.x, c = T.nvec, npr = A.entries_per_row();
  for (int w = threadIdx.x; w < cap; w += blockDim.x) gal_list[w] = -1;
  if (threadIdx.x == 0) { s_w = 0; s_over = 0; }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T.m * npr; idx += blockDim.x) {
    const int i = idx / npr, e = idx - i * npr;
    int j; Cx<double> a;
    if (!gal_entry(A, T.row_of(I, i), e, &j, &a)) continue;
    const int J = T.block_of(j);
    bool placed = false;
    for (int w = 0; w < cap && !placed; ++w) {
      const int old = atomicCAS(&gal_list[w], -1, J);
      placed = (old == -1 || old == J);
    }
    if (!placed) s_over = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int W = 0;
    while (W < cap && gal_list[W] >= 0) ++W;
    for (int a2 = 1; a2 < W; ++a2) {            // insertion sort (W <= cap, a few dozen at most)
      const int v = gal_list[a2]; int b = a2 - 1;
      while (b >= 0 && gal_list[b] > v) { gal_list[b + 1] = gal_list[b]; --b; }
      gal_list[b + 1] = v;
    }
    s_w = W;
    atomicMax(maxw, s_over ? cap + 1 : W);
  }
  __syncthreads();
  const int W = s_w;
  for (int w = threadIdx.x; w < cap; w += blockDim.x) col_out[(size_t)I * cap + w] = w < W ? gal_list[w] : -1;
  for (int o = threadIdx.x; o < W * c * c; o += blockDim.x) {
    const int w = o / (c * c), u = (o / c) % c, v = o % c;
    const int Jw = gal_list[w];
    double sr = 0.0, si = 0.0;
    for (int i = 0; i < T.m; ++i) {
      const int r = T.row_of(I, i);
      double ir = 0.0, ii = 0.0;
      for (int e = 0; e < npr; ++e) {
        int j; Cx<double> a;
        if (!gal_entry(A, r, e, &j, &a) || T.block_of(j) != Jw) continue;
        const Cx<double> p = ldc_ro<double>(T.pv, (size_t)j * c + v);
        ir = fma(a.re, p.re, fma(-a.im, p.im, ir));
        ii = fma(a.re, p.im, fma(a.im, p.re, ii));
      }
      const Cx<double> pu = ldc_ro<double>(T.pv, (size_t)r * c + u);       // conj(pu) * inner
      sr = fma(pu.re, ir, fma(pu.im, ii, sr));
      si = fma(pu.re, ii, fma(-pu.im, ir, si));
    }
    vals_out[(((size_t)I * cap + w) * c + u) * c + v] = cx<double>(sr, si);
  }
}

// In-place Gauss-Jordan inversion with partial (row) pivoting of a dense complex128 matrix M[n][n] (the matrix stays in L2):
// per step one single-block kernel (pivot search over column k, row exchange, scaled pivot row and the eliminated column saved
// to rowk / colk) and one wide kernel (rank-one update of all other rows); at the end every row undoes the row exchanges on its
// own entries (column swaps, last exchange first).  *info = k + 1 when the k-th pivot is exactly zero (cleared by the caller).
// Replaces np.linalg.inv (LAPACK getrf / getri, the same pivoting strategy) of multigrid.py:342-344.
constexpr int GJ_MAX_N = 8192;      // (8 192: 8 192 rank-one updates of a 1 GB matrix, ~3.5 s)
__global__ void __launch_bounds__(1024)
gj_pivot_kernel(Cx<double>* __restrict__ M, int n, int k, int* __restrict__ piv, int* __restrict__ info,
                double2* __restrict__ rowk, double2* __restrict__ colk) {
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ int s_p;
  __shared__ double2 s_piv;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
  // pivot: the largest |M[i][k]|, i >= k (first one on ties)
  double best = -1.0; int bi = k;
  for (int i = k + tid; i < n; i += nt) {
    const Cx<double> v = M[(size_t)i * n + k];
    const double a = v.re * v.re + v.im * v.im;
    if (a > best) { best = a; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { red_v[wid] = best; red_i[wid] = bi; }
  __syncthreads();
  if (tid == 0) {
    double b = red_v[0]; int p = red_i[0];
    for (int w = 1; w < (nt >> 5); ++w) if (red_v[w] > b || (red_v[w] == b && red_i[w] < p)) { b = red_v[w]; p = red_i[w]; }
    s_p = p; piv[k] = p;
    if (!(b > 0.0) && *info == 0) *info = k + 1;
    const Cx<double> pv = M[(size_t)p * n + k];
    s_piv = make_double2(pv.re, pv.im);
  }
  __syncthreads();
  const int p = s_p;
  const double pr = s_piv.x, pi = s_piv.y, pd = pr * pr + pi * pi;
  const double qr = pr / pd, qi = -pi / pd;                               // 1 / pivot
  // row exchange k <-> p; the new row k is divided by the pivot, with a 1 in the pivot position first
  for (int j = tid; j < n; j += nt) {
    const Cx<double> a = M[(size_t)p * n + j];
    if (p != k) M[(size_t)p * n + j] = M[(size_t)k * n + j];
    const double ar = (j == k) ? 1.0 : a.re, ai = (j == k) ? 0.0 : a.im;
    const double2 s = make_double2(ar * qr - ai * qi, ar * qi + ai * qr);
    rowk[j] = s;
    M[(size_t)k * n + j] = cx<double>(s.x, s.y);
  }
  __syncthreads();
  // the column that the update eliminates (after the exchange; row k itself is not updated)
  for (int i = tid; i < n; i += nt) {
    const Cx<double> f = (i == k) ? cx<double>(0.0, 0.0) : M[(size_t)i * n + k];
    colk[i] = make_double2(f.re, f.im);
  }
}

// M[i][:] -= f_i * row k  (with M[i][k] = 0 first) for every row i != k; grid (ceil(n / 128), n)
__global__ void __launch_bounds__(128)
gj_update_kernel(Cx<double>* __restrict__ M, int n, int k, const double2* __restrict__ rowk, const double2* __restrict__ colk) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == k || j >= n) return;
  const double2 f = colk[i], rk = rowk[j];
  Cx<double>* Mi = M + (size_t)i * n;
  Cx<double> a = (j == k) ? cx<double>(0.0, 0.0) : Mi[j];
  a.re = fma(-f.x, rk.x, fma(f.y, rk.y, a.re));
  a.im = fma(-f.x, rk.y, fma(-f.y, rk.x, a.im));
  Mi[j] = a;
}

// undo the row exchanges: in every row swap the entries of columns piv[k] and k, last exchange first (one thread per row)
__global__ void __launch_bounds__(128)
gj_unscramble_kernel(Cx<double>* __restrict__ M, int n, const int* __restrict__ piv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cx<double>* Mi = M + (size_t)i * n;
  for (int k = n - 1; k >= 0; --k) {
    const int p = __ldg(piv + k);
    if (p != k) { const Cx<double> a = Mi[k]; Mi[k] = Mi[p]; Mi[p] = a; }
  }
}

}  // namespace dmlmc
