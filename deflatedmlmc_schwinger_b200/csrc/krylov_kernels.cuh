// krylov_kernels.cuh -- device side of the batched (one Krylov space per column) FGMRES:
// column-wise inner products, Gram-Schmidt updates, Givens/Hessenberg bookkeeping and the
// triangular solves, all complex128 and all on the device (no per-iteration host math).
// Replaces pyamg.krylov.fgmres as called at multigrid.py:362 and the np.vdot / np.dot
// call sites of utils.py:224,249,266,336,353.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace dmlmc {

typedef Cx<double> Z;   // complex128

__device__ __forceinline__ Z zadd(Z a, Z b) { return cx<double>(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ Z zsub(Z a, Z b) { return cx<double>(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ Z zscale(double s, Z a) { return cx<double>(s * a.re, s * a.im); }
__device__ __forceinline__ double zabs2(Z a) { return a.re * a.re + a.im * a.im; }
// acc += conj(a) * b
__device__ __forceinline__ void zfma_conj(Z& acc, Z a, Z b) {
  acc.re = fma(a.re, b.re, fma(a.im, b.im, acc.re));
  acc.im = fma(a.re, b.im, fma(-a.im, b.re, acc.im));
}
__device__ __forceinline__ void zfma(Z& acc, Z a, Z b) {
  acc.re = fma(a.re, b.re, fma(-a.im, b.im, acc.re));
  acc.im = fma(a.re, b.im, fma(a.im, b.re, acc.im));
}

constexpr int DOT_NI = 8;            // vectors per register tile
constexpr int DOT_TX = 32, DOT_TY = 8;

// partial[chunk][i][col] = sum_{r in chunk} conj(V_i[r][col]) * W[r][col],  V_i = Vbase + i*vstride
// grid (ceil(k/32), nchunks), block (32, 8).  Deterministic: fixed row chunking, no atomics, so a
// column's result does not depend on which batch it is part of.
// VT = float (option dot32, OFF): the V_i are read from complex64 copies (arithmetic stays complex128) for the Gram-Schmidt
// coefficients only.  Half the bytes of the pass, but the basis then loses orthogonality at the 1e-7 level and a GMRES
// cycle stagnates near that relative residual: measured 43 outer iterations instead of 8 -- kept as a recorded negative.
// WT: scalar of W (float: the complex64-stored Krylov vectors of the mixed-precision Schur-complement solve; the products are
// accumulated in FP64 either way).
template <typename VT, typename WT = double>
__global__ void __launch_bounds__(256)
multi_dot_kernel(const Cx<VT>* __restrict__ Vbase, size_t vstride, int nv, const Cx<WT>* __restrict__ W,
                 int n, int k, int rows_per_chunk, Z* __restrict__ partial) {
  __shared__ Z red[DOT_TY][DOT_NI][DOT_TX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * DOT_TX + tx;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(n, r0 + rows_per_chunk);
  const bool ok = col < k;
  for (int i0 = 0; i0 < nv; i0 += DOT_NI) {
    Z acc[DOT_NI];
#pragma unroll
    for (int i = 0; i < DOT_NI; ++i) acc[i] = cx<double>(0.0, 0.0);
    if (ok) {
      for (int r = r0 + ty; r < r1; r += DOT_TY) {
        const size_t off = (size_t)r * k + col;
        const Cx<WT> w_ = ldc_ro<WT>(W, off);
        if constexpr (std::is_same<VT, float>::value && std::is_same<WT, float>::value) {
          // complex64 operands: the product in FP32 (its rounding, 6e-8, is that of the stored operands), the sum in FP64 --
          // two FP64 adds per term instead of four FP64 FMAs and four conversions (ncu, run r2_8: the all-FP64 form of this
          // kernel ran at 2.9 TB/s with the FP64 pipe as its top pipe)
#pragma unroll
          for (int i = 0; i < DOT_NI; ++i)
            if (i0 + i < nv) {
              const Cx<float> v = ldc_ro<float>(Vbase, (size_t)(i0 + i) * vstride + off);
              const float pr = fmaf(v.re, w_.re, v.im * w_.im), pi = fmaf(v.re, w_.im, -(v.im * w_.re));
              acc[i].re += (double)pr; acc[i].im += (double)pi;
            }
        } else {
          const Z w = cx<double>((double)w_.re, (double)w_.im);
#pragma unroll
          for (int i = 0; i < DOT_NI; ++i)
            if (i0 + i < nv) {
              const Cx<VT> v = ldc_ro<VT>(Vbase, (size_t)(i0 + i) * vstride + off);
              zfma_conj(acc[i], cx<double>((double)v.re, (double)v.im), w);
            }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DOT_NI; ++i) red[ty][i][tx] = acc[i];
    __syncthreads();
    // thread (tx, ty) reduces vector i = ty over the 8 row lanes
    Z s = cx<double>(0.0, 0.0);
#pragma unroll
    for (int y = 0; y < DOT_TY; ++y) s = zadd(s, red[y][ty][tx]);
    if (ok && i0 + ty < nv) partial[((size_t)blockIdx.y * nv + (i0 + ty)) * k + col] = s;
    __syncthreads();
  }
}

// ---- the same two kernels for complex64 vectors (the mixed-precision Schur-complement solve), two columns per thread: 16-byte
// loads put twice the bytes in flight per thread (ncu of the one-column form at k = 512: 2.9 TB/s, latency-bound; nv <= 3 there).
// Products in FP32, sums in FP64; same row chunks and partial layout as the kernels above (deterministic).  k must be even.
constexpr int D2_NI = 4;
__global__ void __launch_bounds__(256)
multi_dot_c64x2_kernel(const Cx<float>* __restrict__ Vbase, size_t vstride, int nv, const Cx<float>* __restrict__ W,
                       int n, int k, int rows_per_chunk, Z* __restrict__ partial) {
  __shared__ Z red[DOT_TY][D2_NI][2 * DOT_TX + 1];
  typedef Pack<float, 2> P;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = (blockIdx.x * DOT_TX + tx) * 2;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(n, r0 + rows_per_chunk);
  const bool ok = col < k;
  const int kp = k >> 1, cp = col >> 1;
  const P* Wp = reinterpret_cast<const P*>(W);
  const int tid = ty * DOT_TX + tx;
  for (int i0 = 0; i0 < nv; i0 += D2_NI) {
    double ar[D2_NI][2], ai[D2_NI][2];
#pragma unroll
    for (int i = 0; i < D2_NI; ++i) { ar[i][0] = ar[i][1] = ai[i][0] = ai[i][1] = 0.0; }
    if (ok) {
#pragma unroll 2
      for (int r = r0 + ty; r < r1; r += DOT_TY) {
        const size_t off = (size_t)r * kp + cp;
        const P w = ldp_ro<float, 2>(Wp, off);
#pragma unroll
        for (int i = 0; i < D2_NI; ++i)
          if (i0 + i < nv) {
            const P v = ldp_ro<float, 2>(reinterpret_cast<const P*>(Vbase + (size_t)(i0 + i) * vstride), off);
            ar[i][0] += (double)fmaf(v.d[0], w.d[0], v.d[1] * w.d[1]); ai[i][0] += (double)fmaf(v.d[0], w.d[1], -(v.d[1] * w.d[0]));
            ar[i][1] += (double)fmaf(v.d[2], w.d[2], v.d[3] * w.d[3]); ai[i][1] += (double)fmaf(v.d[2], w.d[3], -(v.d[3] * w.d[2]));
          }
      }
    }
#pragma unroll
    for (int i = 0; i < D2_NI; ++i) { red[ty][i][2 * tx] = cx<double>(ar[i][0], ai[i][0]); red[ty][i][2 * tx + 1] = cx<double>(ar[i][1], ai[i][1]); }
    __syncthreads();
    // thread t reduces (vector t / 64, column t % 64) over the 8 row lanes
    {
      const int i = tid >> 6, c = tid & 63;
      Z sacc = cx<double>(0.0, 0.0);
#pragma unroll
      for (int y = 0; y < DOT_TY; ++y) sacc = zadd(sacc, red[y][i][c]);
      const int gcol = blockIdx.x * DOT_TX * 2 + c;
      if (gcol < k && i0 + i < nv) partial[((size_t)blockIdx.y * nv + (i0 + i)) * k + gcol] = sacc;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256)
multi_axpy_norm_c64x2_kernel(const Cx<float>* __restrict__ Vbase, size_t vstride, int nv, const Z* __restrict__ h,
                             Cx<float>* __restrict__ W, int n, int k, int rows_per_chunk, Z* __restrict__ partial) {
  __shared__ double red[DOT_TY][2 * DOT_TX + 1];
  typedef Pack<float, 2> P;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = (blockIdx.x * DOT_TX + tx) * 2;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(n, r0 + rows_per_chunk);
  const int kp = k >> 1, cp = col >> 1;
  P* Wp = reinterpret_cast<P*>(W);
  double sq0 = 0.0, sq1 = 0.0;
  if (col < k) {
#pragma unroll 2
    for (int r = r0 + ty; r < r1; r += DOT_TY) {
      const size_t off = (size_t)r * kp + cp;
      float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
      for (int i = 0; i < nv; ++i) {
        const P v = ldp_ro<float, 2>(reinterpret_cast<const P*>(Vbase + (size_t)i * vstride), off);
        const Z h0 = ldc_ro<double>(h, (size_t)i * k + col), h1 = ldc_ro<double>(h, (size_t)i * k + col + 1);
        const float h0r = (float)h0.re, h0i = (float)h0.im, h1r = (float)h1.re, h1i = (float)h1.im;
        a0r = fmaf(h0r, v.d[0], fmaf(-h0i, v.d[1], a0r)); a0i = fmaf(h0r, v.d[1], fmaf(h0i, v.d[0], a0i));
        a1r = fmaf(h1r, v.d[2], fmaf(-h1i, v.d[3], a1r)); a1i = fmaf(h1r, v.d[3], fmaf(h1i, v.d[2], a1i));
      }
      P w = Wp[off];
      w.d[0] -= a0r; w.d[1] -= a0i; w.d[2] -= a1r; w.d[3] -= a1i;
      Wp[off] = w;
      sq0 += (double)fmaf(w.d[0], w.d[0], w.d[1] * w.d[1]);
      sq1 += (double)fmaf(w.d[2], w.d[2], w.d[3] * w.d[3]);
    }
  }
  red[ty][2 * tx] = sq0; red[ty][2 * tx + 1] = sq1;
  __syncthreads();
  const int tid = ty * DOT_TX + tx;
  if (tid < 2 * DOT_TX) {
    const int gcol = blockIdx.x * DOT_TX * 2 + tid;
    if (gcol < k) {
      double t = red[0][tid];
#pragma unroll
      for (int y = 1; y < DOT_TY; ++y) t += red[y][tid];
      partial[(size_t)blockIdx.y * k + gcol] = cx<double>(t, 0.0);
    }
  }
}

// out[idx] (+)= sum_chunk partial[chunk*count + idx].  Block (32, 8): 8 lanes share the chunks of an output, then a
// fixed-order shared-memory reduction (deterministic; the order depends on nchunks only).
__global__ void __launch_bounds__(256)
sum_partials_kernel(const Z* __restrict__ partial, int nchunks, int count, Z* __restrict__ out, int accumulate) {
  __shared__ Z red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int idx = blockIdx.x * 32 + tx;
  Z s = cx<double>(0.0, 0.0);
  if (idx < count)
    for (int c = ty; c < nchunks; c += 8) s = zadd(s, ldc_ro<double>(partial, (size_t)c * count + idx));
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && idx < count) {
    Z t = red[0][tx];
#pragma unroll
    for (int y = 1; y < 8; ++y) t = zadd(t, red[y][tx]);
    if (accumulate) t = zadd(t, out[idx]);
    out[idx] = t;
  }
}

// W[r][col] += sgn * sum_i h[i][col] * V_i[r][col]
template <typename VT = double>
__global__ void __launch_bounds__(256)
multi_axpy_kernel(const Cx<VT>* __restrict__ Vbase, size_t vstride, int nv, const Z* __restrict__ h,
                  Z* __restrict__ W, size_t nk, int k, double sgn) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nk) return;
  const int col = (int)(idx % k);
  Z acc = cx<double>(0.0, 0.0);
  for (int i = 0; i < nv; ++i) {
    const Cx<VT> v = ldc_ro<VT>(Vbase, (size_t)i * vstride + idx);
    zfma(acc, ldc_ro<double>(h, (size_t)i * k + col), cx<double>((double)v.re, (double)v.im));
  }
  Z w = W[idx];
  w.re = fma(sgn, acc.re, w.re);
  w.im = fma(sgn, acc.im, w.im);
  W[idx] = w;
}

// Gram-Schmidt update fused with the norm of its result:
//   W[r][col] -= sum_i h[i][col] V_i[r][col];   partial[chunk][col] = sum_{r in chunk} |W[r][col]|^2   (as a complex, im = 0)
// same (32, 8) x row-chunk decomposition as multi_dot_kernel, so the norm is deterministic and needs no extra pass over W.
template <typename VT = double>
__global__ void __launch_bounds__(256)
multi_axpy_norm_kernel(const Cx<VT>* __restrict__ Vbase, size_t vstride, int nv, const Z* __restrict__ h,
                       Cx<VT>* __restrict__ W, int n, int k, int rows_per_chunk, Z* __restrict__ partial) {
  __shared__ double red[DOT_TY][DOT_TX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * DOT_TX + tx;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(n, r0 + rows_per_chunk);
  double sq = 0.0;
  if (col < k) {
    for (int r = r0 + ty; r < r1; r += DOT_TY) {
      const size_t off = (size_t)r * k + col;
      if constexpr (std::is_same<VT, float>::value) {
        // complex64 vectors: the update in FP32 (its result is stored in complex64 anyway), the norm accumulated in FP64
        float ar = 0.f, ai = 0.f;
        for (int i = 0; i < nv; ++i) {
          const Cx<float> v = ldc_ro<float>(Vbase, (size_t)i * vstride + off);
          const Z hz = ldc_ro<double>(h, (size_t)i * k + col);
          const float hr = (float)hz.re, hi = (float)hz.im;
          ar = fmaf(hr, v.re, fmaf(-hi, v.im, ar));
          ai = fmaf(hr, v.im, fmaf(hi, v.re, ai));
        }
        const Cx<float> w0 = W[off];
        const Cx<float> ws = cx<float>(w0.re - ar, w0.im - ai);
        W[off] = ws;
        sq += (double)fmaf(ws.re, ws.re, ws.im * ws.im);
      } else {
        Z acc = cx<double>(0.0, 0.0);
        for (int i = 0; i < nv; ++i) {
          const Cx<VT> v = ldc_ro<VT>(Vbase, (size_t)i * vstride + off);
          zfma(acc, ldc_ro<double>(h, (size_t)i * k + col), cx<double>((double)v.re, (double)v.im));
        }
        const Cx<VT> w0 = W[off];
        // the norm is that of the STORED vector (what the next basis vector is made from)
        const Cx<VT> ws = cx<VT>((VT)((double)w0.re - acc.re), (VT)((double)w0.im - acc.im));
        W[off] = ws;
        const Z w = cx<double>((double)ws.re, (double)ws.im);
        sq = fma(w.re, w.re, fma(w.im, w.im, sq));
      }
    }
  }
  red[ty][tx] = sq;
  __syncthreads();
  if (ty == 0 && col < k) {
    double t = red[0][tx];
#pragma unroll
    for (int y = 1; y < DOT_TY; ++y) t += red[y][tx];
    partial[(size_t)blockIdx.y * k + col] = cx<double>(t, 0.0);
  }
}

// Out[r][col] = In[r][col] * scale[col]; Out32 (optional): the same values as complex64 -- the next V-cycle's input,
// written here instead of by a separate conversion pass over Out
__global__ void __launch_bounds__(256)
col_scale_kernel(const Z* __restrict__ In, const double* __restrict__ scale, Z* __restrict__ Out, size_t nk, int k,
                 Cx<float>* __restrict__ Out32) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nk) return;
  const double s = __ldg(scale + (idx % k));
  const Z v = zscale(s, ldc_ro<double>(In, idx));
  Out[idx] = v;
  if (Out32 != nullptr) Out32[idx] = cx<float>((float)v.re, (float)v.im);
}

// deflation with a dense V[n][d] shared by all columns:  C[i][col] = sum_r conj(V[r][i]) X[r][col]
__global__ void __launch_bounds__(256)
defl_dot_kernel(const Z* __restrict__ Vd, int d, const Z* __restrict__ X, int n, int k, int rows_per_chunk,
                Z* __restrict__ partial) {
  __shared__ Z red[DOT_TY][DOT_NI][DOT_TX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * DOT_TX + tx;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(n, r0 + rows_per_chunk);
  const bool ok = col < k;
  for (int i0 = 0; i0 < d; i0 += DOT_NI) {
    Z acc[DOT_NI];
#pragma unroll
    for (int i = 0; i < DOT_NI; ++i) acc[i] = cx<double>(0.0, 0.0);
    if (ok) {
      for (int r = r0 + ty; r < r1; r += DOT_TY) {
        const Z x = ldc_ro<double>(X, (size_t)r * k + col);
#pragma unroll
        for (int i = 0; i < DOT_NI; ++i)
          if (i0 + i < d) zfma_conj(acc[i], ldc_ro<double>(Vd, (size_t)r * d + i0 + i), x);
      }
    }
#pragma unroll
    for (int i = 0; i < DOT_NI; ++i) red[ty][i][tx] = acc[i];
    __syncthreads();
    Z s = cx<double>(0.0, 0.0);
#pragma unroll
    for (int y = 0; y < DOT_TY; ++y) s = zadd(s, red[y][ty][tx]);
    if (ok && i0 + ty < d) partial[((size_t)blockIdx.y * d + (i0 + ty)) * k + col] = s;
    __syncthreads();
  }
}
// X[r][col] -= sum_i V[r][i] * C[i][col]
__global__ void __launch_bounds__(256)
defl_axpy_kernel(const Z* __restrict__ Vd, int d, const Z* __restrict__ C, Z* __restrict__ X, size_t nk, int k) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nk) return;
  const int col = (int)(idx % k);
  const size_t r = idx / k;
  Z acc = cx<double>(0.0, 0.0);
  for (int i = 0; i < d; ++i) zfma(acc, ldc_ro<double>(Vd, r * d + i), ldc_ro<double>(C, (size_t)i * k + col));
  X[idx] = zsub(X[idx], acc);
}

// ------------------------------------------------------------------------------------------
// The deflation projections on the FP64 tensor cores (mma.sync.m8n8k4.f64, "DMMA"): the only dense
// contractions of the estimator itself (utils.py:224,266), complex128 because they enter the estimate.
// A complex product is done as a real GEMM on the interleaved (re,im) arrays:
//   G[2d][2k] = Vr^T Xr   gives  C = V^H X :  C_re[i][c] = G[2i][2c] + G[2i+1][2c+1],  C_im = G[2i][2c+1] - G[2i+1][2c]
//   Xr[n][2k] -= Vr[n][2d] C'[2d][2k]  with  C'[2i][2c] = cr, C'[2i+1][2c] = -ci, C'[2i][2c+1] = ci, C'[2i+1][2c+1] = cr
// Fragment layout of m8n8k4 (row.col): A[m][kk]: m = lane/4, kk = lane%4;  B[kk][n]: kk = lane%4, n = lane/4;
// D[m][n]: m = lane/4, n = 2*(lane%4) + {0,1}.  Shared-memory rows are padded by 4 doubles (conflict-free
// 64-bit fragment reads for row lengths that are multiples of 8).
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int DD_KB = 32;     // rows per shared-memory stage of the dot kernel
constexpr int DD_NT = 32;     // complex columns per CTA (8 warps x 4)

// partial[chunk][i][col] = sum_{r in chunk} conj(V[r][i]) X[r][col];  grid (ceil(k/32), nchunks), 256 threads
template <int MTMAX>          // M tiles available: 2d/8 <= MTMAX
__global__ void __launch_bounds__(256)
defl_dot_dmma_kernel(const Z* __restrict__ Vd, int d, const Z* __restrict__ X, int n, int k, int rows_per_chunk,
                     Z* __restrict__ partial) {
  extern __shared__ double dd_smem[];
  const int ldv = 2 * d + 4, ldx = 2 * DD_NT + 4;
  double* Vs = dd_smem;                       // [DD_KB][ldv]
  double* Xs = dd_smem + DD_KB * ldv;         // [DD_KB][ldx]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int col0 = blockIdx.x * DD_NT;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  const int mtn = (2 * d) / 8;
  double acc[MTMAX][2];
#pragma unroll
  for (int mt = 0; mt < MTMAX; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
  for (int rb = r0; rb < r1; rb += DD_KB) {
    for (int i = tid; i < DD_KB * d; i += 256) {
      const int rr = i / d, ii = i - rr * d;
      Z v = cx<double>(0.0, 0.0);
      if (rb + rr < r1) v = ldc_ro<double>(Vd, (size_t)(rb + rr) * d + ii);
      Vs[rr * ldv + 2 * ii] = v.re; Vs[rr * ldv + 2 * ii + 1] = v.im;
    }
    for (int i = tid; i < DD_KB * DD_NT; i += 256) {
      const int rr = i / DD_NT, cc = i - rr * DD_NT;
      Z v = cx<double>(0.0, 0.0);
      if (rb + rr < r1 && col0 + cc < k) v = ldc_ro<double>(X, (size_t)(rb + rr) * k + col0 + cc);
      Xs[rr * ldx + 2 * cc] = v.re; Xs[rr * ldx + 2 * cc + 1] = v.im;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < DD_KB / 4; ++ks) {
      const int kr = ks * 4 + (lane & 3);
      const double b = Xs[kr * ldx + w * 8 + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < MTMAX; ++mt)
        if (mt < mtn) dmma_884(acc[mt][0], acc[mt][1], Vs[kr * ldv + mt * 8 + (lane >> 2)], b);
    }
    __syncthreads();
  }
  const int col = col0 + w * 4 + (lane & 3);
#pragma unroll
  for (int mt = 0; mt < MTMAX; ++mt) {
    if (mt < mtn) {
      const double p0 = __shfl_xor_sync(0xffffffffu, acc[mt][0], 4);
      const double p1 = __shfl_xor_sync(0xffffffffu, acc[mt][1], 4);
      const int row = mt * 8 + (lane >> 2);            // row of G: 2i (Re part of V) or 2i+1 (Im part)
      if ((row & 1) == 0 && col < k)
        partial[((size_t)blockIdx.y * d + (row >> 1)) * k + col] = cx<double>(acc[mt][0] + p1, acc[mt][1] - p0);
    }
  }
}

// X[r][col] -= sum_i V[r][i] C[i][col];  grid (ceil(k/32), ceil(n/64)), 256 threads (warp = 8 rows)
__global__ void __launch_bounds__(256)
defl_axpy_dmma_kernel(const Z* __restrict__ Vd, int d, const Z* __restrict__ C, Z* __restrict__ X, int n, int k) {
  extern __shared__ double dd_smem[];
  const int ldv = 2 * d + 4, ldc = 2 * DD_NT + 4;
  double* Vs = dd_smem;                       // [64][ldv]
  double* Cs = dd_smem + 64 * ldv;            // [2d][ldc]   the real 2x2-block expansion of C
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int col0 = blockIdx.x * DD_NT, rbase = blockIdx.y * 64;
  for (int i = tid; i < 64 * d; i += 256) {
    const int rr = i / d, ii = i - rr * d;
    Z v = cx<double>(0.0, 0.0);
    if (rbase + rr < n) v = ldc_ro<double>(Vd, (size_t)(rbase + rr) * d + ii);
    Vs[rr * ldv + 2 * ii] = v.re; Vs[rr * ldv + 2 * ii + 1] = v.im;
  }
  for (int i = tid; i < d * DD_NT; i += 256) {
    const int ii = i / DD_NT, cc = i - ii * DD_NT;
    Z c = cx<double>(0.0, 0.0);
    if (col0 + cc < k) c = ldc_ro<double>(C, (size_t)ii * k + col0 + cc);
    Cs[(2 * ii) * ldc + 2 * cc] = c.re;      Cs[(2 * ii + 1) * ldc + 2 * cc] = -c.im;
    Cs[(2 * ii) * ldc + 2 * cc + 1] = c.im;  Cs[(2 * ii + 1) * ldc + 2 * cc + 1] = c.re;
  }
  __syncthreads();
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
  const int ksn = (2 * d) / 4;
  for (int ks = 0; ks < ksn; ++ks) {
    const double a = Vs[(w * 8 + (lane >> 2)) * ldv + ks * 4 + (lane & 3)];
    const double* crow = Cs + (ks * 4 + (lane & 3)) * ldc + (lane >> 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dmma_884(acc[nt][0], acc[nt][1], a, crow[nt * 8]);
  }
  const int row = rbase + w * 8 + (lane >> 2);
  if (row < n) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = col0 + nt * 4 + (lane & 3);
      if (col < k) {
        Z x = X[(size_t)row * k + col];
        x.re -= acc[nt][0]; x.im -= acc[nt][1];
        X[(size_t)row * k + col] = x;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// per-column FGMRES state (all arrays column-contiguous)
struct GmresState {
  int k, m;            // columns, restart length
  Z* hsum;             // [(m+1)][k]  Gram-Schmidt coefficients of the current step
  Z* nrm2;             // [k]         <w,w> after orthogonalisation
  Z* Rm;               // [m][m][k]   rotated Hessenberg (upper triangular), (i,j) at (j*m+i)*k+col
  double* cs;          // [m][k]
  Z* sn;               // [m][k]
  Z* g;                // [(m+1)][k]
  Z* y;                // [m][k]
  double* normb;       // [k]
  double* scale;       // [k]
  double* relres;      // [k]
  double* tolc;        // [k]         the Arnoldi estimate at which the column's current cycle stops (>= tol)
  int* active;         // [k]
  int* done;           // [k]
  int* it_cycle;       // [k]
  int* it_total;       // [k]
  int* n_active;       // [1]
};

// start of a cycle: nrm2 = <r,r> of the residual the cycle starts from.
//   mode 1 (first cycle): r = b (x0 = 0); normb = ||b||.
//   mode 2 (later cycles): r is the TRUE residual b - A x of every column, so this is also the
//   verification of the columns that stopped on the Arnoldi estimate: a column is finished iff its
//   true relative residual is below tol (with a rounding allowance), otherwise it iterates on.
// drop > 0 (mixed-precision refinement): a cycle only runs until the estimate has fallen by that factor relative to the
// residual it started from -- what the complex64-stored basis can resolve -- and the next cycle restarts from the true residual.
__global__ void __launch_bounds__(256)
gmres_init_kernel(GmresState s, double tol, int mode, double drop = 0.0) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= s.k) return;
  const double nr = sqrt(fmax(s.nrm2[col].re, 0.0));
  double accept = tol;
  if (mode == 1) {
    s.normb[col] = (nr == 0.0) ? 1.0 : nr;
    s.it_total[col] = 0;
  } else if (mode == 3) {         // first cycle, normb supplied by the caller (the even-odd solve measures against ||b|| of the full system)
    s.it_total[col] = 0;
  } else {
    accept = 1.25 * tol + 1e-14;
  }
  const double rel = nr / s.normb[col];
  s.it_cycle[col] = 0;
  s.relres[col] = rel;
  s.tolc[col] = fmax(tol, drop * rel);
  const int act = (rel < accept || nr == 0.0) ? 0 : 1;
  s.done[col] = !act;
  s.active[col] = act;
  s.scale[col] = act ? 1.0 / nr : 0.0;
  s.g[col] = cx<double>(act ? nr : 0.0, 0.0);
  if (act) atomicAdd(s.n_active, 1);
}

// after step j: apply the stored rotations to the new Hessenberg column, make the new one,
// update g and the residual estimate, decide convergence.            (Saad, Alg. 6.9 / pyamg fgmres)
__global__ void __launch_bounds__(128)
gmres_step_kernel(GmresState s, int j, double tol) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= s.k) return;
  if (!s.active[col]) { s.scale[col] = 0.0; return; }
  const int k = s.k, m = s.m;
  const double hn = sqrt(fmax(s.nrm2[col].re, 0.0));
  Z hi = s.hsum[col];
  for (int i = 0; i < j; ++i) {
    const Z hn1 = s.hsum[(size_t)(i + 1) * k + col];
    const double c = s.cs[(size_t)i * k + col];
    const Z sn = s.sn[(size_t)i * k + col];
    // t = c*h_i + s*h_{i+1} ; h_{i+1} = -conj(s)*h_i + c*h_{i+1}
    Z t = zscale(c, hi); zfma(t, sn, hn1);
    Z u = zscale(c, hn1);
    Z ms = cx<double>(-sn.re, sn.im);     // -conj(s)
    zfma(u, ms, hi);
    s.Rm[((size_t)j * m + i) * k + col] = t;
    hi = u;
  }
  // hi = h_j (rotated), b = hn (real >= 0)
  const double aa = sqrt(zabs2(hi));
  const double den = sqrt(aa * aa + hn * hn);
  double c; Z sn;
  if (den == 0.0) { c = 1.0; sn = cx<double>(0.0, 0.0); }
  else if (aa == 0.0) { c = 0.0; sn = cx<double>(1.0, 0.0); }
  else { c = aa / den; sn = zscale(hn / (den * aa), hi); }
  Z rjj = zscale(c, hi);
  rjj.re += sn.re * hn; rjj.im += sn.im * hn;
  s.Rm[((size_t)j * m + j) * k + col] = rjj;
  s.cs[(size_t)j * k + col] = c;
  s.sn[(size_t)j * k + col] = sn;
  const Z gj = s.g[(size_t)j * k + col];
  Z gn = cx<double>(0.0, 0.0);
  zfma(gn, cx<double>(-sn.re, sn.im), gj);
  s.g[(size_t)(j + 1) * k + col] = gn;
  s.g[(size_t)j * k + col] = zscale(c, gj);
  const double rel = sqrt(zabs2(gn)) / s.normb[col];
  s.relres[col] = rel;
  s.it_cycle[col] = j + 1;
  s.it_total[col] += 1;
  if (rel < s.tolc[col]) {
    s.active[col] = 0; s.done[col] = 1; s.scale[col] = 0.0;
  } else {
    s.scale[col] = (hn > 0.0) ? 1.0 / hn : 0.0;
    atomicAdd(s.n_active, 1);
  }
}

// y = R^{-1} g for the steps each column actually took; zero beyond
__global__ void __launch_bounds__(128)
gmres_solve_kernel(GmresState s, int jmax) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= s.k) return;
  const int k = s.k, m = s.m;
  const int mc = s.it_cycle[col];
  for (int i = mc; i < jmax; ++i) s.y[(size_t)i * k + col] = cx<double>(0.0, 0.0);
  for (int i = mc - 1; i >= 0; --i) {
    Z acc = s.g[(size_t)i * k + col];
    for (int l = i + 1; l < mc; ++l) {
      const Z r = s.Rm[((size_t)l * m + i) * k + col];
      const Z yl = s.y[(size_t)l * k + col];
      zfma(acc, cx<double>(-r.re, -r.im), yl);
    }
    const Z d = s.Rm[((size_t)i * m + i) * k + col];
    const double d2 = zabs2(d);
    Z yi = cx<double>(0.0, 0.0);
    if (d2 > 0.0) { zfma_conj(yi, d, acc); yi = zscale(1.0 / d2, yi); }   // acc / d
    s.y[(size_t)i * k + col] = yi;
  }
}

// Out[i] = A[i] - B[i]
__global__ void __launch_bounds__(256)
vec_sub_kernel(const Z* __restrict__ A, const Z* __restrict__ Bv, Z* __restrict__ Out, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) Out[i] = zsub(ldc_ro<double>(A, i), ldc_ro<double>(Bv, i));
}
// normb[col] = sqrt(nrm2[col]) (1 for a zero column)
__global__ void set_normb_kernel(const Z* __restrict__ nrm2, double* __restrict__ normb, int k) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < k) { const double v = sqrt(fmax(nrm2[col].re, 0.0)); normb[col] = (v == 0.0) ? 1.0 : v; }
}

// e[col] = a[col] - b[col]
__global__ void zsub_kernel(const Z* a, const Z* b, Z* e, int k) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < k) e[col] = zsub(a[col], b[col]);
}

}  // namespace dmlmc
