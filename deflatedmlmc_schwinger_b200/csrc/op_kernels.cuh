// op_kernels.cuh -- operator kernels on column batches X[n][k] (k contiguous).
//   stencil_kernel : level-0 Wilson-Dirac operator in link form      (multigrid.py:552-557 on A_0)
//   bsr_kernel     : coarse operators A_l = R A P as padded block rows (multigrid.py:276, :552-557)
//   restrict/prolong kernels : R_l = P_l^H and P_l                   (multigrid.py:406,429)
//   dense_kernel   : coarsest_inv * B                                (multigrid.py:413-416)
//   perm_kernel    : Bblock_perm * roll(x, +shift)                   (utils.py:232,288-290)
//   probe_expand_kernel : packed MT19937 bits -> +-1 complex128      (utils.py:213-216)
// All of them are HBM-streaming kernels: one thread owns one 16-byte pack of columns of one
// lattice site / block row, so a warp reads and writes 512 contiguous bytes per row.
#pragma once
#include "common.cuh"

namespace dmlmc {

// what the operator kernels write
enum { M_AX = 0,         // Y = A X
       M_RES = 1,        // Y = B - A X
       M_STEP = 2,       // Y = X - w A X                 (one factor of the smoother polynomial)
       M_STEP_SET = 3,   // Y = c (X - w A X)             (last factor, pre-smoother:  x  = p(A) b)
       M_STEP_ACC = 4 }; // Y += c (X - w A X)            (last factor, post-smoother: x += p(A) r)

template <typename T> struct StencilDev {
  int LX, LT;
  const Cx<T>* Ut;   // [LX*LT]  U_t(x,t)
  const Cx<T>* Ux;   // [LX*LT]  U_x(x,t)
  Cx<T> diag;        // 4 + m
};

template <typename T, int NC, int MODE>
__device__ __forceinline__ void op_epilogue(const Pack<T, NC>& ax, const Pack<T, NC>& xin, size_t idx,
                                            const Pack<T, NC>* __restrict__ B, Pack<T, NC>* __restrict__ Y,
                                            Cx<T> w, Cx<T> c) {
  if constexpr (MODE == M_AX) {
    Y[idx] = ax;
  } else if constexpr (MODE == M_RES) {
    Y[idx] = psub<T, NC>(ldp_ro<T, NC>(B, idx), ax);
  } else {
    Pack<T, NC> rn = xin;
    pfms<T, NC>(rn, w, ax);
    if constexpr (MODE == M_STEP) {
      Y[idx] = rn;
    } else if constexpr (MODE == M_STEP_SET) {
      Y[idx] = pscale<T, NC>(c, rn);
    } else {
      Pack<T, NC> e = Y[idx];
      pfma<T, NC>(e, c, rn);
      Y[idx] = e;
    }
  }
}

// A psi(x) = diag psi(x) - [ (1-s1) Ut(x) psi(x+t) + (1+s1) Ut(x-t)^* psi(x-t)
//                          + (1-s2) Ux(x) psi(x+x) + (1+s2) Ux(x-x)^* psi(x-x) ]
// with the spin projections (1-s1)phi = (a,-a), a = phi0-phi1; (1+s1)phi = (b,b), b = phi0+phi1;
// (1-s2)phi = (c,-ic), c = phi0+i phi1; (1+s2)phi = (d, id), d = phi0-i phi1.
// Thread block = (packs, TT sites in t, TX sites in x): the 2-D site tile makes the x- and t-neighbour
// rows of a site hit in L1 (fetched by the CTA's own threads), so each vector row is pulled from L2
// (TT*TX + 2TT + 2TX)/(TT*TX) times instead of 5.
// MINB = 2: 64 registers, all 14 loads of a thread in flight at once; MINB = 3: 40 registers, loads in
// batches of 6 but 50 % more resident threads.
template <typename T, int NC, int MODE, int MINB>
__global__ void __launch_bounds__(512, MINB)
stencil_kernel(StencilDev<T> op, const Pack<T, NC>* __restrict__ X, const Pack<T, NC>* __restrict__ B,
               Pack<T, NC>* __restrict__ Y, Cx<T> w, Cx<T> cfin, int kp) {
  const int LX = op.LX, LT = op.LT, V = LX * LT;
  const int cp = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z * blockDim.z + threadIdx.z;
  if (cp >= kp || t >= LT || x >= LX) return;
  const int site = x * LT + t;
  const int tp = (t + 1 == LT) ? 0 : t + 1, tm = (t == 0) ? LT - 1 : t - 1;
  const int xp = (x + 1 == LX) ? 0 : x + 1, xm = (x == 0) ? LX - 1 : x - 1;
  const int s_tp = x * LT + tp, s_tm = x * LT + tm, s_xp = xp * LT + t, s_xm = xm * LT + t;
  const size_t kpz = (size_t)kp, Vz = (size_t)V;

  typedef Pack<T, NC> P;
  const P c0 = ldp_ro<T, NC>(X, (size_t)site * kpz + cp),  c1 = ldp_ro<T, NC>(X, (Vz + site) * kpz + cp);
  const P f0 = ldp_ro<T, NC>(X, (size_t)s_tp * kpz + cp),  f1 = ldp_ro<T, NC>(X, (Vz + s_tp) * kpz + cp);
  const P b0 = ldp_ro<T, NC>(X, (size_t)s_tm * kpz + cp),  b1 = ldp_ro<T, NC>(X, (Vz + s_tm) * kpz + cp);
  const P r0 = ldp_ro<T, NC>(X, (size_t)s_xp * kpz + cp),  r1 = ldp_ro<T, NC>(X, (Vz + s_xp) * kpz + cp);
  const P l0 = ldp_ro<T, NC>(X, (size_t)s_xm * kpz + cp),  l1 = ldp_ro<T, NC>(X, (Vz + s_xm) * kpz + cp);
  const Cx<T> ut = ldc_ro<T>(op.Ut, site), utb = cconj(ldc_ro<T>(op.Ut, s_tm));
  const Cx<T> ux = ldc_ro<T>(op.Ux, site), uxb = cconj(ldc_ro<T>(op.Ux, s_xm));

  const P a = psub<T, NC>(f0, f1);                       // (1-s1) psi(x+t) = (a,-a)
  const P b = padd<T, NC>(b0, b1);                       // (1+s1) psi(x-t) = (b, b)
  const P c = padd<T, NC>(r0, pmul_i<T, NC>(r1));        // (1-s2) psi(x+x) = (c,-ic)
  const P d = psub<T, NC>(l0, pmul_i<T, NC>(l1));        // (1+s2) psi(x-x) = (d, id)
  const P ua = pscale<T, NC>(ut, a), ub = pscale<T, NC>(utb, b);
  const P uc = pscale<T, NC>(ux, c), ud = pscale<T, NC>(uxb, d);

  P y0 = pscale<T, NC>(op.diag, c0), y1 = pscale<T, NC>(op.diag, c1);
  y0 = psub<T, NC>(y0, padd<T, NC>(padd<T, NC>(ua, ub), padd<T, NC>(uc, ud)));
  // spin 1: -ua + ub - i uc + i ud
  y1 = psub<T, NC>(y1, padd<T, NC>(psub<T, NC>(ub, ua), pmul_i<T, NC>(psub<T, NC>(ud, uc))));

  op_epilogue<T, NC, MODE>(y0, c0, (size_t)site * kpz + cp, B, Y, w, cfin);
  op_epilogue<T, NC, MODE>(y1, c1, (Vz + site) * kpz + cp, B, Y, w, cfin);
}

// ------------------------------------------------------------------------------------------
template <typename T> struct BsrDev {
  int nb, bpr;
  const int* col;      // [nb][bpr] block column or -1
  const Cx<T>* vals;   // [nb][bpr][BS][BS]
};

template <typename T, int NC, int BS, int MODE>
__global__ void __launch_bounds__(128)
bsr_kernel(BsrDev<T> op, const Pack<T, NC>* __restrict__ X, const Pack<T, NC>* __restrict__ B,
           Pack<T, NC>* __restrict__ Y, Cx<T> w, Cx<T> cfin, int kp) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int I = (int)(gid / kp);
  const int cp = (int)(gid - (long long)I * kp);
  if (I >= op.nb) return;
  const size_t kpz = (size_t)kp;
  typedef Pack<T, NC> P;
  P acc[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) acc[r] = pzero<T, NC>();
  for (int blk = 0; blk < op.bpr; ++blk) {
    const int J = __ldg(op.col + (size_t)I * op.bpr + blk);
    if (J < 0) continue;
    P xv[BS];
#pragma unroll
    for (int c = 0; c < BS; ++c) xv[c] = ldp_ro<T, NC>(X, ((size_t)J * BS + c) * kpz + cp);
    const Cx<T>* vb = op.vals + ((size_t)I * op.bpr + blk) * (BS * BS);
#pragma unroll
    for (int r = 0; r < BS; ++r) {
#pragma unroll
      for (int c = 0; c < BS; ++c) pfma<T, NC>(acc[r], ldc_ro<T>(vb, r * BS + c), xv[c]);
    }
  }
#pragma unroll
  for (int r = 0; r < BS; ++r) {
    const size_t idx = ((size_t)I * BS + r) * kpz + cp;
    P xin = pzero<T, NC>();
    if constexpr (MODE >= M_STEP) xin = ldp_ro<T, NC>(X, idx);
    op_epilogue<T, NC, MODE>(acc[r], xin, idx, B, Y, w, cfin);
  }
}

// ------------------------------------------------------------------------------------------
// complex64 coarse operator, Blackwell path: packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2, sm_100+)
// with the block row's matrix staged in shared memory as (mr, mr, mi, mi) per entry, so the inner
// loop is one broadcast LDS.128 per matrix entry feeding 4*PPT FFMA2 and no shuffles/swaps:
//   a += (mr,mr)*(xr,xi)   b += (mi,mi)*(xr,xi)   =>   (A x).re = a.re - b.im, (A x).im = a.im + b.re
// blockDim = (tpr, RB): RB block rows per CTA, thread tx owns PPT packs (2 columns each) tx + p*tpr.
__device__ __forceinline__ void fma2x(float4& acc, float2 m, const float4& x) {
  const float2 lo = __ffma2_rn(m, make_float2(x.x, x.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(m, make_float2(x.z, x.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int BS, int PPT, int MODE>
__global__ void __launch_bounds__(128)
bsr_f32x2_kernel(int nb, int bpr, const int* __restrict__ col, const float4* __restrict__ vals4,
                 const Pack<float, 2>* __restrict__ X, const Pack<float, 2>* __restrict__ B,
                 Pack<float, 2>* __restrict__ Y, Cx<float> w, Cx<float> cfin, int kp) {
  extern __shared__ float4 bsr_smem[];
  const int tpr = blockDim.x, RB = blockDim.y;
  const int tx = threadIdx.x, rb = threadIdx.y;
  const int nent = bpr * BS * BS;
  float4* vs = bsr_smem;
  int* cs = reinterpret_cast<int*>(bsr_smem + (size_t)RB * nent);
  const int I0 = blockIdx.x * RB;
  const int tid = rb * tpr + tx, nthr = tpr * RB;
  for (int i = tid; i < RB * nent; i += nthr) {
    const int rbi = i / nent, Ii = I0 + rbi;
    vs[i] = (Ii < nb) ? __ldg(vals4 + (size_t)Ii * nent + (i - rbi * nent)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = tid; i < RB * bpr; i += nthr) {
    const int rbi = i / bpr, Ii = I0 + rbi;
    cs[i] = (Ii < nb) ? __ldg(col + (size_t)Ii * bpr + (i - rbi * bpr)) : -1;
  }
  __syncthreads();
  const int I = I0 + rb;
  if (I >= nb) return;
  const size_t kpz = (size_t)kp;
  int cp[PPT]; bool ok[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int c = blockIdx.y * (tpr * PPT) + tx + p * tpr;
    ok[p] = c < kp;
    cp[p] = ok[p] ? c : kp - 1;
  }
  float4 a[BS][PPT], b[BS][PPT];
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int p = 0; p < PPT; ++p) { a[r][p] = make_float4(0.f, 0.f, 0.f, 0.f); b[r][p] = make_float4(0.f, 0.f, 0.f, 0.f); }
  const float4* vrow = vs + (size_t)rb * nent;
  const int* crow = cs + rb * bpr;
  const float4* X4 = reinterpret_cast<const float4*>(X);
  for (int blk = 0; blk < bpr; ++blk) {
    const int J = crow[blk];
    if (J < 0) continue;
    float4 xv[BS][PPT];
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int p = 0; p < PPT; ++p) xv[c][p] = __ldg(X4 + ((size_t)J * BS + c) * kpz + cp[p]);
    const float4* vb = vrow + blk * (BS * BS);
#pragma unroll
    for (int r = 0; r < BS; ++r) {
#pragma unroll
      for (int c = 0; c < BS; ++c) {
        const float4 m = vb[r * BS + c];
        const float2 mr = make_float2(m.x, m.y), mi = make_float2(m.z, m.w);
#pragma unroll
        for (int p = 0; p < PPT; ++p) { fma2x(a[r][p], mr, xv[c][p]); fma2x(b[r][p], mi, xv[c][p]); }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < BS; ++r) {
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (!ok[p]) continue;
      Pack<float, 2> ax;
      ax.d[0] = a[r][p].x - b[r][p].y; ax.d[1] = a[r][p].y + b[r][p].x;
      ax.d[2] = a[r][p].z - b[r][p].w; ax.d[3] = a[r][p].w + b[r][p].z;
      const size_t idx = ((size_t)I * BS + r) * kpz + cp[p];
      Pack<float, 2> xin = pzero<float, 2>();
      if constexpr (MODE >= M_STEP) xin = ldp_ro<float, 2>(X, idx);
      op_epilogue<float, 2, MODE>(ax, xin, idx, B, Y, w, cfin);
    }
  }
}

// ------------------------------------------------------------------------------------------
// aggregation transfer operators.  Row r of P_l: aggregate j = r / aggr, half = ((r % aggr) % dofi) >= h,
// columns (2j + half)*NV + [0,NV).                                   (multigrid.py:203-227)
template <typename T> struct TransferDev {
  int n_f, n_c, aggr, dofi, h, nvec;
  const Cx<T>* pv;     // [n_f][nvec]
};

template <typename T, int NC, int NV>
__global__ void __launch_bounds__(128)
restrict_kernel(TransferDev<T> tr, const Pack<T, NC>* __restrict__ Xf, Pack<T, NC>* __restrict__ Xc, int kp,
                int ldf, int ldc) {   // kp packs per row are processed; rows are ldf / ldc packs apart
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = (int)(gid / kp);                   // (aggregate, half)
  const int cp = (int)(gid - (long long)g * kp);
  if (g >= tr.n_c / NV) return;
  const int j = g >> 1, half = g & 1;
  const size_t ldfz = (size_t)ldf, ldcz = (size_t)ldc;
  typedef Pack<T, NC> P;
  P acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = pzero<T, NC>();
  const int nw = tr.aggr / tr.dofi;
  for (int wq = 0; wq < nw; ++wq) {
    for (int z = 0; z < tr.h; ++z) {
      const int r = j * tr.aggr + wq * tr.dofi + half * tr.h + z;
      const P x = ldp_ro<T, NC>(Xf, (size_t)r * ldfz + cp);
#pragma unroll
      for (int v = 0; v < NV; ++v) pfma_conj<T, NC>(acc[v], ldc_ro<T>(tr.pv, (size_t)r * NV + v), x);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) Xc[((size_t)g * NV + v) * ldcz + cp] = acc[v];
}

template <typename T, int NC, int NV>
__global__ void __launch_bounds__(256)
prolong_add_kernel(TransferDev<T> tr, const Pack<T, NC>* __restrict__ Xc, Pack<T, NC>* __restrict__ Xf, int kp,
                   int ldf, int ldc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(gid / kp);
  const int cp = (int)(gid - (long long)r * kp);
  if (r >= tr.n_f) return;
  const int j = r / tr.aggr;
  const int q = (r - j * tr.aggr) % tr.dofi;
  const int g = 2 * j + (q >= tr.h ? 1 : 0);
  const size_t ldfz = (size_t)ldf, ldcz = (size_t)ldc;
  typedef Pack<T, NC> P;
  P acc = Xf[(size_t)r * ldfz + cp];
#pragma unroll
  for (int v = 0; v < NV; ++v)
    pfma<T, NC>(acc, ldc_ro<T>(tr.pv, (size_t)r * NV + v), ldp_ro<T, NC>(Xc, ((size_t)g * NV + v) * ldcz + cp));
  Xf[(size_t)r * ldfz + cp] = acc;
}

// ------------------------------------------------------------------------------------------
// X[n][k] = M[n][n] B[n][k], M row-major.  32x(32 packs) output tile per CTA, 32-deep k-chunks
// staged in shared memory; thread (tx,ty) owns rows ty, ty+8, ty+16, ty+24 of pack column tx.
template <typename T, int NC>
__global__ void __launch_bounds__(256)
dense_kernel(const Cx<T>* __restrict__ M, int n, const Pack<T, NC>* __restrict__ B, Pack<T, NC>* __restrict__ X, int kp) {
  __shared__ Cx<T> Ms[32][33];
  __shared__ Pack<T, NC> Bs[32][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int cp = blockIdx.x * 32 + tx;
  const int i0 = blockIdx.y * 32;
  typedef Pack<T, NC> P;
  P acc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r] = pzero<T, NC>();
  for (int j0 = 0; j0 < n; j0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r, j = j0 + tx;
      Ms[ty + 8 * r][tx] = (i < n && j < n) ? M[(size_t)i * n + j] : cx<T>(T(0), T(0));
      const int jr = j0 + ty + 8 * r;
      Bs[ty + 8 * r][tx] = (jr < n && cp < kp) ? B[(size_t)jr * kp + cp] : pzero<T, NC>();
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const P b = Bs[jj][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) pfma<T, NC>(acc[r], Ms[ty + 8 * r][jj], b);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r;
    if (i < n && cp < kp) X[(size_t)i * kp + cp] = acc[r];
  }
}

// complex64 dense apply with packed FP32 FMA: M4[n][n] holds (mr,mr,mi,mi) per entry.  32 x (32 packs)
// output tile per CTA, thread (tx,ty) owns rows ty+8r (r<4) of pack tx; per k-step 1 LDS.128 of B and
// 4 broadcast LDS.128 of M feed 16 FFMA2.
__global__ void __launch_bounds__(256)
dense_f32x2_kernel(const float4* __restrict__ M4, int n, const Pack<float, 2>* __restrict__ B,
                   Pack<float, 2>* __restrict__ X, int kp) {
  __shared__ float4 Ms[32][33];
  __shared__ float4 Bs[32][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int cp = blockIdx.x * 32 + tx;
  const int i0 = blockIdx.y * 32;
  const float4* B4 = reinterpret_cast<const float4*>(B);
  float4 a[4], b[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { a[r] = make_float4(0.f, 0.f, 0.f, 0.f); b[r] = make_float4(0.f, 0.f, 0.f, 0.f); }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < n; j0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r, j = j0 + tx;
      Ms[ty + 8 * r][tx] = (i < n && j < n) ? __ldg(M4 + (size_t)i * n + j) : z4;
      const int jr = j0 + ty + 8 * r;
      Bs[ty + 8 * r][tx] = (jr < n && cp < kp) ? __ldg(B4 + (size_t)jr * kp + cp) : z4;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const float4 x = Bs[jj][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 m = Ms[ty + 8 * r][jj];
        fma2x(a[r], make_float2(m.x, m.y), x);
        fma2x(b[r], make_float2(m.z, m.w), x);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r;
    if (i < n && cp < kp) {
      Pack<float, 2> o;
      o.d[0] = a[r].x - b[r].y; o.d[1] = a[r].y + b[r].x; o.d[2] = a[r].z - b[r].w; o.d[3] = a[r].w + b[r].z;
      X[(size_t)i * kp + cp] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// RHS[i] = sum_j vals[i][j] X[(cols[i][j] - shift) mod n]   (nnz == 0: RHS[i] = X[(i - shift) mod n])
template <typename T, int NC>
__global__ void __launch_bounds__(256)
perm_kernel(int n, int shift, int nnz, const int* __restrict__ cols, const Cx<T>* __restrict__ vals,
            const Pack<T, NC>* __restrict__ X, Pack<T, NC>* __restrict__ Y, int kp) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gid / kp);
  const int cp = (int)(gid - (long long)i * kp);
  if (i >= n) return;
  const size_t kpz = (size_t)kp;
  if (nnz == 0) {
    int src = i - shift; if (src < 0) src += n;
    Y[(size_t)i * kpz + cp] = ldp_ro<T, NC>(X, (size_t)src * kpz + cp);
    return;
  }
  Pack<T, NC> acc = pzero<T, NC>();
  for (int j = 0; j < nnz; ++j) {
    const int c = __ldg(cols + (size_t)i * nnz + j);
    if (c < 0) continue;
    int src = c - shift; if (src < 0) src += n;
    pfma<T, NC>(acc, ldc_ro<T>(vals, (size_t)i * nnz + j), ldp_ro<T, NC>(X, (size_t)src * kpz + cp));
  }
  Y[(size_t)i * kpz + cp] = acc;
}

// element i of probe p = 2*bit(p*n+i) - 1                       (utils.py:213-216)
__global__ void __launch_bounds__(256)
probe_expand_kernel(const uint8_t* __restrict__ bits, int n, int k, Cx<double>* __restrict__ X0) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gid / k);
  const int p = (int)(gid - (long long)i * k);
  if (i >= n) return;
  const long long j = (long long)p * n + i;
  const int bit = (bits[j >> 3] >> (j & 7)) & 1;
  X0[(size_t)i * k + p] = cx<double>(bit ? 1.0 : -1.0, 0.0);
}

// Out[r][c] = (Tout) In[r][c], r < n, c < w: copies / converts a block of w columns between two
// row-major batches with leading dimensions ld_in / ld_out (complex elements).  This is how a column
// chunk of the FGMRES basis (complex128, [n][k]) becomes the compact, L2-resident [n][w] working
// array of the V-cycle and back.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
cvt_cols_kernel(const Cx<Tin>* __restrict__ in, size_t ld_in, Cx<Tout>* __restrict__ out, size_t ld_out, int n, int w) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t r = idx / w;
  const int c = (int)(idx - r * w);
  if (r >= (size_t)n) return;
  const Cx<Tin> v = ldc_ro<Tin>(in, r * ld_in + c);
  out[r * ld_out + c] = cx<Tout>((Tout)v.re, (Tout)v.im);
}

// Y = c X (elementwise) / Y += c X
template <typename T, int ACC>
__global__ void __launch_bounds__(256) scale_kernel(Cx<T> c, const Cx<T>* __restrict__ X, Cx<T>* __restrict__ Y, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Pack<T, 1> y = ACC ? *reinterpret_cast<const Pack<T, 1>*>(Y + i) : pzero<T, 1>();
  pfma<T, 1>(y, c, *reinterpret_cast<const Pack<T, 1>*>(X + i));
  *reinterpret_cast<Pack<T, 1>*>(Y + i) = y;
}

}  // namespace dmlmc
