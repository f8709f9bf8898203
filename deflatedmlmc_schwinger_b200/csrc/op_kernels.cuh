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
#include <cuda_bf16.h>

namespace dmlmc {

// what the operator kernels write
enum { M_AX = 0,         // Y = A X
       M_RES = 1,        // Y = B - A X
       M_STEP = 2,       // Y = c (X - w A X)     one factor of the smoother polynomial (c = 1 except at the ends)
       M_STEP_ACC = 3 }; // Y += c (X - w A X)    last factor of the post-smoother: x += p(A) r

template <typename T> struct StencilDev {
  int LX, LT;
  const Cx<T>* Ut;   // [LX*LT]  U_t(x,t)
  const Cx<T>* Ux;   // [LX*LT]  U_x(x,t)
  Cx<T> diag;        // 4 + m
};

// value an operator kernel writes at idx, given (A X)[idx] and X[idx]
template <typename T, int NC, int MODE>
__device__ __forceinline__ Pack<T, NC> op_value(const Pack<T, NC>& ax, const Pack<T, NC>& xin, size_t idx,
                                                const Pack<T, NC>* __restrict__ B, const Pack<T, NC>* Yold,
                                                Cx<T> w, Cx<T> c) {
  if constexpr (MODE == M_AX) {
    return ax;
  } else if constexpr (MODE == M_RES) {
    return psub<T, NC>(ldp_ro<T, NC>(B, idx), ax);
  } else {
    Pack<T, NC> rn = xin;
    pfms<T, NC>(rn, w, ax);
    if constexpr (MODE == M_STEP) {
      return pscale<T, NC>(c, rn);
    } else {
      Pack<T, NC> e = Yold[idx];
      pfma<T, NC>(e, c, rn);
      return e;
    }
  }
}
template <typename T, int NC, int MODE>
__device__ __forceinline__ void op_epilogue(const Pack<T, NC>& ax, const Pack<T, NC>& xin, size_t idx,
                                            const Pack<T, NC>* __restrict__ B, Pack<T, NC>* __restrict__ Y,
                                            Cx<T> w, Cx<T> c) {
  Y[idx] = op_value<T, NC, MODE>(ax, xin, idx, B, Y, w, c);
}

// two complex64 columns stored as BF16 (8 bytes): the storage format of the level-0 smoother's
// intermediate vectors (arithmetic stays FP32; QUDA-style reduced-precision preconditioner storage).
// BF16 rather than FP16 because BF16 -> FP32 is a shift / mask on the integer pipe, while FP16 -> FP32
// (HADD2.F32) competes with the FMAs for the FP32 pipe, which is what bounds the kernel (ncu:
// fmaheavy pipe 65 % active, 40 of 190 instructions were conversions).
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ Pack<float, 2> ldh2_ro(const void* base, size_t idx) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(base) + idx);
  Pack<float, 2> r; r.d[0] = bf_lo(raw.x); r.d[1] = bf_hi(raw.x); r.d[2] = bf_lo(raw.y); r.d[3] = bf_hi(raw.y);
  return r;
}
__device__ __forceinline__ void sth2(void* base, size_t idx, const Pack<float, 2>& v) {
  reinterpret_cast<uint2*>(base)[idx] = make_uint2(bf_pack(v.d[0], v.d[1]), bf_pack(v.d[2], v.d[3]));
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
// HIN / HOUT: X / Y are BF16-stored (complex64 arithmetic, NC = 2 only).
// DOUT (last factor of the post-smoother, MODE = M_STEP_ACC, complex64): the result Yold + c (X - w A X) is written as
// complex128 to the array passed in the (otherwise unused) B argument instead of back to Y -- the V-cycle's output
// lands in the outer solver's complex128 vector without a separate conversion pass.
// EOUT (with HOUT): the BF16 output goes to the checkerboard-compressed half-lattice layout of wilson_hop_eo_kernel
// (parity p = (x + t) & 1 array first the even, then the odd sites: [p][s][x][t/2][kp]).
template <typename T, int NC, int MODE, int MINB, bool HIN = false, bool HOUT = false, int PFD = 0, bool DOUT = false, bool EOUT = false>
__global__ void __launch_bounds__(512, MINB)
stencil_kernel(StencilDev<T> op, const void* __restrict__ Xv, const Pack<T, NC>* __restrict__ B,
               void* __restrict__ Yv, Cx<T> w, Cx<T> cfin, int kp) {
  static_assert(!(HIN || HOUT) || (sizeof(T) == 4 && NC == 2), "BF16 storage is a complex64, 2-column format");
  const Pack<T, NC>* __restrict__ X = reinterpret_cast<const Pack<T, NC>*>(Xv);
  Pack<T, NC>* __restrict__ Y = reinterpret_cast<Pack<T, NC>*>(Yv);
  auto ldx = [&](size_t idx) -> Pack<T, NC> {
    if constexpr (HIN) return ldh2_ro(Xv, idx); else return ldp_ro<T, NC>(X, idx);
  };
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
  // (a byte-offset / 64-bit-add addressing scheme has 25 % fewer instructions but measured 5-14 % slower
  //  on B200 for the complex128 and complex64 Y = A X kernels: profiles/r1_run7_kernel_timings.jsonl)

  typedef Pack<T, NC> P;
  const size_t i0 = (size_t)site * kpz + cp, i1 = (Vz + site) * kpz + cp;
  if constexpr (PFD > 0) {
    // software prefetch into L2 of the rows that the blocks PFD x-slices ahead will read (streaming kernels:
    // the loads of a thread then find their lines in L2 instead of waiting for HBM)
    const int xq = x + PFD;
    if (xq < LX) {
      const size_t q0 = ((size_t)xq * LT + t) * kpz + cp;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(X + q0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(X + q0 + Vz * kpz));
    }
  }
  const P c0 = ldx(i0),  c1 = ldx(i1);
  const P f0 = ldx((size_t)s_tp * kpz + cp),  f1 = ldx((Vz + s_tp) * kpz + cp);
  const P b0 = ldx((size_t)s_tm * kpz + cp),  b1 = ldx((Vz + s_tm) * kpz + cp);
  const P r0 = ldx((size_t)s_xp * kpz + cp),  r1 = ldx((Vz + s_xp) * kpz + cp);
  const P l0 = ldx((size_t)s_xm * kpz + cp),  l1 = ldx((Vz + s_xm) * kpz + cp);
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

  const P o0 = op_value<T, NC, MODE>(y0, c0, i0, B, Y, w, cfin);
  const P o1 = op_value<T, NC, MODE>(y1, c1, i1, B, Y, w, cfin);
  if constexpr (DOUT) {
    static_assert(MODE == M_STEP_ACC && sizeof(T) == 4 && NC == 2 && !HOUT, "complex128 output: last complex64 factor only");
    double2* Yd = reinterpret_cast<double2*>(const_cast<Pack<T, NC>*>(B));
    Yd[2 * i0] = make_double2((double)o0.d[0], (double)o0.d[1]); Yd[2 * i0 + 1] = make_double2((double)o0.d[2], (double)o0.d[3]);
    Yd[2 * i1] = make_double2((double)o1.d[0], (double)o1.d[1]); Yd[2 * i1 + 1] = make_double2((double)o1.d[2], (double)o1.d[3]);
  } else if constexpr (HOUT && EOUT) {
    const size_t par = (size_t)((x + t) & 1), half = (size_t)(V >> 1);
    const size_t e0 = ((par * 2) * half + (size_t)x * (LT >> 1) + (t >> 1)) * kpz + cp;
    sth2(Yv, e0, o0); sth2(Yv, e0 + half * kpz, o1);
  } else if constexpr (HOUT) { sth2(Yv, i0, o0); sth2(Yv, i1, o1); } else { Y[i0] = o0; Y[i1] = o1; }
}

// ------------------------------------------------------------------------------------------
// The hot kernel of the complex64 V-cycle: one factor Y = X - w A X of the level-0 smoother with
// BF16-stored vectors.  The generic stencil_kernel is instruction-issue bound (ncu: 80 % issue-active,
// 313 instructions per thread, DRAM and L2 below 20 % because the 16-bit ping-pong vectors live in L2),
// so this version minimises instructions:
//   * the two columns of a thread are held as (re0,re1) / (im0,im1) register pairs, so every
//     complex operation is packed FP32 (FADD2 / FMUL2 / FFMA2, Blackwell) for both columns at once;
//   * links come pre-splatted and pre-conjugated, L4[dir][site] = (ur,ur,ui,ui) for
//     dir = U_t(x), U_t(x-t)^*, U_x(x), U_x(x-x)^*: one LDG.128 per direction, no splat/negate code;
//   * 32-bit element indices (one IMAD.WIDE per address).
struct C2 { float2 re, im; };     // two complex64 columns

__device__ __forceinline__ C2 ldh_c2(const uint2* __restrict__ base, uint32_t idx) {
  const uint2 raw = __ldg(base + idx);
  C2 r;
  r.re = make_float2(bf_lo(raw.x), bf_lo(raw.y));
  r.im = make_float2(bf_hi(raw.x), bf_hi(raw.y));
  return r;
}
__device__ __forceinline__ void sth_c2(uint2* __restrict__ base, uint32_t idx, const C2& v) {
  base[idx] = make_uint2(bf_pack(v.re.x, v.im.x), bf_pack(v.re.y, v.im.y));
}
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// u * v with u = (ur,ur,ui,ui)
__device__ __forceinline__ C2 cmul_splat(const float4& u, const C2& v) {
  const float2 ur = make_float2(u.x, u.y), ui = make_float2(u.z, u.w);
  C2 p;
  p.re = __ffma2_rn(neg2(ui), v.im, __fmul2_rn(ur, v.re));
  p.im = __ffma2_rn(ui, v.re, __fmul2_rn(ur, v.im));
  return p;
}

template <int MINB>     // 2 / 3: 512-thread blocks (64 / 40 registers); 5: 256-thread blocks, 48 registers
__global__ void __launch_bounds__(MINB == 5 ? 256 : 512, MINB)
stencil_step_bf16_kernel(int LX, int LT, const float4* __restrict__ L4, float diag,
                        const uint2* __restrict__ X, uint2* __restrict__ Y, float wr, float wi, uint32_t kp,
                        long long rowb, long long ltb, long long spb) {   // bytes per row / per x-slice / per spin component
  const uint32_t cp = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t t = blockIdx.y * blockDim.y + threadIdx.y;
  const uint32_t x = blockIdx.z * blockDim.z + threadIdx.z;
  if (cp >= kp || t >= (uint32_t)LT || x >= (uint32_t)LX) return;
  const uint32_t V = (uint32_t)LX * LT;
  const uint32_t site = x * LT + t;
  // links first: ncu's source view showed the first use of each link (the FMUL2 of cmul_splat) as a top
  // long-scoreboard stall when these loads were issued after the ten vector loads
  const float4* lp = L4 + site;
  const float4 ut = __ldg(lp), utb = __ldg(lp + V), ux = __ldg(lp + 2 * V), uxb = __ldg(lp + 3 * V);
  // neighbour rows as byte offsets from the centre row (adds on the integer pipe; the FP32 pipe, which also
  // executes IMAD, is the one that bounds this kernel)
  const long long d_tp = (t + 1 == (uint32_t)LT) ? rowb - ltb : rowb;
  const long long d_tm = (t == 0) ? ltb - rowb : -rowb;
  const long long d_xp = (x + 1 == (uint32_t)LX) ? ltb - spb : ltb;
  const long long d_xm = (x == 0) ? spb - ltb : -ltb;
  const uint32_t ic = site * kp + cp;
  const char* pc = reinterpret_cast<const char*>(X) + (size_t)ic * 8;
  auto ld = [](const char* p) { return ldh_c2(reinterpret_cast<const uint2*>(p), 0u); };
  const C2 c0 = ld(pc),        c1 = ld(pc + spb);
  const C2 f0 = ld(pc + d_tp), f1 = ld(pc + d_tp + spb);
  const C2 b0 = ld(pc + d_tm), b1 = ld(pc + d_tm + spb);
  const C2 r0 = ld(pc + d_xp), r1 = ld(pc + d_xp + spb);
  const C2 l0 = ld(pc + d_xm), l1 = ld(pc + d_xm + spb);
  C2 a, b, c, d;                                     // spin projections (see stencil_kernel)
  a.re = __fadd2_rn(f0.re, neg2(f1.re)); a.im = __fadd2_rn(f0.im, neg2(f1.im));      // f0 - f1
  b.re = __fadd2_rn(b0.re, b1.re);       b.im = __fadd2_rn(b0.im, b1.im);            // b0 + b1
  c.re = __fadd2_rn(r0.re, neg2(r1.im)); c.im = __fadd2_rn(r0.im, r1.re);            // r0 + i r1
  d.re = __fadd2_rn(l0.re, l1.im);       d.im = __fadd2_rn(l0.im, neg2(l1.re));      // l0 - i l1
  const C2 ua = cmul_splat(ut, a), ub = cmul_splat(utb, b), uc = cmul_splat(ux, c), ud = cmul_splat(uxb, d);

  // (A x)_0 = diag c0 - (ua + ub + uc + ud);   (A x)_1 = diag c1 + (ua - ub) + i (uc - ud)
  const float2 dg = make_float2(diag, diag);
  C2 s, q, tt, y0, y1;
  s.re = __fadd2_rn(__fadd2_rn(ua.re, ub.re), __fadd2_rn(uc.re, ud.re));
  s.im = __fadd2_rn(__fadd2_rn(ua.im, ub.im), __fadd2_rn(uc.im, ud.im));
  q.re = __fadd2_rn(ua.re, neg2(ub.re));  q.im = __fadd2_rn(ua.im, neg2(ub.im));
  tt.re = __fadd2_rn(uc.re, neg2(ud.re)); tt.im = __fadd2_rn(uc.im, neg2(ud.im));
  y0.re = __ffma2_rn(dg, c0.re, neg2(s.re));
  y0.im = __ffma2_rn(dg, c0.im, neg2(s.im));
  y1.re = __ffma2_rn(dg, c1.re, __fadd2_rn(q.re, neg2(tt.im)));
  y1.im = __ffma2_rn(dg, c1.im, __fadd2_rn(q.im, tt.re));

  // x - w (A x)
  const float2 w_r = make_float2(wr, wr), w_i = make_float2(wi, wi);
  C2 o0, o1;
  o0.re = __ffma2_rn(w_i, y0.im, __ffma2_rn(neg2(w_r), y0.re, c0.re));
  o0.im = __ffma2_rn(neg2(w_i), y0.re, __ffma2_rn(neg2(w_r), y0.im, c0.im));
  o1.re = __ffma2_rn(w_i, y1.im, __ffma2_rn(neg2(w_r), y1.re, c1.re));
  o1.im = __ffma2_rn(neg2(w_i), y1.re, __ffma2_rn(neg2(w_r), y1.im, c1.im));
  char* py = reinterpret_cast<char*>(Y) + (size_t)ic * 8;
  sth_c2(reinterpret_cast<uint2*>(py), 0u, o0);
  sth_c2(reinterpret_cast<uint2*>(py + spb), 0u, o1);
}

// site arithmetic shared by the factor-kernel variants: o = c - w (A c) for both spin components
__device__ __forceinline__ void wilson_step_site(const C2& c0, const C2& c1, const C2& f0, const C2& f1, const C2& b0, const C2& b1,
                                                 const C2& r0, const C2& r1, const C2& l0, const C2& l1,
                                                 const float4& ut, const float4& utb, const float4& ux, const float4& uxb,
                                                 float2 dg, float2 w_r, float2 w_i, C2& o0, C2& o1) {
  C2 a, b, c, d;
  a.re = __fadd2_rn(f0.re, neg2(f1.re)); a.im = __fadd2_rn(f0.im, neg2(f1.im));
  b.re = __fadd2_rn(b0.re, b1.re);       b.im = __fadd2_rn(b0.im, b1.im);
  c.re = __fadd2_rn(r0.re, neg2(r1.im)); c.im = __fadd2_rn(r0.im, r1.re);
  d.re = __fadd2_rn(l0.re, l1.im);       d.im = __fadd2_rn(l0.im, neg2(l1.re));
  const C2 ua = cmul_splat(ut, a), ub = cmul_splat(utb, b), uc = cmul_splat(ux, c), ud = cmul_splat(uxb, d);
  C2 s, q, tt, y0, y1;
  s.re = __fadd2_rn(__fadd2_rn(ua.re, ub.re), __fadd2_rn(uc.re, ud.re));
  s.im = __fadd2_rn(__fadd2_rn(ua.im, ub.im), __fadd2_rn(uc.im, ud.im));
  q.re = __fadd2_rn(ua.re, neg2(ub.re));  q.im = __fadd2_rn(ua.im, neg2(ub.im));
  tt.re = __fadd2_rn(uc.re, neg2(ud.re)); tt.im = __fadd2_rn(uc.im, neg2(ud.im));
  y0.re = __ffma2_rn(dg, c0.re, neg2(s.re));
  y0.im = __ffma2_rn(dg, c0.im, neg2(s.im));
  y1.re = __ffma2_rn(dg, c1.re, __fadd2_rn(q.re, neg2(tt.im)));
  y1.im = __ffma2_rn(dg, c1.im, __fadd2_rn(q.im, tt.re));
  o0.re = __ffma2_rn(w_i, y0.im, __ffma2_rn(neg2(w_r), y0.re, c0.re));
  o0.im = __ffma2_rn(neg2(w_i), y0.re, __ffma2_rn(neg2(w_r), y0.im, c0.im));
  o1.re = __ffma2_rn(w_i, y1.im, __ffma2_rn(neg2(w_r), y1.re, c1.re));
  o1.im = __ffma2_rn(neg2(w_i), y1.re, __ffma2_rn(neg2(w_r), y1.im, c1.im));
}

// Two t-adjacent sites per thread (option "stencil_t2", default): site t's forward neighbour is site t+1's centre and
// vice versa, so the pair needs 8 site rows instead of 10 (-20 % loads and BF16 conversions) and shares the address
// arithmetic: 20.5 -> ~17 us per launch at k = 256, the bench step 42.7 -> 38.1 ms.  (A 2 x 2 sites-per-thread variant,
// 12 rows for 4 sites but 138 registers, measured the same 39.0 ms and was dropped; profiles/r1_run19_*.)
__global__ void __launch_bounds__(256, 2)
stencil_step_bf16_t2_kernel(int LX, int LT, const float4* __restrict__ L4, float diag,
                            const uint2* __restrict__ X, uint2* __restrict__ Y, float wr, float wi, uint32_t kp,
                            long long rowb, long long ltb, long long spb) {
  const uint32_t cp = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t t = 2 * (blockIdx.y * blockDim.y + threadIdx.y);
  const uint32_t x = blockIdx.z * blockDim.z + threadIdx.z;
  if (cp >= kp || t >= (uint32_t)LT || x >= (uint32_t)LX) return;
  const uint32_t V = (uint32_t)LX * LT;
  const uint32_t site = x * LT + t;
  const float4* lp = L4 + site;
  const float4 ut0 = __ldg(lp), utb0 = __ldg(lp + V), ux0 = __ldg(lp + 2 * V), uxb0 = __ldg(lp + 3 * V);
  const float4 ut1 = __ldg(lp + 1), utb1 = __ldg(lp + V + 1), ux1 = __ldg(lp + 2 * V + 1), uxb1 = __ldg(lp + 3 * V + 1);
  const long long d_tp2 = (t + 2 == (uint32_t)LT) ? 2 * rowb - ltb : 2 * rowb;     // row of site t+2
  const long long d_tm = (t == 0) ? ltb - rowb : -rowb;
  const long long d_xp = (x + 1 == (uint32_t)LX) ? ltb - spb : ltb;
  const long long d_xm = (x == 0) ? spb - ltb : -ltb;
  const uint32_t ic = site * kp + cp;
  const char* pc = reinterpret_cast<const char*>(X) + (size_t)ic * 8;
  auto ld = [](const char* p) { return ldh_c2(reinterpret_cast<const uint2*>(p), 0u); };
  const C2 a0 = ld(pc),          a1 = ld(pc + spb);                  // site t
  const C2 e0 = ld(pc + rowb),   e1 = ld(pc + rowb + spb);           // site t+1
  const C2 m0 = ld(pc + d_tm),   m1 = ld(pc + d_tm + spb);           // site t-1
  const C2 p0 = ld(pc + d_tp2),  p1 = ld(pc + d_tp2 + spb);          // site t+2
  const C2 ra0 = ld(pc + d_xp),  ra1 = ld(pc + d_xp + spb);          // x+1 of site t
  const C2 re0 = ld(pc + d_xp + rowb), re1 = ld(pc + d_xp + rowb + spb);
  const C2 la0 = ld(pc + d_xm),  la1 = ld(pc + d_xm + spb);          // x-1 of site t
  const C2 le0 = ld(pc + d_xm + rowb), le1 = ld(pc + d_xm + rowb + spb);
  const float2 dg = make_float2(diag, diag), w_r = make_float2(wr, wr), w_i = make_float2(wi, wi);
  C2 o0, o1, q0, q1;
  wilson_step_site(a0, a1, e0, e1, m0, m1, ra0, ra1, la0, la1, ut0, utb0, ux0, uxb0, dg, w_r, w_i, o0, o1);
  wilson_step_site(e0, e1, p0, p1, a0, a1, re0, re1, le0, le1, ut1, utb1, ux1, uxb1, dg, w_r, w_i, q0, q1);
  char* py = reinterpret_cast<char*>(Y) + (size_t)ic * 8;
  sth_c2(reinterpret_cast<uint2*>(py), 0u, o0);
  sth_c2(reinterpret_cast<uint2*>(py + spb), 0u, o1);
  sth_c2(reinterpret_cast<uint2*>(py + rowb), 0u, q0);
  sth_c2(reinterpret_cast<uint2*>(py + rowb + spb), 0u, q1);
}

// ------------------------------------------------------------------------------------------
// Even-odd (red-black) form of the level-0 smoother.  A = c I + H with H coupling sites of opposite parity only, so
//   A^{-1} r :  r^_e = r_e - H_eo r_o / c ;  x_e = S^{-1} r^_e ,  S = c - H_eo H_oe / c ;  x_o = (r_o - H_oe x_e) / c
// and the polynomial approximates the inverse of the Schur complement S on the even sites: the spectrum of S is the
// quadratic image lambda (2c - lambda) / c of A's, and a polynomial of degree d in S does what one of degree 2d in A does
// at the cost of d applications of A (two half-lattice sweeps each): 8 outer iterations at degree 16 in S against
// degree 36 in A (CPU experiment with the exact two-grid method, then measured on B200).
// Half-lattice vectors are checkerboard-compressed and BF16-stored: V_p[s][x][th][kp], site t = 2 th + ((x + p) & 1).
// One kernel does every sweep:   Out_p[site] = a * In2_p[site] + b * (H In_q)[site],   q = 1 - p
//   * w_o  = H_oe y_e                                (a = 0: HAS2 = false)
//   * y_e' = (1 - nu c) y_e + (nu / c) H_eo w_o      (one factor of the polynomial in S)
//   * r^_e and x_o                                   (the two ends)
// ZOUT: the result plus the coarse-grid correction Xc (complex64, full lattice) at the same site is written as
// complex128 into the full-lattice vector Z (the outer solver's Z_j); HOUT: the BF16 half-lattice output is written.
// NP = packs (of two columns) per thread: 2 -> 16-byte loads / stores and the four link loads shared by four columns.
template <int NP> __device__ __forceinline__ void ldh_packs(const uint2* __restrict__ p, C2 (&out)[NP]) {
  if constexpr (NP == 2) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    out[0].re = make_float2(bf_lo(raw.x), bf_lo(raw.y)); out[0].im = make_float2(bf_hi(raw.x), bf_hi(raw.y));
    out[1].re = make_float2(bf_lo(raw.z), bf_lo(raw.w)); out[1].im = make_float2(bf_hi(raw.z), bf_hi(raw.w));
  } else {
    out[0] = ldh_c2(p, 0u);
  }
}
template <int NP> __device__ __forceinline__ void sth_packs(uint2* __restrict__ p, const C2 (&v)[NP]) {
  if constexpr (NP == 2) {
    *reinterpret_cast<uint4*>(p) = make_uint4(bf_pack(v[0].re.x, v[0].im.x), bf_pack(v[0].re.y, v[0].im.y),
                                              bf_pack(v[1].re.x, v[1].im.x), bf_pack(v[1].re.y, v[1].im.y));
  } else {
    sth_c2(p, 0u, v[0]);
  }
}

template <bool HAS2, bool HOUT, bool ZOUT, int NP>
__global__ void __launch_bounds__(256, NP == 2 ? 2 : 3)
wilson_hop_eo_kernel(int LX, int LT, int p, const float4* __restrict__ L4, const uint2* __restrict__ Inq,
                     const uint2* __restrict__ In2, uint2* __restrict__ Outp, float ar, float ai, float br, float bi,
                     uint32_t kp, const Pack<float, 2>* __restrict__ Xc, double2* __restrict__ Z, int zhalf) {   // zhalf: 1 = Z is a half-lattice complex128 array, 2 = complex64
  const uint32_t cp = (blockIdx.x * blockDim.x + threadIdx.x) * NP;
  const uint32_t th = blockIdx.y * blockDim.y + threadIdx.y;
  const uint32_t x = blockIdx.z * blockDim.z + threadIdx.z;
  const uint32_t LH = (uint32_t)LT >> 1;
  if (cp >= kp || th >= LH || x >= (uint32_t)LX) return;
  const uint32_t V = (uint32_t)LX * LT, VH = V >> 1;
  const uint32_t a = (x + (uint32_t)p) & 1u;               // t offset of parity p in row x (the other parity: 1 - a)
  const uint32_t t = 2 * th + a;
  const uint32_t site = x * LT + t;
  const float4* lp = L4 + site;
  const float4 ut = __ldg(lp), utb = __ldg(lp + V), ux = __ldg(lp + 2 * V), uxb = __ldg(lp + 3 * V);
  // neighbours, all of parity q.  Same row: th or th +- 1; rows x +- 1: the same th
  const uint32_t th_f = a ? ((th + 1 == LH) ? 0u : th + 1) : th;          // t + 1
  const uint32_t th_b = a ? th : ((th == 0) ? LH - 1 : th - 1);           // t - 1
  const uint32_t xp = (x + 1 == (uint32_t)LX) ? 0u : x + 1, xm = (x == 0) ? (uint32_t)LX - 1 : x - 1;
  const size_t kpz = kp, sp = (size_t)VH * kpz;            // spin stride
  const uint2* qf = Inq + ((size_t)x * LH + th_f) * kpz + cp;
  const uint2* qb = Inq + ((size_t)x * LH + th_b) * kpz + cp;
  const uint2* qr = Inq + ((size_t)xp * LH + th) * kpz + cp;
  const uint2* ql = Inq + ((size_t)xm * LH + th) * kpz + cp;
  C2 f0[NP], f1[NP], b0[NP], b1[NP], r0[NP], r1[NP], l0[NP], l1[NP], c0[NP], c1[NP];
  ldh_packs<NP>(qf, f0); ldh_packs<NP>(qf + sp, f1);
  ldh_packs<NP>(qb, b0); ldh_packs<NP>(qb + sp, b1);
  ldh_packs<NP>(qr, r0); ldh_packs<NP>(qr + sp, r1);
  ldh_packs<NP>(ql, l0); ldh_packs<NP>(ql + sp, l1);
  const size_t ic = ((size_t)x * LH + th) * kpz + cp;      // this site in the parity-p arrays
  if constexpr (HAS2) { ldh_packs<NP>(In2 + ic, c0); ldh_packs<NP>(In2 + ic + sp, c1); }
  const float2 b_r = make_float2(br, br), b_i = make_float2(bi, bi);
  const float2 a_r = make_float2(ar, ar), a_i = make_float2(ai, ai);
  C2 o0[NP], o1[NP];
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    // spin projections and link products as in wilson_step_site
    C2 pa, pb, pc, pd;
    pa.re = __fadd2_rn(f0[u].re, neg2(f1[u].re)); pa.im = __fadd2_rn(f0[u].im, neg2(f1[u].im));
    pb.re = __fadd2_rn(b0[u].re, b1[u].re);       pb.im = __fadd2_rn(b0[u].im, b1[u].im);
    pc.re = __fadd2_rn(r0[u].re, neg2(r1[u].im)); pc.im = __fadd2_rn(r0[u].im, r1[u].re);
    pd.re = __fadd2_rn(l0[u].re, l1[u].im);       pd.im = __fadd2_rn(l0[u].im, neg2(l1[u].re));
    const C2 ua = cmul_splat(ut, pa), ub = cmul_splat(utb, pb), uc = cmul_splat(ux, pc), ud = cmul_splat(uxb, pd);
    // (H v)_0 = -(ua + ub + uc + ud);   (H v)_1 = (ua - ub) + i (uc - ud)
    C2 h0, h1, q, tt;
    h0.re = neg2(__fadd2_rn(__fadd2_rn(ua.re, ub.re), __fadd2_rn(uc.re, ud.re)));
    h0.im = neg2(__fadd2_rn(__fadd2_rn(ua.im, ub.im), __fadd2_rn(uc.im, ud.im)));
    q.re = __fadd2_rn(ua.re, neg2(ub.re));  q.im = __fadd2_rn(ua.im, neg2(ub.im));
    tt.re = __fadd2_rn(uc.re, neg2(ud.re)); tt.im = __fadd2_rn(uc.im, neg2(ud.im));
    h1.re = __fadd2_rn(q.re, neg2(tt.im));
    h1.im = __fadd2_rn(q.im, tt.re);
    // o = a c + b h
    o0[u].re = __ffma2_rn(neg2(b_i), h0.im, __fmul2_rn(b_r, h0.re));
    o0[u].im = __ffma2_rn(b_i, h0.re, __fmul2_rn(b_r, h0.im));
    o1[u].re = __ffma2_rn(neg2(b_i), h1.im, __fmul2_rn(b_r, h1.re));
    o1[u].im = __ffma2_rn(b_i, h1.re, __fmul2_rn(b_r, h1.im));
    if constexpr (HAS2) {
      o0[u].re = __ffma2_rn(neg2(a_i), c0[u].im, __ffma2_rn(a_r, c0[u].re, o0[u].re));
      o0[u].im = __ffma2_rn(a_i, c0[u].re, __ffma2_rn(a_r, c0[u].im, o0[u].im));
      o1[u].re = __ffma2_rn(neg2(a_i), c1[u].im, __ffma2_rn(a_r, c1[u].re, o1[u].re));
      o1[u].im = __ffma2_rn(a_i, c1[u].re, __ffma2_rn(a_r, c1[u].im, o1[u].im));
    }
  }
  if constexpr (HOUT) { sth_packs<NP>(Outp + ic, o0); sth_packs<NP>(Outp + ic + sp, o1); }
  if constexpr (ZOUT) {
    const size_t j0 = (size_t)site * kpz + cp, j1 = ((size_t)V + site) * kpz + cp;     // full-lattice rows of the site
    if (zhalf == 2) {
      // complex64 half-lattice output (the Z_j of the mixed-precision Schur-complement solve)
      Pack<float, 2>* Zf = reinterpret_cast<Pack<float, 2>*>(Z);
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const Pack<float, 2> e0 = Xc[j0 + u], e1 = Xc[j1 + u];
        Pack<float, 2> q0, q1;
        q0.d[0] = e0.d[0] + o0[u].re.x; q0.d[1] = e0.d[1] + o0[u].im.x; q0.d[2] = e0.d[2] + o0[u].re.y; q0.d[3] = e0.d[3] + o0[u].im.y;
        q1.d[0] = e1.d[0] + o1[u].re.x; q1.d[1] = e1.d[1] + o1[u].im.x; q1.d[2] = e1.d[2] + o1[u].re.y; q1.d[3] = e1.d[3] + o1[u].im.y;
        Zf[ic + u] = q0; Zf[ic + sp + u] = q1;
      }
    } else {
      const size_t z0 = zhalf ? ic : j0, z1 = zhalf ? ic + sp : j1;
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const Pack<float, 2> e0 = Xc[j0 + u], e1 = Xc[j1 + u];
        Z[2 * (z0 + u)]     = make_double2((double)e0.d[0] + (double)o0[u].re.x, (double)e0.d[1] + (double)o0[u].im.x);
        Z[2 * (z0 + u) + 1] = make_double2((double)e0.d[2] + (double)o0[u].re.y, (double)e0.d[3] + (double)o0[u].im.y);
        Z[2 * (z1 + u)]     = make_double2((double)e1.d[0] + (double)o1[u].re.x, (double)e1.d[1] + (double)o1[u].im.x);
        Z[2 * (z1 + u) + 1] = make_double2((double)e1.d[2] + (double)o1[u].re.y, (double)e1.d[3] + (double)o1[u].im.y);
      }
    }
  }
}

// ---- sweeps on checkerboard half-lattice vectors [s][x][t/2][k] in the outer solver's own precision (the outer solve on the
// even-odd Schur complement): complex128, or complex64 (two columns per thread) for the complex64-stored Krylov vectors of the
// mixed-precision refinement.   Out_p[site] = a * In2_p[site] + b * (H In_q)[site]   (same conventions as wilson_hop_eo_kernel)
template <typename T, int NC, bool HAS2>
__global__ void __launch_bounds__(256)
wilson_hop_eo_z_kernel(StencilDev<T> op, int p, const Cx<T>* __restrict__ Inq, const Cx<T>* __restrict__ In2,
                       Cx<T>* __restrict__ Outp, Cx<T> a, Cx<T> b, int kp) {    // kp = packs of NC columns per row
  typedef Pack<T, NC> P;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int th = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z * blockDim.z + threadIdx.z;
  const int LX = op.LX, LT = op.LT, LH = LT >> 1;
  if (col >= kp || th >= LH || x >= LX) return;
  const int off = (x + p) & 1;
  const int t = 2 * th + off;
  const int site = x * LT + t;
  const int tm = (t == 0) ? LT - 1 : t - 1;
  const int xp = (x + 1 == LX) ? 0 : x + 1, xm = (x == 0) ? LX - 1 : x - 1;
  const int th_f = off ? ((th + 1 == LH) ? 0 : th + 1) : th;
  const int th_b = off ? th : ((th == 0) ? LH - 1 : th - 1);
  const Cx<T> ut = ldc_ro<T>(op.Ut, site), utb = cconj(ldc_ro<T>(op.Ut, x * LT + tm));
  const Cx<T> ux = ldc_ro<T>(op.Ux, site), uxb = cconj(ldc_ro<T>(op.Ux, xm * LT + t));
  const size_t kz = (size_t)kp, sp = (size_t)(LX * LH) * kz;
  const P* Q = reinterpret_cast<const P*>(Inq);
  const size_t i_f = ((size_t)x * LH + th_f) * kz + col, i_b = ((size_t)x * LH + th_b) * kz + col;
  const size_t i_r = ((size_t)xp * LH + th) * kz + col,  i_l = ((size_t)xm * LH + th) * kz + col;
  const P f0 = ldp_ro<T, NC>(Q, i_f), f1 = ldp_ro<T, NC>(Q, i_f + sp);
  const P b0 = ldp_ro<T, NC>(Q, i_b), b1 = ldp_ro<T, NC>(Q, i_b + sp);
  const P r0 = ldp_ro<T, NC>(Q, i_r), r1 = ldp_ro<T, NC>(Q, i_r + sp);
  const P l0 = ldp_ro<T, NC>(Q, i_l), l1 = ldp_ro<T, NC>(Q, i_l + sp);
  const P pa = psub<T, NC>(f0, f1);
  const P pb = padd<T, NC>(b0, b1);
  const P pc = padd<T, NC>(r0, pmul_i<T, NC>(r1));
  const P pd = psub<T, NC>(l0, pmul_i<T, NC>(l1));
  const P ua = pscale<T, NC>(ut, pa), ub = pscale<T, NC>(utb, pb);
  const P uc = pscale<T, NC>(ux, pc), ud = pscale<T, NC>(uxb, pd);
  // (H v)_0 = -(ua + ub + uc + ud);  (H v)_1 = (ua - ub) + i (uc - ud)
  P h0 = pzero<T, NC>();
  h0 = psub<T, NC>(h0, padd<T, NC>(padd<T, NC>(ua, ub), padd<T, NC>(uc, ud)));
  const P h1 = padd<T, NC>(psub<T, NC>(ua, ub), pmul_i<T, NC>(psub<T, NC>(uc, ud)));
  P o0 = pscale<T, NC>(b, h0), o1 = pscale<T, NC>(b, h1);
  const size_t ic = ((size_t)x * LH + th) * kz + col;
  if constexpr (HAS2) {
    const P* C = reinterpret_cast<const P*>(In2);
    pfma<T, NC>(o0, a, ldp_ro<T, NC>(C, ic));
    pfma<T, NC>(o1, a, ldp_ro<T, NC>(C, ic + sp));
  }
  P* O = reinterpret_cast<P*>(Outp);
  O[ic] = o0; O[ic + sp] = o1;
}

// The true residual of the Schur-complement system at the end of a cycle, in one pass:  R = Bh - (a * X + b * H Wo)  on the
// even sites (complex128, one column per thread; a = c, b = -1/c gives R = b^ - S x with Wo = H_oe x), and the squared norm of
// every column of R as per-block partial sums  partial[(blockIdx.z * gridDim.y + blockIdx.y) * k + col]  (fixed order:
// deterministic, independent of the batch) -- replaces the operator kernel + vec_sub_kernel + multi_dot_kernel of that step.
__global__ void __launch_bounds__(256)
wilson_schur_residual_kernel(StencilDev<double> op, int p, const Cx<double>* __restrict__ Inq, const Cx<double>* __restrict__ In2,
                             const Cx<double>* __restrict__ In3, Cx<double>* __restrict__ Outp, Cx<double> a, Cx<double> b, int k,
                             Cx<double>* __restrict__ partial) {
  typedef Pack<double, 1> P;
  __shared__ double red[8][33];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int th = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z * blockDim.z + threadIdx.z;
  const int LX = op.LX, LT = op.LT, LH = LT >> 1;
  const bool valid = col < k && th < LH && x < LX;
  double sq = 0.0;
  if (valid) {
    const int off = (x + p) & 1;
    const int t = 2 * th + off;
    const int site = x * LT + t;
    const int tm = (t == 0) ? LT - 1 : t - 1;
    const int xp = (x + 1 == LX) ? 0 : x + 1, xm = (x == 0) ? LX - 1 : x - 1;
    const int th_f = off ? ((th + 1 == LH) ? 0 : th + 1) : th;
    const int th_b = off ? th : ((th == 0) ? LH - 1 : th - 1);
    const Cx<double> ut = ldc_ro<double>(op.Ut, site), utb = cconj(ldc_ro<double>(op.Ut, x * LT + tm));
    const Cx<double> ux = ldc_ro<double>(op.Ux, site), uxb = cconj(ldc_ro<double>(op.Ux, xm * LT + t));
    const size_t kz = (size_t)k, sp = (size_t)(LX * LH) * kz;
    const P* Q = reinterpret_cast<const P*>(Inq);
    const size_t i_f = ((size_t)x * LH + th_f) * kz + col, i_b = ((size_t)x * LH + th_b) * kz + col;
    const size_t i_r = ((size_t)xp * LH + th) * kz + col,  i_l = ((size_t)xm * LH + th) * kz + col;
    const P f0 = ldp_ro<double, 1>(Q, i_f), f1 = ldp_ro<double, 1>(Q, i_f + sp);
    const P b0 = ldp_ro<double, 1>(Q, i_b), b1 = ldp_ro<double, 1>(Q, i_b + sp);
    const P r0 = ldp_ro<double, 1>(Q, i_r), r1 = ldp_ro<double, 1>(Q, i_r + sp);
    const P l0 = ldp_ro<double, 1>(Q, i_l), l1 = ldp_ro<double, 1>(Q, i_l + sp);
    const P pa = psub<double, 1>(f0, f1);
    const P pb = padd<double, 1>(b0, b1);
    const P pc = padd<double, 1>(r0, pmul_i<double, 1>(r1));
    const P pd = psub<double, 1>(l0, pmul_i<double, 1>(l1));
    const P ua = pscale<double, 1>(ut, pa), ub = pscale<double, 1>(utb, pb);
    const P uc = pscale<double, 1>(ux, pc), ud = pscale<double, 1>(uxb, pd);
    P h0 = pzero<double, 1>();
    h0 = psub<double, 1>(h0, padd<double, 1>(padd<double, 1>(ua, ub), padd<double, 1>(uc, ud)));
    const P h1 = padd<double, 1>(psub<double, 1>(ua, ub), pmul_i<double, 1>(psub<double, 1>(uc, ud)));
    P o0 = pscale<double, 1>(b, h0), o1 = pscale<double, 1>(b, h1);
    const size_t ic = ((size_t)x * LH + th) * kz + col;
    const P* C = reinterpret_cast<const P*>(In2);
    pfma<double, 1>(o0, a, ldp_ro<double, 1>(C, ic));
    pfma<double, 1>(o1, a, ldp_ro<double, 1>(C, ic + sp));
    const P* B3 = reinterpret_cast<const P*>(In3);
    const P q0 = psub<double, 1>(ldp_ro<double, 1>(B3, ic), o0), q1 = psub<double, 1>(ldp_ro<double, 1>(B3, ic + sp), o1);
    P* O = reinterpret_cast<P*>(Outp);
    O[ic] = q0; O[ic + sp] = q1;
    sq = fma(q0.d[0], q0.d[0], fma(q0.d[1], q0.d[1], fma(q1.d[0], q1.d[0], q1.d[1] * q1.d[1])));
  }
  const int lane = threadIdx.z * blockDim.y + threadIdx.y;                // 0 .. 7 (blockDim = (bx, 4, 2))
  red[lane][threadIdx.x] = sq;
  __syncthreads();
  if (lane == 0 && col < k) {
    double t = red[0][threadIdx.x];
    const int nl = blockDim.y * blockDim.z;
    for (int y = 1; y < nl; ++y) t += red[y][threadIdx.x];
    partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * k + col] = cx<double>(t, 0.0);
  }
}

// full-lattice X[s][x][t][k]  <->  checkerboard halves E, O [s][x][t/2][k]   (complex128)
// dir = 0: split X into E and O;  dir = 1: merge E and O into X
__global__ void __launch_bounds__(256)
eo_split_merge_kernel(int LX, int LT, int k, int dir, Cx<double>* __restrict__ X, Cx<double>* __restrict__ E, Cx<double>* __restrict__ O) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nk = (size_t)2 * LX * LT * k;
  if (idx >= nk) return;
  const size_t row = idx / k; const int col = (int)(idx - row * k);
  const int V = LX * LT;
  const int s = (int)(row / V); const int site = (int)(row - (size_t)s * V);
  const int x = site / LT, t = site - x * LT;
  const int p = (x + t) & 1;
  const size_t hrow = (size_t)s * (V >> 1) + (size_t)x * (LT >> 1) + (t >> 1);
  Cx<double>* Hh = p ? O : E;
  if (dir == 0) Hh[hrow * k + col] = X[idx]; else X[idx] = Hh[hrow * k + col];
}

// Out[r][col] = In[r][col] * scale[col] on even-parity half-lattice vectors (In complex128 or complex64, Out in the Krylov
// vectors' storage precision), and the same values as complex64 at the even sites of the full-lattice array Out32 (whose odd
// sites stay zero): the V-cycle's input (v_e, 0)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
col_scale_eo_kernel(int LX, int LT, const Cx<TI>* __restrict__ In, const double* __restrict__ scale, Cx<TO>* __restrict__ Out,
                    int k, Cx<float>* __restrict__ Out32) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nkh = (size_t)LX * LT * k;                  // 2 spins * V/2 sites
  if (idx >= nkh) return;
  const size_t hrow = idx / k; const int col = (int)(idx - hrow * k);
  const int VH = (LX * LT) >> 1, LH = LT >> 1;
  const int s = (int)(hrow / VH); const int r = (int)(hrow - (size_t)s * VH);
  const int x = r / LH, th = r - x * LH;
  const int t = 2 * th + (x & 1);
  const double sc = __ldg(scale + col);
  const Cx<TI> v = ldc_ro<TI>(In, idx);
  const Cx<TO> o = cx<TO>((TO)((double)v.re * sc), (TO)((double)v.im * sc));
  Out[idx] = o;
  if (Out32 != nullptr) Out32[((size_t)s * (LX * LT) + (size_t)x * LT + t) * k + col] = cx<float>((float)o.re, (float)o.im);
}

// the complex64 -> complex64 case of col_scale_eo_kernel with two columns per thread (16-byte loads and stores; k even)
__global__ void __launch_bounds__(256)
col_scale_eo_c64x2_kernel(int LX, int LT, const Cx<float>* __restrict__ In, const double* __restrict__ scale, Cx<float>* __restrict__ Out,
                          int k, Cx<float>* __restrict__ Out32) {
  typedef Pack<float, 2> P;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // pack index
  const int kp = k >> 1;
  const size_t npk = (size_t)LX * LT * kp;
  if (idx >= npk) return;
  const size_t hrow = idx / kp; const int cp = (int)(idx - hrow * kp);
  const int VH = (LX * LT) >> 1, LH = LT >> 1;
  const int s = (int)(hrow / VH); const int r = (int)(hrow - (size_t)s * VH);
  const int x = r / LH, th = r - x * LH;
  const int t = 2 * th + (x & 1);
  const double s0 = __ldg(scale + 2 * cp), s1 = __ldg(scale + 2 * cp + 1);
  const P v = ldp_ro<float, 2>(reinterpret_cast<const P*>(In), idx);
  P o;
  o.d[0] = (float)((double)v.d[0] * s0); o.d[1] = (float)((double)v.d[1] * s0);
  o.d[2] = (float)((double)v.d[2] * s1); o.d[3] = (float)((double)v.d[3] * s1);
  reinterpret_cast<P*>(Out)[idx] = o;
  if (Out32 != nullptr) reinterpret_cast<P*>(Out32)[((size_t)s * (LX * LT) + (size_t)x * LT + t) * kp + cp] = o;
}

// Shared-memory-tiled variant of the same factor kernel (option "stencil_smem", OFF by default: measured on B200 at
// k = 256 it needs 26.5 us per launch against 20.5 us for the L1-cached direct-load kernel above -- the load / barrier /
// compute phases of a CTA do not overlap as well as 48 independent warps per SM do; profiles/r1_run17_*).
// A CTA stages the (TX+2) x (TT+2) halo'd site tile of
// 32 columns (both spin components, BF16: 128 bytes per site row) with three coalesced LDG.128 per thread -- all
// global loads of the CTA are in flight at once and every vector row is read from L2 1.5 times instead of ~3 --
// and computes from shared memory (LDS.64, conflict-free: a warp reads two 128-byte rows).
template <int TX, int TT>
__global__ void __launch_bounds__(512, 3)
stencil_step_bf16_smem_kernel(int LX, int LT, const float4* __restrict__ L4, float diag,
                              const uint2* __restrict__ X, uint2* __restrict__ Y, float wr, float wi, uint32_t kp) {
  constexpr int SX = TX + 2, ST = TT + 2;
  __shared__ __align__(16) uint2 S[2][SX][ST][16];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.z * TX, t0 = blockIdx.y * TT;
  const uint32_t p0 = blockIdx.x * 16;
  const uint32_t V = (uint32_t)LX * LT;
  for (int c = tid; c < 2 * SX * ST * 8; c += 512) {
    const int q = c & 7, row = c >> 3;
    const int tt = row % ST, r2 = row / ST;
    const int xx = r2 % SX, sp = r2 / SX;
    const bool corner = (xx == 0 || xx == SX - 1) && (tt == 0 || tt == ST - 1);
    if (corner) continue;
    int x = x0 + xx - 1; x = (x < 0) ? x + LX : ((x >= LX) ? x - LX : x);
    int t = t0 + tt - 1; t = (t < 0) ? t + LT : ((t >= LT) ? t - LT : t);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + 2 * q + 1 < kp && x < LX && t < LT)
      v = __ldg(reinterpret_cast<const uint4*>(X + ((size_t)(sp * V + (uint32_t)x * LT + t) * kp + p0 + 2 * q)));
    *reinterpret_cast<uint4*>(&S[sp][xx][tt][2 * q]) = v;
  }
  __syncthreads();
  const float2 dg = make_float2(diag, diag);
  const float2 w_r = make_float2(wr, wr), w_i = make_float2(wi, wi);
  auto cvt = [](const uint2 raw) { C2 r; r.re = make_float2(bf_lo(raw.x), bf_lo(raw.y)); r.im = make_float2(bf_hi(raw.x), bf_hi(raw.y)); return r; };
#pragma unroll
  for (int it = tid; it < TX * TT * 16; it += 512) {
    const int p = it & 15, sidx = it >> 4;
    const int tl = sidx % TT, xl = sidx / TT;
    const int x = x0 + xl, t = t0 + tl;
    if (x >= LX || t >= LT || p0 + p >= kp) continue;
    const uint32_t site = (uint32_t)x * LT + t;
    const float4* lp = L4 + site;
    const float4 ut = __ldg(lp), utb = __ldg(lp + V), ux = __ldg(lp + 2 * V), uxb = __ldg(lp + 3 * V);
    const C2 c0 = cvt(S[0][xl + 1][tl + 1][p]), c1 = cvt(S[1][xl + 1][tl + 1][p]);
    const C2 f0 = cvt(S[0][xl + 1][tl + 2][p]), f1 = cvt(S[1][xl + 1][tl + 2][p]);
    const C2 b0 = cvt(S[0][xl + 1][tl][p]),     b1 = cvt(S[1][xl + 1][tl][p]);
    const C2 r0 = cvt(S[0][xl + 2][tl + 1][p]), r1 = cvt(S[1][xl + 2][tl + 1][p]);
    const C2 l0 = cvt(S[0][xl][tl + 1][p]),     l1 = cvt(S[1][xl][tl + 1][p]);
    C2 a, b, c, d;
    a.re = __fadd2_rn(f0.re, neg2(f1.re)); a.im = __fadd2_rn(f0.im, neg2(f1.im));
    b.re = __fadd2_rn(b0.re, b1.re);       b.im = __fadd2_rn(b0.im, b1.im);
    c.re = __fadd2_rn(r0.re, neg2(r1.im)); c.im = __fadd2_rn(r0.im, r1.re);
    d.re = __fadd2_rn(l0.re, l1.im);       d.im = __fadd2_rn(l0.im, neg2(l1.re));
    const C2 ua = cmul_splat(ut, a), ub = cmul_splat(utb, b), uc = cmul_splat(ux, c), ud = cmul_splat(uxb, d);
    C2 s_, q_, tt_, y0, y1;
    s_.re = __fadd2_rn(__fadd2_rn(ua.re, ub.re), __fadd2_rn(uc.re, ud.re));
    s_.im = __fadd2_rn(__fadd2_rn(ua.im, ub.im), __fadd2_rn(uc.im, ud.im));
    q_.re = __fadd2_rn(ua.re, neg2(ub.re));  q_.im = __fadd2_rn(ua.im, neg2(ub.im));
    tt_.re = __fadd2_rn(uc.re, neg2(ud.re)); tt_.im = __fadd2_rn(uc.im, neg2(ud.im));
    y0.re = __ffma2_rn(dg, c0.re, neg2(s_.re));
    y0.im = __ffma2_rn(dg, c0.im, neg2(s_.im));
    y1.re = __ffma2_rn(dg, c1.re, __fadd2_rn(q_.re, neg2(tt_.im)));
    y1.im = __ffma2_rn(dg, c1.im, __fadd2_rn(q_.im, tt_.re));
    C2 o0, o1;
    o0.re = __ffma2_rn(w_i, y0.im, __ffma2_rn(neg2(w_r), y0.re, c0.re));
    o0.im = __ffma2_rn(neg2(w_i), y0.re, __ffma2_rn(neg2(w_r), y0.im, c0.im));
    o1.re = __ffma2_rn(w_i, y1.im, __ffma2_rn(neg2(w_r), y1.re, c1.re));
    o1.im = __ffma2_rn(neg2(w_i), y1.re, __ffma2_rn(neg2(w_r), y1.im, c1.im));
    uint2* py = Y + ((size_t)site * kp + p0 + p);
    sth_c2(py, 0u, o0);
    sth_c2(py + (size_t)V * kp, 0u, o1);
  }
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

// HIN / HOUT: X / Y are BF16-stored (the smoother's intermediate vectors, as on level 0).
template <int BS, int PPT, int MODE, bool HIN = false, bool HOUT = false>
__global__ void __launch_bounds__(128)
bsr_f32x2_kernel(int nb, int bpr, const int* __restrict__ col, const float4* __restrict__ vals4,
                 const void* __restrict__ Xv, const Pack<float, 2>* __restrict__ B,
                 void* __restrict__ Yv, Cx<float> w, Cx<float> cfin, int kp) {
  const Pack<float, 2>* __restrict__ X = reinterpret_cast<const Pack<float, 2>*>(Xv);
  Pack<float, 2>* __restrict__ Y = reinterpret_cast<Pack<float, 2>*>(Yv);
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
      for (int p = 0; p < PPT; ++p) {
        const size_t xi = ((size_t)J * BS + c) * kpz + cp[p];
        if constexpr (HIN) {
          const Pack<float, 2> v = ldh2_ro(Xv, xi);
          xv[c][p] = make_float4(v.d[0], v.d[1], v.d[2], v.d[3]);
        } else {
          xv[c][p] = __ldg(X4 + xi);
        }
      }
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
      if constexpr (MODE >= M_STEP) { if constexpr (HIN) xin = ldh2_ro(Xv, idx); else xin = ldp_ro<float, 2>(X, idx); }
      const Pack<float, 2> o = op_value<float, 2, MODE>(ax, xin, idx, B, Y, w, cfin);
      if constexpr (HOUT) sth2(Yv, idx, o); else Y[idx] = o;
    }
  }
}

// Variant for BF16-stored input (the smoother's intermediate vectors): (re0,re1)/(im0,im1) register pairs come for free
// out of the BF16 -> FP32 conversion, one accumulator pair per output -> 80 registers instead of 126 (measured: 28 -> 23.4 us
// per factor on level 1 of 128^2 at k = 256; for FP32 input the (a, b) kernel above is faster: 25 vs 35 us).
template <int BS, int PPT, int MODE, bool HIN = false, bool HOUT = false>
__global__ void __launch_bounds__(256, 3)
bsr_f32x2_soa_kernel(int nb, int bpr, const int* __restrict__ col, const float4* __restrict__ vals4,
                 const void* __restrict__ Xv, const Pack<float, 2>* __restrict__ B,
                 void* __restrict__ Yv, Cx<float> w, Cx<float> cfin, int kp) {
  const Pack<float, 2>* __restrict__ X = reinterpret_cast<const Pack<float, 2>*>(Xv);
  Pack<float, 2>* __restrict__ Y = reinterpret_cast<Pack<float, 2>*>(Yv);
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
  // two columns of a pack are held as (re0,re1) / (im0,im1) register pairs: a complex multiply-accumulate is 4 FFMA2
  // for both columns and needs one accumulator pair per output (half the registers of an (a, b) split)
  C2 acc[BS][PPT];
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int p = 0; p < PPT; ++p) { acc[r][p].re = make_float2(0.f, 0.f); acc[r][p].im = make_float2(0.f, 0.f); }
  const float4* vrow = vs + (size_t)rb * nent;
  const int* crow = cs + rb * bpr;
  const float4* X4 = reinterpret_cast<const float4*>(X);
  auto ldx2 = [&](size_t xi) -> C2 {
    C2 r;
    if constexpr (HIN) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(Xv) + xi);
      r.re = make_float2(bf_lo(raw.x), bf_lo(raw.y)); r.im = make_float2(bf_hi(raw.x), bf_hi(raw.y));
    } else {
      const float4 v = __ldg(X4 + xi);
      r.re = make_float2(v.x, v.z); r.im = make_float2(v.y, v.w);
    }
    return r;
  };
  for (int blk = 0; blk < bpr; ++blk) {
    const int J = crow[blk];
    if (J < 0) continue;
    C2 xv[BS][PPT];
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int p = 0; p < PPT; ++p) xv[c][p] = ldx2(((size_t)J * BS + c) * kpz + cp[p]);
    const float4* vb = vrow + blk * (BS * BS);
#pragma unroll
    for (int r = 0; r < BS; ++r) {
#pragma unroll
      for (int c = 0; c < BS; ++c) {
        const float4 m = vb[r * BS + c];
        const float2 mr = make_float2(m.x, m.y), mi = make_float2(m.z, m.w);
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          acc[r][p].re = __ffma2_rn(neg2(mi), xv[c][p].im, __ffma2_rn(mr, xv[c][p].re, acc[r][p].re));
          acc[r][p].im = __ffma2_rn(mi, xv[c][p].re, __ffma2_rn(mr, xv[c][p].im, acc[r][p].im));
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < BS; ++r) {
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (!ok[p]) continue;
      Pack<float, 2> ax;
      ax.d[0] = acc[r][p].re.x; ax.d[1] = acc[r][p].im.x;
      ax.d[2] = acc[r][p].re.y; ax.d[3] = acc[r][p].im.y;
      const size_t idx = ((size_t)I * BS + r) * kpz + cp[p];
      Pack<float, 2> xin = pzero<float, 2>();
      if constexpr (MODE >= M_STEP) { if constexpr (HIN) xin = ldh2_ro(Xv, idx); else xin = ldp_ro<float, 2>(X, idx); }
      const Pack<float, 2> o = op_value<float, 2, MODE>(ax, xin, idx, B, Y, w, cfin);
      if constexpr (HOUT) sth2(Yv, idx, o); else Y[idx] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// aggregation transfer operators.  Row r of P_l: aggregate j = r / aggr, half = ((r % aggr) % dofi) >= h,
// columns (2j + half)*NV + [0,NV).                                   (multigrid.py:203-227)
// Indexed form (geometric aggregates of the preconditioner hierarchy, dmlmc_set_transfer_indexed): the coarse block
// of fine row r is cblk[r] (columns cblk[r]*NV + [0,NV)), and coarse block g owns the m fine rows rows[g*m + (0..m)).
template <typename T> struct TransferDev {
  int n_f, n_c, aggr, dofi, h, nvec;
  const Cx<T>* pv;     // [n_f][nvec]
  const int* rows = nullptr; const int* cblk = nullptr; int m = 0;
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
  if (tr.rows != nullptr) {
    const int* rr = tr.rows + (size_t)g * tr.m;
    for (int q = 0; q < tr.m; ++q) {
      const int r = __ldg(rr + q);
      const P x = ldp_ro<T, NC>(Xf, (size_t)r * ldfz + cp);
#pragma unroll
      for (int v = 0; v < NV; ++v) pfma_conj<T, NC>(acc[v], ldc_ro<T>(tr.pv, (size_t)r * NV + v), x);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) Xc[((size_t)g * NV + v) * ldcz + cp] = acc[v];
    return;
  }
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
                   int ldf, int ldc, int add) {   // add = 0: Xf = P Xc (Xf is not read)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(gid / kp);
  const int cp = (int)(gid - (long long)r * kp);
  if (r >= tr.n_f) return;
  int g;
  if (tr.cblk != nullptr) {
    g = __ldg(tr.cblk + r);
  } else {
    const int j = r / tr.aggr;
    const int q = (r - j * tr.aggr) % tr.dofi;
    g = 2 * j + (q >= tr.h ? 1 : 0);
  }
  const size_t ldfz = (size_t)ldf, ldcz = (size_t)ldc;
  typedef Pack<T, NC> P;
  P acc = add ? Xf[(size_t)r * ldfz + cp] : pzero<T, NC>();
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

// ------------------------------------------------------------------------------------------
// The probe stream on the device.  The reference draws every probe element with
// np.random.randint(2) from the global legacy generator (utils.py:213-216, 255-258): element j of the
// stream is the least significant bit of the j-th tempered 32-bit output of MT19937.  This kernel
// advances that same generator on the GPU: state[0..623] is the MT19937 key, state[624] the position
// (exactly np.random.get_state()[1:3]), so that the host generator can be re-synchronised afterwards.
// It skips `skip_before` outputs, writes the LSBs of the next `count` outputs to out[0..count)
// (one byte each), skips `skip_after` more and stores the new state; `backup` (may be null) receives
// the state before the call (the rewind point of the sequential stopping rule).
// One CTA: the twist x[i+624] = x[i+397] ^ f(x[i], x[i+1]) is parallel over i in three runs of <= 227
// words (double-buffered in shared memory, 3 barriers per 624 outputs).
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__global__ void __launch_bounds__(256)
mt19937_bits_kernel(uint32_t* __restrict__ state, uint32_t* __restrict__ backup, long long skip_before, long long count,
                    long long skip_after, uint8_t* __restrict__ out) {
  __shared__ uint32_t s[2][624];
  const int tid = threadIdx.x;
  for (int i = tid; i < 624; i += 256) { const uint32_t v = state[i]; s[0][i] = v; if (backup) backup[i] = v; }
  int mti = (int)state[624];
  if (backup && tid == 0) backup[624] = (uint32_t)mti;
  int cur = 0;
  __syncthreads();
  const long long total = skip_before + count + skip_after;
  const long long out_end = skip_before + count;
  long long done = 0;
  while (done < total) {
    if (mti >= 624) {
      const uint32_t* o = s[cur];
      uint32_t* nw = s[cur ^ 1];
      if (tid < 227) nw[tid] = o[tid + 397] ^ mt_mix(o[tid], o[tid + 1]);
      __syncthreads();
      if (tid < 227) nw[227 + tid] = nw[tid] ^ mt_mix(o[227 + tid], o[228 + tid]);
      __syncthreads();
      if (tid < 169) nw[454 + tid] = nw[227 + tid] ^ mt_mix(o[454 + tid], o[455 + tid]);
      if (tid == 255) nw[623] = nw[396] ^ mt_mix(o[623], nw[0]);
      __syncthreads();
      cur ^= 1;
      mti = 0;
    }
    const long long left = total - done;
    const int take = (int)((left < (long long)(624 - mti)) ? left : (long long)(624 - mti));
    if (done + take > skip_before && done < out_end) {       // this run of outputs overlaps the wanted range
      for (int i = mti + tid; i < mti + take; i += 256) {
        const long long g = done + (i - mti);
        if (g >= skip_before && g < out_end) out[g - skip_before] = (uint8_t)(mt_temper(s[cur][i]) & 1u);
      }
    }
    done += take;
    mti += take;
  }
  __syncthreads();
  for (int i = tid; i < 624; i += 256) state[i] = s[cur][i];
  if (tid == 0) state[624] = (uint32_t)mti;
}

// ---- the same stream with jump-ahead: many CTAs per call, none of them walks over another rank's probes ----------------
// The MT19937 transition is linear over GF(2), so the raw word sequence X[t] obeys  X[J + w] = XOR_{i: g_i = 1} X[i + w]
// with g(t) = t^J mod phi(t) (phi: the characteristic polynomial, degree 19937).  `tab` holds g for J = 2^b, b = 0..47
// (deflatedmlmc_schwinger_b200/mtjump.py builds and pins it against np.random).  CTA c < gridDim.x - 1 produces the outputs
// [c * chunk, (c + 1) * chunk) of the wanted range: it jumps from the call's start state by its own distance -- one
// polynomial per set bit >= 10 of the distance, the low 10 bits are walked -- and then generates.  One polynomial =
// 32 twists (20592 consecutive raw words in shared memory) + a 19937 x 624 bit-by-word convolution, one output word per
// thread (all reads from shared memory, ~0.15 ms per CTA).  The last CTA writes the state the sequential generator would
// leave behind (numpy's alignment: origin a multiple of 624 words from the start state's, position in [1, 624]) to
// `state_out`, and the call's start state to `backup`.
constexpr int MTJ_THREADS = 640;
constexpr int MTJ_BLOCKS = 8;                        // twists per fill: the buffer holds 624 * 9 = 5616 consecutive raw words
constexpr int MTJ_BUF = 624 * (MTJ_BLOCKS + 1);      // words (22 KB: static shared memory, the default carve-out)
constexpr int MTJ_SLABW = 624 * MTJ_BLOCKS / 32;     // polynomial words per slab (156): 4992 coefficients need X[i0 .. i0 + 5616)
constexpr int MTJ_DEG = 19937;

// X[624 (r + 1) + i] for r < nb, i < 624, from X[0..623]:  X[t + 624] = X[t + 397] ^ mix(X[t], X[t + 1]), 227 words per barrier
__device__ __forceinline__ void mtj_fill(uint32_t* __restrict__ X, int nb, int tid) {
  for (int r = 0; r < nb; ++r) {
    uint32_t* o = X + 624 * r;
    if (tid < 227) o[624 + tid] = o[397 + tid] ^ mt_mix(o[tid], o[tid + 1]);
    __syncthreads();
    if (tid < 227) o[851 + tid] = o[624 + tid] ^ mt_mix(o[227 + tid], o[228 + tid]);
    __syncthreads();
    if (tid < 170) o[1078 + tid] = o[851 + tid] ^ mt_mix(o[454 + tid], o[455 + tid]);
    __syncthreads();
  }
}
// X[0..623] <- the 624 words J = 2^b positions further on (g = row b of the table).  The 19937 coefficients are consumed in four
// slabs of 4992: each slab needs the raw words X[i0 .. i0 + 5616), which is what the buffer holds after 8 twists; its last 624
// words then become the front of the next slab's buffer.  Two output words per thread (one 64-bit shared-memory load serves two
// polynomial bits), the slab's coefficient words split between the two halves of the CTA; the coefficient word is warp-uniform.
__device__ __forceinline__ void mtj_apply(uint32_t* __restrict__ X, uint32_t* __restrict__ P, const uint32_t* __restrict__ g, int tid) {
  for (int i = tid; i < 624; i += MTJ_THREADS) P[i] = __ldg(g + i);
  const int grp = tid / 320, t = tid - grp * 320;
  uint32_t a0 = 0, a1 = 0;
  for (int slab = 0; slab < 4; ++slab) {
    mtj_fill(X, MTJ_BLOCKS, tid);                                        // (its barriers also publish P and the moved front)
    if (t < 312) {
      const int w0 = grp * (MTJ_SLABW / 2), w1 = w0 + MTJ_SLABW / 2;
      for (int iw = w0; iw < w1; ++iw) {
        const uint32_t gw = P[slab * MTJ_SLABW + iw];
        const uint2* xp = reinterpret_cast<const uint2*>(X + iw * 32 + 2 * t);
        uint2 cur = xp[0];
#pragma unroll
        for (int b = 0; b < 32; b += 2) {
          const uint2 nxt = xp[b / 2 + 1];
          if (gw & (1u << b)) { a0 ^= cur.x; a1 ^= cur.y; }
          if (gw & (2u << b)) { a0 ^= cur.y; a1 ^= nxt.x; }
          cur = nxt;
        }
      }
    }
    __syncthreads();
    if (slab < 3) {                                                      // the newest 624 words start the next slab
      const uint32_t v = tid < 624 ? X[624 * MTJ_BLOCKS + tid] : 0u;
      __syncthreads();
      if (tid < 624) X[tid] = v;
      __syncthreads();
    }
  }
  if (grp == 1 && t < 312) { P[2 * t] = a0; P[2 * t + 1] = a1; }
  __syncthreads();
  if (grp == 0 && t < 312) { X[2 * t] = a0 ^ P[2 * t]; X[2 * t + 1] = a1 ^ P[2 * t + 1]; }
  __syncthreads();
}
__global__ void __launch_bounds__(MTJ_THREADS)
mt19937_jump_bits_kernel(const uint32_t* __restrict__ state, uint32_t* __restrict__ state_out, uint32_t* __restrict__ backup,
                         const uint32_t* __restrict__ tab, long long skip_before, long long count, long long total,
                         long long chunk, uint8_t* __restrict__ out) {
  __shared__ uint32_t mtj_smem[MTJ_BUF + 624];
  uint32_t* X = mtj_smem;
  uint32_t* P = mtj_smem + MTJ_BUF;
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  const bool last = c == (int)gridDim.x - 1;
  const int pos = (int)state[624];
  for (int i = tid; i < 624; i += MTJ_THREADS) { const uint32_t v = state[i]; X[i] = v; if (last && backup) backup[i] = v; }
  if (last && backup && tid == 0) backup[624] = (uint32_t)pos;
  __syncthreads();
  long long skip, q = 0, T = 0;
  if (last) {
    q = (long long)pos + total;
    T = total == 0 ? 0 : 624 * ((q - 1) / 624);
    if (T == 0) {                                   // still inside the start state's own block of 624 words
      for (int i = tid; i < 624; i += MTJ_THREADS) state_out[i] = X[i];
      if (tid == 0) state_out[624] = (uint32_t)q;
      return;
    }
    skip = T - pos;
  } else {
    skip = skip_before + (long long)c * chunk;
  }
  // the low 31 bits of word 0 of a freshly seeded state are not part of the sequence: work on the window one word further on
  const int base = pos >= 1 ? 1 : 0;
  if (base) {
    mtj_fill(X, 1, tid);
    const uint32_t v = tid < 624 ? X[tid + 1] : 0u;
    __syncthreads();
    if (tid < 624) X[tid] = v;
    __syncthreads();
  }
  const long long J = (long long)(pos - base) + skip;
  const int lo = (int)(J & 1023);
  {
    long long hi = J >> 10;
    for (int b = 10; hi != 0; ++b, hi >>= 1)
      if (hi & 1) mtj_apply(X, P, tab + (size_t)b * 624, tid);
  }
  if (last) {
    mtj_fill(X, 2, tid);                            // lo + 624 <= 1647 < 1872
    for (int i = tid; i < 624; i += MTJ_THREADS) state_out[i] = X[lo + i];
    if (tid == 0) state_out[624] = (uint32_t)(q - T);
    return;
  }
  long long n_out = count - (long long)c * chunk;
  if (n_out > chunk) n_out = chunk;
  uint8_t* o = out + (size_t)c * chunk;
  long long done = 0;
  int p = lo;
  while (done < n_out) {
    const long long need = (long long)p + (n_out - done) - 624;          // words wanted beyond the 624 in the buffer
    const int nb = need <= 0 ? 0 : (need >= (long long)624 * MTJ_BLOCKS ? MTJ_BLOCKS : (int)((need + 623) / 624));
    mtj_fill(X, nb, tid);
    const int have = 624 * (nb + 1) - p;
    const int take = (n_out - done) < (long long)have ? (int)(n_out - done) : have;
    for (int i = tid; i < take; i += MTJ_THREADS) o[done + i] = (uint8_t)(mt_temper(X[p + i]) & 1u);
    done += take;
    if (done < n_out) {                             // the newest 624 words become the front of the next pass (all of them used)
      const uint32_t v = tid < 624 ? X[624 * nb + tid] : 0u;
      __syncthreads();
      if (tid < 624) X[tid] = v;
      __syncthreads();
      p = 624;
    }
  }
}

// element i of probe p = 2*lsb[p*n+i] - 1 (one byte per element, as written by mt19937_bits_kernel).  The generator's order is
// probe-major, X0 is element-major: a 32-probe x 128-element tile goes through shared memory so that both the byte reads (along
// i) and the complex128 writes (along p) are coalesced.  block (32, 8), grid (ceil(k / 32), ceil(n / 128)).
__global__ void __launch_bounds__(256)
probe_expand_bytes_kernel(const uint8_t* __restrict__ lsb, int n, int k, Cx<double>* __restrict__ X0) {
  __shared__ uint8_t tile[32][132];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int p0 = blockIdx.x * 32, i0 = blockIdx.y * 128;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r;
    if (p >= k) continue;
    const size_t base = (size_t)p * n + i0;
    if (i0 + 128 <= n && ((base & 3) == 0)) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(lsb + base) + tx);
      *reinterpret_cast<uint32_t*>(&tile[r][4 * tx]) = v;
    } else {
      for (int q = 0; q < 4; ++q) { const int i = i0 + 4 * tx + q; tile[r][4 * tx + q] = i < n ? __ldg(lsb + (size_t)p * n + i) : 0; }
    }
  }
  __syncthreads();
  const int p = p0 + tx;
  if (p >= k) return;
  for (int ii = ty; ii < 128; ii += 8) {
    const int i = i0 + ii;
    if (i < n) X0[(size_t)i * k + p] = cx<double>(tile[tx][ii] ? 1.0 : -1.0, 0.0);
  }
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

// ---- set-up: the values of the aggregation prolongator (multigrid.py:232-259) -----------------------------------------------
// One warp per (aggregate, half): the nvec test-vector pieces on the half's rows (row q of the half = aggregate row
// (q / hd) * dofi + q % hd + half * hd, hd = dofi / 2) are orthonormalised by CLASSICAL Gram-Schmidt exactly as the reference
// does it -- the projections of column c on the finished columns w < c are all taken from the unmodified column, subtracted in
// the order w = 0, 1, ..., then the column is divided by its norm -- in complex128.  Only the order of the additions inside an
// inner product differs from numpy's (lanes, then a butterfly), i.e. the values agree with the host builder to ~1e-15.
// ev: [n][ld] row-major (first nvec columns used), pv: [n][nvec].  Dynamic shared memory: warps_per_block * mrows * nvec complex.
__global__ void __launch_bounds__(128)
prolongator_values_kernel(const Cx<double>* __restrict__ ev, int ld, int n, int aggr_size, int dofi, int nvec,
                          Cx<double>* __restrict__ pv, const int* __restrict__ rows, int passes) {
  extern __shared__ double2 pvk_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long item = (long long)blockIdx.x * wpb + wib;            // (aggregate, half)
  const int hd = dofi / 2, mrows = aggr_size / 2;
  const long long nitems = 2ll * (n / aggr_size);
  if (item >= nitems) return;
  const int half = (int)(item & 1);
  const size_t base = (size_t)(item >> 1) * aggr_size;
  double2* B = pvk_smem + (size_t)wib * mrows * nvec;                  // B[q * nvec + v]
  // rows != nullptr: the block's rows are rows[item * mrows + q] (indexed aggregates of the geometric hierarchies)
  auto row_of = [&](int q) -> size_t {
    return rows != nullptr ? (size_t)rows[(size_t)item * mrows + q] : base + (size_t)(q / hd) * dofi + (q % hd) + half * hd;
  };
  for (int idx = lane; idx < mrows * nvec; idx += 32) {
    const int q = idx / nvec, v = idx - q * nvec;
    const Cx<double> e = ev[row_of(q) * ld + v];
    B[idx] = make_double2(e.re, e.im);
  }
  __syncwarp();
  for (int c = 0; c < nvec; ++c) {
   // passes = 2: the projections are taken and subtracted a second time before the normalisation (blocks whose test-vector
   // pieces are nearly dependent; the reference's aggregates use one pass, multigrid.py:232-259)
   for (int pass = 0; pass < passes; ++pass) {
    // rs[w] = <col_w, col_c> for w < c, from the unmodified column c
    double2 rs[16];
    for (int w = 0; w < c; ++w) {
      double sr = 0.0, si = 0.0;
      for (int q = lane; q < mrows; q += 32) {
        const double2 a = B[q * nvec + w], b = B[q * nvec + c];
        sr += a.x * b.x + a.y * b.y;                                   // conj(a) * b
        si += a.x * b.y - a.y * b.x;
      }
      for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
      rs[w] = make_double2(sr, si);
    }
    for (int w = 0; w < c; ++w) {
      for (int q = lane; q < mrows; q += 32) {
        const double2 a = B[q * nvec + w];
        double2 b = B[q * nvec + c];
        b.x -= rs[w].x * a.x - rs[w].y * a.y;
        b.y -= rs[w].x * a.y + rs[w].y * a.x;
        B[q * nvec + c] = b;
      }
      __syncwarp();
    }
   }
    double sq = 0.0;
    for (int q = lane; q < mrows; q += 32) { const double2 b = B[q * nvec + c]; sq += b.x * b.x + b.y * b.y; }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const double nrm = sqrt(sq);
    for (int q = lane; q < mrows; q += 32) { double2 b = B[q * nvec + c]; b.x /= nrm; b.y /= nrm; B[q * nvec + c] = b; }
    __syncwarp();
  }
  for (int idx = lane; idx < mrows * nvec; idx += 32) {
    const int q = idx / nvec, v = idx - q * nvec;
    pv[row_of(q) * nvec + v] = cx<double>(B[idx].x, B[idx].y);
  }
}

// the complex64 and the splatted (re, re, im, im) copies of a complex128 dense inverse (dmlmc_set_dense_inverse_device_full)
__global__ void __launch_bounds__(256) dense_formats_kernel(const Cx<double>* __restrict__ M, size_t count, Cx<float>* __restrict__ F,
                                                            float4* __restrict__ F4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Cx<double> m = M[i];
  const float mr = (float)m.re, mi = (float)m.im;
  F[i] = cx<float>(mr, mi);
  F4[i] = make_float4(mr, mr, mi, mi);
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
