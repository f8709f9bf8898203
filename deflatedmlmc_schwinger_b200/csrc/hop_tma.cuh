// hop_tma.cuh -- the half-lattice sweep of the even-odd smoother with the neighbour halo staged in shared memory by TMA.
//
//   Out_p[site] = a * In2_p[site] + b * (H In_q)[site]          (same operation, layouts and arithmetic as
//                                                                wilson_hop_eo_kernel<HAS2, true, false, 2> in op_kernels.cuh)
//
// Why: ncu of the direct kernel at k = 512 (profiles/r2_run2_ncu_full_k512.md) -- long-scoreboard stalls 4.1 per issued
// instruction, 36 % of the warp slots, L1 hit rate 25 %: every site row is fetched from L2 by the four sites that neighbour
// it, and the kernel waits for those loads.  Here a CTA walks over tiles of HT_X x HT_TH sites x 64 columns; for each tile ONE
// warp issues bulk asynchronous copies (cp.async.bulk, the TMA unit; one 256-byte row segment each, the periodic wrap is in the
// source address) of the HT_X (HT_TH + 1) + 2 HT_TH neighbour rows and the HT_X HT_TH centre rows into the next stage of a
// two-stage ring while the CTA computes the current one out of shared memory; completion is counted in bytes on an mbarrier.
// L2 -> SM traffic per site falls from ~6 + 1 row fetches to 1.6 + 1, and no thread waits on a global load of vector data.
//
// MEASURED (B200, 128^2, k = 512, `bench.py --opt hop_tma=1`, profiles/r2_run9_*, r2_run10_*): bit-identical output
// (tests/test_gpu_kernels.py::test_tma_staged_hop_kernel_is_bit_exact) but SLOWER than the direct kernel -- 46.3 us per sweep
// with 256-byte rows (4 x 8 sites x 64 columns per tile, 168 copies), 27.8 us with 1 KB rows (2 x 4 sites x 256 columns, 52
// copies) against 18.0 us: with two stages and two CTAs per SM ~100 KB are in flight per SM, no more than the direct kernel's
// 28 warps x 10 outstanding 16-byte loads, and every tile pays the copy latency once more.  The option (hop_tma) is OFF.
//
// Tile rows in shared memory (per stage, per spin):  Q[r][64 columns], r =
//   xl (HT_TH + 1) + j,  j = 0..HT_TH      row x0 + xl of the q-parity array, th = th0 - (1 - a_xl) + j,   a_xl = (x0 + xl + p) & 1
//                                           (forward neighbour of site thl: j = thl + 1, backward: j = thl, for either a)
//   HT_X (HT_TH + 1) + thl                  row x0 - 1, th = th0 + thl
//   HT_X (HT_TH + 1) + HT_TH + thl          row x0 + HT_X, th = th0 + thl
// and C[xl HT_TH + thl][64 columns] for In2.
#pragma once
#include "common.cuh"
#include "op_kernels.cuh"
#include "dense_umma.cuh"      // smem_u32, mbar_init, mbar_arrive_expect_tx, mbar_try_wait

namespace dmlmc {

#ifndef DMLMC_HT_X
#define DMLMC_HT_X 2
#define DMLMC_HT_TH 4
#define DMLMC_HT_CW 128
#endif
// tile: HT_X x HT_TH sites, HT_CW packs of two columns.  Measured at 128^2, k = 512 (sweep with / without In2 averaged):
// 4 x 8 x 32 (168 copies of 256 B per tile): 46 us against 18 us of the direct kernel -- the TMA unit is request-rate bound
// (~1 request per 38 cycles per SM); hence 1 KB rows: 2 x 4 x 128 (52 copies of 1 KB).
constexpr int HT_X = DMLMC_HT_X, HT_TH = DMLMC_HT_TH, HT_CW = DMLMC_HT_CW;
constexpr int HT_QROWS = HT_X * (HT_TH + 1) + 2 * HT_TH;             // 52
constexpr int HT_CROWS = HT_X * HT_TH;                               // 32
constexpr int HT_ROWBYTES = HT_CW * 8;                               // 256 B per row and spin
constexpr int HT_STAGE_BYTES = (HT_QROWS + HT_CROWS) * 2 * HT_ROWBYTES;   // 43008
constexpr int HT_THREADS = 256;
constexpr size_t HT_SMEM = 2 * (size_t)HT_STAGE_BYTES + 64;

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lds_packs2(const uint2* p, C2 (&out)[2]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  out[0].re = make_float2(bf_lo(raw.x), bf_lo(raw.y)); out[0].im = make_float2(bf_hi(raw.x), bf_hi(raw.y));
  out[1].re = make_float2(bf_lo(raw.z), bf_lo(raw.w)); out[1].im = make_float2(bf_hi(raw.z), bf_hi(raw.w));
}

template <bool HAS2>
__global__ void __launch_bounds__(HT_THREADS, 2)
wilson_hop_eo_tma_kernel(int LX, int LT, int p, const float4* __restrict__ L4, const uint2* __restrict__ Inq,
                         const uint2* __restrict__ In2, uint2* __restrict__ Outp, float ar, float ai, float br, float bi,
                         uint32_t kp, int ntiles) {
  extern __shared__ __align__(128) unsigned char ht_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ht_smem + 2 * (size_t)HT_STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t LH = (uint32_t)LT >> 1;
  const uint32_t V = (uint32_t)LX * LT, VH = V >> 1;
  const size_t kpz = kp, sp = (size_t)VH * kpz;                       // spin stride (packs)
  const int tiles_c = (int)(kp / HT_CW), tiles_th = (int)(LH / HT_TH);
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // warp 0: all copies of one tile into stage s
  auto issue = [&](int tile, int s) {
    const int ct = tile % tiles_c, rest = tile / tiles_c;
    const int tht = rest % tiles_th, xt = rest / tiles_th;
    const uint32_t x0 = xt * HT_X, th0 = tht * HT_TH, c0 = ct * HT_CW;
    unsigned char* base = ht_smem + (size_t)s * HT_STAGE_BYTES;
    if (lane == 0) mbar_arrive_expect_tx(&bars[s], (uint32_t)((HT_QROWS + (HAS2 ? HT_CROWS : 0)) * 2 * HT_ROWBYTES));
    __syncwarp();
    const int nq = HT_QROWS * 2, nc = HAS2 ? HT_CROWS * 2 : 0;
    for (int i = lane; i < nq + nc; i += 32) {
      const bool isq = i < nq;
      const int ii = isq ? i : i - nq;
      const int spin = ii & 1, r = ii >> 1;
      uint32_t xg, thg;
      if (!isq) {
        xg = x0 + r / HT_TH; thg = th0 + r % HT_TH;
      } else if (r < HT_X * (HT_TH + 1)) {
        const int xl = r / (HT_TH + 1), j = r % (HT_TH + 1);
        xg = x0 + xl;
        const uint32_t a = (xg + (uint32_t)p) & 1u;
        thg = (th0 + LH - (1u - a) + j) % LH;
      } else {
        const int q = r - HT_X * (HT_TH + 1);
        xg = q < HT_TH ? (x0 + LX - 1) % LX : (x0 + HT_X) % LX;
        thg = th0 + (q % HT_TH);
      }
      const uint2* src = (isq ? Inq : In2) + ((size_t)xg * LH + thg) * kpz + c0 + (spin ? sp : 0);
      unsigned char* dst = base + ((size_t)(isq ? 0 : HT_QROWS * 2) + (size_t)r * 2 + spin) * HT_ROWBYTES;
      bulk_g2s(dst, src, HT_ROWBYTES, &bars[s]);
    }
  };

  const float2 b_r = make_float2(br, br), b_i = make_float2(bi, bi);
  const float2 a_r = make_float2(ar, ar), a_i = make_float2(ai, ai);
  constexpr int LPS = HT_CW / 2;                                      // lanes per site (two packs = four columns each)
  constexpr int SPP = HT_THREADS / LPS;                               // sites per pass
  constexpr int NPASS = HT_X * HT_TH / SPP;
  static_assert(HT_THREADS % LPS == 0 && (HT_X * HT_TH) % SPP == 0, "tile shape");
  const int l16 = tid % LPS, st = tid / LPS;
  uint32_t phase[2] = {0u, 0u};
  int s = 0;
  int tile = blockIdx.x;
  if (warp == 0 && tile < ntiles) issue(tile, 0);
  for (; tile < ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (warp == 0 && next < ntiles) issue(next, s ^ 1);               // (stage s ^ 1 was released by the barrier below)
    while (!mbar_try_wait(&bars[s], phase[s])) { }
    phase[s] ^= 1u;
    const int ct = tile % tiles_c, rest = tile / tiles_c;
    const int tht = rest % tiles_th, xt = rest / tiles_th;
    const uint32_t x0 = xt * HT_X, th0 = tht * HT_TH, c0 = ct * HT_CW;
    const uint2* Q = reinterpret_cast<const uint2*>(ht_smem + (size_t)s * HT_STAGE_BYTES);
    const uint2* C = Q + (size_t)HT_QROWS * 2 * HT_CW;
    auto qrow = [&](int r, int spin) { return Q + ((size_t)r * 2 + spin) * HT_CW + 2 * l16; };
#pragma unroll
    for (int it = 0; it < NPASS; ++it) {
      const int sidx = st + SPP * it;
      const int xl = sidx / HT_TH, thl = sidx % HT_TH;
      const uint32_t x = x0 + xl, th = th0 + thl;
      const uint32_t a = (x + (uint32_t)p) & 1u;
      const uint32_t t = 2 * th + a;
      const uint32_t site = x * LT + t;
      const float4* lp = L4 + site;
      const float4 ut = __ldg(lp), utb = __ldg(lp + V), ux = __ldg(lp + 2 * V), uxb = __ldg(lp + 3 * V);
      const int rf = xl * (HT_TH + 1) + thl + 1, rb = xl * (HT_TH + 1) + thl;
      const int rr = (xl + 1 < HT_X) ? (xl + 1) * (HT_TH + 1) + thl + (int)a : HT_X * (HT_TH + 1) + HT_TH + thl;
      const int rl = (xl >= 1) ? (xl - 1) * (HT_TH + 1) + thl + (int)a : HT_X * (HT_TH + 1) + thl;
      C2 f0[2], f1[2], b0[2], b1[2], r0[2], r1[2], l0[2], l1[2], c0v[2], c1v[2];
      lds_packs2(qrow(rf, 0), f0); lds_packs2(qrow(rf, 1), f1);
      lds_packs2(qrow(rb, 0), b0); lds_packs2(qrow(rb, 1), b1);
      lds_packs2(qrow(rr, 0), r0); lds_packs2(qrow(rr, 1), r1);
      lds_packs2(qrow(rl, 0), l0); lds_packs2(qrow(rl, 1), l1);
      if constexpr (HAS2) {
        const int rc = xl * HT_TH + thl;
        lds_packs2(C + ((size_t)rc * 2) * HT_CW + 2 * l16, c0v);
        lds_packs2(C + ((size_t)rc * 2 + 1) * HT_CW + 2 * l16, c1v);
      }
      C2 o0[2], o1[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        C2 pa, pb, pc, pd;
        pa.re = __fadd2_rn(f0[u].re, neg2(f1[u].re)); pa.im = __fadd2_rn(f0[u].im, neg2(f1[u].im));
        pb.re = __fadd2_rn(b0[u].re, b1[u].re);       pb.im = __fadd2_rn(b0[u].im, b1[u].im);
        pc.re = __fadd2_rn(r0[u].re, neg2(r1[u].im)); pc.im = __fadd2_rn(r0[u].im, r1[u].re);
        pd.re = __fadd2_rn(l0[u].re, l1[u].im);       pd.im = __fadd2_rn(l0[u].im, neg2(l1[u].re));
        const C2 ua = cmul_splat(ut, pa), ub = cmul_splat(utb, pb), uc = cmul_splat(ux, pc), ud = cmul_splat(uxb, pd);
        C2 h0, h1, q, tt;
        h0.re = neg2(__fadd2_rn(__fadd2_rn(ua.re, ub.re), __fadd2_rn(uc.re, ud.re)));
        h0.im = neg2(__fadd2_rn(__fadd2_rn(ua.im, ub.im), __fadd2_rn(uc.im, ud.im)));
        q.re = __fadd2_rn(ua.re, neg2(ub.re));  q.im = __fadd2_rn(ua.im, neg2(ub.im));
        tt.re = __fadd2_rn(uc.re, neg2(ud.re)); tt.im = __fadd2_rn(uc.im, neg2(ud.im));
        h1.re = __fadd2_rn(q.re, neg2(tt.im));
        h1.im = __fadd2_rn(q.im, tt.re);
        o0[u].re = __ffma2_rn(neg2(b_i), h0.im, __fmul2_rn(b_r, h0.re));
        o0[u].im = __ffma2_rn(b_i, h0.re, __fmul2_rn(b_r, h0.im));
        o1[u].re = __ffma2_rn(neg2(b_i), h1.im, __fmul2_rn(b_r, h1.re));
        o1[u].im = __ffma2_rn(b_i, h1.re, __fmul2_rn(b_r, h1.im));
        if constexpr (HAS2) {
          o0[u].re = __ffma2_rn(neg2(a_i), c0v[u].im, __ffma2_rn(a_r, c0v[u].re, o0[u].re));
          o0[u].im = __ffma2_rn(a_i, c0v[u].re, __ffma2_rn(a_r, c0v[u].im, o0[u].im));
          o1[u].re = __ffma2_rn(neg2(a_i), c1v[u].im, __ffma2_rn(a_r, c1v[u].re, o1[u].re));
          o1[u].im = __ffma2_rn(a_i, c1v[u].re, __ffma2_rn(a_r, c1v[u].im, o1[u].im));
        }
      }
      const size_t ic = ((size_t)x * LH + th) * kpz + c0 + 2 * l16;
      sth_packs<2>(Outp + ic, o0); sth_packs<2>(Outp + ic + sp, o1);
    }
    __syncthreads();                                                  // every thread is done with stage s
    s ^= 1;
  }
}

}  // namespace dmlmc
