// dense_umma.cuh -- the dense coarse solve of the V-cycle on the 5th-generation tensor cores.
//
//   X = A_l^{-1} B     (multigrid.py:413-416 moved up to the first level small enough for a dense
//                       inverse; inside the V-cycle it is a preconditioner, so BF16 operands with FP32
//                       accumulation are enough -- the outer FGMRES is complex128 and verifies the true
//                       residual.  The coarsest-level solve of the MLMC estimator itself stays FP64.)
//
// The complex product is one real GEMM: every complex entry m of A^{-1} becomes the 2x2 block
// [[mr, -mi], [mi, mr]] of Mt[2n][2n] (BF16, K contiguous), a complex vector is its interleaved
// (re, im) real vector, and the interleaved result rows come out of the same GEMM:
//   D[2n][k] = Mt[2n][2n] * Bt[k][2n]^T       (both operands K-major, D in FP32)
//
// Kernel anatomy (one CTA = one 128-row tile of D, 6 warps):
//   warp 0   TMA producer: cp.async.bulk.tensor 128x64 (A) and 256x64 (B) BF16 boxes, SWIZZLE_128B,
//            into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1   allocates TMEM (256 columns = the 128 x 256 FP32 accumulator) and issues
//            tcgen05.mma.cta_group::1.kind::f16 (M128 x N<=256 x K16) from one elected lane;
//            tcgen05.commit releases the ring slots and finally signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32x32b (lane = row of D), re/im rows are neighbouring lanes, one
//            shuffle pairs them into interleaved complex64 and each lane stores 128 contiguous bytes
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmlmc {

constexpr int UM_BM = 128;        // rows of D per CTA (UMMA M)
constexpr int UM_BN = 256;        // columns of D per CTA (UMMA N <= 256)
constexpr int UM_BK = 64;         // BF16 elements per 128-byte swizzled row
constexpr int UM_STAGES = 4;
constexpr int UM_THREADS = 192;
constexpr uint32_t UM_A_BYTES = UM_BM * UM_BK * 2;    // 16 KB
constexpr uint32_t UM_B_BYTES = UM_BN * UM_BK * 2;    // 32 KB
constexpr size_t UM_SMEM = (size_t)UM_STAGES * (UM_A_BYTES + UM_B_BYTES) + 1024 /* alignment slack */ + 256 /* barriers */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded spin: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major operand, rows of 128 bytes,
// SWIZZLE_128B, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: BF16 x BF16 -> FP32, both K-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4)                       // c_format  = F32
       | (1u << 7)                       // a_format  = BF16
       | (1u << 10)                      // b_format  = BF16
       | ((uint32_t)(N >> 3) << 17)      // n_dim
       | ((uint32_t)(M >> 4) << 24);     // m_dim
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// D = Mt * Bt^T, written as interleaved complex64 Y[n][ldy]:  D row 2i (+1) = Re (Im) of row i.
//   tmA: Mt[two_n][two_n] BF16, box 64 x 128;  tmB: Bt[k][two_n] BF16, box 64 x 256 (rows >= k read as 0)
__global__ void __launch_bounds__(UM_THREADS, 1)
dense_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ Y, int ldy, int two_n, int k, int kdim) {   // kdim: length of the contraction (two_n, or
                                                                                  // 3 * two_n for the split-BF16 operands)
  extern __shared__ uint8_t um_smem_raw[];
  const uint32_t base = (smem_u32(um_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = um_smem_raw + (base - smem_u32(um_smem_raw));
  uint8_t* sA = smem;
  uint8_t* sB = smem + UM_STAGES * UM_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + UM_STAGES * UM_B_BYTES);
  uint64_t* empty = full + UM_STAGES;
  uint64_t* accum_full = empty + UM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (column tiles, row tiles): the CTAs that share a 128-row slab of the matrix are neighbours in launch order, run at
  // the same time and the slab comes from DRAM once (the matrix, 537 MB at n = 8192, does not fit in L2)
  const int row0 = blockIdx.y * UM_BM;                  // first row of D of this CTA
  const int col0 = blockIdx.x * UM_BN;                  // first column
  const int ncols = min(UM_BN, k - col0);
  const int bn = (ncols + 15) & ~15;                    // UMMA N: multiple of 16 for M = 128
  const int num_kb = (kdim + UM_BK - 1) / UM_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < UM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % UM_STAGES;
        const uint32_t ph = (uint32_t)(kb / UM_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);                  // slot free (first pass: passes immediately)
        mbar_arrive_expect_tx(&full[s], UM_A_BYTES + UM_B_BYTES);
        tma_load_2d(sA + (size_t)s * UM_A_BYTES, &tmA, &full[s], kb * UM_BK, row0);
        tma_load_2d(sB + (size_t)s * UM_B_BYTES, &tmB, &full[s], kb * UM_BK, col0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(UM_BM, bn);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % UM_STAGES;
      const uint32_t ph = (uint32_t)(kb / UM_STAGES) & 1u;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(sA + (size_t)s * UM_A_BYTES));
        const uint64_t db = umma_desc_k_sw128(smem_u32(sB + (size_t)s * UM_B_BYTES));
#pragma unroll
        for (int kk = 0; kk < UM_BK / 16; ++kk)         // UMMA K = 16 BF16 = 32 bytes: +2 in the 16-byte address field
          umma_bf16(tmem_d, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);                         // frees the slot when these MMAs have read it
        if (kb == num_kb - 1) umma_commit(accum_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> interleaved complex64 =====
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    mbar_wait(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = row0 + q * 32 + lane;                 // row of D: 2i + part
    const int i = r >> 1;
    const bool odd = lane & 1;
    for (int c0 = 0; c0 < bn; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      // even lane (Re row) keeps columns [c0, c0+16), odd lane (Im row) keeps [c0+16, c0+32)
      float2 out[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t send = odd ? v[j] : v[16 + j];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        out[j] = odd ? make_float2(__uint_as_float(recv), __uint_as_float(v[16 + j]))
                     : make_float2(__uint_as_float(v[j]), __uint_as_float(recv));
      }
      const int cbase = col0 + c0 + (odd ? 16 : 0);
      if (r < two_n) {
        float2* dst = reinterpret_cast<float2*>(Y) + (size_t)i * ldy + cbase;
        if (cbase + 16 <= k && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            *reinterpret_cast<float4*>(dst + j) = make_float4(out[j].x, out[j].y, out[j + 1].x, out[j + 1].y);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cbase + j < k) dst[j] = out[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
  }
}

// Mt[2i + a][2j + b] from the complex128 inverse Minv[n][n] (row-major):  [[mr, -mi], [mi, mr]]
__global__ void __launch_bounds__(256)
umma_expand_matrix_kernel(const Cx<double>* __restrict__ Minv, int n, __nv_bfloat16* __restrict__ Mt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const size_t i = idx / n, j = idx - i * n;
  const Cx<double> m = ldc_ro<double>(Minv, idx);
  const size_t two_n = 2 * (size_t)n;
  __nv_bfloat162 top, bot;
  top.x = __float2bfloat16_rn((float)m.re); top.y = __float2bfloat16_rn((float)-m.im);
  bot.x = __float2bfloat16_rn((float)m.im); bot.y = __float2bfloat16_rn((float)m.re);
  *reinterpret_cast<__nv_bfloat162*>(Mt + (2 * i) * two_n + 2 * j) = top;
  *reinterpret_cast<__nv_bfloat162*>(Mt + (2 * i + 1) * two_n + 2 * j) = bot;
}

// Split-BF16 operands (FP32-class accuracy from BF16 tensor cores): m = hi + lo with hi = bf16(m), lo = bf16(m - hi),
// likewise the right-hand side, and  M B ~= M_hi B_hi + M_lo B_hi + M_hi B_lo  (the dropped lo x lo term is 2^-16
// relative) as ONE GEMM over the concatenated contraction:  Mt3[2n][6n] = [hi | lo | hi],  Bt3[k][6n] = [hi | hi | lo].
// Used where a small level's dense inverse is itself the preconditioner of that level's solve (the MLMC coarse
// solve at n = 2048): 3 outer iterations like the FP32 SIMT kernel, at a third of its time.
__device__ __forceinline__ void bf16_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__global__ void __launch_bounds__(256)
umma_expand_matrix_split_kernel(const Cx<double>* __restrict__ Minv, int n, __nv_bfloat16* __restrict__ Mt3) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const size_t i = idx / n, j = idx - i * n;
  const Cx<double> m = ldc_ro<double>(Minv, idx);
  const size_t two_n = 2 * (size_t)n, ld = 3 * two_n;
  __nv_bfloat16 rh, rl, ih, il;
  bf16_split((float)m.re, rh, rl); bf16_split((float)m.im, ih, il);
  const __nv_bfloat16 nih = __hneg(ih), nil = __hneg(il);
  __nv_bfloat16* top = Mt3 + (2 * i) * ld + 2 * j;
  __nv_bfloat16* bot = Mt3 + (2 * i + 1) * ld + 2 * j;
  __nv_bfloat162 t;
  t.x = rh; t.y = nih; *reinterpret_cast<__nv_bfloat162*>(top) = t; *reinterpret_cast<__nv_bfloat162*>(top + 2 * two_n) = t;
  t.x = rl; t.y = nil; *reinterpret_cast<__nv_bfloat162*>(top + two_n) = t;
  t.x = ih; t.y = rh;  *reinterpret_cast<__nv_bfloat162*>(bot) = t; *reinterpret_cast<__nv_bfloat162*>(bot + 2 * two_n) = t;
  t.x = il; t.y = rl;  *reinterpret_cast<__nv_bfloat162*>(bot + two_n) = t;
}
// Bt3[c][0:2n] = Bt3[c][2n:4n] = hi, Bt3[c][4n:6n] = lo of the interleaved (Re, Im) column c of X
template <typename T>
__global__ void __launch_bounds__(256)
umma_pack_rhs_split_kernel(const Cx<T>* __restrict__ X, int ldx, int n, int k, __nv_bfloat16* __restrict__ Bt3) {
  __shared__ uint32_t thi[32][33], tlo[32][33];
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;          // block (32, 8)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty + 8 * r, c = c0 + tx;
    uint32_t ph = 0, pl = 0;
    if (j < n && c < k) {
      const Cx<T> v = ldc_ro<T>(X, (size_t)j * ldx + c);
      __nv_bfloat162 bh, bl;
      bf16_split((float)v.re, bh.x, bl.x); bf16_split((float)v.im, bh.y, bl.y);
      ph = *reinterpret_cast<uint32_t*>(&bh); pl = *reinterpret_cast<uint32_t*>(&bl);
    }
    thi[ty + 8 * r][tx] = ph; tlo[ty + 8 * r][tx] = pl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, j = j0 + tx;
    if (j < n && c < k) {
      uint32_t* row = reinterpret_cast<uint32_t*>(Bt3) + (size_t)c * 3 * n;
      row[j] = thi[tx][ty + 8 * r]; row[n + j] = thi[tx][ty + 8 * r]; row[2 * n + j] = tlo[tx][ty + 8 * r];
    }
  }
}

// Bt[c][2j], Bt[c][2j+1] = BF16(Re, Im) of X[j][c]: transpose through a 32 x 32 shared-memory tile
template <typename T>
__global__ void __launch_bounds__(256)
umma_pack_rhs_kernel(const Cx<T>* __restrict__ X, int ldx, int n, int k, __nv_bfloat16* __restrict__ Bt) {
  __shared__ uint32_t tile[32][33];
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;          // block (32, 8)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty + 8 * r, c = c0 + tx;
    uint32_t packed = 0;
    if (j < n && c < k) {
      const Cx<T> v = ldc_ro<T>(X, (size_t)j * ldx + c);
      __nv_bfloat162 b; b.x = __float2bfloat16_rn((float)v.re); b.y = __float2bfloat16_rn((float)v.im);
      packed = *reinterpret_cast<uint32_t*>(&b);
    }
    tile[ty + 8 * r][tx] = packed;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, j = j0 + tx;
    if (j < n && c < k) reinterpret_cast<uint32_t*>(Bt)[(size_t)c * n + j] = tile[tx][ty + 8 * r];
  }
}

// ---- host side: tensor maps -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// BF16 matrix [rows][cols] row-major (cols contiguous), box = 64 columns x box_rows rows, SWIZZLE_128B
inline int make_tmap_bf16(CUtensorMap* m, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)UM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace dmlmc
