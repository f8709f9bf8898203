// dmlmc.cu -- libdmlmc_sm100.so: device-resident multigrid hierarchy, V-cycle, batched FGMRES,
// the fused per-batch sample of utils.one_defl_Hutch_step and the C ABI of include/dmlmc.h.
// Host code here only orders kernel launches on one CUDA stream; all arithmetic is in the
// kernels of op_kernels.cuh / krylov_kernels.cuh.  There is no CPU fallback.
#include "../../include/dmlmc.h"
#include "common.cuh"
#include "op_kernels.cuh"
#include "krylov_kernels.cuh"
#include "dense_umma.cuh"
#include "hop_tma.cuh"
#include "setup_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

using namespace dmlmc;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail((int)e_, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)
#define CHECK(cond, msg) do { if (!(cond)) return fail(-1, std::string("dmlmc: ") + msg); } while (0)
#define RET(call) do { int rc_ = (call); if (rc_ != 0) return rc_; } while (0)

namespace {

constexpr int ROWS_PER_CHUNK = 256;
constexpr int MIN_ROWS_PER_CHUNK = 32;     // smallest row chunk of the complex64 Gram-Schmidt kernels (option "gs_rows")
constexpr int MAX_LEVELS = 12;

template <typename T> struct LevelT {
  // stencil (level 0) or BSR
  Cx<T>* Ut = nullptr; Cx<T>* Ux = nullptr; Cx<T> diag;
  Cx<T>* bsr_vals = nullptr;
  Cx<T>* pv = nullptr;
  Cx<T>* perm_vals = nullptr;
};

struct Level {
  int n = 0;
  int kind = -1;            // 0 stencil, 1 bsr
  int LX = 0, LT = 0;
  int bs = 0, bpr = 0, nb = 0; int* bsr_col = nullptr; float4* bsr_vals4 = nullptr;   // (mr,mr,mi,mi) per entry
  float4* links4 = nullptr;          // [4][V] (ur,ur,ui,ui): U_t(x), U_t(x-t)^*, U_x(x), U_x(x-x)^*
  bool has_transfer = false; int aggr = 0, dofi = 0, nvec = 0, n_c = 0;
  int* tr_rows = nullptr; int* tr_cblk = nullptr; int tr_m = 0;     // indexed (geometric) aggregates, else closed form
  // smoother polynomial in product form: p(A) = p0 * prod_i (I - nu_i A)
  bool has_smoother = false; std::vector<Cx<double>> nu; Cx<double> p0; bool smoother16 = true;
  // the same for the even-odd Schur complement S = c - H_eo H_oe / c of a stencil level (dmlmc_set_smoother_eo)
  bool has_eo = false; std::vector<Cx<double>> nu_eo; Cx<double> p0_eo;
  bool has_perm = false; int shift = 0, perm_nnz = 0; int* perm_cols = nullptr;
  int defl_d = 0; Cx<double>* defl_V = nullptr;
  bool has_dense = false; Cx<double>* minv_d = nullptr; Cx<float>* minv_f = nullptr; float4* minv4 = nullptr;
  // tensor-core operand of the dense inverse: Mt[2n][2n] BF16 (2x2 real block per complex entry) + its TMA map
  bool has_umma = false; __nv_bfloat16* minv_bf16 = nullptr; CUtensorMap tmA;
  const void* tmB_ptr = nullptr; int tmB_k = 0; CUtensorMap tmB;
  // split-BF16 operand [2n][6n] = [hi | lo | hi] of a small level's inverse (FP32-class accuracy on the tensor cores)
  bool has_umma3 = false; __nv_bfloat16* minv_bf16x3 = nullptr; CUtensorMap tmA3;
  const void* tmB3_ptr = nullptr; int tmB3_k = 0; CUtensorMap tmB3;
  LevelT<double> d;
  LevelT<float> f;
};

template <typename T> struct Sel;
template <> struct Sel<double> { static LevelT<double>& get(Level& L) { return L.d; } };
template <> struct Sel<float>  { static LevelT<float>&  get(Level& L) { return L.f; } };

}  // namespace

struct dmlmc_hier {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_levels = 0;
  Level lv[MAX_LEVELS];
  int inner_prec = DMLMC_C64;
  int reorth = 0;
  int stencil_by = 4, stencil_bz = 4;     // site tile (t, x) of the stencil kernel's thread block
  int bsr_threads = 128;                  // threads per CTA of the packed-FP32 BSR kernel (block rows per CTA = this / threads per row)
  int stencil_t2 = 1;                     // two t-adjacent sites per thread in the factor kernel
  int eo_packs = 2;                           // column packs per thread of that kernel (2: 16-byte loads, four columns share the links)
  int eo_by = 2, eo_bz = 2;                   // thread-block tile (t/2, x) of the even-odd hop kernel
  int stencil_t2_by = 2, stencil_t2_bz = 2;   // its thread-block tile: (32 packs, 2 thread rows = 4 sites in t, 2 in x)
  int stencil_smem = 0;                   // shared-memory-tiled variant of the packed-FP32 factor kernel
  int stencil_fast = 1;                   // packed-FP32 (FFMA2) kernel for the BF16-stored smoother factors
  int prefetch_slices = 16;               // Y = A X / B - A X on level 0: L2 prefetch distance in x-slices (0, 8, 16)
  int pre_smooth = 0;                     // 0: V-cycle = coarse correction + post-smoothing (default), 1: pre- and post-smoothing
  int smoother_half = 1;                  // BF16 storage of the level-0 smoother's intermediate vectors (c64 cycle)
  int dense_direct_exact = 1;             // a dense level solved directly (not from a finer level) uses its FP32 copy
  int outer_eo = 1;                       // outer FGMRES of a stencil level on the even-odd Schur complement (fgmres_eo)
  int outer_c64 = 1;                      // ... with the Krylov vectors V_j, Z_j, w STORED in complex64 (coefficients, solution and true
                                          // residual in complex128; a cycle runs until its estimate has dropped by outer_drop)
  double outer_drop = 1e-4;               // (measured, run r2_7: 1e-4 -> cycles of 3 + 3 + 2 iterations, 38.5k probes/s; 1e-5 / 1e-6 -> 8-9 iterations, 34.3k / 34.9k)
  std::vector<int> expect_cyc[MAX_LEVELS];   // iterations per cycle of the previous Schur-complement solve of (expect_tol, expect_k)
  int eo_zhalf = 0;                       // (internal) the even-odd smoother writes only Z_e, into a half-lattice array
  int smoother_eo = 1;                    // even-odd (Schur complement) form of the level-0 post-smoother when one is set
  int dot32 = 0;                          // Gram-Schmidt coefficients from complex64 copies of the basis vectors: OFF -- measured on
                                          // B200 (profiles/r1_run32_tune_dot32.jsonl): the basis loses orthogonality at the 1e-7
                                          // level, a cycle then stagnates near 1e-7 and the solve needs 43 iterations instead of 8
  int fuse_res = 1;                       // the V-cycle's residual before the post-smoother is written as BF16 (level-0 stencil)
  int adaptive_poll = 1;                  // skip the per-iteration convergence poll until one iteration before the count the
                                          // previous solve of the level needed (columns that are done do no work, so this is safe)
  int expect_it[MAX_LEVELS] = {}; double expect_tol[MAX_LEVELS] = {}; int expect_k[MAX_LEVELS] = {};
  int fuse_io = 1;                        // V-cycle input / output conversions fused into the neighbouring kernels (bit-identical)
  int dense_split_bf16 = 1;               // such a level (1024 <= n <= 4096) is applied as a split-BF16 tensor-core GEMM instead
  int dense_tensor_min_n = 1024;          // dense inverses at least this large are applied on the tensor cores
  bool umma_attr_set = false, dmma_attr_set = false;
  int defl_tensor = 1;                    // deflation projections on the FP64 tensor cores (d % 4 == 0, d <= 64)
  int stencil_minb = 3;                   // resident 512-thread blocks per SM the stencil kernel is compiled for
  int chunk_cols = 0;                     // V-cycle column chunk (0 = sized from l2_budget_mb)
  double l2_budget_mb = 0.0;              // MB the per-chunk working vectors may occupy (0 = no chunking)
  char* ws = nullptr; size_t ws_bytes = 0, ws_off = 0;
  int* h_nactive = nullptr;        // pinned
  cudaStream_t rng_stream = nullptr; cudaEvent_t rng_done = nullptr;   // the probe stream runs beside the solver
  cudaStream_t rng_stream_lo = nullptr; cudaEvent_t rng_order = nullptr; // (jump-ahead kernel: many CTAs, no priority over the solver)
  uint32_t* mt_tab = nullptr; int mt_tab_rows = 0;    // jump polynomials t^(2^b) mod phi (dmlmc_set_mt_jump_table)
  uint32_t* mt_state_out = nullptr;                   // [625] state written by the jump kernel, copied over the caller's afterwards
  int mt_jump = 1;                                    // option: 0 = the sequential one-CTA generator
  int fuse_residual = 1;                              // option: true residual of the Schur system + its norms in one kernel per cycle
  int gs_x2 = 1;                                      // option: two-column (16-byte) Gram-Schmidt kernels for complex64 vectors
  int gs_rows = 0;                                    // option: rows per chunk (= per thread block) of the complex64 Gram-Schmidt kernels;
                                                      // 0 = the largest power of two <= 256 that gives >= 2 048 blocks (256 gave 512 blocks
                                                      // of 80 registers for 444 resident slots at k = 512: two waves, the second 15 % full,
                                                      // ncu r2_run29: 2.6 TB/s)
  int mt_prio = 0;                                    // option: 1 = the jump-ahead kernel on the high-priority stream as well
  int hop_tma = 0;                                    // option: the even-odd sweeps with the halo staged in shared memory by TMA bulk copies
                                                      // (OFF: measured 27.8 us per sweep against 18.0 us of the direct kernel, runs r2_9 / r2_10)
  bool hop_tma_attr_set = false;
  int n_sms = 148;
  int smoother_only = 0;                              // option "precond_smoother_only": M^{-1} = the level's smoother polynomial
                                                      // (the bootstrap solver of the set-up phase, before a hierarchy exists)
  bool mtj_attr_set = false;
  cudaStream_t cap_stream = nullptr;                                   // capture stream of the V-cycle graphs
  long long launches = 0;
  int unconverged = 0;                    // columns the solves of the last fgmres / level_sample call left above tol at maxiter
  std::vector<void*> owned;
  // CUDA graphs of the V-cycle for small batches (launch-bound there: ~90 kernels of a few microseconds each)
  struct GraphEntry { int level, k, prec; char* ws; size_t ws_off; int j, m, reorth; double tol; cudaGraphExec_t exec; long long launches; };
  std::vector<GraphEntry> graphs;          // j < 0: marker "the first (eager) iteration of this configuration has run"
  int use_graphs = 1, graph_max_k = 1024;
  // optional separate preconditioner hierarchy per level (dmlmc_set_preconditioner): the V-cycle that preconditions
  // this level's FGMRES runs on (prec_hier[level], prec_level[level]) instead of on this hierarchy's own levels
  dmlmc_hier* prec_hier[MAX_LEVELS] = {}; int prec_level[MAX_LEVELS] = {};
};

namespace {

// Device memory of the operators comes from the device's stream-ordered pool (cudaMallocAsync on the handle's stream; the pool
// keeps what is freed: dmlmc_hier_create sets its release threshold): set-up builds and discards whole hierarchies (the rough
// one of the two-stage eigensolve), and cudaFree / cudaMalloc of hundreds of buffers cost seconds there (1.5 s per discarded
// hierarchy at 1024^2, profiles/r2_run27_*).
template <typename T> cudaError_t dev_alloc(dmlmc_hier* h, T** p, size_t bytes) {
  return cudaMallocAsync(reinterpret_cast<void**>(p), std::max<size_t>(bytes, 16), h->stream);
}
inline void dev_free(dmlmc_hier* h, void* p) { if (p) cudaFreeAsync(p, h->stream); }

template <typename T> int upload(dmlmc_hier* h, const T* host, size_t count, T** out) {
  T* p = nullptr;
  CU(dev_alloc(h, &p, std::max<size_t>(count, 1) * sizeof(T)));
  h->owned.push_back(p);
  if (count) CU(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = p;
  return 0;
}
// complex128 host array -> device copies in both precisions
int upload_cx(dmlmc_hier* h, const double* host, size_t count, Cx<double>** d, Cx<float>** f) {
  RET(upload<Cx<double>>(h, reinterpret_cast<const Cx<double>*>(host), count, d));
  std::vector<Cx<float>> tmp(count);
  for (size_t i = 0; i < count; ++i) tmp[i] = cx<float>((float)host[2 * i], (float)host[2 * i + 1]);
  RET(upload<Cx<float>>(h, tmp.data(), count, f));
  return 0;
}

inline unsigned nblocks(size_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

#define LAUNCH_CHECK(h)                                                                       \
  do { (h)->launches++; cudaError_t e_ = cudaGetLastError();                                  \
       if (e_ != cudaSuccess) return fail((int)e_, std::string("kernel launch: ") + cudaGetErrorString(e_)); } while (0)

// ---- work-space bump allocator ------------------------------------------------------------
struct WsMark { size_t off; };
inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
int ws_alloc(dmlmc_hier* h, size_t bytes, void** out) {
  size_t off = align_up(h->ws_off);
  if (h->ws == nullptr || off + bytes > h->ws_bytes)
    return fail(-2, "dmlmc: work space too small (call dmlmc_workspace_bytes / dmlmc_set_workspace)");
  *out = h->ws + off;
  h->ws_off = off + bytes;
  return 0;
}
// restores the bump allocator on every exit path of a function (error returns included)
struct WsScope {
  dmlmc_hier* h; size_t mark;
  explicit WsScope(dmlmc_hier* h_) : h(h_), mark(h_->ws_off) {}
  ~WsScope() { h->ws_off = mark; }
  WsScope(const WsScope&) = delete; WsScope& operator=(const WsScope&) = delete;
};
template <typename T> int ws_get(dmlmc_hier* h, size_t count, T** out) {
  void* p; RET(ws_alloc(h, count * sizeof(T), &p)); *out = reinterpret_cast<T*>(p); return 0;
}

// ---- operator dispatch ----------------------------------------------------------------------
template <typename T, int NC, int MODE, bool HIN = false, bool HOUT = false>
int launch_op_nc(dmlmc_hier* h, int level, const void* X, const void* B, void* Y, Cx<double> w, Cx<double> c, int k) {
  Level& L = h->lv[level];
  LevelT<T>& D = Sel<T>::get(L);
  const int kp = k / NC;
  typedef Pack<T, NC> P;
  const Cx<T> wt = cx<T>((T)w.re, (T)w.im), ct = cx<T>((T)c.re, (T)c.im);
  if (L.kind == 0) {
    StencilDev<T> op; op.LX = L.LX; op.LT = L.LT; op.Ut = D.Ut; op.Ux = D.Ux; op.diag = D.diag;
    int bx = 1; while (bx < 32 && bx < kp) bx *= 2;
    int by = std::max(1, std::min(h->stencil_by, L.LT)), bz = std::max(1, std::min(h->stencil_bz, L.LX));
    while (bx * by * bz > 512) { if (bz > 1) bz /= 2; else by /= 2; }
    while (bx * by * bz < 256 && by < L.LT) by *= 2;
    dim3 blk(bx, by, bz), grd((kp + bx - 1) / bx, (L.LT + by - 1) / by, (L.LX + bz - 1) / bz);
    if constexpr (!HIN && !HOUT && (MODE == M_AX || MODE == M_RES)) {
      if (h->prefetch_slices == 8) { stencil_kernel<T, NC, MODE, 3, false, false, 8><<<grd, blk, 0, h->stream>>>(op, X, (const P*)B, Y, wt, ct, kp); LAUNCH_CHECK(h); return 0; }
      if (h->prefetch_slices == 16) { stencil_kernel<T, NC, MODE, 3, false, false, 16><<<grd, blk, 0, h->stream>>>(op, X, (const P*)B, Y, wt, ct, kp); LAUNCH_CHECK(h); return 0; }
    }
    if (h->stencil_minb == 3) stencil_kernel<T, NC, MODE, 3, HIN, HOUT><<<grd, blk, 0, h->stream>>>(op, X, (const P*)B, Y, wt, ct, kp);
    else                      stencil_kernel<T, NC, MODE, 2, HIN, HOUT><<<grd, blk, 0, h->stream>>>(op, X, (const P*)B, Y, wt, ct, kp);
    LAUNCH_CHECK(h);
    return 0;
  }
  if constexpr (HIN || HOUT) {
    // BF16-stored vectors on a coarse level: the packed-FP32 BSR kernel only
    if constexpr (std::is_same<T, float>::value && NC == 2) {
      if (L.kind == 1 && L.bs >= 2) {
        const int PPT = (L.bs <= 4) ? 2 : 1;
        const int need = (kp + PPT - 1) / PPT;
        int tpr = 1;
        if (need >= 32) tpr = std::min(128, ((need + 31) / 32) * 32); else while (tpr < need) tpr *= 2;
        const size_t per_row = (size_t)L.bpr * L.bs * L.bs * sizeof(float4) + (size_t)L.bpr * sizeof(int);
        int RB = std::max(1, (HIN ? h->bsr_threads : 128) / tpr);
        RB = (int)std::max<size_t>(1, std::min<size_t>(RB, 40960 / per_row));
        const size_t smem = RB * per_row + 16;
        if (smem <= 48 * 1024) {
          dim3 blk(tpr, RB), grd((L.nb + RB - 1) / RB, (kp + tpr * PPT - 1) / (tpr * PPT));
#define BSR2H(BS_, PPT_) do { if constexpr (HIN) bsr_f32x2_soa_kernel<BS_, PPT_, MODE, HIN, HOUT><<<grd, blk, smem, h->stream>>>(L.nb, L.bpr, L.bsr_col, L.bsr_vals4, \
              X, (const P*)B, Y, wt, ct, kp); else bsr_f32x2_kernel<BS_, PPT_, MODE, HIN, HOUT><<<grd, blk, smem, h->stream>>>(L.nb, L.bpr, L.bsr_col, L.bsr_vals4, \
              X, (const P*)B, Y, wt, ct, kp); } while (0)
          if (L.bs == 2) BSR2H(2, 2); else if (L.bs == 4) BSR2H(4, 2); else BSR2H(8, 1);
#undef BSR2H
          LAUNCH_CHECK(h);
          return 0;
        }
      }
    }
    return fail(-1, "dmlmc: BF16 vector storage is not implemented for this operator format");
  } else {
  if (L.kind == 1) {
    if constexpr (std::is_same<T, float>::value && NC == 2) {
      // Blackwell FFMA2 + shared-memory path for the complex64 V-cycle
      const int PPT = (L.bs <= 4) ? 2 : 1;
      const int need = (kp + PPT - 1) / PPT;
      int tpr = 1;
      if (need >= 32) tpr = std::min(128, ((need + 31) / 32) * 32); else while (tpr < need) tpr *= 2;
      const size_t per_row = (size_t)L.bpr * L.bs * L.bs * sizeof(float4) + (size_t)L.bpr * sizeof(int);
      int RB = std::max(1, 128 / tpr);
      RB = (int)std::max<size_t>(1, std::min<size_t>(RB, 40960 / per_row));
      const size_t smem = RB * per_row + 16;
      if (smem <= 48 * 1024 && L.bs >= 2) {
        dim3 blk(tpr, RB), grd((L.nb + RB - 1) / RB, (kp + tpr * PPT - 1) / (tpr * PPT));
#define BSR2(BS_, PPT_) bsr_f32x2_kernel<BS_, PPT_, MODE><<<grd, blk, smem, h->stream>>>(L.nb, L.bpr, L.bsr_col, L.bsr_vals4, \
            (const P*)X, (const P*)B, (P*)Y, wt, ct, kp)
        if (L.bs == 2) BSR2(2, 2); else if (L.bs == 4) BSR2(4, 2); else BSR2(8, 1);
#undef BSR2
        LAUNCH_CHECK(h);
        return 0;
      }
    }
    BsrDev<T> op; op.nb = L.nb; op.bpr = L.bpr; op.col = L.bsr_col; op.vals = D.bsr_vals;
    const size_t total = (size_t)L.nb * kp;
    const unsigned g = nblocks(total, 128);
    switch (L.bs) {
      case 1: bsr_kernel<T, NC, 1, MODE><<<g, 128, 0, h->stream>>>(op, (const P*)X, (const P*)B, (P*)Y, wt, ct, kp); break;
      case 2: bsr_kernel<T, NC, 2, MODE><<<g, 128, 0, h->stream>>>(op, (const P*)X, (const P*)B, (P*)Y, wt, ct, kp); break;
      case 4: bsr_kernel<T, NC, 4, MODE><<<g, 128, 0, h->stream>>>(op, (const P*)X, (const P*)B, (P*)Y, wt, ct, kp); break;
      case 8: bsr_kernel<T, NC, 8, MODE><<<g, 128, 0, h->stream>>>(op, (const P*)X, (const P*)B, (P*)Y, wt, ct, kp); break;
      default: return fail(-1, "dmlmc: unsupported BSR block size");
    }
  } else {
    return fail(-1, "dmlmc: operator of this level not set");
  }
  LAUNCH_CHECK(h);
  return 0;
  }
}
template <typename T> constexpr int max_nc() { return sizeof(T) == 4 ? 2 : 1; }

// X, B, Y: compact [n_level][k]
template <typename T, int MODE>
int launch_op(dmlmc_hier* h, int level, const void* X, const void* B, void* Y, Cx<double> w, Cx<double> c, int k) {
  if (max_nc<T>() == 2 && (k % 2) == 0) return launch_op_nc<T, max_nc<T>(), MODE>(h, level, X, B, Y, w, c, k);
  return launch_op_nc<T, 1, MODE>(h, level, X, B, Y, w, c, k);
}
const Cx<double> ZERO = {0.0, 0.0};

// transfers on k columns; rows of the fine / coarse arrays are ldf / ldc complex elements apart
template <typename T, int NC>
int launch_restrict_nc(dmlmc_hier* h, int level, const void* Xf, void* Xc, int k, int ldf, int ldc) {
  Level& L = h->lv[level];
  TransferDev<T> tr; tr.n_f = L.n; tr.n_c = L.n_c; tr.aggr = L.aggr; tr.dofi = L.dofi; tr.h = L.dofi / 2; tr.nvec = L.nvec;
  tr.pv = Sel<T>::get(L).pv; tr.rows = L.tr_rows; tr.cblk = L.tr_cblk; tr.m = L.tr_m;
  const int kp = k / NC; typedef Pack<T, NC> P;
  const size_t total = (size_t)(L.n_c / L.nvec) * kp;
  const unsigned g = nblocks(total, 128);
  switch (L.nvec) {
    case 1: restrict_kernel<T, NC, 1><<<g, 128, 0, h->stream>>>(tr, (const P*)Xf, (P*)Xc, kp, ldf / NC, ldc / NC); break;
    case 2: restrict_kernel<T, NC, 2><<<g, 128, 0, h->stream>>>(tr, (const P*)Xf, (P*)Xc, kp, ldf / NC, ldc / NC); break;
    case 4: restrict_kernel<T, NC, 4><<<g, 128, 0, h->stream>>>(tr, (const P*)Xf, (P*)Xc, kp, ldf / NC, ldc / NC); break;
    case 8: restrict_kernel<T, NC, 8><<<g, 128, 0, h->stream>>>(tr, (const P*)Xf, (P*)Xc, kp, ldf / NC, ldc / NC); break;
    default: return fail(-1, "dmlmc: unsupported number of test vectors");
  }
  LAUNCH_CHECK(h);
  return 0;
}
inline bool pack2_ok(const void* a, const void* b, int k, int ldf, int ldc) {
  return (k % 2) == 0 && (ldf % 2) == 0 && (ldc % 2) == 0 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0;
}
template <typename T> int launch_restrict(dmlmc_hier* h, int level, const void* Xf, void* Xc, int k, int ldf, int ldc) {
  if (!h->lv[level].has_transfer) return fail(-1, "dmlmc: transfer operator of this level not set");
  if (max_nc<T>() == 2 && pack2_ok(Xf, Xc, k, ldf, ldc)) return launch_restrict_nc<T, max_nc<T>()>(h, level, Xf, Xc, k, ldf, ldc);
  return launch_restrict_nc<T, 1>(h, level, Xf, Xc, k, ldf, ldc);
}
template <typename T> int launch_restrict(dmlmc_hier* h, int level, const void* Xf, void* Xc, int k) {
  return launch_restrict<T>(h, level, Xf, Xc, k, k, k);
}
template <typename T, int NC>
int launch_prolong_nc(dmlmc_hier* h, int level, const void* Xc, void* Xf, int k, int ldf, int ldc, int add) {
  Level& L = h->lv[level];
  TransferDev<T> tr; tr.n_f = L.n; tr.n_c = L.n_c; tr.aggr = L.aggr; tr.dofi = L.dofi; tr.h = L.dofi / 2; tr.nvec = L.nvec;
  tr.pv = Sel<T>::get(L).pv; tr.rows = L.tr_rows; tr.cblk = L.tr_cblk; tr.m = L.tr_m;
  const int kp = k / NC; typedef Pack<T, NC> P;
  const size_t total = (size_t)L.n * kp;
  const unsigned g = nblocks(total, 256);
  switch (L.nvec) {
    case 1: prolong_add_kernel<T, NC, 1><<<g, 256, 0, h->stream>>>(tr, (const P*)Xc, (P*)Xf, kp, ldf / NC, ldc / NC, add); break;
    case 2: prolong_add_kernel<T, NC, 2><<<g, 256, 0, h->stream>>>(tr, (const P*)Xc, (P*)Xf, kp, ldf / NC, ldc / NC, add); break;
    case 4: prolong_add_kernel<T, NC, 4><<<g, 256, 0, h->stream>>>(tr, (const P*)Xc, (P*)Xf, kp, ldf / NC, ldc / NC, add); break;
    case 8: prolong_add_kernel<T, NC, 8><<<g, 256, 0, h->stream>>>(tr, (const P*)Xc, (P*)Xf, kp, ldf / NC, ldc / NC, add); break;
    default: return fail(-1, "dmlmc: unsupported number of test vectors");
  }
  LAUNCH_CHECK(h);
  return 0;
}
template <typename T> int launch_prolong(dmlmc_hier* h, int level, const void* Xc, void* Xf, int k, int ldf, int ldc, int add = 1) {
  if (!h->lv[level].has_transfer) return fail(-1, "dmlmc: transfer operator of this level not set");
  if (max_nc<T>() == 2 && pack2_ok(Xf, Xc, k, ldf, ldc)) return launch_prolong_nc<T, max_nc<T>()>(h, level, Xc, Xf, k, ldf, ldc, add);
  return launch_prolong_nc<T, 1>(h, level, Xc, Xf, k, ldf, ldc, add);
}
template <typename T> int launch_prolong(dmlmc_hier* h, int level, const void* Xc, void* Xf, int k) {
  return launch_prolong<T>(h, level, Xc, Xf, k, k, k);
}

template <typename T> Cx<T>* minv_of(Level& L);
template <> Cx<double>* minv_of<double>(Level& L) { return L.minv_d; }
template <> Cx<float>*  minv_of<float>(Level& L)  { return L.minv_f; }

// X = A_level^{-1} B on the tensor cores (complex64 in/out, BF16 operands, FP32 accumulation)
int launch_dense_umma(dmlmc_hier* h, int level, const Cx<float>* B, Cx<float>* X, int k) {
  Level& L = h->lv[level];
  const int n = L.n;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  __nv_bfloat16* Bt;
  RET(ws_get<__nv_bfloat16>(h, (size_t)k * 2 * n, &Bt));
  dim3 pblk(32, 8), pgrd((n + 31) / 32, (k + 31) / 32);
  umma_pack_rhs_kernel<float><<<pgrd, pblk, 0, h->stream>>>(B, k, n, k, Bt); LAUNCH_CHECK(h);
  if (L.tmB_ptr != Bt || L.tmB_k != k) {
    if (make_tmap_bf16(&L.tmB, Bt, (uint64_t)k, (uint64_t)2 * n, UM_BN) != 0) return fail(-4, "dmlmc: cuTensorMapEncodeTiled failed (rhs)");
    L.tmB_ptr = Bt; L.tmB_k = k;
  }
  if (!h->umma_attr_set) {
    CU(cudaFuncSetAttribute(dense_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM));
    h->umma_attr_set = true;
  }
  dim3 grd((k + UM_BN - 1) / UM_BN, (2 * n + UM_BM - 1) / UM_BM);
  dense_umma_kernel<<<grd, UM_THREADS, UM_SMEM, h->stream>>>(L.tmA, L.tmB, reinterpret_cast<float*>(X), k, 2 * n, k, 2 * n);
  LAUNCH_CHECK(h);
  h->ws_off = mark;
  return 0;
}

// the same with split-BF16 operands (hi/lo of the matrix and of the right-hand side, one GEMM over 6n): FP32-class accuracy
int launch_dense_umma_split(dmlmc_hier* h, int level, const Cx<float>* B, Cx<float>* X, int k) {
  Level& L = h->lv[level];
  const int n = L.n;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  __nv_bfloat16* Bt;
  RET(ws_get<__nv_bfloat16>(h, (size_t)k * 6 * n, &Bt));
  dim3 pblk(32, 8), pgrd((n + 31) / 32, (k + 31) / 32);
  umma_pack_rhs_split_kernel<float><<<pgrd, pblk, 0, h->stream>>>(B, k, n, k, Bt); LAUNCH_CHECK(h);
  if (L.tmB3_ptr != Bt || L.tmB3_k != k) {
    if (make_tmap_bf16(&L.tmB3, Bt, (uint64_t)k, (uint64_t)6 * n, UM_BN) != 0) return fail(-4, "dmlmc: cuTensorMapEncodeTiled failed (split rhs)");
    L.tmB3_ptr = Bt; L.tmB3_k = k;
  }
  if (!h->umma_attr_set) {
    CU(cudaFuncSetAttribute(dense_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM));
    h->umma_attr_set = true;
  }
  dim3 grd((k + UM_BN - 1) / UM_BN, (2 * n + UM_BM - 1) / UM_BM);
  dense_umma_kernel<<<grd, UM_THREADS, UM_SMEM, h->stream>>>(L.tmA3, L.tmB3, reinterpret_cast<float*>(X), k, 2 * n, k, 6 * n);
  LAUNCH_CHECK(h);
  h->ws_off = mark;
  return 0;
}

// X = A_level^{-1} B with the level's dense inverse; B, X compact [n][k]
template <typename T> int launch_dense(dmlmc_hier* h, int level, const void* B, void* X, int k, bool prefer_exact = false) {
  Level& L = h->lv[level];
  if (!L.has_dense) return fail(-1, "dmlmc: dense inverse of this level not set");
  const int n = L.n;
  if constexpr (std::is_same<T, float>::value) {
    if (L.has_umma && (L.minv4 == nullptr || (!prefer_exact && n >= h->dense_tensor_min_n)))
      return launch_dense_umma(h, level, (const Cx<float>*)B, (Cx<float>*)X, k);
    if (prefer_exact && L.has_umma3 && h->dense_split_bf16)
      return launch_dense_umma_split(h, level, (const Cx<float>*)B, (Cx<float>*)X, k);
  }
  if (minv_of<T>(L) == nullptr) return fail(-1, "dmlmc: this level's dense inverse exists only as the tensor-core (complex64) operand");
  dim3 blk(32, 8);
  if (max_nc<T>() == 2 && (k % 2) == 0) {
    constexpr int NC = max_nc<T>(); const int kp = k / NC;
    dim3 grd((kp + 31) / 32, (n + 31) / 32);
    if constexpr (std::is_same<T, float>::value)
      dense_f32x2_kernel<<<grd, blk, 0, h->stream>>>(L.minv4, n, (const Pack<float, 2>*)B, (Pack<float, 2>*)X, kp);
    else
      dense_kernel<T, NC><<<grd, blk, 0, h->stream>>>(minv_of<T>(L), n, (const Pack<T, NC>*)B, (Pack<T, NC>*)X, kp);
  } else {
    dim3 grd((k + 31) / 32, (n + 31) / 32);
    dense_kernel<T, 1><<<grd, blk, 0, h->stream>>>(minv_of<T>(L), n, (const Pack<T, 1>*)B, (Pack<T, 1>*)X, k);
  }
  LAUNCH_CHECK(h);
  return 0;
}

// build the tensor-core operand of a level's dense inverse from the complex128 matrix on the device
int build_umma_operand(dmlmc_hier* h, int level, const Cx<double>* minv_dev) {
  Level& L = h->lv[level];
  const size_t n = L.n;
  if (n % 8 != 0) return 0;                       // TMA needs 16-byte row strides; such tiny levels use the SIMT kernel
  __nv_bfloat16* mt = nullptr;
  CU(dev_alloc(h, &mt, 4 * n * n * sizeof(__nv_bfloat16)));
  h->owned.push_back(mt);
  umma_expand_matrix_kernel<<<nblocks(n * n, 256), 256, 0, h->stream>>>(minv_dev, (int)n, mt); LAUNCH_CHECK(h);
  CU(cudaStreamSynchronize(h->stream));
  if (make_tmap_bf16(&L.tmA, mt, 2 * n, 2 * n, UM_BM) != 0) return fail(-4, "dmlmc: cuTensorMapEncodeTiled failed (matrix)");
  L.minv_bf16 = mt; L.has_umma = true; L.tmB_ptr = nullptr;
  return 0;
}

int build_umma_split_operand(dmlmc_hier* h, int level, const Cx<double>* minv_dev) {
  Level& L = h->lv[level];
  const size_t n = L.n;
  if (n % 8 != 0) return 0;
  __nv_bfloat16* mt = nullptr;
  CU(dev_alloc(h, &mt, 12 * n * n * sizeof(__nv_bfloat16)));
  h->owned.push_back(mt);
  umma_expand_matrix_split_kernel<<<nblocks(n * n, 256), 256, 0, h->stream>>>(minv_dev, (int)n, mt); LAUNCH_CHECK(h);
  CU(cudaStreamSynchronize(h->stream));
  if (make_tmap_bf16(&L.tmA3, mt, 2 * n, 6 * n, UM_BM) != 0) return fail(-4, "dmlmc: cuTensorMapEncodeTiled failed (split matrix)");
  L.minv_bf16x3 = mt; L.has_umma3 = true; L.tmB3_ptr = nullptr;
  return 0;
}

// Out[:, 0:w] = (Tout) In[:, 0:w] between batches with leading dimensions ld_in / ld_out
template <typename Tin, typename Tout>
int cvt_cols(dmlmc_hier* h, const Cx<Tin>* in, size_t ld_in, Cx<Tout>* out, size_t ld_out, int n, int w) {
  if (std::is_same<Tin, Tout>::value && (const void*)in == (const void*)out) return 0;
  cvt_cols_kernel<Tin, Tout><<<nblocks((size_t)n * w, 256), 256, 0, h->stream>>>(in, ld_in, out, ld_out, n, w);
  LAUNCH_CHECK(h);
  return 0;
}

// thread-block / grid shape of the generic stencil kernel for kp packs per row
void stencil_dims(dmlmc_hier* h, const Level& L, int kp, dim3& blk, dim3& grd) {
  int bx = 1; while (bx < 32 && bx < kp) bx *= 2;
  int by = std::max(1, std::min(h->stencil_by, L.LT)), bz = std::max(1, std::min(h->stencil_bz, L.LX));
  while (bx * by * bz > 512) { if (bz > 1) bz /= 2; else by /= 2; }
  while (bx * by * bz < 256 && by < L.LT) by *= 2;
  blk = dim3(bx, by, bz); grd = dim3((kp + bx - 1) / bx, (L.LT + by - 1) / by, (L.LX + bz - 1) / bz);
}

// R16 = B - A X with X, B complex64 and the result stored as BF16: the post-smoother's input on the level-0 stencil, so that
// ALL of its factors but the last run in the packed-FP32 BF16 -> BF16 kernel (option fuse_res)
int launch_residual_half(dmlmc_hier* h, int level, const void* X, const void* B, void* R16, int k) {
  Level& L = h->lv[level];
  StencilDev<float> op; op.LX = L.LX; op.LT = L.LT; op.Ut = L.f.Ut; op.Ux = L.f.Ux; op.diag = L.f.diag;
  const int kp = k / 2;
  dim3 blk, grd; stencil_dims(h, L, kp, blk, grd);
  stencil_kernel<float, 2, M_RES, 3, false, true><<<grd, blk, 0, h->stream>>>(op, X, (const Pack<float, 2>*)B, R16, cx<float>(0.f, 0.f),
                                                                               cx<float>(0.f, 0.f), kp);
  LAUNCH_CHECK(h);
  return 0;
}

// ---- even-odd post-smoother of a stencil level (see wilson_hop_eo_kernel) ---------------------------------------------
// Z[n][k] (complex128) = Xc + S_eo(B - A Xc):  Xc, B complex64 full-lattice vectors, t0 / t1 the level's two work vectors
// (n k complex64 each = eight BF16 half-lattice buffers).
template <bool HAS2, bool HOUT, bool ZOUT>
int launch_hop_eo(dmlmc_hier* h, const Level& L, int p, const uint2* Inq, const uint2* In2, uint2* Outp, Cx<double> a, Cx<double> b,
                  int kp, const void* Xc, void* Z) {
  if constexpr (HOUT && !ZOUT) {
    if (h->hop_tma && (kp % HT_CW) == 0 && (L.LX % HT_X) == 0 && ((L.LT / 2) % HT_TH) == 0) {
      if (!h->hop_tma_attr_set) {
        CU(cudaFuncSetAttribute(wilson_hop_eo_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_SMEM));
        CU(cudaFuncSetAttribute(wilson_hop_eo_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_SMEM));
        h->hop_tma_attr_set = true;
      }
      const int ntiles = (L.LX / HT_X) * ((L.LT / 2) / HT_TH) * (kp / HT_CW);
      const int grid = std::min(ntiles, 2 * h->n_sms);
      wilson_hop_eo_tma_kernel<HAS2><<<grid, HT_THREADS, HT_SMEM, h->stream>>>(L.LX, L.LT, p, L.links4, Inq, In2, Outp, (float)a.re, (float)a.im,
                                                                             (float)b.re, (float)b.im, (uint32_t)kp, ntiles);
      LAUNCH_CHECK(h);
      return 0;
    }
  }
  const bool two = h->eo_packs == 2 && (kp % 2) == 0;
  const int kt = two ? kp / 2 : kp;                      // threads along the columns
  int bx = 1; while (bx < 32 && bx < kt) bx *= 2;
  int by = std::max(1, std::min(h->eo_by, L.LT / 2)), bz = std::max(1, std::min(h->eo_bz, L.LX));
  while (bx * by * bz > 256) { if (bz > 1) bz /= 2; else by /= 2; }
  dim3 blk(bx, by, bz), grd((kt + bx - 1) / bx, (L.LT / 2 + by - 1) / by, (L.LX + bz - 1) / bz);
  if (two)
    wilson_hop_eo_kernel<HAS2, HOUT, ZOUT, 2><<<grd, blk, 0, h->stream>>>(L.LX, L.LT, p, L.links4, Inq, In2, Outp, (float)a.re, (float)a.im,
                                                                          (float)b.re, (float)b.im, (uint32_t)kp,
                                                                          (const Pack<float, 2>*)Xc, (double2*)Z, h->eo_zhalf);
  else
    wilson_hop_eo_kernel<HAS2, HOUT, ZOUT, 1><<<grd, blk, 0, h->stream>>>(L.LX, L.LT, p, L.links4, Inq, In2, Outp, (float)a.re, (float)a.im,
                                                                          (float)b.re, (float)b.im, (uint32_t)kp,
                                                                          (const Pack<float, 2>*)Xc, (double2*)Z, h->eo_zhalf);
  LAUNCH_CHECK(h);
  return 0;
}
bool smoother_eo_ok(dmlmc_hier* h, int level, int k) {
  Level& L = h->lv[level];
  return h->smoother_eo && h->fuse_io && h->smoother_half && L.has_eo && L.smoother16 && L.kind == 0 && (k % 2) == 0 &&
         !L.nu_eo.empty() && (L.LT % 2) == 0 && (L.LX % 2) == 0 && L.d.diag.im == 0.0 && L.links4 != nullptr;
}
int smooth_eo(dmlmc_hier* h, int level, const void* Xc, const void* B, void* t0, void* t1, Cx<double>* Z, int k) {
  Level& L = h->lv[level];
  const int kp = k / 2, m = (int)L.nu_eo.size();
  const size_t H = (size_t)(L.n / 2) * kp;                // uint2 elements of one half-lattice buffer
  uint2* T0 = reinterpret_cast<uint2*>(t0); uint2* T1 = reinterpret_cast<uint2*>(t1);
  uint2 *Re = T0, *Ro = T0 + H, *W = T0 + 2 * H, *Y[2] = {T1, T1 + H};
  const double c = L.d.diag.re;
  const Cx<double> ONE = {1.0, 0.0}, NIC = {-1.0 / c, 0.0}, IC = {1.0 / c, 0.0};
  // r = B - A Xc, BF16, checkerboard layout [even | odd]
  {
    StencilDev<float> op; op.LX = L.LX; op.LT = L.LT; op.Ut = L.f.Ut; op.Ux = L.f.Ux; op.diag = L.f.diag;
    dim3 blk, grd; stencil_dims(h, L, kp, blk, grd);
    stencil_kernel<float, 2, M_RES, 3, false, true, 0, false, true><<<grd, blk, 0, h->stream>>>(op, Xc, (const Pack<float, 2>*)B, Re,
                                                                                                 cx<float>(0.f, 0.f), cx<float>(0.f, 0.f), kp);
    LAUNCH_CHECK(h);
  }
  // r^_e = r_e - H_eo r_o / c
  RET((launch_hop_eo<true, true, false>(h, L, 0, Ro, Re, Y[0], ONE, NIC, kp, nullptr, nullptr)));
  int cur = 0;
  for (int i = 0; i < m; ++i) {
    const Cx<double> nu = L.nu_eo[i];
    Cx<double> a = {1.0 - nu.re * c, -nu.im * c}, b = {nu.re / c, nu.im / c};
    RET((launch_hop_eo<false, true, false>(h, L, 1, Y[cur], nullptr, W, ONE, ONE, kp, nullptr, nullptr)));     // w_o = H_oe y_e
    if (i < m - 1) {
      RET((launch_hop_eo<true, true, false>(h, L, 0, W, Y[cur], Y[cur ^ 1], a, b, kp, nullptr, nullptr)));      // y_e' = a y_e + b H_eo w_o
    } else {
      // last factor carries p0: x_e (BF16, for x_o) and Z_e = Xc_e + x_e
      const Cx<double> p0 = L.p0_eo;
      const Cx<double> ap = {a.re * p0.re - a.im * p0.im, a.re * p0.im + a.im * p0.re};
      const Cx<double> bp = {b.re * p0.re - b.im * p0.im, b.re * p0.im + b.im * p0.re};
      RET((launch_hop_eo<true, true, true>(h, L, 0, W, Y[cur], Y[cur ^ 1], ap, bp, kp, Xc, Z)));
    }
    cur ^= 1;
  }
  // Z_o = Xc_o + (r_o - H_oe x_e) / c   (not needed when only the even part of the cycle's output is used)
  if (!h->eo_zhalf) RET((launch_hop_eo<true, false, true>(h, L, 1, Y[cur], Ro, nullptr, IC, NIC, kp, Xc, Z)));
  return 0;
}

// the complex64 BF16-storage path of smooth_apply on the level-0 stencil can write its (accumulated) result as complex128
bool smoother_dout_ok(dmlmc_hier* h, int level, int k) {
  Level& L = h->lv[level];
  return h->fuse_io && h->smoother_half && L.has_smoother && L.smoother16 && L.kind == 0 && (k % 2) == 0 && L.nu.size() >= 2;
}

// ---- smoother: E (+)= p(A) R with p in product form, p(A) = p0 prod_i (I - nu_i A) ---------------
// One fused kernel per factor (operator + update: read x, write x'), no reductions, no host
// synchronisation.  R, E, t0, t1: compact [n_level][k]; R may alias t0; E must not alias R, t0, t1.
template <typename T>
int smooth_apply(dmlmc_hier* h, int level, const void* R, void* E, bool acc, void* t0, void* t1, int k, Cx<double>* Ed = nullptr,
                 bool r_half = false) {   // r_half (only with smoother_dout_ok): R is already BF16-stored (launch_residual_half)
  // Ed (only with smoother_dout_ok): the accumulated result E + p(A) R goes to Ed[n][k] as complex128, E is left unchanged
  Level& L = h->lv[level];
  if (!L.has_smoother) return fail(-1, "dmlmc: smoother of this level not set");
  const int m = (int)L.nu.size();
  if (m == 0) {
    const size_t cnt = (size_t)L.n * k;
    const Cx<T> c = cx<T>((T)L.p0.re, (T)L.p0.im);
    if (acc) scale_kernel<T, 1><<<nblocks(cnt, 256), 256, 0, h->stream>>>(c, (const Cx<T>*)R, (Cx<T>*)E, cnt);
    else     scale_kernel<T, 0><<<nblocks(cnt, 256), 256, 0, h->stream>>>(c, (const Cx<T>*)R, (Cx<T>*)E, cnt);
    LAUNCH_CHECK(h);
    return 0;
  }
  void* pp[2] = {t0, t1};
  if (R == t0) { pp[0] = t1; pp[1] = t0; }
  const void* in = R;
  const Cx<double> ONE = {1.0, 0.0};
  if constexpr (std::is_same<T, float>::value) {
    // complex64 cycle on the level-0 stencil: the intermediate vectors of the product are stored as BF16
    // (FP32 arithmetic), which halves the bytes every factor kernel moves and keeps them L2-resident.
    // (HS is a harmless power-of-two pre-scale kept from the FP16 variant of this path.)
    const bool bsr_half = L.kind == 1 && L.bs >= 2 &&
                          ((size_t)L.bpr * L.bs * L.bs * sizeof(float4) + (size_t)L.bpr * sizeof(int) + 16) <= 48 * 1024;
    if (h->smoother_half && L.smoother16 && (L.kind == 0 || bsr_half) && (k % 2) == 0 && m >= 2) {
      const double HS = r_half ? 1.0 : 64.0;
      const Cx<double> cfirst = {HS, 0.0}, clast = {L.p0.re / HS, L.p0.im / HS};
      for (int i = 0; i < m; ++i) {
        if (i == 0 && !r_half) {
          RET((launch_op_nc<float, 2, M_STEP, false, true>(h, level, in, nullptr, pp[0], L.nu[i], cfirst, k)));
          in = pp[0];
        } else if (i == m - 1) {
          if (acc && Ed != nullptr) {
            CHECK(L.kind == 0, "complex128 smoother output: level-0 stencil only");
            StencilDev<float> op; op.LX = L.LX; op.LT = L.LT; op.Ut = L.f.Ut; op.Ux = L.f.Ux; op.diag = L.f.diag;
            const int kp = k / 2;
            dim3 blk, grd; stencil_dims(h, L, kp, blk, grd);
            stencil_kernel<float, 2, M_STEP_ACC, 3, true, false, 0, true><<<grd, blk, 0, h->stream>>>(
                op, in, (const Pack<float, 2>*)Ed, E, cx<float>((float)L.nu[i].re, (float)L.nu[i].im),
                cx<float>((float)clast.re, (float)clast.im), kp);
            LAUNCH_CHECK(h);
          }
          else if (acc) RET((launch_op_nc<float, 2, M_STEP_ACC, true, false>(h, level, in, nullptr, E, L.nu[i], clast, k)));
          else     RET((launch_op_nc<float, 2, M_STEP, true, false>(h, level, in, nullptr, E, L.nu[i], clast, k)));
        } else {
          void* out = (in == pp[i & 1]) ? pp[(i + 1) & 1] : pp[i & 1];
          if (L.kind == 0 && h->stencil_fast && L.d.diag.im == 0.0 && (size_t)L.n * (k / 2) < (1ull << 32)) {
            const int kp = k / 2;
            if (h->stencil_smem && (kp % 2) == 0) {
              dim3 grd((kp + 15) / 16, (L.LT + 7) / 8, (L.LX + 7) / 8);
              stencil_step_bf16_smem_kernel<8, 8><<<grd, 512, 0, h->stream>>>(L.LX, L.LT, L.links4, (float)L.d.diag.re, (const uint2*)in,
                                                                           (uint2*)out, (float)L.nu[i].re, (float)L.nu[i].im, (uint32_t)kp);
              LAUNCH_CHECK(h);
              in = out;
              continue;
            }
            const long long rowb = (long long)kp * 8;
            if (h->stencil_t2 && (L.LT % 2) == 0) {
              int bx2 = 1; while (bx2 < 32 && bx2 < kp) bx2 *= 2;
              int by2 = std::max(1, std::min(h->stencil_t2_by, L.LT / 2)), bz2 = std::max(1, std::min(h->stencil_t2_bz, L.LX));
              while (bx2 * by2 * bz2 > 256) { if (bz2 > 1) bz2 /= 2; else by2 /= 2; }
              dim3 blk2(bx2, by2, bz2), grd2((kp + bx2 - 1) / bx2, (L.LT / 2 + by2 - 1) / by2, (L.LX + bz2 - 1) / bz2);
              stencil_step_bf16_t2_kernel<<<grd2, blk2, 0, h->stream>>>(L.LX, L.LT, L.links4, (float)L.d.diag.re, (const uint2*)in,
                                                                      (uint2*)out, (float)L.nu[i].re, (float)L.nu[i].im, (uint32_t)kp,
                                                                      rowb, rowb * L.LT, rowb * L.LT * L.LX);
              LAUNCH_CHECK(h);
              in = out;
              continue;
            }
            int bx = 1; while (bx < 32 && bx < kp) bx *= 2;
            int by = std::max(1, std::min(h->stencil_by, L.LT)), bz = std::max(1, std::min(h->stencil_bz, L.LX));
            while (bx * by * bz > 512) { if (bz > 1) bz /= 2; else by /= 2; }
            while (bx * by * bz < 256 && by < L.LT) by *= 2;
            dim3 blk(bx, by, bz), grd((kp + bx - 1) / bx, (L.LT + by - 1) / by, (L.LX + bz - 1) / bz);
            if (h->stencil_minb == 3)
              stencil_step_bf16_kernel<3><<<grd, blk, 0, h->stream>>>(L.LX, L.LT, L.links4, (float)L.d.diag.re, (const uint2*)in,
                                                                     (uint2*)out, (float)L.nu[i].re, (float)L.nu[i].im, (uint32_t)kp,
                                                                     rowb, rowb * L.LT, rowb * L.LT * L.LX);
            else
              stencil_step_bf16_kernel<2><<<grd, blk, 0, h->stream>>>(L.LX, L.LT, L.links4, (float)L.d.diag.re, (const uint2*)in,
                                                                     (uint2*)out, (float)L.nu[i].re, (float)L.nu[i].im, (uint32_t)kp,
                                                                     rowb, rowb * L.LT, rowb * L.LT * L.LX);
            LAUNCH_CHECK(h);
          } else {
            RET((launch_op_nc<float, 2, M_STEP, true, true>(h, level, in, nullptr, out, L.nu[i], ONE, k)));
          }
          in = out;
        }
      }
      return 0;
    }
  }
  for (int i = 0; i < m; ++i) {
    if (i == m - 1) {
      if (acc) RET((launch_op<T, M_STEP_ACC>(h, level, in, nullptr, E, L.nu[i], L.p0, k)));
      else     RET((launch_op<T, M_STEP>(h, level, in, nullptr, E, L.nu[i], L.p0, k)));
    } else {
      void* out = pp[i & 1];
      RET((launch_op<T, M_STEP>(h, level, in, nullptr, out, L.nu[i], ONE, k)));
      in = out;
    }
  }
  return 0;
}

// columns per chunk on `level`.  With option "l2_budget_mb" > 0 the chunk's working vectors (b, x and
// the smoother's ping-pong pair) are sized to stay resident in the 126 MB L2 while the level's
// ~2(d+1) kernels run over them.  Measured on B200 (profiles/r1_run4_*): the level-0 kernel is not
// DRAM-bound enough for this to pay (8.9 us per 64-column launch vs 29 us per 256-column launch), so
// the default is one chunk.
int chunk_cols(dmlmc_hier* h, int level, int k, size_t elem) {
  if (h->chunk_cols > 0) return std::min(k, std::max(2, h->chunk_cols & ~1));
  if (h->l2_budget_mb <= 0.0) return k;
  const double per_col = 4.0 * (double)h->lv[level].n * (double)elem;
  long kc = (long)(h->l2_budget_mb * 1048576.0 / per_col);
  long c = 64; while (c * 2 <= kc) c *= 2;      // 64 * 2^a: chunks of a level nest in those of the next
  return c >= k ? k : (int)c;
}

// ---- V-cycle (multigrid.py:369-447) -----------------------------------------------------------
// B, X: [n_level0][k] of scalar TIO; the cycle computes in T on compact per-level arrays stored as
// consecutive column chunks [chunk][n_l][w].  Per level: for every chunk (pre-smooth, residual,
// restrict), then the coarser level on all k columns, then for every chunk (prolong, residual,
// post-smooth).
template <typename T> struct VcBuf { Cx<T>* b = nullptr; Cx<T>* x = nullptr; Cx<T>* t0 = nullptr; Cx<T>* t1 = nullptr; int kc = 0; };

template <typename T, typename TIO>
int vcycle_level(dmlmc_hier* h, int l, int level0, int lb, std::vector<VcBuf<T>>& vb, const Cx<TIO>* Bio, Cx<TIO>* Xio, int k,
                 bool b_ready = false) {   // b_ready: vb[level0].b already holds the input in T (no conversion pass)
  Level& L = h->lv[l];
  const int n = L.n;
  VcBuf<T>& me = vb[l];
  if (l == lb) {
    for (int col0 = 0; col0 < k; col0 += me.kc) {
      const int w = std::min(me.kc, k - col0);
      RET(launch_dense<T>(h, l, me.b + (size_t)col0 * n, me.x + (size_t)col0 * n, w));
    }
    return 0;
  }
  VcBuf<T>& co = vb[l + 1];
  const int nc = h->lv[l + 1].n;
  for (int phase = 0; phase < 2; ++phase) {
    for (int col0 = 0; col0 < k; col0 += me.kc) {
      const int w = std::min(me.kc, k - col0);
      Cx<T>* bc = me.b + (size_t)col0 * n;
      Cx<T>* xc = me.x + (size_t)col0 * n;
      const int C0 = (col0 / co.kc) * co.kc, wC = std::min(co.kc, k - C0);
      const size_t coff = (size_t)C0 * nc + (col0 - C0);
      if (phase == 0) {
        if (l == level0 && !b_ready) RET((cvt_cols<TIO, T>(h, Bio + col0, (size_t)k, bc, (size_t)w, n, w)));
        if (h->pre_smooth) {
          RET(smooth_apply<T>(h, l, bc, xc, false, me.t0, me.t1, w));
          RET((launch_op<T, M_RES>(h, l, xc, bc, me.t0, ZERO, ZERO, w)));
          RET(launch_restrict<T>(h, l, me.t0, co.b + coff, w, w, wC));
        } else {
          // post-smoothing only: x = 0, the coarse level sees R b directly
          RET(launch_restrict<T>(h, l, bc, co.b + coff, w, w, wC));
          if (!h->fuse_io) CU(cudaMemsetAsync(xc, 0, (size_t)n * w * sizeof(Cx<T>), h->stream));
        }
      } else {
        // x = 0 before the coarse correction when there is no pre-smoothing: x = P x_c without reading x
        RET(launch_prolong<T>(h, l, co.x + coff, xc, w, w, wC, (h->pre_smooth || !h->fuse_io) ? 1 : 0));
        bool dout = false, rhalf = false;
        if constexpr (std::is_same<T, float>::value) {
          rhalf = h->fuse_res && smoother_dout_ok(h, l, w) && h->lv[l].nu.size() >= 3;
          if constexpr (std::is_same<TIO, double>::value) dout = (l == level0 && w == k && smoother_dout_ok(h, l, w));
        }
        bool eo = false;
        if constexpr (std::is_same<T, float>::value && std::is_same<TIO, double>::value) eo = dout && smoother_eo_ok(h, l, w);
        if (eo) {
          if constexpr (std::is_same<TIO, double>::value) RET(smooth_eo(h, l, xc, bc, me.t0, me.t1, Xio, w));
          continue;
        }
        if (rhalf) RET(launch_residual_half(h, l, xc, bc, me.t0, w));
        else       RET((launch_op<T, M_RES>(h, l, xc, bc, me.t0, ZERO, ZERO, w)));
        if (dout) {
          if constexpr (std::is_same<TIO, double>::value)
            RET(smooth_apply<T>(h, l, me.t0, xc, true, me.t0, me.t1, w, Xio, rhalf));
        } else if (rhalf) {
          RET(smooth_apply<T>(h, l, me.t0, xc, true, me.t0, me.t1, w, nullptr, true));
          if (l == level0) RET((cvt_cols<T, TIO>(h, xc, (size_t)w, Xio + col0, (size_t)k, n, w)));
        } else {
          RET(smooth_apply<T>(h, l, me.t0, xc, true, me.t0, me.t1, w));
          if (l == level0) RET((cvt_cols<T, TIO>(h, xc, (size_t)w, Xio + col0, (size_t)k, n, w)));
        }
      }
    }
    if (phase == 0) RET((vcycle_level<T, TIO>(h, l + 1, level0, lb, vb, Bio, Xio, k, b_ready)));
  }
  return 0;
}

// B32 (optional): the input already converted to complex64 [n][k] by the caller (fgmres: written by the kernel that
// produced the basis vector); used in place of the conversion pass when the cycle computes in complex64 on one chunk
template <typename T, typename TIO>
int vcycle(dmlmc_hier* h, int level0, const void* Bin, void* Xout, int k, const Cx<float>* B32 = nullptr) {
  const int nl = h->n_levels;
  CHECK(level0 >= 0 && level0 < nl, "vcycle: bad level");
  // the cycle bottoms out at the first level (from level0 down) that owns a dense inverse usable in
  // precision T (tensor-core-only inverses serve the complex64 cycle)
  auto usable = [&](const Level& L) {
    return L.has_dense && (std::is_same<T, float>::value ? (L.has_umma || L.minv4 != nullptr) : L.minv_d != nullptr);
  };
  int lb = level0;
  while (lb < nl - 1 && !usable(h->lv[lb])) ++lb;
  CHECK(usable(h->lv[lb]), "vcycle: no dense inverse at the bottom of the cycle");
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  if (lb == level0) {
    // the dense inverse IS the preconditioner of this level's own solve: use the most accurate copy
    // (a BF16 inverse costs ~20 FGMRES iterations at 1e-12 instead of 3)
    int rc;
    const bool exact = h->dense_direct_exact != 0;
    if constexpr (std::is_same<T, TIO>::value) {
      rc = launch_dense<T>(h, level0, Bin, Xout, k, exact);
    } else {
      const size_t cnt = (size_t)h->lv[level0].n * k;
      Cx<T>*bt, *xt; RET(ws_get<Cx<T>>(h, cnt, &bt)); RET(ws_get<Cx<T>>(h, cnt, &xt));
      RET((cvt_cols<TIO, T>(h, (const Cx<TIO>*)Bin, (size_t)k, bt, (size_t)k, h->lv[level0].n, k)));
      rc = launch_dense<T>(h, level0, bt, xt, k, exact);
      if (rc == 0) rc = cvt_cols<T, TIO>(h, xt, (size_t)k, (Cx<TIO>*)Xout, (size_t)k, h->lv[level0].n, k);
    }
    h->ws_off = mark;
    return rc;
  }
  std::vector<VcBuf<T>> vb(nl);
  int kc_prev = 0;
  for (int l = level0; l <= lb; ++l) {
    const size_t n = h->lv[l].n;
    vb[l].kc = std::max(kc_prev, chunk_cols(h, l, k, sizeof(Cx<T>)));
    kc_prev = vb[l].kc;
    RET(ws_get<Cx<T>>(h, n * k, &vb[l].b)); RET(ws_get<Cx<T>>(h, n * k, &vb[l].x));
    if (l < lb) { RET(ws_get<Cx<T>>(h, n * vb[l].kc, &vb[l].t0)); RET(ws_get<Cx<T>>(h, n * vb[l].kc, &vb[l].t1)); }
  }
  bool b_ready = false;
  if constexpr (std::is_same<T, float>::value) {
    if (B32 != nullptr && h->fuse_io && vb[level0].kc >= k) { vb[level0].b = const_cast<Cx<float>*>(B32); b_ready = true; }
  }
  int rc = vcycle_level<T, TIO>(h, level0, level0, lb, vb, (const Cx<TIO>*)Bin, (Cx<TIO>*)Xout, k, b_ready);
  h->ws_off = mark;
  return rc;
}

// ---- reductions (complex128 accumulation; the vectors complex128, or complex64 in the mixed-precision Schur solve) ---------
// rows per chunk of the complex64 Gram-Schmidt kernels for n rows and k columns (see dmlmc_hier::gs_rows)
int gs_rows_c64(const dmlmc_hier* h, int n, int k) {
  if (h->gs_rows > 0) return h->gs_rows;
  const long long colblocks = (k / 2 + DOT_TX - 1) / DOT_TX;
  int rows = ROWS_PER_CHUNK;
  while (rows > MIN_ROWS_PER_CHUNK && (long long)((n + rows - 1) / rows) * colblocks < 2048) rows >>= 1;
  return rows;
}
template <typename VT = double, typename WT = double>
int multi_dot(dmlmc_hier* h, const Cx<VT>* Vbase, size_t vstride, int nv, const Cx<WT>* W, int n, int k, Z* partial, Z* out, int accumulate) {
  int nchunks = (n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
  dim3 blk(DOT_TX, DOT_TY), grd((k + DOT_TX - 1) / DOT_TX, nchunks);
  if constexpr (std::is_same<VT, float>::value && std::is_same<WT, float>::value) {
    if ((k % 2) == 0 && (vstride % 2) == 0 && h->gs_x2) {
      const int rows = gs_rows_c64(h, n, k);
      nchunks = (n + rows - 1) / rows;
      dim3 g2((k / 2 + DOT_TX - 1) / DOT_TX, nchunks);
      multi_dot_c64x2_kernel<<<g2, blk, 0, h->stream>>>(Vbase, vstride, nv, W, n, k, rows, partial);
      LAUNCH_CHECK(h);
      sum_partials_kernel<<<nblocks((size_t)nv * k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, nv * k, out, accumulate);
      LAUNCH_CHECK(h);
      return 0;
    }
  }
  multi_dot_kernel<VT, WT><<<grd, blk, 0, h->stream>>>(Vbase, vstride, nv, W, n, k, ROWS_PER_CHUNK, partial);
  LAUNCH_CHECK(h);
  sum_partials_kernel<<<nblocks((size_t)nv * k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, nv * k, out, accumulate);
  LAUNCH_CHECK(h);
  return 0;
}
// the same with the V_i read from their complex64 copies
int multi_dot32(dmlmc_hier* h, const Cx<float>* Vbase, size_t vstride, int nv, const Z* W, int n, int k, Z* partial, Z* out) {
  return multi_dot<float, double>(h, Vbase, vstride, nv, W, n, k, partial, out, 0);
}
// complex numbers of the partial-sum buffer of a reduction over n (or, in the Schur-complement solve, n / 2) rows
size_t partial_count(const dmlmc_hier* h, int n, int nv, int k) {
  size_t chunks = (size_t)(n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
  for (int rows_n : {n, (n + 1) / 2}) {
    const int rows = std::min(ROWS_PER_CHUNK, gs_rows_c64(h, rows_n, k));
    chunks = std::max(chunks, (size_t)(rows_n + rows - 1) / rows);
  }
  return chunks * nv * k;
}

template <typename VT = double>
int multi_axpy(dmlmc_hier* h, const Cx<VT>* Vbase, size_t vstride, int nv, const Z* hc, Z* W, int n, int k, double sgn) {
  const size_t nk = (size_t)n * k;
  multi_axpy_kernel<VT><<<nblocks(nk, 256), 256, 0, h->stream>>>(Vbase, vstride, nv, hc, W, nk, k, sgn);
  LAUNCH_CHECK(h);
  return 0;
}

// W -= sum_i hc[i] V_i and nrm2[col] = ||W[:, col]||^2 (deterministic chunked reduction)
template <typename VT = double>
int multi_axpy_norm(dmlmc_hier* h, const Cx<VT>* Vbase, size_t vstride, int nv, const Z* hc, Cx<VT>* W, int n, int k, Z* partial, Z* nrm2) {
  int nchunks = (n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
  dim3 blk(DOT_TX, DOT_TY), grd((k + DOT_TX - 1) / DOT_TX, nchunks);
  if constexpr (std::is_same<VT, float>::value) {
    if ((k % 2) == 0 && (vstride % 2) == 0 && h->gs_x2) {
      const int rows = gs_rows_c64(h, n, k);
      nchunks = (n + rows - 1) / rows;
      dim3 g2((k / 2 + DOT_TX - 1) / DOT_TX, nchunks);
      multi_axpy_norm_c64x2_kernel<<<g2, blk, 0, h->stream>>>(Vbase, vstride, nv, hc, W, n, k, rows, partial);
      LAUNCH_CHECK(h);
      sum_partials_kernel<<<nblocks((size_t)k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, k, nrm2, 0);
      LAUNCH_CHECK(h);
      return 0;
    }
  }
  multi_axpy_norm_kernel<VT><<<grd, blk, 0, h->stream>>>(Vbase, vstride, nv, hc, W, n, k, ROWS_PER_CHUNK, partial);
  LAUNCH_CHECK(h);
  sum_partials_kernel<<<nblocks((size_t)k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, k, nrm2, 0);
  LAUNCH_CHECK(h);
  return 0;
}

void invalidate_graphs(dmlmc_hier* h) {
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

template <typename T> int smooth_chunked(dmlmc_hier* h, int level, const Cx<T>* R, Cx<T>* E, int k);

int precond_eager(dmlmc_hier* h, int level, const Z* V, Z* Zout, int k, const Cx<float>* V32) {
  if (h->smoother_only) return smooth_chunked<double>(h, level, V, Zout, k);
  if (dmlmc_hier* hp = h->prec_hier[level]) {
    // the preconditioner hierarchy works on the caller's stream (the capture stream while a graph is recorded)
    // and in the caller's work space
    cudaStream_t keep_stream = hp->stream; char* keep_ws = hp->ws; const size_t keep_bytes = hp->ws_bytes, keep_off = hp->ws_off;
    const long long l0 = hp->launches;
    hp->stream = h->stream; hp->ws = h->ws; hp->ws_bytes = h->ws_bytes; hp->ws_off = h->ws_off;
    const int pl = h->prec_level[level];
    const int rc = hp->inner_prec == DMLMC_C128 ? vcycle<double, double>(hp, pl, V, Zout, k) : vcycle<float, double>(hp, pl, V, Zout, k, V32);
    h->launches += hp->launches - l0;
    hp->stream = keep_stream; hp->ws = keep_ws; hp->ws_bytes = keep_bytes; hp->ws_off = keep_off;
    return rc;
  }
  if (h->inner_prec == DMLMC_C128) return vcycle<double, double>(h, level, V, Zout, k);
  return vcycle<float, double>(h, level, V, Zout, k, V32);
}

int precond(dmlmc_hier* h, int level, const Z* V, Z* Zout, int k, const Cx<float>* V32 = nullptr) { return precond_eager(h, level, V, Zout, k, V32); }

// One FGMRES iteration (V-cycle, operator, Gram-Schmidt, Givens step, next basis vector) is a fixed sequence of ~105
// launches whose arguments depend only on (level, k, precision, work space, Krylov index j, restart, tol, reorth).
// Small batches are launch-bound (8.4 us per launch at k = 1) and even at k = 256 the gaps between launches cost
// ~6 %, so each such sequence is captured once into a CUDA graph and replayed by every later solve.  The first
// iteration of a configuration runs eagerly (one-time attribute / descriptor set-up); capture happens on an internal
// stream (the caller's may be the legacy default stream, which cannot capture) and the instantiated graph is
// launched on the caller's stream like every other kernel.  Any set_* / option call drops the graphs.
template <typename Body>
int run_iteration(dmlmc_hier* h, bool graphs, int level, int k, size_t mark, int j, int m, double tol, Body body) {
  if (!graphs) return body();
  dmlmc_hier::GraphEntry* e = nullptr;
  bool warmed = false;
  for (auto& g : h->graphs) {
    if (!(g.level == level && g.k == k && g.prec == h->inner_prec && g.ws == h->ws && g.ws_off == mark && g.m == m &&
          g.tol == tol && g.reorth == h->reorth)) continue;
    if (g.j < 0) warmed = true;
    else if (g.j == j) e = &g;
  }
  if (!warmed) {
    h->graphs.push_back({level, k, h->inner_prec, h->ws, mark, -1, m, h->reorth, tol, nullptr, 0});
    return body();
  }
  if (e != nullptr) {
    CU(cudaGraphLaunch(e->exec, h->stream));
    h->launches += e->launches;
    return 0;
  }
  h->graphs.push_back({level, k, h->inner_prec, h->ws, mark, j, m, h->reorth, tol, nullptr, 0});
  e = &h->graphs.back();
  const long long l0 = h->launches;
  const size_t ws_keep = h->ws_off;
  cudaGraph_t graph = nullptr;
  int rc = 0;
  cudaStream_t user = h->stream;
  cudaError_t ce = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed);
  if (ce == cudaSuccess) {
    h->stream = h->cap_stream;
    rc = body();
    h->stream = user;
    ce = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc == 0 && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&e->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
  }
  if (rc != 0 || ce != cudaSuccess || e->exec == nullptr) {      // capture not possible here: stay eager from now on
    cudaGetLastError();
    h->use_graphs = 0;
    if (e->exec) cudaGraphExecDestroy(e->exec);
    h->graphs.pop_back();
    h->launches = l0;
    h->ws_off = ws_keep;
    return body();
  }
  e->launches = h->launches - l0;
  CU(cudaGraphLaunch(e->exec, h->stream));
  return 0;
}

int read_nactive(dmlmc_hier* h, int* dev_counter, int* out) {
  CU(cudaMemcpyAsync(h->h_nactive, dev_counter, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = *h->h_nactive;
  return 0;
}

bool outer_eo_ok(dmlmc_hier* h, int level, int k);
int fgmres_eo(dmlmc_hier* h, int level, const Z* B, Z* X, int k, double tol, int restart, int maxiter,
              int32_t* iters_host, double* relres_host);

// ---- batched FGMRES (multigrid.py:347-366 / pyamg.krylov.fgmres) ------------------------------
int fgmres(dmlmc_hier* h, int level, const Z* B, Z* X, int k, double tol, int restart, int maxiter,
           int32_t* iters_host, double* relres_host) {
  CHECK(level >= 0 && level < h->n_levels, "fgmres: bad level");
  CHECK(k >= 1 && restart >= 1 && maxiter >= 1, "fgmres: bad k / restart / maxiter");
  if (outer_eo_ok(h, level, k)) return fgmres_eo(h, level, B, X, k, tol, restart, maxiter, iters_host, relres_host);
  Level& L = h->lv[level];
  const int n = L.n, m = restart;
  const size_t nk = (size_t)n * k;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  GmresState s; s.k = k; s.m = m;
  Z *Vb, *Zb, *W, *Rb, *partial;
  RET(ws_get<Z>(h, nk * (m + 1), &Vb));
  RET(ws_get<Z>(h, nk * m, &Zb));
  RET(ws_get<Z>(h, nk, &W));
  RET(ws_get<Z>(h, nk, &Rb));
  RET(ws_get<Z>(h, partial_count(h, n, m + 1, k), &partial));
  RET(ws_get<Z>(h, (size_t)(m + 1) * k, &s.hsum));
  RET(ws_get<Z>(h, (size_t)k, &s.nrm2));
  RET(ws_get<Z>(h, (size_t)m * m * k, &s.Rm));
  RET(ws_get<double>(h, (size_t)m * k, &s.cs));
  RET(ws_get<Z>(h, (size_t)m * k, &s.sn));
  RET(ws_get<Z>(h, (size_t)(m + 1) * k, &s.g));
  RET(ws_get<Z>(h, (size_t)m * k, &s.y));
  RET(ws_get<double>(h, (size_t)k, &s.normb));
  RET(ws_get<double>(h, (size_t)k, &s.scale));
  RET(ws_get<double>(h, (size_t)k, &s.relres));
  RET(ws_get<double>(h, (size_t)k, &s.tolc));
  RET(ws_get<int>(h, (size_t)k, &s.active));
  RET(ws_get<int>(h, (size_t)k, &s.done));
  RET(ws_get<int>(h, (size_t)k, &s.it_cycle));
  RET(ws_get<int>(h, (size_t)k, &s.it_total));
  RET(ws_get<int>(h, 1, &s.n_active));
  // complex64 copy of the current basis vector = the V-cycle's input, written by the kernel that normalises it
  Cx<float>* V32 = nullptr;
  {
    dmlmc_hier* hv = h->prec_hier[level] ? h->prec_hier[level] : h;
    if (h->fuse_io && hv->fuse_io && hv->inner_prec == DMLMC_C64) RET(ws_get<Cx<float>>(h, h->dot32 ? nk * (m + 1) : nk, &V32));
  }
  const bool dot32 = V32 != nullptr && h->dot32;       // complex64 copies of ALL basis vectors, read by the Gram-Schmidt dots
  auto v32 = [&](int j) -> Cx<float>* { return V32 == nullptr ? nullptr : (dot32 ? V32 + (size_t)j * nk : V32); };

  CU(cudaMemsetAsync(X, 0, nk * sizeof(Z), h->stream));
  const Z* Rsrc = B;
  const int expect = (h->expect_tol[level] == tol && h->expect_k[level] == k) ? h->expect_it[level] : 0;
  int total_it = 0, nact = 0;
  int mode = 1;                      // 1: first cycle (r = b), 2: later cycles (true residual of every column)
  const unsigned gk = nblocks(k, 128);
  while (true) {
    // Every cycle after the first starts from the TRUE residual b - A x of every column: a column
    // leaves the solve only when that residual meets the tolerance, whatever the Arnoldi recurrence
    // estimated (this is what makes single-pass Gram-Schmidt safe).
    RET(multi_dot(h, Rsrc, 0, 1, Rsrc, n, k, partial, s.nrm2, 0));
    CU(cudaMemsetAsync(s.n_active, 0, sizeof(int), h->stream));
    gmres_init_kernel<<<nblocks(k, 256), 256, 0, h->stream>>>(s, tol, mode); LAUNCH_CHECK(h);
    mode = 2;
    RET(read_nactive(h, s.n_active, &nact));
    if (nact == 0 || total_it >= maxiter) { h->unconverged += nact; break; }
    col_scale_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(Rsrc, s.scale, Vb, nk, k, v32(0)); LAUNCH_CHECK(h);
    int j = 0;
    for (; j < m; ++j) {
      Z* Vj = Vb + (size_t)j * nk;
      Z* Zj = Zb + (size_t)j * nk;
      auto body = [&]() -> int {
        RET(precond(h, level, Vj, Zj, k, v32(j)));
        RET((launch_op<double, M_AX>(h, level, Zj, nullptr, W, ZERO, ZERO, k)));
        // classical Gram-Schmidt, optionally with one re-orthogonalisation pass
        if (dot32) RET(multi_dot32(h, V32, nk, j + 1, W, n, k, partial, s.hsum));
        else       RET(multi_dot(h, Vb, nk, j + 1, W, n, k, partial, s.hsum, 0));
        if (h->reorth) {
          RET(multi_axpy(h, Vb, nk, j + 1, s.hsum, W, n, k, -1.0));
          RET(multi_dot(h, Vb, nk, j + 1, W, n, k, partial, s.y, 0));       // s.y used as scratch [(j+1)][k]
          RET(multi_axpy_norm(h, Vb, nk, j + 1, s.y, W, n, k, partial, s.nrm2));
          // hsum += second-pass coefficients
          const int cnt = (j + 1) * k;
          sum_partials_kernel<<<nblocks(cnt, 32), dim3(32, 8), 0, h->stream>>>(s.y, 1, cnt, s.hsum, 1); LAUNCH_CHECK(h);
        } else {
          // w -= V h and ||w||^2 in one pass over W
          RET(multi_axpy_norm(h, Vb, nk, j + 1, s.hsum, W, n, k, partial, s.nrm2));
        }
        CU(cudaMemsetAsync(s.n_active, 0, sizeof(int), h->stream));
        gmres_step_kernel<<<gk, 128, 0, h->stream>>>(s, j, tol); LAUNCH_CHECK(h);
        if (j + 1 < m) { col_scale_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(W, s.scale, Vb + (size_t)(j + 1) * nk, nk, k, v32(j + 1)); LAUNCH_CHECK(h); }
        return 0;
      };
      RET(run_iteration(h, h->use_graphs && k <= h->graph_max_k, level, k, mark, j, m, tol, body));
      ++total_it;
      // (also every 4th iteration, so that a stale expectation wastes at most 3 iterations)
      if (!h->adaptive_poll || total_it + 1 >= expect || (total_it & 3) == 0 || j + 1 == m || total_it >= maxiter)
        RET(read_nactive(h, s.n_active, &nact));
      if (nact == 0 || total_it >= maxiter) { ++j; break; }
    }
    const int steps = std::min(j, m);
    gmres_solve_kernel<<<gk, 128, 0, h->stream>>>(s, steps); LAUNCH_CHECK(h);
    RET(multi_axpy(h, Zb, nk, steps, s.y, X, n, k, +1.0));
    RET((launch_op<double, M_RES>(h, level, X, B, Rb, ZERO, ZERO, k)));
    Rsrc = Rb;
  }
  h->expect_it[level] = total_it; h->expect_tol[level] = tol; h->expect_k[level] = k;
  if (iters_host) CU(cudaMemcpyAsync(iters_host, s.it_total, sizeof(int) * k, cudaMemcpyDeviceToHost, h->stream));
  if (relres_host) CU(cudaMemcpyAsync(relres_host, s.relres, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
  if (iters_host || relres_host) CU(cudaStreamSynchronize(h->stream));
  h->ws_off = mark;
  return 0;
}

// ---- outer solve on the even-odd Schur complement (option "outer_eo", default 1; profiles/r2_run1_*: 27.97k -> 34.21k probes/s) ----
// Batched FGMRES of a stencil level on the even-odd Schur complement:  S x_e = b^_e = b_e - H_eo b_o / c,  S = c - H_eo H_oe / c,
// x_o = (b_o - H_oe x_e) / c.  All Krylov vectors are half-lattice (checkerboard) arrays, so Gram-Schmidt, normalisation and
// the final update move half the bytes; the preconditioner is the even part of the V-cycle applied to (v_e, 0).  CPU experiment
// (profiles/exp_schur_outer_solve.py): the same 8 outer iterations to 1e-12 as the solve on A.  The residual of the full
// system is (r^_e, 0), so convergence is measured as ||r^_e|| / ||b||.
template <typename T, bool HAS2>
int launch_hop_z(dmlmc_hier* h, const Level& L, int p, const Cx<T>* Inq, const Cx<T>* In2, Cx<T>* Out, Cx<double> a, Cx<double> b, int k) {
  StencilDev<T> op; op.LX = L.LX; op.LT = L.LT; op.diag = cx<T>((T)L.d.diag.re, (T)L.d.diag.im);
  if constexpr (std::is_same<T, double>::value) { op.Ut = L.d.Ut; op.Ux = L.d.Ux; } else { op.Ut = L.f.Ut; op.Ux = L.f.Ux; }
  constexpr int NC = std::is_same<T, double>::value ? 1 : 2;       // complex64: two columns per thread (k is even on this path)
  const int kp = k / NC;
  int bx = 1; while (bx < 32 && bx < kp) bx *= 2;
  dim3 blk(bx, 4, 2), grd((kp + bx - 1) / bx, (L.LT / 2 + 3) / 4, (L.LX + 1) / 2);
  wilson_hop_eo_z_kernel<T, NC, HAS2><<<grd, blk, 0, h->stream>>>(op, p, Inq, In2, Out, cx<T>((T)a.re, (T)a.im), cx<T>((T)b.re, (T)b.im), kp);
  LAUNCH_CHECK(h);
  return 0;
}
// r = b^ - S x and ||r||^2 per column in one pass (see wilson_schur_residual_kernel); false if the partial buffer is too small
bool launch_schur_residual(dmlmc_hier* h, const Level& L, const Z* Wo, const Z* Xe, const Z* Bhat, Z* Rb, double c, int k,
                           Z* partial, size_t partial_cap, Z* nrm2, int* rc) {
  StencilDev<double> op; op.LX = L.LX; op.LT = L.LT; op.Ut = L.d.Ut; op.Ux = L.d.Ux; op.diag = L.d.diag;
  int bx = 1; while (bx < 32 && bx < k) bx *= 2;
  dim3 blk(bx, 4, 2), grd((k + bx - 1) / bx, (L.LT / 2 + 3) / 4, (L.LX + 1) / 2);
  const size_t nchunks = (size_t)grd.y * grd.z;
  if (!h->fuse_residual || nchunks * k > partial_cap) return false;
  *rc = 0;
  wilson_schur_residual_kernel<<<grd, blk, 0, h->stream>>>(op, 0, Wo, Xe, Bhat, Rb, cx<double>(c, 0.0), cx<double>(-1.0 / c, 0.0), k, partial);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *rc = fail((int)e, std::string("kernel launch: ") + cudaGetErrorString(e)); return true; }
  sum_partials_kernel<<<nblocks((size_t)k, 32), dim3(32, 8), 0, h->stream>>>(partial, (int)nchunks, k, nrm2, 0);
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) *rc = fail((int)e, std::string("kernel launch: ") + cudaGetErrorString(e));
  return true;
}
bool outer_eo_ok(dmlmc_hier* h, int level, int k) {
  if (!h->outer_eo || !h->fuse_io || h->smoother_only) return false;
  Level& L = h->lv[level];
  if (L.kind != 0 || (L.LT % 2) || (L.LX % 2) || L.d.diag.im != 0.0) return false;
  dmlmc_hier* hv = h->prec_hier[level] ? h->prec_hier[level] : h;
  const int pl = h->prec_hier[level] ? h->prec_level[level] : level;
  return hv->inner_prec == DMLMC_C64 && smoother_dout_ok(hv, pl, k) && smoother_eo_ok(hv, pl, k) && !hv->pre_smooth &&
         chunk_cols(hv, pl, k, sizeof(Cx<float>)) >= k;
}
// VT = double: every Krylov vector complex128.  VT = float (option outer_c64, the default): V_j, Z_j and w are STORED in complex64
// -- Gram-Schmidt, the normalisation, S z_j and the solution update move half the bytes again -- while the coefficients are
// accumulated in FP64 and x_e, b^_e and the true residual r = b^_e - S x_e stay complex128.  The Arnoldi relation then holds to
// ~1e-7 of the cycle's starting residual only, so a cycle stops once its estimate has dropped by `outer_drop` (1e-4) and the next
// starts from the true residual: iterative refinement with FGMRES as the inner solver.  CPU experiment
// (profiles/exp_mixed_precision_ir.py): the same 8 preconditioner applications to 1e-12 in 2-3 cycles.
template <typename VT>
int fgmres_eo_t(dmlmc_hier* h, int level, const Z* B, Z* X, int k, double tol, int restart, int maxiter,
                int32_t* iters_host, double* relres_host) {
  typedef Cx<VT> C;
  constexpr bool MIXED = std::is_same<VT, float>::value;
  Level& L = h->lv[level];
  dmlmc_hier* hv = h->prec_hier[level] ? h->prec_hier[level] : h;
  const int n = L.n, nh = n / 2, m = restart;
  const size_t nk = (size_t)n * k, nkh = (size_t)nh * k;
  const double c = L.d.diag.re;
  const Cx<double> ONE = {1.0, 0.0}, CC = {c, 0.0}, NIC = {-1.0 / c, 0.0}, IC = {1.0 / c, 0.0};
  const double drop = MIXED ? h->outer_drop : 0.0;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  GmresState s; s.k = k; s.m = m;
  C *Vb, *Zb, *W, *Wv;
  Z *Wo, *Wd, *Rb, *Be, *Bo, *Bhat, *Xe, *Xo, *partial;
  RET(ws_get<C>(h, nkh * (m + 1), &Vb));
  RET(ws_get<C>(h, nkh * m, &Zb));
  RET(ws_get<C>(h, nkh, &W)); RET(ws_get<C>(h, nkh, &Wv));
  RET(ws_get<Z>(h, nkh, &Wo)); RET(ws_get<Z>(h, nkh, &Wd)); RET(ws_get<Z>(h, nkh, &Rb));
  RET(ws_get<Z>(h, nkh, &Be)); RET(ws_get<Z>(h, nkh, &Bo)); RET(ws_get<Z>(h, nkh, &Bhat));
  RET(ws_get<Z>(h, nkh, &Xe)); RET(ws_get<Z>(h, nkh, &Xo));
  RET(ws_get<Z>(h, partial_count(h, n, m + 1, k), &partial));
  RET(ws_get<Z>(h, (size_t)(m + 1) * k, &s.hsum));
  RET(ws_get<Z>(h, (size_t)k, &s.nrm2));
  RET(ws_get<Z>(h, (size_t)m * m * k, &s.Rm));
  RET(ws_get<double>(h, (size_t)m * k, &s.cs));
  RET(ws_get<Z>(h, (size_t)m * k, &s.sn));
  RET(ws_get<Z>(h, (size_t)(m + 1) * k, &s.g));
  RET(ws_get<Z>(h, (size_t)m * k, &s.y));
  RET(ws_get<double>(h, (size_t)k, &s.normb));
  RET(ws_get<double>(h, (size_t)k, &s.scale));
  RET(ws_get<double>(h, (size_t)k, &s.relres));
  RET(ws_get<double>(h, (size_t)k, &s.tolc));
  RET(ws_get<int>(h, (size_t)k, &s.active));
  RET(ws_get<int>(h, (size_t)k, &s.done));
  RET(ws_get<int>(h, (size_t)k, &s.it_cycle));
  RET(ws_get<int>(h, (size_t)k, &s.it_total));
  RET(ws_get<int>(h, 1, &s.n_active));
  Cx<float>* V32;                                        // (v_e, 0) as complex64 on the full lattice: the V-cycle's input
  RET(ws_get<Cx<float>>(h, nk, &V32));
  CU(cudaMemsetAsync(V32, 0, nk * sizeof(Cx<float>), h->stream));
  const unsigned gk = nblocks(k, 128);
  // ||b|| of the full system, b_e, b_o, b^_e
  RET(multi_dot(h, B, 0, 1, B, n, k, partial, s.nrm2, 0));
  set_normb_kernel<<<gk, 128, 0, h->stream>>>(s.nrm2, s.normb, k); LAUNCH_CHECK(h);
  eo_split_merge_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(L.LX, L.LT, k, 0, const_cast<Z*>(B), Be, Bo); LAUNCH_CHECK(h);
  RET((launch_hop_z<double, true>(h, L, 0, Bo, Be, Bhat, ONE, NIC, k)));
  CU(cudaMemsetAsync(Xe, 0, nkh * sizeof(Z), h->stream));
  const Z* Rsrc = Bhat;
  // Host polls (a blocking read of the active-column count) only where the previous solve of the same shape ended a cycle: its
  // iteration counts per cycle are the expectation; a stale one costs wasted iterations on finished (zeroed) columns, not errors
  const bool have_exp = h->adaptive_poll && h->expect_tol[level] == tol && h->expect_k[level] == k && !h->expect_cyc[level].empty();
  const std::vector<int> exp_cyc = have_exp ? h->expect_cyc[level] : std::vector<int>();
  std::vector<int> cyc_len;
  int total_it = 0, nact = 0, mode = 3, cycles = 0;
  bool have_norm = false;                 // s.nrm2 already holds ||Rsrc||^2 per column (written by the fused residual kernel)
  const size_t partial_cap = partial_count(h, n, m + 1, k);
  while (true) {
    if (!have_norm) RET(multi_dot(h, Rsrc, 0, 1, Rsrc, nh, k, partial, s.nrm2, 0));
    have_norm = false;
    CU(cudaMemsetAsync(s.n_active, 0, sizeof(int), h->stream));
    gmres_init_kernel<<<nblocks(k, 256), 256, 0, h->stream>>>(s, tol, mode, drop); LAUNCH_CHECK(h);
    mode = 2;
    if (!have_exp || cycles >= (int)exp_cyc.size() || total_it >= maxiter) {
      RET(read_nactive(h, s.n_active, &nact));
      if (nact == 0 || total_it >= maxiter) { h->unconverged += nact; break; }
    }
    const int this_exp = (have_exp && cycles < (int)exp_cyc.size()) ? exp_cyc[cycles] : 1;
    col_scale_eo_kernel<double, VT><<<nblocks(nkh, 256), 256, 0, h->stream>>>(L.LX, L.LT, Rsrc, s.scale, Vb, k, V32); LAUNCH_CHECK(h);
    int j = 0;
    for (; j < m; ++j) {
      C* Vj = Vb + (size_t)j * nkh;
      C* Zj = Zb + (size_t)j * nkh;
      auto body = [&]() -> int {
        hv->eo_zhalf = MIXED ? 2 : 1;
        const int rcp = precond(h, level, reinterpret_cast<const Z*>(Vj), reinterpret_cast<Z*>(Zj), k, V32);   // Z_j = even part of M^{-1} (v_e, 0), half-lattice layout
        hv->eo_zhalf = 0;
        RET(rcp);
        RET((launch_hop_z<VT, false>(h, L, 1, Zj, nullptr, Wv, ONE, ONE, k)));        // w_o = H_oe z
        RET((launch_hop_z<VT, true>(h, L, 0, Wv, Zj, W, CC, NIC, k)));                // w = c z - H_eo w_o / c
        RET((multi_dot<VT, VT>(h, Vb, nkh, j + 1, W, nh, k, partial, s.hsum, 0)));
        RET((multi_axpy_norm<VT>(h, Vb, nkh, j + 1, s.hsum, W, nh, k, partial, s.nrm2)));
        CU(cudaMemsetAsync(s.n_active, 0, sizeof(int), h->stream));
        gmres_step_kernel<<<gk, 128, 0, h->stream>>>(s, j, tol); LAUNCH_CHECK(h);
        if (j + 1 < m) {
          bool done2 = false;
          if constexpr (MIXED) {
            if (h->gs_x2 && (k % 2) == 0) {
              col_scale_eo_c64x2_kernel<<<nblocks(nkh / 2, 256), 256, 0, h->stream>>>(L.LX, L.LT, W, s.scale, Vb + (size_t)(j + 1) * nkh, k, V32);
              LAUNCH_CHECK(h);
              done2 = true;
            }
          }
          if (!done2) { col_scale_eo_kernel<VT, VT><<<nblocks(nkh, 256), 256, 0, h->stream>>>(L.LX, L.LT, W, s.scale, Vb + (size_t)(j + 1) * nkh, k, V32); LAUNCH_CHECK(h); }
        }
        return 0;
      };
      // (the kernel arguments of iteration j are the same in every cycle: one graph per j serves all cycles)
      RET(run_iteration(h, h->use_graphs && k <= h->graph_max_k, level + 64 + (MIXED ? 64 : 0), k, mark, j, m, tol, body));
      ++total_it;
      if (j + 1 >= this_exp || j + 1 == m || total_it >= maxiter) RET(read_nactive(h, s.n_active, &nact));
      else nact = 1;
      if (nact == 0 || total_it >= maxiter) { ++j; break; }
    }
    const int steps = std::min(j, m);
    cyc_len.push_back(steps);
    gmres_solve_kernel<<<gk, 128, 0, h->stream>>>(s, steps); LAUNCH_CHECK(h);
    RET((multi_axpy<VT>(h, Zb, nkh, steps, s.y, Xe, nh, k, +1.0)));
    // true residual of the Schur system (= the residual of the full system, whose odd part is zero by construction)
    RET((launch_hop_z<double, false>(h, L, 1, Xe, nullptr, Wo, ONE, ONE, k)));
    int rc_res = 0;
    if (launch_schur_residual(h, L, Wo, Xe, Bhat, Rb, c, k, partial, partial_cap, s.nrm2, &rc_res)) {
      RET(rc_res);
      have_norm = true;
    } else {
      RET((launch_hop_z<double, true>(h, L, 0, Wo, Xe, Wd, CC, NIC, k)));
      vec_sub_kernel<<<nblocks(nkh, 256), 256, 0, h->stream>>>(Bhat, Wd, Rb, nkh); LAUNCH_CHECK(h);
    }
    Rsrc = Rb;
    ++cycles;
  }
  h->expect_cyc[level] = cyc_len;
  // x_o = (b_o - H_oe x_e) / c, then the full-lattice layout
  RET((launch_hop_z<double, true>(h, L, 1, Xe, Bo, Xo, IC, NIC, k)));
  eo_split_merge_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(L.LX, L.LT, k, 1, X, Xe, Xo); LAUNCH_CHECK(h);
  h->expect_it[level] = total_it; h->expect_tol[level] = tol; h->expect_k[level] = k;
  if (iters_host) CU(cudaMemcpyAsync(iters_host, s.it_total, sizeof(int) * k, cudaMemcpyDeviceToHost, h->stream));
  if (relres_host) CU(cudaMemcpyAsync(relres_host, s.relres, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
  if (iters_host || relres_host) CU(cudaStreamSynchronize(h->stream));
  h->ws_off = mark;
  return 0;
}
int fgmres_eo(dmlmc_hier* h, int level, const Z* B, Z* X, int k, double tol, int restart, int maxiter,
              int32_t* iters_host, double* relres_host) {
  if (h->outer_c64 && (k % 2) == 0) return fgmres_eo_t<float>(h, level, B, X, k, tol, restart, maxiter, iters_host, relres_host);
  return fgmres_eo_t<double>(h, level, B, X, k, tol, restart, maxiter, iters_host, relres_host);
}

int apply_perm(dmlmc_hier* h, int level, const Z* X, Z* Y, int k) {
  Level& L = h->lv[level];
  const size_t nk = (size_t)L.n * k;
  if (!L.has_perm) { CU(cudaMemcpyAsync(Y, X, nk * sizeof(Z), cudaMemcpyDeviceToDevice, h->stream)); return 0; }
  perm_kernel<double, 1><<<nblocks(nk, 256), 256, 0, h->stream>>>(L.n, L.shift, L.perm_nnz, L.perm_cols, L.d.perm_vals,
                                                                 (const Pack<double, 1>*)X, (Pack<double, 1>*)Y, k);
  LAUNCH_CHECK(h);
  return 0;
}

int deflate(dmlmc_hier* h, int level, Z* X, int k) {
  Level& L = h->lv[level];
  if (L.defl_d == 0) return 0;
  const int n = L.n, d = L.defl_d;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  Z *partial, *C;
  RET(ws_get<Z>(h, partial_count(h, n, d, k), &partial));
  RET(ws_get<Z>(h, (size_t)d * k, &C));
  const int nchunks = (n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
  const size_t nk = (size_t)n * k;
  if (h->defl_tensor && d % 4 == 0 && d <= 64) {
    // FP64 tensor cores (DMMA m8n8k4): both projections as real GEMMs on the interleaved arrays
    const size_t sm1 = (size_t)DD_KB * (2 * d + 4 + 2 * DD_NT + 4) * sizeof(double);
    const size_t sm2 = ((size_t)64 * (2 * d + 4) + (size_t)2 * d * (2 * DD_NT + 4)) * sizeof(double);
    if (!h->dmma_attr_set) {
      CU(cudaFuncSetAttribute(defl_dot_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      CU(cudaFuncSetAttribute(defl_dot_dmma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      CU(cudaFuncSetAttribute(defl_dot_dmma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      CU(cudaFuncSetAttribute(defl_axpy_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      h->dmma_attr_set = true;
    }
    dim3 g1((k + DD_NT - 1) / DD_NT, nchunks);
    if (d <= 16)      defl_dot_dmma_kernel<4><<<g1, 256, sm1, h->stream>>>(L.defl_V, d, X, n, k, ROWS_PER_CHUNK, partial);
    else if (d <= 32) defl_dot_dmma_kernel<8><<<g1, 256, sm1, h->stream>>>(L.defl_V, d, X, n, k, ROWS_PER_CHUNK, partial);
    else              defl_dot_dmma_kernel<16><<<g1, 256, sm1, h->stream>>>(L.defl_V, d, X, n, k, ROWS_PER_CHUNK, partial);
    LAUNCH_CHECK(h);
    sum_partials_kernel<<<nblocks((size_t)d * k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, d * k, C, 0); LAUNCH_CHECK(h);
    dim3 g2((k + DD_NT - 1) / DD_NT, (n + 63) / 64);
    defl_axpy_dmma_kernel<<<g2, 256, sm2, h->stream>>>(L.defl_V, d, C, X, n, k); LAUNCH_CHECK(h);
  } else {
    dim3 blk(DOT_TX, DOT_TY), grd((k + DOT_TX - 1) / DOT_TX, nchunks);
    defl_dot_kernel<<<grd, blk, 0, h->stream>>>(L.defl_V, d, X, n, k, ROWS_PER_CHUNK, partial); LAUNCH_CHECK(h);
    sum_partials_kernel<<<nblocks((size_t)d * k, 32), dim3(32, 8), 0, h->stream>>>(partial, nchunks, d * k, C, 0); LAUNCH_CHECK(h);
    defl_axpy_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(L.defl_V, d, C, X, nk, k); LAUNCH_CHECK(h);
  }
  h->ws_off = mark;
  return 0;
}

// ---- one batch of utils.one_defl_Hutch_step (utils.py:207-361) ----------------------------------
int level_sample(dmlmc_hier* h, int method, int lf, int lc, const Z* X0, int k, double tol, int restart,
                 int maxiter, Z* e_dev, int32_t* iters_host) {
  const int nl = h->n_levels;
  CHECK(lf >= 0 && lf < nl - 1, "level_sample: bad fine level");
  CHECK(method == 0 || method == 1, "level_sample: bad method");
  if (method == 0) CHECK(lf == 0, "level_sample: hutchinson runs on level 0");
  if (method == 1) CHECK(lc == lf + 1 || lc == lf + 2, "level_sample: coarse level must be fine+1 or fine+2");
  if (method == 1) CHECK(lc < nl, "level_sample: coarse level out of range");
  Level& Lf = h->lv[lf];
  const int nf = Lf.n;
  const size_t nkf = (size_t)nf * k;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  Z *Xdef, *RHS, *Zs, *partial, *e1;
  RET(ws_get<Z>(h, nkf, &Xdef));
  RET(ws_get<Z>(h, nkf, &RHS));
  RET(ws_get<Z>(h, nkf, &Zs));
  RET(ws_get<Z>(h, partial_count(h, nf, 1, k), &partial));
  RET(ws_get<Z>(h, (size_t)k, &e1));
  // x_def = x0 - V (V^H x0)            utils.py:224,266
  const Z* xd = X0;
  if (Lf.defl_d > 0) {
    CU(cudaMemcpyAsync(Xdef, X0, nkf * sizeof(Z), cudaMemcpyDeviceToDevice, h->stream));
    RET(deflate(h, lf, Xdef, k));
    xd = Xdef;
  }
  // rhs = Bblock_perm * roll(x_def, +shift)   utils.py:232,288-290
  const Z* rhs = xd;
  if (Lf.has_perm) { RET(apply_perm(h, lf, xd, RHS, k)); rhs = RHS; }
  RET(fgmres(h, lf, rhs, Zs, k, tol, restart, maxiter, iters_host, nullptr));
  RET(multi_dot(h, X0, 0, 1, Zs, nf, k, partial, e1, 0));            // e1 = x0^H z   utils.py:249,336
  if (method == 0) {
    CU(cudaMemcpyAsync(e_dev, e1, sizeof(Z) * k, cudaMemcpyDeviceToDevice, h->stream));
    h->ws_off = mark;
    return 0;
  }
  // xc = R (R) rhs                      utils.py:299-303
  const bool skip = (lc == lf + 2);
  const int nc = h->lv[lc].n;
  Z *Xm = nullptr, *Xc, *Y, *e2;
  if (skip) RET(ws_get<Z>(h, (size_t)h->lv[lf + 1].n * k, &Xm));
  RET(ws_get<Z>(h, (size_t)nc * k, &Xc));
  RET(ws_get<Z>(h, (size_t)nc * k, &Y));
  RET(ws_get<Z>(h, (size_t)k, &e2));
  if (skip) { RET(launch_restrict<double>(h, lf, rhs, Xm, k)); RET(launch_restrict<double>(h, lf + 1, Xm, Xc, k)); }
  else      { RET(launch_restrict<double>(h, lf, rhs, Xc, k)); }
  // y = A_c^{-1} xc                     utils.py:306-329
  if (lc == nl - 1) RET(launch_dense<double>(h, lc, Xc, Y, k));
  else RET(fgmres(h, lc, Xc, Y, k, tol, restart, maxiter, iters_host ? iters_host + k : nullptr, nullptr));
  // w = P (P) y ; e2 = x0^H w           utils.py:337-353  (w overwrites the solution buffer's sibling RHS)
  Z* Wf = RHS;     // RHS is no longer needed (rhs may alias X0/Xdef when no perm: then RHS is free as well)
  CU(cudaMemsetAsync(Wf, 0, nkf * sizeof(Z), h->stream));
  if (skip) {
    CU(cudaMemsetAsync(Xm, 0, (size_t)h->lv[lf + 1].n * k * sizeof(Z), h->stream));
    RET(launch_prolong<double>(h, lf + 1, Y, Xm, k));
    RET(launch_prolong<double>(h, lf, Xm, Wf, k));
  } else {
    RET(launch_prolong<double>(h, lf, Y, Wf, k));
  }
  RET(multi_dot(h, X0, 0, 1, Wf, nf, k, partial, e2, 0));
  zsub_kernel<<<nblocks(k, 128), 128, 0, h->stream>>>(e1, e2, e_dev, k); LAUNCH_CHECK(h);
  h->ws_off = mark;
  return 0;
}

template <typename T>
int smooth_chunked(dmlmc_hier* h, int level, const Cx<T>* R, Cx<T>* E, int k) {
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  const int n = h->lv[level].n;
  const int kc = chunk_cols(h, level, k, sizeof(Cx<T>));
  Cx<T>*rc, *ec, *t0, *t1;
  RET(ws_get<Cx<T>>(h, (size_t)n * kc, &rc)); RET(ws_get<Cx<T>>(h, (size_t)n * kc, &ec));
  RET(ws_get<Cx<T>>(h, (size_t)n * kc, &t0)); RET(ws_get<Cx<T>>(h, (size_t)n * kc, &t1));
  for (int col0 = 0; col0 < k; col0 += kc) {
    const int w = std::min(kc, k - col0);
    RET((cvt_cols<T, T>(h, R + col0, (size_t)k, rc, (size_t)w, n, w)));
    RET(smooth_apply<T>(h, level, rc, ec, false, t0, t1, w));
    RET((cvt_cols<T, T>(h, ec, (size_t)w, E + col0, (size_t)k, n, w)));
  }
  h->ws_off = mark;
  return 0;
}
size_t vcycle_bytes(dmlmc_hier* h, int level, int k, size_t elem) {
  size_t b = 0;
  for (int l = level; l < h->n_levels; ++l) b += 4 * align_up((size_t)h->lv[l].n * k * elem);
  return b + 4096;
}
size_t fgmres_bytes(dmlmc_hier* h, int level, int k, int m) {
  const size_t n = h->lv[level].n, nk = n * (size_t)k, z = sizeof(Z);
  size_t b = 0;
  if (outer_eo_ok(h, level, k)) {
    // solve on the even-odd Schur complement: half-lattice Krylov vectors V_j, Z_j, w, w_o in the storage precision, eight
    // complex128 half-lattice work vectors (fgmres_eo_t) -- 5.4 GB instead of 22 GB at n = 32768, k = 512, m = 40
    const size_t nkh = nk / 2, e = (h->outer_c64 && (k % 2) == 0) ? sizeof(Cx<float>) : z;
    b += align_up(nkh * (m + 1) * e) + align_up(nkh * m * e) + 2 * align_up(nkh * e) + 8 * align_up(nkh * z);
  } else {
    b += align_up(nk * (m + 1) * z) + align_up(nk * m * z) + 2 * align_up(nk * z);
  }
  b += align_up(partial_count(h, (int)n, m + 1, k) * z);
  b += align_up((size_t)m * m * k * z) + 6 * align_up((size_t)(m + 1) * k * z) + 16 * align_up((size_t)k * 16);
  size_t vb = vcycle_bytes(h, level, k, sizeof(Z));
  if (h->prec_hier[level]) vb = std::max(vb, vcycle_bytes(h->prec_hier[level], h->prec_level[level], k, sizeof(Z)) + (size_t)(1 << 16));
  b += vb;
  b += align_up(nk * (h->dot32 ? m + 1 : 1) * sizeof(Cx<float>));   // complex64 copy of the current basis vector (all: dot32)
  b += 2 * align_up(nk * z);          // staging buffers of the graph-replayed V-cycle
  return b + (1 << 16);
}

}  // namespace

// =================================================================================================
extern "C" {

int dmlmc_abi_version(void) { return DMLMC_ABI_VERSION; }
const char* dmlmc_last_error(void) { return g_err.c_str(); }

int dmlmc_hier_create(int device, void* cuda_stream, int n_levels, dmlmc_hier** out) {
  CHECK(out != nullptr, "hier_create: out is NULL");
  CHECK(n_levels >= 1 && n_levels <= MAX_LEVELS, "hier_create: bad number of levels");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(e != cudaSuccess ? (int)e : -3, "dmlmc: no CUDA device available (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  dmlmc_hier* h = new dmlmc_hier();
  h->device = device; h->stream = (cudaStream_t)cuda_stream; h->n_levels = n_levels;
  { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->n_sms = sms; }
  {
    // the stream-ordered pool keeps up to 16 GB of freed operator memory for the next hierarchy instead of returning it to
    // the driver at every synchronisation (see dev_alloc)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool != nullptr) {
      uint64_t cur = 0, want = 16ull << 30;
      if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < want)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
    }
    cudaGetLastError();
  }
  cudaError_t e2 = cudaMallocHost(&h->h_nactive, sizeof(int));
  if (e2 != cudaSuccess) { delete h; return fail((int)e2, "cudaMallocHost failed"); }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->rng_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->rng_stream_lo, cudaStreamNonBlocking, lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->rng_order, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->rng_done, cudaEventDisableTiming) != cudaSuccess) {
    delete h; return fail(-3, "dmlmc: cannot create the probe-stream CUDA stream");
  }
  *out = h;
  return 0;
}

int dmlmc_hier_destroy(dmlmc_hier* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  invalidate_graphs(h);
  for (void* p : h->owned) dev_free(h, p);
  if (h->h_nactive) cudaFreeHost(h->h_nactive);
  if (h->rng_stream) { cudaStreamSynchronize(h->rng_stream); cudaStreamDestroy(h->rng_stream); }
  if (h->rng_stream_lo) { cudaStreamSynchronize(h->rng_stream_lo); cudaStreamDestroy(h->rng_stream_lo); }
  if (h->rng_order) cudaEventDestroy(h->rng_order);
  dev_free(h, h->mt_tab);
  dev_free(h, h->mt_state_out);
  cudaStreamSynchronize(h->stream);
  if (h->rng_done) cudaEventDestroy(h->rng_done);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
  return 0;
}

int dmlmc_set_stencil(dmlmc_hier* h, int level, int LX, int LT, const double* links_host, double diag_re, double diag_im) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_stencil: bad handle/level");
  CHECK(LX >= 2 && LT >= 2 && links_host, "set_stencil: bad lattice");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  L.kind = 0; L.LX = LX; L.LT = LT; L.n = 2 * LX * LT;
  const size_t V = (size_t)LX * LT;
  RET(upload_cx(h, links_host, V, &L.d.Ut, &L.f.Ut));
  RET(upload_cx(h, links_host + 2 * V, V, &L.d.Ux, &L.f.Ux));
  L.d.diag = cx<double>(diag_re, diag_im); L.f.diag = cx<float>((float)diag_re, (float)diag_im);
  {
    // pre-splatted, pre-conjugated links of the packed-FP32 kernel (stencil_step_bf16_kernel)
    std::vector<float4> l4(4 * V);
    const double* ut = links_host; const double* ux = links_host + 2 * V;
    for (int x = 0; x < LX; ++x)
      for (int t = 0; t < LT; ++t) {
        const size_t s = (size_t)x * LT + t;
        const size_t stm = (size_t)x * LT + (t == 0 ? LT - 1 : t - 1), sxm = (size_t)(x == 0 ? LX - 1 : x - 1) * LT + t;
        auto splat = [](double re, double im) { return make_float4((float)re, (float)re, (float)im, (float)im); };
        l4[s] = splat(ut[2 * s], ut[2 * s + 1]);
        l4[V + s] = splat(ut[2 * stm], -ut[2 * stm + 1]);
        l4[2 * V + s] = splat(ux[2 * s], ux[2 * s + 1]);
        l4[3 * V + s] = splat(ux[2 * sxm], -ux[2 * sxm + 1]);
      }
    RET(upload<float4>(h, l4.data(), 4 * V, &L.links4));
  }
  return 0;
}

int dmlmc_set_bsr(dmlmc_hier* h, int level, int n, int bs, int bpr, const int32_t* colidx_host, const double* vals_host) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_bsr: bad handle/level");
  CHECK(bs == 1 || bs == 2 || bs == 4 || bs == 8, "set_bsr: block size must be 1, 2, 4 or 8");
  CHECK(n > 0 && n % bs == 0 && bpr >= 1 && colidx_host && vals_host, "set_bsr: bad arguments");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  L.kind = 1; L.n = n; L.bs = bs; L.bpr = bpr; L.nb = n / bs;
  for (size_t i = 0; i < (size_t)L.nb * bpr; ++i) CHECK(colidx_host[i] >= -1 && colidx_host[i] < L.nb, "set_bsr: block column out of range");
  RET(upload<int>(h, colidx_host, (size_t)L.nb * bpr, &L.bsr_col));
  RET(upload_cx(h, vals_host, (size_t)L.nb * bpr * bs * bs, &L.d.bsr_vals, &L.f.bsr_vals));
  {
    const size_t cnt = (size_t)L.nb * bpr * bs * bs;
    std::vector<float4> v4(cnt);
    for (size_t i = 0; i < cnt; ++i) {
      const float mr = (float)vals_host[2 * i], mi = (float)vals_host[2 * i + 1];
      v4[i] = make_float4(mr, mr, mi, mi);
    }
    RET(upload<float4>(h, v4.data(), cnt, &L.bsr_vals4));
  }
  return 0;
}

int dmlmc_set_transfer(dmlmc_hier* h, int level, int n_f, int aggr_size, int dofi, int nvec, const double* pvals_host) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels - 1, "set_transfer: bad handle/level");
  CHECK(nvec == 1 || nvec == 2 || nvec == 4 || nvec == 8, "set_transfer: nvec must be 1, 2, 4 or 8");
  CHECK(n_f > 0 && aggr_size > 0 && dofi >= 2 && dofi % 2 == 0 && aggr_size % dofi == 0 && n_f % aggr_size == 0 && pvals_host,
        "set_transfer: inconsistent aggregation");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.n == 0 || L.n == n_f, "set_transfer: n_f does not match the level's operator");
  L.n = n_f; L.has_transfer = true; L.aggr = aggr_size; L.dofi = dofi; L.nvec = nvec; L.n_c = (n_f / aggr_size) * 2 * nvec;
  RET(upload_cx(h, pvals_host, (size_t)n_f * nvec, &L.d.pv, &L.f.pv));
  return 0;
}

int dmlmc_set_transfer_indexed(dmlmc_hier* h, int level, int n_f, int n_blocks, int nvec, const double* pvals_host,
                               const int32_t* cblk_host) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels - 1, "set_transfer_indexed: bad handle/level");
  CHECK(nvec == 1 || nvec == 2 || nvec == 4 || nvec == 8, "set_transfer_indexed: nvec must be 1, 2, 4 or 8");
  CHECK(n_f > 0 && n_blocks > 0 && n_f % n_blocks == 0 && pvals_host && cblk_host, "set_transfer_indexed: inconsistent aggregation");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.n == 0 || L.n == n_f, "set_transfer_indexed: n_f does not match the level's operator");
  // every coarse block must own exactly m = n_f / n_blocks fine rows; rows[] lists them in increasing order
  const int m = n_f / n_blocks;
  std::vector<int> fill(n_blocks, 0), rows((size_t)n_f);
  for (int r = 0; r < n_f; ++r) {
    const int g = cblk_host[r];
    CHECK(g >= 0 && g < n_blocks && fill[g] < m, "set_transfer_indexed: coarse blocks must all have n_f / n_blocks rows");
    rows[(size_t)g * m + fill[g]++] = r;
  }
  L.n = n_f; L.has_transfer = true; L.aggr = 0; L.dofi = 2; L.nvec = nvec; L.n_c = n_blocks * nvec; L.tr_m = m;
  RET(upload<int>(h, rows.data(), rows.size(), &L.tr_rows));
  RET(upload<int>(h, cblk_host, (size_t)n_f, &L.tr_cblk));
  RET(upload_cx(h, pvals_host, (size_t)n_f * nvec, &L.d.pv, &L.f.pv));
  return 0;
}

int dmlmc_set_preconditioner(dmlmc_hier* h, int level, dmlmc_hier* hp, int level_p) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_preconditioner: bad handle/level");
  if (hp == nullptr) { h->prec_hier[level] = nullptr; return 0; }
  CHECK(hp != h && level_p >= 0 && level_p < hp->n_levels, "set_preconditioner: bad preconditioner hierarchy/level");
  CHECK(hp->device == h->device, "set_preconditioner: both hierarchies must live on the same device");
  CHECK(hp->lv[level_p].n == h->lv[level].n && h->lv[level].n > 0, "set_preconditioner: level sizes differ");
  h->prec_hier[level] = hp; h->prec_level[level] = level_p;
  return 0;
}

int dmlmc_set_dense_inverse(dmlmc_hier* h, int level, int n, const double* minv_host) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels && n > 0 && minv_host, "set_dense_inverse: bad arguments");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.n == 0 || L.n == n, "set_dense_inverse: size does not match the level");
  L.n = n;
  RET(upload_cx(h, minv_host, (size_t)n * n, &L.minv_d, &L.minv_f));
  {
    const size_t cnt = (size_t)n * n;
    std::vector<float4> v4(cnt);
    for (size_t i = 0; i < cnt; ++i) {
      const float mr = (float)minv_host[2 * i], mi = (float)minv_host[2 * i + 1];
      v4[i] = make_float4(mr, mr, mi, mi);
    }
    RET(upload<float4>(h, v4.data(), cnt, &L.minv4));
  }
  L.has_dense = true;
  if (n >= 256) RET(build_umma_operand(h, level, L.minv_d));
  if (n >= 1024 && n <= 4096) RET(build_umma_split_operand(h, level, L.minv_d));
  return 0;
}

int dmlmc_set_dense_inverse_device(dmlmc_hier* h, int level, int n, const void* minv_dev) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels && n > 0 && minv_dev, "set_dense_inverse_device: bad arguments");
  CHECK(n % 8 == 0, "set_dense_inverse_device: n must be a multiple of 8");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.n == 0 || L.n == n, "set_dense_inverse_device: size does not match the level");
  L.n = n;
  RET(build_umma_operand(h, level, (const Cx<double>*)minv_dev));
  L.has_dense = true;
  return 0;
}

int dmlmc_prolongator_values(dmlmc_hier* h, const void* eig_vecs_dev, int ld, int n, int aggr_size, int dofi, int nvec, void* pvals_dev) {
  CHECK(h != nullptr, "NULL handle"); CU(cudaSetDevice(h->device));
  CHECK(eig_vecs_dev && pvals_dev && n > 0 && nvec >= 1 && nvec <= 16 && ld >= nvec, "prolongator_values: bad arguments");
  CHECK(dofi >= 2 && dofi % 2 == 0 && aggr_size % dofi == 0 && n % aggr_size == 0, "prolongator_values: inconsistent dof / aggregate size");
  const int mrows = aggr_size / 2;
  const size_t per_warp = (size_t)mrows * nvec * sizeof(double2);
  CHECK(per_warp <= 48 * 1024, "prolongator_values: aggregate too large for one warp's shared memory");
  int wpb = 4; while (wpb > 1 && wpb * per_warp > 48 * 1024) wpb /= 2;
  const long long nitems = 2ll * (n / aggr_size);
  prolongator_values_kernel<<<(unsigned)((nitems + wpb - 1) / wpb), 32 * wpb, wpb * per_warp, h->stream>>>(
      (const Cx<double>*)eig_vecs_dev, ld, n, aggr_size, dofi, nvec, (Cx<double>*)pvals_dev, nullptr, 1);
  LAUNCH_CHECK(h);
  return 0;
}

int dmlmc_block_orthonormal_values(dmlmc_hier* h, const void* vecs_dev, int ld, int n, int m, int nvec, const int32_t* rows_dev,
                                   int passes, void* pvals_dev) {
  CHECK(h != nullptr, "NULL handle"); CU(cudaSetDevice(h->device));
  CHECK(vecs_dev && pvals_dev && rows_dev && n > 0 && nvec >= 1 && nvec <= 16 && ld >= nvec, "block_orthonormal_values: bad arguments");
  CHECK(m >= nvec && n % m == 0 && (passes == 1 || passes == 2), "block_orthonormal_values: blocks of m >= nvec rows, n a multiple of m");
  const size_t per_warp = (size_t)m * nvec * sizeof(double2);
  CHECK(per_warp <= 48 * 1024, "block_orthonormal_values: block too large for one warp's shared memory");
  int wpb = 4; while (wpb > 1 && wpb * per_warp > 48 * 1024) wpb /= 2;
  const long long nitems = n / m;
  prolongator_values_kernel<<<(unsigned)((nitems + wpb - 1) / wpb), 32 * wpb, wpb * per_warp, h->stream>>>(
      (const Cx<double>*)vecs_dev, ld, n, 2 * m, 2, nvec, (Cx<double>*)pvals_dev, rows_dev, passes);
  LAUNCH_CHECK(h);
  return 0;
}

int dmlmc_galerkin(dmlmc_hier* h, int level, int cap, int32_t* col_dev, void* vals_dev, int* slots_host) {
  CHECK(h && level >= 0 && level < h->n_levels - 1, "galerkin: bad handle/level");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.kind >= 0 && L.has_transfer, "galerkin: the level needs its operator and its transfer operator first");
  CHECK(cap >= 1 && cap <= 1024 && col_dev && vals_dev && slots_host, "galerkin: bad arguments");
  GalOp A; A.kind = L.kind; A.LX = L.LX; A.LT = L.LT; A.Ut = L.d.Ut; A.Ux = L.d.Ux; A.diag = L.d.diag;
  A.bs = L.bs; A.bpr = L.bpr; A.col = L.bsr_col; A.vals = L.d.bsr_vals;
  GalTr T; T.aggr = L.aggr; T.dofi = L.dofi; T.nvec = L.nvec; T.rows = L.tr_rows; T.cblk = L.tr_cblk; T.pv = L.d.pv;
  T.m = L.tr_rows ? L.tr_m : L.aggr / 2;
  const int nbc = L.n_c / L.nvec;
  int* maxw = nullptr;
  CU(dev_alloc(h, &maxw, sizeof(int)));
  CU(cudaMemsetAsync(maxw, 0, sizeof(int), h->stream));
  CU(cudaMemsetAsync(vals_dev, 0, (size_t)nbc * cap * L.nvec * L.nvec * sizeof(Cx<double>), h->stream));
  galerkin_kernel<<<nbc, 256, cap * sizeof(int), h->stream>>>(A, T, cap, col_dev, (Cx<double>*)vals_dev, maxw);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(slots_host, maxw, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dev_free(h, maxw);
  if (e != cudaSuccess) return fail((int)e, std::string("galerkin: ") + cudaGetErrorString(e));
  return 0;
}

int dmlmc_set_bsr_device(dmlmc_hier* h, int level, int n, int bs, int bpr, const int32_t* colidx_dev, const void* vals_dev) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_bsr_device: bad handle/level");
  CHECK(bs == 1 || bs == 2 || bs == 4 || bs == 8, "set_bsr_device: block size must be 1, 2, 4 or 8");
  CHECK(n > 0 && n % bs == 0 && bpr >= 1 && colidx_dev && vals_dev, "set_bsr_device: bad arguments");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  L.kind = 1; L.n = n; L.bs = bs; L.bpr = bpr; L.nb = n / bs;
  const size_t nc = (size_t)L.nb * bpr, cnt = nc * bs * bs;
  int* col = nullptr; Cx<double>* vd = nullptr; Cx<float>* vf = nullptr; float4* v4 = nullptr;
  CU(dev_alloc(h, &col, nc * sizeof(int))); h->owned.push_back(col);
  CU(dev_alloc(h, &vd, cnt * sizeof(Cx<double>))); h->owned.push_back(vd);
  CU(dev_alloc(h, &vf, cnt * sizeof(Cx<float>))); h->owned.push_back(vf);
  CU(dev_alloc(h, &v4, cnt * sizeof(float4))); h->owned.push_back(v4);
  CU(cudaMemcpyAsync(col, colidx_dev, nc * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(vd, vals_dev, cnt * sizeof(Cx<double>), cudaMemcpyDeviceToDevice, h->stream));
  dense_formats_kernel<<<nblocks(cnt, 256), 256, 0, h->stream>>>(vd, cnt, vf, v4); LAUNCH_CHECK(h);
  CU(cudaStreamSynchronize(h->stream));
  L.bsr_col = col; L.d.bsr_vals = vd; L.f.bsr_vals = vf; L.bsr_vals4 = v4;
  return 0;
}

int dmlmc_dense_inverse(dmlmc_hier* h, int n, void* m_dev) {
  CHECK(h != nullptr, "NULL handle"); CU(cudaSetDevice(h->device));
  CHECK(m_dev && n >= 1 && n <= GJ_MAX_N, "dense_inverse: n must be in [1, 8192]");
  char* aux = nullptr;                                                  // piv[n], info, rowk[n], colk[n]
  const size_t ints = ((size_t)(n + 1) * sizeof(int) + 15) & ~(size_t)15;
  CU(dev_alloc(h, &aux, ints + 2 * (size_t)n * sizeof(double2)));
  int* piv = reinterpret_cast<int*>(aux); int* info_d = piv + n;
  double2* rowk = reinterpret_cast<double2*>(aux + ints); double2* colk = rowk + n;
  cudaError_t e = cudaMemsetAsync(info_d, 0, sizeof(int), h->stream);
  Cx<double>* M = (Cx<double>*)m_dev;
  for (int k = 0; k < n && e == cudaSuccess; ++k) {
    gj_pivot_kernel<<<1, 1024, 0, h->stream>>>(M, n, k, piv, info_d, rowk, colk);
    gj_update_kernel<<<dim3((n + 127) / 128, n), 128, 0, h->stream>>>(M, n, k, rowk, colk);
    h->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) { gj_unscramble_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(M, n, piv); h->launches++; e = cudaGetLastError(); }
  int info = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&info, info_d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dev_free(h, aux);
  if (e != cudaSuccess) return fail((int)e, std::string("dense_inverse: ") + cudaGetErrorString(e));
  CHECK(info == 0, "dense_inverse: the matrix is singular (zero pivot)");
  return 0;
}

int dmlmc_set_dense_inverse_device_full(dmlmc_hier* h, int level, int n, const void* minv_dev) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels && n > 0 && minv_dev, "set_dense_inverse_device_full: bad arguments");
  CU(cudaSetDevice(h->device));
  Level& L = h->lv[level];
  CHECK(L.n == 0 || L.n == n, "set_dense_inverse_device_full: size does not match the level");
  L.n = n;
  const size_t cnt = (size_t)n * n;
  Cx<double>* md = nullptr; Cx<float>* mf = nullptr; float4* m4 = nullptr;
  CU(dev_alloc(h, &md, cnt * sizeof(Cx<double>))); h->owned.push_back(md);
  CU(dev_alloc(h, &mf, cnt * sizeof(Cx<float>))); h->owned.push_back(mf);
  CU(dev_alloc(h, &m4, cnt * sizeof(float4))); h->owned.push_back(m4);
  CU(cudaMemcpyAsync(md, minv_dev, cnt * sizeof(Cx<double>), cudaMemcpyDeviceToDevice, h->stream));
  dense_formats_kernel<<<nblocks(cnt, 256), 256, 0, h->stream>>>(md, cnt, mf, m4); LAUNCH_CHECK(h);
  L.minv_d = md; L.minv_f = mf; L.minv4 = m4;
  L.has_dense = true;
  if (n >= 256 && n % 8 == 0) RET(build_umma_operand(h, level, L.minv_d));
  if (n >= 1024 && n <= 4096 && n % 8 == 0) RET(build_umma_split_operand(h, level, L.minv_d));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int dmlmc_set_coarsest_inverse(dmlmc_hier* h, int n, const double* minv_host) {
  CHECK(h != nullptr, "NULL handle");
  return dmlmc_set_dense_inverse(h, h->n_levels - 1, n, minv_host);
}

int dmlmc_set_smoother(dmlmc_hier* h, int level, int nfactors, const double* nu_host, double p0_re, double p0_im) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_smoother: bad handle/level");
  CHECK(nfactors >= 0 && (nfactors == 0 || nu_host), "set_smoother: bad factors");
  Level& L = h->lv[level];
  L.nu.resize(nfactors);
  for (int i = 0; i < nfactors; ++i) L.nu[i] = cx<double>(nu_host[2 * i], nu_host[2 * i + 1]);
  L.p0 = cx<double>(p0_re, p0_im);
  L.has_smoother = true;
  return 0;
}

int dmlmc_set_smoother_eo(dmlmc_hier* h, int level, int nfactors, const double* nu_host, double p0_re, double p0_im) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_smoother_eo: bad handle/level");
  CHECK(nfactors >= 0 && (nfactors == 0 || nu_host), "set_smoother_eo: bad factors");
  Level& L = h->lv[level];
  CHECK(nfactors == 0 || L.kind == 0, "set_smoother_eo: the level's operator must be a Wilson stencil (dmlmc_set_stencil)");
  L.nu_eo.resize(nfactors);
  for (int i = 0; i < nfactors; ++i) L.nu_eo[i] = cx<double>(nu_host[2 * i], nu_host[2 * i + 1]);
  L.p0_eo = cx<double>(p0_re, p0_im);
  L.has_eo = nfactors > 0;
  return 0;
}

int dmlmc_set_smoother_storage(dmlmc_hier* h, int level, int allow16) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_smoother_storage: bad handle/level");
  h->lv[level].smoother16 = allow16 != 0;
  return 0;
}

int dmlmc_set_perm(dmlmc_hier* h, int level, int shift, int nnz_per_row, const int32_t* cols_host, const double* vals_host) {
  if (h) invalidate_graphs(h);
  CHECK(h && level >= 0 && level < h->n_levels, "set_perm: bad handle/level");
  Level& L = h->lv[level];
  CHECK(L.n > 0, "set_perm: set the level's operator first");
  CHECK(shift >= 0 && shift < L.n && nnz_per_row >= 0, "set_perm: bad shift");
  CU(cudaSetDevice(h->device));
  L.has_perm = true; L.shift = shift; L.perm_nnz = nnz_per_row;
  if (nnz_per_row > 0) {
    CHECK(cols_host && vals_host, "set_perm: missing Bblock_perm arrays");
    RET(upload<int>(h, cols_host, (size_t)L.n * nnz_per_row, &L.perm_cols));
    RET(upload_cx(h, vals_host, (size_t)L.n * nnz_per_row, &L.d.perm_vals, &L.f.perm_vals));
  }
  return 0;
}

int dmlmc_set_deflation(dmlmc_hier* h, int level, int d, const double* v_host) {
  CHECK(h && level >= 0 && level < h->n_levels, "set_deflation: bad handle/level");
  Level& L = h->lv[level];
  CHECK(L.n > 0, "set_deflation: set the level's operator first");
  CU(cudaSetDevice(h->device));
  L.defl_d = 0;
  if (d > 0) {
    CHECK(v_host, "set_deflation: missing vectors");
    RET(upload<Cx<double>>(h, reinterpret_cast<const Cx<double>*>(v_host), (size_t)L.n * d, &L.defl_V));
    L.defl_d = d;
  }
  return 0;
}

#define ENTER(h) do { CHECK((h) != nullptr, "NULL handle"); CU(cudaSetDevice((h)->device)); } while (0)
#define CHECK_LEVEL(h, level) CHECK((level) >= 0 && (level) < (h)->n_levels && (h)->lv[level].n > 0, "bad level")
#define CHECK_PREC(prec) CHECK((prec) == DMLMC_C128 || (prec) == DMLMC_C64, "bad precision")

int dmlmc_spmm(dmlmc_hier* h, int level, int prec, const void* X, void* Y, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK_PREC(prec); CHECK(X && Y && k >= 1, "spmm: bad arguments");
  if (level == h->n_levels - 1 && h->lv[level].kind < 0) return fail(-1, "spmm: the coarsest level has no sparse operator");
  if (prec == DMLMC_C128) return launch_op<double, M_AX>(h, level, X, nullptr, Y, ZERO, ZERO, k);
  return launch_op<float, M_AX>(h, level, X, nullptr, Y, ZERO, ZERO, k);
}
int dmlmc_restrict(dmlmc_hier* h, int level, int prec, const void* Xf, void* Xc, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK_PREC(prec); CHECK(Xf && Xc && k >= 1, "restrict: bad arguments");
  return prec == DMLMC_C128 ? launch_restrict<double>(h, level, Xf, Xc, k) : launch_restrict<float>(h, level, Xf, Xc, k);
}
int dmlmc_prolong_add(dmlmc_hier* h, int level, int prec, const void* Xc, void* Xf, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK_PREC(prec); CHECK(Xf && Xc && k >= 1, "prolong_add: bad arguments");
  return prec == DMLMC_C128 ? launch_prolong<double>(h, level, Xc, Xf, k) : launch_prolong<float>(h, level, Xc, Xf, k);
}
int dmlmc_coarsest_apply(dmlmc_hier* h, int prec, const void* B, void* X, int k) {
  ENTER(h); CHECK_PREC(prec); CHECK(B && X && k >= 1, "coarsest_apply: bad arguments");
  const int lc = h->n_levels - 1;
  return prec == DMLMC_C128 ? launch_dense<double>(h, lc, B, X, k) : launch_dense<float>(h, lc, B, X, k);
}
int dmlmc_smooth(dmlmc_hier* h, int level, int prec, const void* R, void* E, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK_PREC(prec); CHECK(R && E && R != E && k >= 1, "smooth: bad arguments");
  return prec == DMLMC_C128 ? smooth_chunked<double>(h, level, (const Cx<double>*)R, (Cx<double>*)E, k)
                            : smooth_chunked<float>(h, level, (const Cx<float>*)R, (Cx<float>*)E, k);
}
int dmlmc_vcycle(dmlmc_hier* h, int level, int prec, const void* B, void* X, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK_PREC(prec); CHECK(B && X && B != X && k >= 1, "vcycle: bad arguments");
  return prec == DMLMC_C128 ? vcycle<double, double>(h, level, B, X, k) : vcycle<float, float>(h, level, B, X, k);
}
int dmlmc_precondition(dmlmc_hier* h, int level, const void* V, void* Zout, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK(V && Zout && V != Zout && k >= 1, "precondition: bad arguments");
  return precond_eager(h, level, (const Z*)V, (Z*)Zout, k, nullptr);
}
int dmlmc_dotc(dmlmc_hier* h, const void* X, const void* Y, int n, int k, void* out_dev) {
  ENTER(h); CHECK(X && Y && out_dev && n >= 1 && k >= 1, "dotc: bad arguments");
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  Z* partial; RET(ws_get<Z>(h, partial_count(h, n, 1, k), &partial));
  int rc = multi_dot(h, (const Z*)X, 0, 1, (const Z*)Y, n, k, partial, (Z*)out_dev, 0);
  h->ws_off = mark;
  return rc;
}
int dmlmc_deflate(dmlmc_hier* h, int level, void* X, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK(X && k >= 1, "deflate: bad arguments");
  return deflate(h, level, (Z*)X, k);
}
int dmlmc_probe_expand(dmlmc_hier* h, const uint8_t* bits_dev, int n, int k, void* X0) {
  ENTER(h); CHECK(bits_dev && X0 && n >= 1 && k >= 1, "probe_expand: bad arguments");
  const size_t nk = (size_t)n * k;
  probe_expand_kernel<<<nblocks(nk, 256), 256, 0, h->stream>>>(bits_dev, n, k, (Cx<double>*)X0);
  LAUNCH_CHECK(h);
  return 0;
}
int dmlmc_set_mt_jump_table(dmlmc_hier* h, const uint32_t* tab_host, int rows) {
  ENTER(h); CHECK(tab_host && rows >= 11 && rows <= 64, "set_mt_jump_table: bad arguments");
  if (h->mt_tab) {
    CU(cudaStreamSynchronize(h->rng_stream)); CU(cudaStreamSynchronize(h->rng_stream_lo));      // a generator may still read it
    dev_free(h, h->mt_tab); h->mt_tab = nullptr;
  }
  CU(dev_alloc(h, &h->mt_tab, (size_t)rows * 624 * sizeof(uint32_t)));
  CU(cudaStreamSynchronize(h->stream));                 // (the copy below and the generator run on other streams)
  CU(cudaMemcpy(h->mt_tab, tab_host, (size_t)rows * 624 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (!h->mt_state_out) { CU(dev_alloc(h, &h->mt_state_out, 625 * sizeof(uint32_t))); CU(cudaStreamSynchronize(h->stream)); }
  h->mt_tab_rows = rows;
  return 0;
}
int dmlmc_mt19937_bits(dmlmc_hier* h, uint32_t* state_dev, uint32_t* backup_dev, long long skip_before, long long count,
                       long long skip_after, uint8_t* lsb_dev) {
  ENTER(h); CHECK(state_dev && skip_before >= 0 && count >= 0 && skip_after >= 0 && (count == 0 || lsb_dev), "mt19937_bits: bad arguments");
  // ordered after everything already queued on the solver stream (the buffers may still be in use there) and after the
  // previous generator call, then asynchronous beside the solver: the one-CTA sequential kernel on a high-priority stream
  // (it is the critical path of the next round), the jump-ahead kernel (~130 CTAs for ~2 ms) without priority over the solver
  const long long total = skip_before + count + skip_after;
  const bool jump = h->mt_jump && h->mt_tab != nullptr && total + 1024 < (1ll << (h->mt_tab_rows - 1));
  cudaStream_t rs = (jump && !h->mt_prio) ? h->rng_stream_lo : h->rng_stream;
  CU(cudaEventRecord(h->rng_order, h->stream));
  CU(cudaStreamWaitEvent(rs, h->rng_order, 0));
  CU(cudaStreamWaitEvent(rs, h->rng_done, 0));
  if (jump) {
    // one CTA per chunk of the wanted outputs + one for the end state; chunk = a power of two (few set bits in the jump
    // distances).  At most 32 chunks: the kernel shares the GPU with the solve of the previous round and is issue-bound like
    // the solver's own kernels, so what matters is its total work (one polynomial per set bit of every CTA's distance:
    // ~80 applications with 32 chunks against ~450 with 128; measured 2.5 ms -> 0.x ms lost per round), not its latency
    long long chunk = 1 << 14;
    while (chunk * 32 < count) chunk <<= 1;
    const int nchunk = (int)((count + chunk - 1) / chunk);
    mt19937_jump_bits_kernel<<<nchunk + 1, MTJ_THREADS, 0, rs>>>(state_dev, h->mt_state_out, backup_dev, h->mt_tab,
                                                                   skip_before, count, total, chunk, lsb_dev);
    LAUNCH_CHECK(h);
    CU(cudaMemcpyAsync(state_dev, h->mt_state_out, 625 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, rs));
  } else {
    mt19937_bits_kernel<<<1, 256, 0, rs>>>(state_dev, backup_dev, skip_before, count, skip_after, lsb_dev);
    LAUNCH_CHECK(h);
  }
  CU(cudaEventRecord(h->rng_done, rs));
  return 0;
}
int dmlmc_probe_expand_bytes(dmlmc_hier* h, const uint8_t* lsb_dev, int n, int k, void* X0) {
  ENTER(h); CHECK(lsb_dev && X0 && n >= 1 && k >= 1, "probe_expand_bytes: bad arguments");
  CU(cudaStreamWaitEvent(h->stream, h->rng_done, 0));          // the generator that filled lsb_dev
  probe_expand_bytes_kernel<<<dim3((k + 31) / 32, (n + 127) / 128), dim3(32, 8), 0, h->stream>>>(lsb_dev, n, k, (Cx<double>*)X0);
  LAUNCH_CHECK(h);
  return 0;
}
int dmlmc_rng_sync(dmlmc_hier* h) {
  ENTER(h);
  CU(cudaStreamSynchronize(h->rng_stream));
  CU(cudaStreamSynchronize(h->rng_stream_lo));
  return 0;
}
int dmlmc_hop_eo(dmlmc_hier* h, int level, int parity, const void* in_q, const void* in2, void* out_p,
                 double a_re, double a_im, double b_re, double b_im, int k, const void* xc, void* z) {
  ENTER(h); CHECK_LEVEL(h, level);
  Level& L = h->lv[level];
  CHECK(L.kind == 0 && L.links4 != nullptr && (L.LT % 2) == 0 && (L.LX % 2) == 0, "hop_eo: the level must be a Wilson stencil on an even lattice");
  CHECK(in_q && k >= 2 && (k % 2) == 0 && (parity == 0 || parity == 1), "hop_eo: bad arguments (k must be even)");
  CHECK((z == nullptr) == (xc == nullptr), "hop_eo: xc and z go together");
  const Cx<double> a = {a_re, a_im}, b = {b_re, b_im};
  const uint2* Q = (const uint2*)in_q; const uint2* I2 = (const uint2*)in2; uint2* O = (uint2*)out_p;
  const int kp = k / 2;
  if (in2 && out_p && !z)  return launch_hop_eo<true, true, false>(h, L, parity, Q, I2, O, a, b, kp, nullptr, nullptr);
  if (!in2 && out_p && !z) return launch_hop_eo<false, true, false>(h, L, parity, Q, nullptr, O, a, b, kp, nullptr, nullptr);
  if (in2 && out_p && z)   return launch_hop_eo<true, true, true>(h, L, parity, Q, I2, O, a, b, kp, xc, z);
  if (in2 && !out_p && z)  return launch_hop_eo<true, false, true>(h, L, parity, Q, I2, nullptr, a, b, kp, xc, z);
  return fail(-1, "hop_eo: this combination of inputs / outputs is not one the solver uses");
}
int dmlmc_apply_perm(dmlmc_hier* h, int level, const void* X, void* RHS, int k) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK(X && RHS && X != RHS && k >= 1, "apply_perm: bad arguments");
  return apply_perm(h, level, (const Z*)X, (Z*)RHS, k);
}

size_t dmlmc_workspace_bytes(dmlmc_hier* h, int level, int k, int restart) {
  if (!h || level < 0 || level >= h->n_levels || k < 1 || restart < 1) return 0;
  // level_sample: fine solve + (sequentially) coarse solve share the FGMRES arena; the coarse level may need MORE than the fine
  // one now that the fine stencil level solves on half-lattice complex64 vectors
  size_t b = 0;
  for (int l = level; l < h->n_levels - 1 && l <= level + 2; ++l)
    if (h->lv[l].n > 0 && h->lv[l].kind >= 0) b = std::max(b, fgmres_bytes(h, l, k, std::min(restart, std::max(1, h->lv[l].n))));
  const size_t nk = (size_t)h->lv[level].n * k * sizeof(Z);
  b += 8 * align_up(nk) + align_up(partial_count(h, h->lv[level].n, 64, k) * sizeof(Z));
  b += align_up((size_t)h->lv[level].n * k / 8 + 64);
  return b + (1 << 20);
}
int dmlmc_set_workspace(dmlmc_hier* h, void* ws_dev, size_t bytes) {
  if (h) invalidate_graphs(h);
  CHECK(h != nullptr, "NULL handle");
  h->ws = (char*)ws_dev; h->ws_bytes = bytes; h->ws_off = 0;
  return 0;
}
int dmlmc_set_inner_precision(dmlmc_hier* h, int prec) {
  if (h) invalidate_graphs(h);
  CHECK(h != nullptr, "NULL handle"); CHECK_PREC(prec);
  h->inner_prec = prec;
  return 0;
}
int dmlmc_fgmres(dmlmc_hier* h, int level, const void* B, void* X, int k, double tol, int restart, int maxiter,
                 int32_t* iters_host, double* relres_host) {
  ENTER(h); CHECK_LEVEL(h, level); CHECK(B && X && B != X, "fgmres: bad arguments");
  if (level == h->n_levels - 1) return fail(-1, "fgmres: the coarsest level is solved by dmlmc_coarsest_apply");
  if (h->lv[level].kind < 0) return fail(-1, "fgmres: operator of this level not set");
  h->unconverged = 0;
  return fgmres(h, level, (const Z*)B, (Z*)X, k, tol, restart, maxiter, iters_host, relres_host);
}
int dmlmc_level_sample(dmlmc_hier* h, int method, int level_f, int level_c, const void* X0, int k, double tol,
                       int restart, int maxiter, void* e_dev, int32_t* iters_host) {
  ENTER(h); CHECK(X0 && e_dev && k >= 1, "level_sample: bad arguments");
  h->unconverged = 0;
  return level_sample(h, method, level_f, level_c, (const Z*)X0, k, tol, restart, maxiter, (Z*)e_dev, iters_host);
}
int dmlmc_level_sample_host(dmlmc_hier* h, int method, int level_f, int level_c, const uint8_t* bits_host, int k,
                            double tol, int restart, int maxiter, double* e_host, int32_t* iters_host) {
  ENTER(h); CHECK(bits_host && e_host && k >= 1, "level_sample_host: bad arguments");
  CHECK(level_f >= 0 && level_f < h->n_levels && h->lv[level_f].n > 0, "level_sample_host: bad level");
  const int n = h->lv[level_f].n;
  h->unconverged = 0;
  const size_t nbits = (size_t)n * k, nbytes = (nbits + 7) / 8;
  WsScope ws_scope(h); const size_t mark = ws_scope.mark;
  uint8_t* bits_dev; Z* X0; Z* e_dev;
  RET(ws_get<uint8_t>(h, nbytes, &bits_dev));
  RET(ws_get<Z>(h, (size_t)n * k, &X0));
  RET(ws_get<Z>(h, (size_t)k, &e_dev));
  CU(cudaMemcpyAsync(bits_dev, bits_host, nbytes, cudaMemcpyHostToDevice, h->stream));
  probe_expand_kernel<<<nblocks(nbits, 256), 256, 0, h->stream>>>(bits_dev, n, k, X0); LAUNCH_CHECK(h);
  int rc = level_sample(h, method, level_f, level_c, X0, k, tol, restart, maxiter, e_dev, iters_host);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(e_host, e_dev, sizeof(Z) * k, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail((int)e, std::string("level_sample_host: ") + cudaGetErrorString(e));
  }
  h->ws_off = mark;
  return rc;
}
int dmlmc_set_option(dmlmc_hier* h, const char* name, double value) {
  if (h) invalidate_graphs(h);
  CHECK(h != nullptr && name != nullptr, "set_option: bad arguments");
  if (std::strcmp(name, "outer_eo") == 0) { h->outer_eo = value != 0.0; return 0; }
  if (std::strcmp(name, "outer_c64") == 0) { h->outer_c64 = value != 0.0; return 0; }
  if (std::strcmp(name, "outer_drop") == 0) { h->outer_drop = value; return 0; }
  if (std::strcmp(name, "mt_jump") == 0) { h->mt_jump = value != 0.0; return 0; }
  if (std::strcmp(name, "mt_prio") == 0) { h->mt_prio = value != 0.0; return 0; }
  if (std::strcmp(name, "gs_x2") == 0) { h->gs_x2 = value != 0.0; return 0; }
  if (std::strcmp(name, "gs_rows") == 0) {
    const int v = (int)value;
    CHECK(v == 0 || (v >= MIN_ROWS_PER_CHUNK && v <= 1024 && v % 8 == 0), "set_option: gs_rows must be 0 (automatic) or a multiple of 8 in [32, 1024]");
    h->gs_rows = v; return 0;
  }
  if (std::strcmp(name, "fuse_residual") == 0) { h->fuse_residual = value != 0.0; return 0; }
  if (std::strcmp(name, "hop_tma") == 0) { h->hop_tma = value != 0.0; return 0; }
  if (std::strcmp(name, "precond_smoother_only") == 0) { h->smoother_only = value != 0.0; return 0; }
  if (std::strcmp(name, "smoother_eo") == 0) { h->smoother_eo = value != 0.0; return 0; }
  if (std::strcmp(name, "dot32") == 0) { h->dot32 = value != 0.0; return 0; }
  if (std::strcmp(name, "fuse_res") == 0) { h->fuse_res = value != 0.0; return 0; }
  if (std::strcmp(name, "adaptive_poll") == 0) { h->adaptive_poll = value != 0.0; return 0; }
  if (std::strcmp(name, "fuse_io") == 0) { h->fuse_io = value != 0.0; return 0; }
  if (std::strcmp(name, "dense_split_bf16") == 0) { h->dense_split_bf16 = value != 0.0; return 0; }
  if (std::strcmp(name, "use_graphs") == 0) { h->use_graphs = value != 0.0; return 0; }
  if (std::strcmp(name, "graph_max_k") == 0) { h->graph_max_k = (int)value; return 0; }
  if (std::strcmp(name, "reorth") == 0) { h->reorth = value != 0.0 ? 1 : 0; return 0; }
  if (std::strcmp(name, "chunk_cols") == 0) { h->chunk_cols = (int)value; return 0; }
  if (std::strcmp(name, "l2_budget_mb") == 0) { CHECK(value >= 0, "l2_budget_mb must be >= 0"); h->l2_budget_mb = value; return 0; }
  if (std::strcmp(name, "stencil_by") == 0) { CHECK(value >= 1, "stencil_by must be >= 1"); h->stencil_by = (int)value; return 0; }
  if (std::strcmp(name, "defl_tensor") == 0) { h->defl_tensor = value != 0.0; return 0; }
  if (std::strcmp(name, "bsr_threads") == 0) { CHECK(value >= 32 && value <= 256, "bsr_threads must be in [32, 256]"); h->bsr_threads = (int)value; return 0; }
  if (std::strcmp(name, "stencil_t2") == 0) { h->stencil_t2 = value != 0.0; return 0; }
  if (std::strcmp(name, "eo_packs") == 0) { CHECK(value == 1 || value == 2, "eo_packs must be 1 or 2"); h->eo_packs = (int)value; return 0; }
  if (std::strcmp(name, "eo_by") == 0) { CHECK(value >= 1, "eo_by must be >= 1"); h->eo_by = (int)value; return 0; }
  if (std::strcmp(name, "eo_bz") == 0) { CHECK(value >= 1, "eo_bz must be >= 1"); h->eo_bz = (int)value; return 0; }
  if (std::strcmp(name, "stencil_t2_by") == 0) { CHECK(value >= 1, "stencil_t2_by must be >= 1"); h->stencil_t2_by = (int)value; return 0; }
  if (std::strcmp(name, "stencil_t2_bz") == 0) { CHECK(value >= 1, "stencil_t2_bz must be >= 1"); h->stencil_t2_bz = (int)value; return 0; }
  if (std::strcmp(name, "stencil_smem") == 0) { h->stencil_smem = value != 0.0; return 0; }
  if (std::strcmp(name, "stencil_fast") == 0) { h->stencil_fast = value != 0.0; return 0; }
  if (std::strcmp(name, "prefetch_slices") == 0) { h->prefetch_slices = (int)value; return 0; }
  if (std::strcmp(name, "pre_smooth") == 0) { h->pre_smooth = value != 0.0; return 0; }
  if (std::strcmp(name, "smoother_half") == 0) { h->smoother_half = value != 0.0; return 0; }
  if (std::strcmp(name, "dense_direct_exact") == 0) { h->dense_direct_exact = value != 0.0; return 0; }
  if (std::strcmp(name, "dense_tensor_min_n") == 0) { h->dense_tensor_min_n = (int)value; return 0; }
  if (std::strcmp(name, "stencil_minb") == 0) { CHECK(value == 2 || value == 3, "stencil_minb must be 2 or 3"); h->stencil_minb = (int)value; return 0; }
  if (std::strcmp(name, "stencil_bz") == 0) { CHECK(value >= 1, "stencil_bz must be >= 1"); h->stencil_bz = (int)value; return 0; }
  return fail(-1, std::string("dmlmc: unknown option ") + name);
}
long long dmlmc_launch_count(dmlmc_hier* h) { return h ? h->launches : 0; }
int dmlmc_unconverged_columns(dmlmc_hier* h) { return h ? h->unconverged : 0; }
int dmlmc_vcycle_chunk_cols(dmlmc_hier* h, int level, int prec, int k) {
  if (!h || level < 0 || level >= h->n_levels || k < 1) return 0;
  return chunk_cols(h, level, k, prec == DMLMC_C128 ? sizeof(Cx<double>) : sizeof(Cx<float>));
}

}  // extern "C"
