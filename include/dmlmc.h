/*
 * dmlmc.h -- C ABI of libdmlmc_sm100.so: the sm_100a (B200) hot path of the deflated
 * multilevel-Monte-Carlo Hutchinson estimator for tr(D^-1) of the 2-D Schwinger
 * Wilson-Dirac operator.
 *
 * The reference (Gustavroot/DeflatedMLMC_Schwinger) is pure Python and has no FFI; the
 * boundary below is what its hot-path call sites bind to when they are moved to the GPU.
 * Each entry point cites the reference call site (file:line) it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for an argument / layout error, >0 for a
 *     CUDA runtime error code; dmlmc_last_error() gives the message (thread local).
 *   - "vectors" are column batches X[n][k] of k probes: row-major, k contiguous,
 *     complex interleaved (re,im).  prec selects the scalar: DMLMC_C128 (double) or
 *     DMLMC_C64 (float).
 *   - `*_dev` / "device pointer" arguments are caller-owned device memory (torch tensors);
 *     `*_host` arguments are host memory.  Everything runs asynchronously on the stream
 *     given at creation; the only host synchronisations are the documented convergence
 *     polls inside dmlmc_fgmres / dmlmc_level_sample* and the *_host entry points.
 *   - the library allocates device memory (stream-ordered pool, cudaMallocAsync) only for the operators it is handed at setup
 *     (dmlmc_set_*); solver work space is supplied by the caller (dmlmc_set_workspace).
 *   - one hierarchy handle is single-threaded; different handles are independent.
 *   - there is NO CPU fallback: without a CUDA device every call that touches data fails.
 */
#ifndef DMLMC_H
#define DMLMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMLMC_ABI_VERSION 6

enum { DMLMC_C128 = 0, DMLMC_C64 = 1 };

typedef struct dmlmc_hier dmlmc_hier;

int         dmlmc_abi_version(void);
const char* dmlmc_last_error(void);

/* ---- hierarchy: replaces the containers LevelML / SimpleML (multigrid.py:26-48) and the
 *      attributes MG.setup leaves behind (multigrid.py:339-344) ------------------------ */
int dmlmc_hier_create(int device, void* cuda_stream, int n_levels, dmlmc_hier** out);
int dmlmc_hier_destroy(dmlmc_hier* h);

/* level-0 operator A = S + m I in link form (matrix.py:14-31; stencil of SURVEY.md sec. 0):
 * links_host = [2][LX][LT] complex128 (U_t then U_x), diag = 4 + m. Row i = s*V + x*LT + t. */
int dmlmc_set_stencil(dmlmc_hier* h, int level, int LX, int LT, const double* links_host,
                      double diag_re, double diag_im);
/* coarse operator A_l = R A P (multigrid.py:276) as padded block-sparse rows:
 * colidx_host[n/bs][bpr] (block column, -1 = padding), vals_host[n/bs][bpr][bs][bs] complex128 */
int dmlmc_set_bsr(dmlmc_hier* h, int level, int n, int bs, int bpr, const int32_t* colidx_host,
                  const double* vals_host);
/* aggregation prolongator P_l (multigrid.py:192-262): row r of P has its nvec entries in
 * columns j*2*nvec + half*nvec + [0,nvec), j = r / aggr_size, half = ((r % aggr_size) % dofi) >= dofi/2;
 * pvals_host[n_f][nvec] complex128.  R_l = P_l^H (multigrid.py:267-274). */
int dmlmc_set_transfer(dmlmc_hier* h, int level, int n_f, int aggr_size, int dofi, int nvec,
                       const double* pvals_host);
/* the same kind of prolongator on arbitrary equal-sized aggregates: fine row r belongs to coarse block
 * cblk_host[r] in [0, n_blocks) and has its nvec entries in columns cblk_host[r]*nvec + [0,nvec).
 * Used by the geometric (4 x 4 sites, spin-split) hierarchy that preconditions the level-0 solve; the
 * estimator's own transfers (utils.py:299-303,337-341) stay the reference's, dmlmc_set_transfer. */
int dmlmc_set_transfer_indexed(dmlmc_hier* h, int level, int n_f, int n_blocks, int nvec,
                               const double* pvals_host, const int32_t* cblk_host);
/* precondition the FGMRES solves of `level` (multigrid.py:347-366, the M= argument of pyamg fgmres at :362)
 * with the V-cycle of another hierarchy `hp` started at its level `level_p` (same operator size, same device;
 * hp == NULL removes it).  hp runs on h's stream and in h's work space; it must outlive h's use of it. */
int dmlmc_set_preconditioner(dmlmc_hier* h, int level, dmlmc_hier* hp, int level_p);
/* dense inverse of the coarsest operator (multigrid.py:342-344), row-major n x n complex128 */
int dmlmc_set_coarsest_inverse(dmlmc_hier* h, int n, const double* minv_host);
/* optional dense inverse of an intermediate level: the V-cycle then bottoms out there (exact coarse
 * solve) instead of recursing to the coarsest level.  Same layout as dmlmc_set_coarsest_inverse. */
int dmlmc_set_dense_inverse(dmlmc_hier* h, int level, int n, const double* minv_host);
/* same from a DEVICE matrix (complex128 [n][n], e.g. the batched solver's answer to A X = I); only the
 * tensor-core operand is kept: BF16 [2n][2n], every complex entry as the real block [[re,-im],[im,re]],
 * applied by the tcgen05 kernel with FP32 accumulation.  Such a level serves the complex64 V-cycle only. */
int dmlmc_set_dense_inverse_device(dmlmc_hier* h, int level, int n, const void* minv_dev);
/* Set-up, multigrid.py:232-259: the values of the aggregation prolongator P_l from the test vectors, on the device.
 * eig_vecs_dev: complex128 [n][ld] row-major (the first nvec columns are the test vectors), pvals_dev: complex128 [n][nvec].
 * Every (aggregate, half) block -- aggr_size / 2 rows, closed-form row map of multigrid.py:203-227 -- is orthonormalised by
 * classical Gram-Schmidt in the reference's order of operations (one warp per block); agreement with the host builder
 * (which is bit-identical to the reference) ~1e-15, the summation order inside an inner product being the only difference. */
int dmlmc_prolongator_values(dmlmc_hier* h, const void* eig_vecs_dev, int ld, int n, int aggr_size, int dofi, int nvec,
                             void* pvals_dev);
/* The same orthonormalisation on INDEXED blocks (the geometric aggregates of the preconditioner hierarchies, where the
 * reference has nothing to compare with): block b owns the m rows rows_dev[b*m .. b*m+m) (n / m blocks, m >= nvec);
 * passes = 2 repeats the projection step once (classical Gram-Schmidt with re-orthogonalisation). */
int dmlmc_block_orthonormal_values(dmlmc_hier* h, const void* vecs_dev, int ld, int n, int m, int nvec, const int32_t* rows_dev,
                                   int passes, void* pvals_dev);
/* Set-up, multigrid.py:276: the Galerkin product A_{level+1} = R A P on the device, from the operator (dmlmc_set_stencil /
 * dmlmc_set_bsr[_device]) and the transfer operator (dmlmc_set_transfer[_indexed]) of `level`, written straight in the padded
 * block-sparse layout of dmlmc_set_bsr with `cap` slots per block row: col_dev[n_c/nvec][cap] int32 (sorted block columns,
 * -1 = padding), vals_dev[n_c/nvec][cap][nvec][nvec] complex128.  *slots_host = the number of slots the fullest row needs
 * (the bpr to keep); cap + 1 if a row did not fit (call again with a larger cap).  One thread sums each output number in a
 * fixed order: identical on every rank and launch; agreement with scipy's R*A*P ~1e-16 relative (summation order). */
int dmlmc_galerkin(dmlmc_hier* h, int level, int cap, int32_t* col_dev, void* vals_dev, int* slots_host);
/* dmlmc_set_bsr from DEVICE arrays (the output of dmlmc_galerkin, trimmed to bpr slots); the arrays are copied. */
int dmlmc_set_bsr_device(dmlmc_hier* h, int level, int n, int bs, int bpr, const int32_t* colidx_dev, const void* vals_dev);
/* Set-up, multigrid.py:342-344 (np.linalg.inv of the coarsest operator): in-place inverse of the dense complex128 device
 * matrix m_dev[n][n], n <= 8192, by Gauss-Jordan elimination with partial pivoting (two kernels per pivot, matrix in L2). */
int dmlmc_dense_inverse(dmlmc_hier* h, int n, void* m_dev);
/* The same hand-over for an inverse kept in ALL precisions (what dmlmc_set_dense_inverse makes from a host array: the
 * complex128 and complex64 copies, the splatted FP32 operand and the tensor-core operands), from a complex128 device
 * array [n][n] -- the inverse never visits the host (multigrid.py:342-344 of the reference inverts on the host). */
int dmlmc_set_dense_inverse_device_full(dmlmc_hier* h, int level, int n, const void* minv_dev);
/* smoother on `level`: e = p(A_level) r with the fixed polynomial p in product form,
 *   p(A) = p0 * prod_{i<nfactors} (I - nu_i A)       (nu_host: nfactors complex128, applied in order),
 * one fused operator+update kernel per factor, no inner products.  Replaces the lgmres call of
 * multigrid.py:393-394 / 438-439 (FGMRES is flexible: parity is on the converged solve). */
int dmlmc_set_smoother(dmlmc_hier* h, int level, int nfactors, const double* nu_host,
                       double p0_re, double p0_im);
/* the same polynomial form for the even-odd Schur complement S = c - H_eo H_oe / c of a stencil level (A = c I + H, H couples
 * sites of opposite parity): p(S) = p0 * prod_i (I - nu_i S).  When set (and option "smoother_eo" = 1, default) the complex64
 * V-cycle's post-smoother on that level is  x += [x_e ; x_o],  r^_e = r_e - H_eo r_o / c,  x_e = p(S) r^_e,
 * x_o = (r_o - H_oe x_e) / c  -- a polynomial of degree d in S does the work of one of degree 2d in A at the cost of d
 * operator applications (replaces lgmres, multigrid.py:393-394,438-439, like dmlmc_set_smoother). nfactors = 0 removes it. */
int dmlmc_set_smoother_eo(dmlmc_hier* h, int level, int nfactors, const double* nu_host, double p0_re, double p0_im);
/* allow16 = 0: this level's smoother keeps its intermediate vectors in FP32 even when option "smoother_half"
 * is on (the setup sets it when the polynomial is not stable enough for BF16 storage on that level) */
int dmlmc_set_smoother_storage(dmlmc_hier* h, int level, int allow16);
/* permutation data of a level (multigrid.py:142-155, 320-331): x_perm = roll(x, +shift),
 * then Bblock_perm (nnz_per_row == 0: identity) in padded row form cols/vals[n][nnz_per_row] */
int dmlmc_set_perm(dmlmc_hier* h, int level, int shift, int nnz_per_row,
                   const int32_t* cols_host, const double* vals_host);
/* deflation vectors of a level, V[n][d] row-major complex128 (utils.py:145-157); d = 0 clears */
int dmlmc_set_deflation(dmlmc_hier* h, int level, int d, const double* v_host);

/* ---- single operators (device pointers) -------------------------------------------- */
/* Y = A_level X          MG.matvec (multigrid.py:552-557), residual sites :388,402,433 */
int dmlmc_spmm(dmlmc_hier* h, int level, int prec, const void* X, void* Y, int k);
/* Xc = R_level Xf        multigrid.py:406; utils.py:301-303 */
int dmlmc_restrict(dmlmc_hier* h, int level, int prec, const void* Xf, void* Xc, int k);
/* Xf += P_level Xc       multigrid.py:429; utils.py:339-341 */
int dmlmc_prolong_add(dmlmc_hier* h, int level, int prec, const void* Xc, void* Xf, int k);
/* X = coarsest_inv B     multigrid.py:413-416; utils.py:309,321 */
int dmlmc_coarsest_apply(dmlmc_hier* h, int prec, const void* B, void* X, int k);
/* E = p(A_level) R       the smoother, multigrid.py:393-394 */
int dmlmc_smooth(dmlmc_hier* h, int level, int prec, const void* R, void* E, int k);
/* X = V-cycle(B) from `level` down to the coarsest   MG.one_mg_step (multigrid.py:369-447) */
int dmlmc_vcycle(dmlmc_hier* h, int level, int prec, const void* B, void* X, int k);
/* Z = M^{-1} V, complex128 in and out: exactly the preconditioner the level's FGMRES applies (the M= argument of pyamg
 * fgmres at multigrid.py:362, i.e. one_mg_step, :369-447) -- the V-cycle in the inner precision on this hierarchy, or on
 * the hierarchy attached with dmlmc_set_preconditioner. */
int dmlmc_precondition(dmlmc_hier* h, int level, const void* V, void* Z, int k);
/* out[c] = sum_r conj(X[r][c]) Y[r][c]   np.vdot (utils.py:249,336,353); out = k complex128 (device) */
int dmlmc_dotc(dmlmc_hier* h, const void* X, const void* Y, int n, int k, void* out_dev);
/* X -= V (V^H X) with the level's deflation vectors   utils.py:224,266 */
int dmlmc_deflate(dmlmc_hier* h, int level, void* X, int k);
/* Rademacher probes from packed bits (utils.py:213-216,255-258): element i of probe p is
 * 2*bit(p*n+i)-1 where bit(j) = (bits[j>>3] >> (j&7)) & 1 (numpy packbits, bitorder='little');
 * X0[n][k] complex128 */
int dmlmc_probe_expand(dmlmc_hier* h, const uint8_t* bits_dev, int n, int k, void* X0);
/* The reference's probe stream on the device: advance the legacy numpy generator (MT19937) whose state is
 * state_dev[625] = np.random.get_state() key[624] + position.  Skips skip_before 32-bit outputs, writes the
 * least significant bit of each of the next `count` outputs to lsb_dev[count] (one byte each -- element j of
 * np.random.randint(2, size=...) is exactly that bit), skips skip_after more.  backup_dev[625] (or NULL)
 * receives the state before the call.  Runs on an internal high-priority stream beside the solver;
 * dmlmc_probe_expand_bytes / dmlmc_rng_sync order after it. */
int dmlmc_mt19937_bits(dmlmc_hier* h, uint32_t* state_dev, uint32_t* backup_dev, long long skip_before,
                       long long count, long long skip_after, uint8_t* lsb_dev);
/* Jump-ahead table of that generator: tab_host[rows][624] = the polynomials t^(2^b) mod phi(t), b = 0..rows-1, packed 32
 * coefficients per word (phi: characteristic polynomial of the MT19937 transition; mtjump.py computes the table from the
 * generator's own output).  Once set, dmlmc_mt19937_bits runs one CTA per 2^14..-word chunk of the wanted outputs, each
 * jumping over everything before its chunk (skip_before included: the blocks of the other ranks, stoch_trace.py:288 has one
 * stream for all probes) instead of one CTA generating and discarding it; same outputs, same final state (option
 * "mt_jump" = 0 restores the sequential kernel). */
int dmlmc_set_mt_jump_table(dmlmc_hier* h, const uint32_t* tab_host, int rows);
/* X0[i][p] = 2*lsb[p*n+i] - 1, complex128 [n][k] (utils.py:213-216) */
int dmlmc_probe_expand_bytes(dmlmc_hier* h, const uint8_t* lsb_dev, int n, int k, void* X0);
/* wait for the probe-stream generator (before reading its state back to the host) */
int dmlmc_rng_sync(dmlmc_hier* h);
/* one half-lattice sweep of the even-odd smoother on a stencil level, as the V-cycle launches it (the operator inside the
 * lgmres smoother call of multigrid.py:393-394, restricted to one checkerboard colour):
 *   Out_p = a * In2_p + b * (H In_q),   H = A_level - diag,  q = 1 - parity
 * In_q / In2_p / Out_p: BF16 checkerboard half-lattice arrays [2 spins][LX][LT/2][k] of (re, im) BF16 pairs, site
 * t = 2 th + ((x + p) & 1); in2 == NULL: a = 0; out_p == NULL: no half-lattice output; z != NULL: additionally
 * Z[site][:] = Xc[site][:] + result at the sites of parity p (Xc complex64, Z complex128, full-lattice [n][k]).
 * k must be even.  The four combinations the solver uses are available: (in2, out_p, z) = (y,y,n), (n,y,n), (y,y,y), (y,n,y). */
int dmlmc_hop_eo(dmlmc_hier* h, int level, int parity, const void* in_q, const void* in2, void* out_p,
                 double a_re, double a_im, double b_re, double b_im, int k, const void* xc, void* z);
/* RHS = Bblock_perm_level * roll(X, +shift_level)   utils.py:232,288-290 (identity if no perm set) */
int dmlmc_apply_perm(dmlmc_hier* h, int level, const void* X, void* RHS, int k);

/* ---- solver ---------------------------------------------------------------------- */
/* bytes of caller-supplied device work space needed by dmlmc_fgmres / dmlmc_level_sample
 * for batches of k probes on `level` with restart length `restart` */
size_t dmlmc_workspace_bytes(dmlmc_hier* h, int level, int k, int restart);
int    dmlmc_set_workspace(dmlmc_hier* h, void* ws_dev, size_t bytes);
/* inner (V-cycle) precision: DMLMC_C64 (default) or DMLMC_C128; the outer FGMRES is always
 * complex128, so per-probe solves meet the complex128 tolerance either way */
int    dmlmc_set_inner_precision(dmlmc_hier* h, int prec);
/* batched right-preconditioned flexible GMRES, x0 = 0, one Krylov space per column, stops a
 * column when ||r|| < tol ||b||  (MG.solve, multigrid.py:347-366; pyamg.krylov.fgmres).
 * B, X: [n_level][k] complex128 device.  iters_host[k], relres_host[k] may be NULL. */
int dmlmc_fgmres(dmlmc_hier* h, int level, const void* B, void* X, int k, double tol,
                 int restart, int maxiter, int32_t* iters_host, double* relres_host);

/* one batch of k samples of utils.one_defl_Hutch_step (utils.py:207-361):
 *   method 0 ("hutchinson"): e = x0^H A_0^{-1} C x_def
 *   method 1 ("mlmc")      : e = x0^H (A_f^{-1} - P A_c^{-1} R) C_f x_def, level_c = level_f+1 or +2 (skip)
 * X0[n_f][k] complex128 device (probes); e_dev[k] complex128 device; iters_host[2*k] (fine, coarse) or NULL */
int dmlmc_level_sample(dmlmc_hier* h, int method, int level_f, int level_c, const void* X0, int k,
                       double tol, int restart, int maxiter, void* e_dev, int32_t* iters_host);
/* same with HOST buffers (the end-to-end call): packed probe bits in, estimates out;
 * host<->device copies happen inside */
int dmlmc_level_sample_host(dmlmc_hier* h, int method, int level_f, int level_c,
                            const uint8_t* bits_host, int k, double tol, int restart, int maxiter,
                            double* e_host, int32_t* iters_host);

/* solver options by name.  Unknown names are an error.
 *   "use_graphs", "graph_max_k"  1 / 1024 (defaults): for batches of at most graph_max_k columns every FGMRES iteration
 *                  (V-cycle, operator, Gram-Schmidt, Givens step: ~105 launches) is captured once per Krylov index into
 *                  a CUDA graph and replayed by all later solves (launch-bound for small batches, ~6 % at k = 256)
 *   "reorth"       1 = classical Gram-Schmidt with a re-orthogonalisation pass, 0 = single pass (default;
 *                  every column is verified against its true residual before it leaves the solve)
 *   "chunk_cols"   columns per V-cycle chunk; 0 (default) = derive from "l2_budget_mb"
 *   "l2_budget_mb" MB that the four working vectors of a chunk may occupy (so that they stay L2-resident
 *                  over the smoother's kernels); 0 (default) = one chunk
 *   "dense_tensor_min_n"  dense inverses with n >= this (default 1024) are applied on the tensor cores
 *                  inside the complex64 V-cycle (BF16 operands, FP32 accumulation); smaller ones by the FP32 kernel
 *   "prefetch_slices"  0 | 8 | 16 (default): the level-0 Y = A X / B - A X kernels prefetch into L2 the rows that the
 *                  thread blocks this many x-slices ahead will read (+3 % on the complex128 SpMM)
 *   "pre_smooth"   0 (default): the V-cycle is coarse-grid correction followed by the polynomial post-smoother;
 *                  1: polynomial pre- and post-smoothing as in multigrid.py:369-447 (measured: 13 outer iterations
 *                  at degree 64+64 against 14 at degree 0+96, which is 30 % less smoothing work)
 *   "smoother_half" 1 (default): inside the complex64 V-cycle the level-0 smoother keeps the intermediate
 *                  vectors of the polynomial product in BF16 (FP32 arithmetic); 0: FP32 storage
 *   "stencil_fast"  1 (default): packed-FP32 (FFMA2) kernel for those BF16-stored factors; 0: generic kernel
 *   "defl_tensor"  1 (default): the deflation projections V^H x and x - V(.) run on the FP64 tensor cores
 *                  (mma.sync m8n8k4.f64) when d is a multiple of 4 and <= 64; 0: SIMT kernels
 *   "dense_direct_exact"  1 (default): a V-cycle that STARTS on a dense level (that level's own solve) uses
 *                  the FP32 copy of the inverse when there is one; 0: tensor cores there as well
 *   "stencil_t2"   1 (default): the BF16 factor kernel computes two t-adjacent sites per thread (8 row loads per pair
 *                  instead of 10); "stencil_t2_by", "stencil_t2_bz": its thread-block tile (default 2 x 2)
 *   "stencil_by", "stencil_bz"   site tile (t, x) of the level-0 kernel's thread block (default 4 x 4)
 *   "stencil_minb" 2 | 3 (default): resident 512-thread blocks per SM the level-0 kernel is compiled for
 *   "dense_split_bf16"  1 (default): where a small level's (1024 <= n <= 4096) dense inverse is the preconditioner of that
 *                  level's own solve it is applied as ONE tcgen05 GEMM on split-BF16 operands ([hi|lo|hi] x [hi;hi;lo],
 *                  error ~1e-5) instead of the FP32 SIMT kernel (75 us against 305 us at n = 2048, k = 256)
 *   "fuse_io"      1 (default): the complex64 copy of a Krylov vector (the V-cycle's input) is written by the kernel that
 *                  normalises it, the last smoother factor writes complex128 straight into Z_j, the prolongation after the
 *                  coarse solve does not read the zero vector (bit-identical to 0)
 *   "fuse_res"     1 (default, needs fuse_io): the V-cycle's residual before the post-smoother is stored as BF16, so that
 *                  every factor but the last runs in the packed BF16 -> BF16 kernel
 *   "adaptive_poll" 1 (default): the per-iteration convergence poll (a host read) is skipped until one iteration before the
 *                  count the previous solve with the same (level, k, tol) needed, and done every 4th iteration at least
 *   "smoother_eo"  1 (default): use the even-odd post-smoother on a stencil level that has one (dmlmc_set_smoother_eo);
 *                  "eo_packs" 2 (default) | 1: column packs per thread of its kernel; "eo_by", "eo_bz": thread-block tile (2 x 2)
 *   "dot32"        0 (default; 1 measured harmful): Gram-Schmidt coefficients from complex64 copies of the basis
 *   "gs_x2"        1 (default): two-column (16-byte) Gram-Schmidt kernels for complex64 Krylov vectors
 *   "gs_rows"      rows per thread block of those kernels (fixed-order partial sums per chunk of rows); 0 (default) = the
 *                  largest power of two <= 256 that gives at least 2 048 thread blocks
 *   "fuse_residual" 1 (default): true residual of the Schur-complement system, its storage and its column norms in one kernel */
int dmlmc_set_option(dmlmc_hier* h, const char* name, double value);

/* columns that the solves of the last dmlmc_fgmres / dmlmc_level_sample[_host] call left above the tolerance when
 * maxiter was reached (the reference ignores pyamg's exit code at multigrid.py:362; here it can be checked) */
int dmlmc_unconverged_columns(dmlmc_hier* h);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
long long dmlmc_launch_count(dmlmc_hier* h);
/* columns per chunk the V-cycle uses on `level` for a batch of k columns of precision prec */
int dmlmc_vcycle_chunk_cols(dmlmc_hier* h, int level, int prec, int k);

#ifdef __cplusplus
}
#endif
#endif /* DMLMC_H */
