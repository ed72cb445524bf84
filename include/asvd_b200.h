/* asvd_b200.h — C ABI of the B200-native ASVD hot path.
 *
 * The upstream project (hahnyuan/ASVD4LLM) is pure Python and has no FFI; its boundary for this path is a
 * Python module contract.  This header is what a binding for that contract calls: plain
 * pointers and sizes, a cudaStream_t passed as void*, `int` status (0 = ok), no torch types, no exceptions.
 * All pointers are DEVICE pointers unless the name ends in `_host`.  The caller owns every buffer including
 * the workspace; the library keeps no global state besides a thread-local error string.
 *
 * Each entry point cites the upstream code it replaces (paths relative to the upstream tree).
 */
#ifndef ASVD_B200_H
#define ASVD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASVD_B200_VERSION 200

/* element types of caller tensors */
enum { ASVD_F32 = 0, ASVD_F16 = 1, ASVD_BF16 = 2 };
/* sigma_fuse modes — modules/svd_linear.py:16-24 */
enum { ASVD_FUSE_UV = 0, ASVD_FUSE_U = 1, ASVD_FUSE_V = 2 };
/* calibration statistic — act_aware_utils.py:64-74 */
enum { ASVD_STAT_ABS_MEAN = 0, ASVD_STAT_ABS_MAX = 1, ASVD_STAT_SQ_MEAN = 2 };
/* status codes */
enum {
  ASVD_OK = 0,
  ASVD_ERR_INVALID = 1,      /* bad argument */
  ASVD_ERR_WORKSPACE = 2,    /* workspace too small */
  ASVD_ERR_CUDA = 3,         /* a CUDA call failed; see asvd_last_error() */
  ASVD_ERR_NONFINITE = 4,    /* NaN/Inf in the input or in the factors (upstream: "nan in S/U/V", svd_linear.py:80-98) */
  ASVD_ERR_NOT_CONVERGED = 5 /* sweep limit hit above tolerance (factors are still written) */
};

int asvd_version(void);
/* thread-local description of the last non-zero status */
const char* asvd_last_error(void);

/* ---- measurement hooks (bench.py): kernel-launch counters and optional per-class CUDA-event timing.
 * Classes, in order: prep, gram, solve, update, finalize, extract, forward, absstat (8 entries).
 * asvd_profile_enable(1) resets the timers and makes every launch record an event pair (adds overhead:
 * never enabled inside a timed region); asvd_profile_read fills ms_out[8] / launches_out[8] (either may be
 * NULL) and returns the number of classes.  asvd_launch_count() = all launches since the library loaded. */
void asvd_profile_enable(int on);
int asvd_profile_read(double* ms_out, uint64_t* launches_out);
uint64_t asvd_launch_count(void);

/* ---- a2: rank formula — modules/svd_linear.py:39-44.  Host arithmetic, here so every binding agrees. */
int asvd_rank_for_ratio(int64_t out_features, int64_t in_features, double param_ratio, int rank_align);

/* ---- a3: scaling vector — modules/svd_linear.py:48-59.
 * scale[j] = rnd(rnd(sdm[j]^alpha) * rnd(fisher[j]^alpha)) + 1e-6, every step rounded to `stat_dtype`
 * exactly as the upstream in-place tensor expressions do; written as fp32.  sdm / fisher may be NULL. */
int asvd_scaling_vector(const void* sdm, const void* fisher, int stat_dtype, int n, double alpha,
                        float* scale_out, void* stream);

/* ---- a3+a4: activation-scaled SVD of `batch` same-shape weights — replaces modules/svd_linear.py:47-65
 * (w.float() * scale, torch.svd_lowrank) with an exact SVD (one-sided block Jacobi) of W*diag(scale).
 *
 *   W_host_ptrs[b]     device pointer (16-byte aligned) to weight b, row-major [m, n], leading dimension ldw
 *   scale_host_ptrs[b] device pointer to fp32 [n] from asvd_scaling_vector, or NULL entry / NULL array for
 *                      act_aware=False
 *   workspace          asvd_svd_workspace_bytes(m, n, batch) bytes, 256-byte aligned.  After the call it
 *                      holds the full factorisation of every weight (all min(m,n) triplets) and is the
 *                      handle asvd_svd_extract / asvd_svd_sigma read from.
 *   tol                rotation threshold on max |<x_p,x_q>| / (|x_p||x_q|) of a block pair (pairs below it are left
 *                      alone); a weight is finished after a sweep whose largest such cosine, measured before the
 *                      sweep's own rotations, was below 5 tol.  <=0 selects the default (4e-6)
 *   max_sweeps         <=0 selects the default (30)
 *   sweeps_out_host    optional host int[batch]: sweeps used (for the 2:1 and flatter-than-that rectangles that go
 *                      through the Gram pre-conditioner: sweeps of the square stage + sweeps of the main stage)
 * sigma and the second factor are recovered from the ORIGINAL weight after convergence (exact bf16-plane GEMM on the
 * tensor cores for 16-bit weights, fp32 SIMT GEMM otherwise), so they carry no accumulated rotation error.
 * The inner eigen-solve of a block pair is the triangular kernel pair (solve_tri_g_kernel: upper triangle of the Gram
 * matrix, rotation parameters on a warp of their own, three pairs per SM; solve_tri_r_kernel: replay of the rotation
 * record on R).  A weight's factors are bitwise reproducible run to run; they are bitwise independent of its batch-mates
 * whenever the batches compared cut the Gram pass into the same column chunks (always from batch * pairs >= SMs on);
 * ASVD_B200_GRAM_CHUNKS=n pins the chunk count where that must hold at any batch size.
 * Environment switches for A/B runs (read at every call): ASVD_B200_SOLVE=tri|quad|oddeven|lean (tri is the default;
 * quad / oddeven: the second- / first-generation single-kernel solves; lean: the quad kernel split in a G-only kernel at
 * two CTAs per SM + replay), ASVD_B200_POLISH=NS, ASVD_B200_GRAMPRE=0, ASVD_B200_RECOVER=simt, ASVD_B200_PRESORT=0,
 * ASVD_B200_SIMT=1, ASVD_B200_TRACE=1, ASVD_B200_NEAR_PCT=p (share of a sweep's pair visits below 1e-2 after which a
 * matrix takes the accurate Gram pass; default 50), ASVD_B200_OVERLAP=1 (two half-batches on two internal streams; same
 * results bitwise, measured not faster).
 * Blocks the calling thread until the factorisation is complete on `stream` (it polls a convergence flag
 * once per sweep). */
size_t asvd_svd_workspace_bytes(int m, int n, int batch);
int asvd_scaled_svd(const void* const* W_host_ptrs, int w_dtype, int64_t ldw, int m, int n, int batch,
                    const float* const* scale_host_ptrs, void* workspace, size_t workspace_bytes,
                    float tol, int max_sweeps, int* sweeps_out_host, void* stream);

/* all min(m,n) singular values of weight b of a finished workspace, descending, fp32 (device buffer) */
int asvd_svd_sigma(const void* workspace, int m, int n, int batch, int b, float* sigma_out, void* stream);

/* ---- a5+a6: rank-r truncation, un-scaling, sigma fusion, cast — replaces modules/svd_linear.py:69-70,
 * 8-24 and :102.  Re-slices a finished workspace, so the six ratios of the sensitivity sweep
 * (sensitivity.py:39-52) cost one SVD.
 *   A_out  [m, r] row-major, lda elements  (ALinear.weight)
 *   B_out  [r, n] row-major, ldb elements  (BLinear.weight)
 * (m, n, batch) must repeat the values given to asvd_scaled_svd: the workspace layout is a pure function
 * of them.  r is clamped by the caller to min(m, n) (upstream quirk: a larger request yields min(m,n)). */
int asvd_svd_extract(const void* workspace, int m, int n, int batch, int b, int r, int sigma_fuse, int out_dtype,
                     void* A_out, int64_t lda, void* B_out, int64_t ldb, void* stream);

/* ---- a7: SVDLinear.forward — modules/svd_linear.py:105-109 and ASVDLinear.forward
 * (huggingface_repos/modeling_asvd_llama.py:11-12).  y = (x B^T) A^T + bias.
 *   x [M, n] ldx;  B [r, n] ldb;  A [m, r] lda;  bias [m] or NULL;  y [M, m] ldy;  all `dtype`
 *   scratch: asvd_lowrank_forward_scratch_bytes(M, r, m) bytes, 256-byte aligned: the [M, round_up(r, 64)] intermediate
 *            (padded pitch: every rank runs on the tensor cores) plus room for a padded copy of A, made only when A's own
 *            rows are not 16-byte aligned (contiguous [m, r] with r % 8 != 0; pass lda = round_up(r, 64), or any multiple of 8, to avoid it).
 * F16 / BF16: tcgen05 GEMMs (CTA pairs, cta_group::2); F32 modules and x / B / y whose rows are not 16-byte aligned
 * (in / out features not a multiple of 8) run on the fp32 SIMT kernel.  ASVD_B200_FWD=1cta selects the single-CTA
 * multicast kernel (A/B runs). */
size_t asvd_lowrank_forward_scratch_bytes(int64_t M, int r, int m);
int asvd_lowrank_forward(const void* x, int64_t ldx, int64_t M, int n, const void* B, int64_t ldb, int r,
                         const void* A, int64_t lda, int m, const void* bias, void* y, int64_t ldy, int dtype,
                         void* scratch, size_t scratch_bytes, void* stream);

/* ---- a1: calibration statistic of one hook call — act_aware_utils.py:64-74.
 * x [L, n] ldx in `dtype`; acc [n] in the same dtype (upstream accumulates in the activation dtype).
 *   ABS_MEAN: acc[j] = rnd(acc[j] + rnd(mean_i |x[i,j]|));  ABS_MAX: acc[j] = max(acc[j], max_i |x[i,j]|)
 *   SQ_MEAN:  acc[j] = rnd(acc[j] + rnd(mean_i rnd(x[i,j]^2))) — the Fisher statistic of calib_fisher_info,
 *             act_aware_utils.py:31 (`fisher_info += weight.grad.pow(2).mean(0)`), x = the [m, n] weight gradient
 * scratch: asvd_absstat_scratch_bytes(n) bytes. */
size_t asvd_absstat_scratch_bytes(int n);
int asvd_absstat_accum(const void* x, int64_t ldx, int64_t L, int n, int dtype, int mode, void* acc,
                       void* scratch, size_t scratch_bytes, void* stream);

/* ---- a1 fused into its producer: one calibration step of one nn.Linear -- the layer's own forward
 * y = x W^T + bias AND the hook's statistic of its input x (act_aware_utils.py:64-74) in the same tcgen05 GEMM kernel: a
 * spare warp of every CTA reduces |x| over the rows while the tensor pipe works, so the activation is not read a second
 * time from HBM and the hook's separate launches disappear.
 *   x [M, n] ldx;  W [m, n] ldw (nn.Linear.weight);  bias [m] or NULL;  y [M, m] ldy;  F16 / BF16 only, rows 16-byte aligned
 *   mode ABS_MEAN / ABS_MAX;  acc [n] in the activation dtype, updated exactly as asvd_absstat_accum does (the fp32 column
 *   sums are combined with atomics, so the last bit of an ABS_MEAN update is not run-to-run deterministic)
 *   scratch: 4 n bytes. */
int asvd_linear_forward_stat(const void* x, int64_t ldx, int64_t M, int n, const void* W, int64_t ldw, int m,
                             const void* bias, void* y, int64_t ldy, int dtype, int mode, void* acc, void* scratch,
                             size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASVD_B200_H */
