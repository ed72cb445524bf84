// Activation-scaled SVD by one-sided block Jacobi (Hestenes) — replaces modules/svd_linear.py:47-70 of the
// upstream tree (w.float() * scale -> torch.svd_lowrank -> V / scale) with an exact factorisation.
//
// Data layout (HBM): the n' = min(m,n) vectors to orthogonalise (columns of W*diag(s) when m >= n, rows
// otherwise) are the ROWS of X [nv_pad, len_pad] fp32; a block of JB=64 vectors is a contiguous slab and
// the long dimension is contiguous, so every pass below streams whole 128-byte lines.
//
// One round (nv_pad/JB - 1 rounds per sweep, nv_pad/(2*JB) disjoint block pairs per round, all pairs of all
// matrices of the batch in the same launch):
//   gram_kernel   : partial Gram matrices G_c = P_c P_c^T of each 128-vector panel over 512-column chunks
//   solve_kernel  : G = sum_c G_c in shared memory; one odd-even sweep of two-sided Jacobi on the 128x128 G,
//                   rotations accumulated in a register-resident R; sort by norm; Newton-Schulz polish of R
//   update_kernel : panel <- R^T panel  (in place, cp.async double-buffered column tiles)
// After convergence (max |cos| < tol at visit time over one whole sweep) the rows of X are sigma_j * u_j.
// The second factor is NOT accumulated: it is recovered from the ORIGINAL weight with one fp32 GEMM
// (Y = Xhat * W*diag(s)), which also anchors sigma_j = |Y_j| to the input and removes accumulated drift.
#include <stdarg.h>
#include <type_traits>
#include <float.h>
#include <vector>
#include <algorithm>
#include <atomic>
#include <utility>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.h"
#include "svd_tc.h"
#include "umma.cuh"
#include <stdlib.h>

namespace asvd {

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ profiling
static std::atomic<unsigned long long> g_launches[K_COUNT];
static bool g_prof_on = false;
static double g_prof_ms[K_COUNT];
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_ev[K_COUNT];
void prof_begin(int kind, cudaStream_t st) {
  g_launches[kind].fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, st);
  g_prof_ev[kind].push_back({a, b});
}
void prof_end(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof_ev[kind].back().second, st);
}
void prof_collect() {
  if (!g_prof_on) return;
  for (int k = 0; k < K_COUNT; ++k) {
    for (auto& e : g_prof_ev[k]) {
      float ms = 0.f;
      cudaEventSynchronize(e.second);
      cudaEventElapsedTime(&ms, e.first, e.second);
      g_prof_ms[k] += ms;
      cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    g_prof_ev[k].clear();
  }
}
void prof_reset(bool on) {
  prof_collect();
  g_prof_on = on;
  for (int k = 0; k < K_COUNT; ++k) g_prof_ms[k] = 0.0;
}
double prof_ms(int k) { return g_prof_ms[k]; }
unsigned long long launches(int k) { return g_launches[k].load(); }

// ------------------------------------------------------------------------------------------------ plan
constexpr int RECOVER_MAX_SPLITS = 16;     // K splits of the tensor-core recovery GEMM (runs of <= 1024 columns)
constexpr int RECOVER_RUN = 768;
SvdPlan make_plan(int m, int n, int batch, bool allow_inner) {
  SvdPlan p;
  memset(&p, 0, sizeof(p));
  p.m = m; p.n = n; p.batch = batch;
  p.tall = (m >= n);
  p.nv = p.tall ? n : m;
  p.len = p.tall ? m : n;
  p.nv_pad = (int)round_up(p.nv, JK);
  p.len_pad = (int)round_up(p.len, 128);
  p.ldy = (int)round_up(p.nv, 4);
  p.nb = p.nv_pad / JB;
  p.rounds = p.nb - 1;
  p.pairs = p.nb / 2;
  {
    // Gram work items = batch * pairs * chunks.  One wave of items that each reduce a long run of columns beats many
    // short ones: every item ends with a 64 KB partial Gram that the solve kernel has to read back and sum (measured
    // at batch 4 x 4096^2: 512-column chunks 260.7 ms, 1024-column chunks 259.5 ms).  So: just enough chunks to
    // give every SM an item, never narrower than 512 columns.
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
      cudaGetLastError();
    }
    int want = sms / (batch * p.pairs > 0 ? batch * p.pairs : 1);
    // The chunking decides the summation order of a Gram entry, so a weight's factors are bitwise independent of its
    // batch-mates exactly when the batch sizes compared cut the columns alike (always, from batch * pairs >= SMs on: one
    // chunk).  ASVD_B200_GRAM_CHUNKS=n pins the count for runs that must reproduce each other at any batch size.
    if (const char* ce = getenv("ASVD_B200_GRAM_CHUNKS")) { const int v = atoi(ce); if (v > 0) want = v; }
    const int max_chunks = (p.len_pad + 511) / 512;
    if (want < 1) want = 1;
    if (want > max_chunks) want = max_chunks;
    p.chunk_cols = (int)round_up((p.len_pad + want - 1) / want, 32);
    p.chunks = (p.len_pad + p.chunk_cols - 1) / p.chunk_cols;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (size_t)round_up((int64_t)(off + bytes), 256); return o; };
  p.off_ptrs = take(sizeof(void*) * 2 * (size_t)batch);
  p.off_pairs = take(sizeof(int2) * (size_t)p.rounds * p.pairs);
  p.off_X = take(sizeof(float) * (size_t)batch * p.nv_pad * p.len_pad);    // block-tiled (xt_off), the Jacobi working set
  p.off_Xr = take(sizeof(float) * (size_t)batch * p.nv_pad * p.len_pad);   // row-major copy made once after convergence
  p.off_Y = take(sizeof(float) * (size_t)batch * p.nv_pad * p.ldy);
  p.off_R = take(sizeof(float) * (size_t)batch * p.pairs * JK * JK);
  p.off_flag = take(sizeof(int) * (size_t)batch * p.pairs);
  p.off_maxoff = take(sizeof(unsigned) * 2 * (size_t)batch);   // [batch] max cosine bits, [batch] near-converged pair counts
  p.off_done = take(sizeof(int) * (size_t)batch);
  p.off_prec = take(sizeof(int) * (size_t)batch);      // Gram mode of every matrix (0: single TF32 pass, 1: fp32-accurate split)
  p.off_sigma = take(sizeof(float) * (size_t)batch * p.nv_pad);
  p.off_perm = take(sizeof(int) * (size_t)batch * p.nv_pad);
  p.off_status = take(sizeof(int) * (size_t)batch);
  p.off_scale = take(sizeof(float) * (size_t)batch * n);
  p.off_norm = take(sizeof(float) * (size_t)batch * p.nv_pad);
  p.off_track = take(sizeof(int) * (size_t)batch * (p.nb + (size_t)p.nb * p.nb));
  // scratch of the tensor-core recovery GEMM, one matrix at a time: bf16 planes [a1|a2|a3] of the normalised vectors
  // and [b1|b2] of the weight (transposed when the vectors are its columns)
  p.off_As = take(sizeof(__nv_bfloat16) * 3 * (size_t)p.nv_pad * p.len_pad);
  p.off_Bs = take(sizeof(__nv_bfloat16) * 2 * (size_t)p.nv_pad * p.len_pad);
  // Gram pre-conditioner (see gram_precondition): shapes at least 2:1 whose short side is worth it and whose long side
  // the plane GEMM can contract
  p.gram_pre = (allow_inner && p.len_pad >= 2 * p.nv_pad && p.nv_pad >= 1024 && p.len_pad <= RECOVER_MAX_SPLITS * 1024) ? 1 : 0;
  {
    size_t part = (size_t)RECOVER_MAX_SPLITS * p.nv_pad * p.ldy;                     // fp32 partial results of the K splits
    if (p.gram_pre) {
      const size_t inner_splits = (size_t)((p.nv_pad + RECOVER_RUN - 1) / RECOVER_RUN);
      const size_t a = inner_splits * p.nv_pad * p.len_pad, g = (size_t)RECOVER_MAX_SPLITS * p.nv_pad * p.nv_pad;
      if (a > part) part = a;
      if (g > part) part = g;
    }
    p.off_Yp = take(sizeof(float) * part);
  }
  if (p.gram_pre) {
    p.off_Gm = take(sizeof(float) * (size_t)batch * p.nv_pad * p.nv_pad);
    p.off_inner = take(make_plan(p.nv, p.nv, batch, false).bytes);
  }
  // per-pair record of the lean solve (ASVD_B200_SOLVE=lean): rotation history, scales, column order (svd_solve_quad.cuh)
  p.off_aux = take(sizeof(float) * (size_t)batch * p.pairs * QAUX_FLOATS_PLAN);
  // LAST: the only buffer whose size depends on the Gram chunking (hence, with ASVD_B200_GRAM_CHUNKS, on the environment
  // at call time); asvd_svd_sigma / asvd_svd_extract re-derive the plan later and must find everything else in place
  p.off_G = take(sizeof(float) * (size_t)batch * p.pairs * p.chunks * JK * JK);
  p.bytes = off;
  return p;
}

// ------------------------------------------------------------------------------------------------ prep (a3)
// X = fp32(W) * diag(s), transposed when the vectors are the columns of W (tall case).
template <typename T>
__global__ void __launch_bounds__(256) prep_kernel(const void* const* __restrict__ Wptrs, const float* __restrict__ scale,
                                                   int64_t ldw, int m, int n, int tall, float* __restrict__ X,
                                                   int64_t mat_stride, int ldx, const int* __restrict__ slot, int nv_pad) {
  // slot (optional): row of X that vector v goes to (the ascending-norm order of presort_*), per matrix [nv_pad]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int* sl = slot ? slot + (int64_t)b * nv_pad : nullptr;
  const T* W = reinterpret_cast<const T*>(Wptrs[b]);
  const float* s = scale + (int64_t)b * n;
  float* Xb = X + b * mat_stride;
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (tall) {
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, j = j0 + tx;
      tile[r][tx] = (i < m && j < n) ? to_f32<T>(W[(int64_t)i * ldw + j]) * s[j] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int j = j0 + r, i = i0 + tx;
      if (j < n && i < m) Xb[xt_off(sl ? sl[j] : j, i, ldx >> 5)] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, j = j0 + tx;
      if (i < m && j < n) Xb[xt_off(sl ? sl[i] : i, j, ldx >> 5)] = to_f32<T>(W[(int64_t)i * ldw + j]) * s[j];
    }
  }
}

// Sort keys for the initial order of the vectors: key = 1 / |vector|^2, so the descending sort_kernel below puts the
// vectors in ASCENDING norm order (zero padding keys sort last).  Which vector sits in which row of X is immaterial to
// the result; measured on the scaled Gaussian workload this order saves about one sweep in 15 (4096^2: 254.6 -> 246.2
// ms per batch of four; 11008x4096: 410.9 -> 399.1 ms).  Fixed summation order: results stay run-to-run bitwise equal.
template <typename T>
__global__ void __launch_bounds__(256) presort_key_kernel(const void* const* __restrict__ Wptrs, const float* __restrict__ scale,
                                                          int64_t ldw, int m, int n, int tall, float* __restrict__ key,
                                                          int nv_pad) {
  const int b = blockIdx.y;
  const T* W = reinterpret_cast<const T*>(Wptrs[b]);
  const float* s = scale + (int64_t)b * n;
  float* kb = key + (int64_t)b * nv_pad;
  __shared__ float part[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (tall) {                       // vector j = column j of W * s: 32 columns per CTA, rows split over the 8 warps
    const int j = blockIdx.x * 32 + tx;
    float ss = 0.f;
    if (j < n)
      for (int i = ty; i < m; i += 8) { const float x = to_f32<T>(W[(int64_t)i * ldw + j]); ss = fmaf(x, x, ss); }
    part[ty][tx] = ss;
    __syncthreads();
    if (ty == 0 && j < n) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][tx];
      t *= s[j] * s[j];
      kb[j] = t > 0.f ? 1.f / t : 0.f;
    }
  } else {                          // vector i = row i of W * diag(s): one warp per row, 8 rows per CTA
    const int i = blockIdx.x * 8 + ty;
    float ss = 0.f;
    if (i < m)
      for (int l = tx; l < n; l += 32) { const float x = to_f32<T>(W[(int64_t)i * ldw + l]) * s[l]; ss = fmaf(x, x, ss); }
    ss = warp_sum(ss);
    if (tx == 0 && i < m) kb[i] = ss > 0.f ? 1.f / ss : 0.f;
  }
}
// perm[r] = vector placed in row r  ->  slot[vector] = r
__global__ void presort_slot_kernel(const int* __restrict__ perm, int* __restrict__ slot, int nv_pad) {
  const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nv_pad) slot[(int64_t)b * nv_pad + perm[(int64_t)b * nv_pad + r]] = r;
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ Gram
// Gpart[b][p][c] = P P^T over columns [c*GRAM_CHUNK, ...) of the 128-vector panel of pair p.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ X, int64_t mat_stride, int ldx,
                                                   const int2* __restrict__ pairs, int len_pad, int chunks,
                                                   int pairs_per_mat, float* __restrict__ Gpart,
                                                   const int* __restrict__ done, const int* __restrict__ track, int nb,
                                                   int chunk_cols) {
  __shared__ __align__(16) float As[2][GK][GLD];
  const int b = blockIdx.z, p = blockIdx.y, c = blockIdx.x;
  if (done[b]) return;
  const int2 pr = pairs[p];
  if (pair_is_clean(track, nb, b, pr.x, pr.y)) return;
  const float* Xb = X + b * mat_stride;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int row = t >> 1, kq = (t & 1) * 8;
  const int vec = (row < JB) ? pr.x * JB + row : pr.y * JB + (row - JB);
  const int nct = ldx >> 5;
  const int kbeg = c * chunk_cols, kend = min(len_pad, kbeg + chunk_cols);
  const int nk = (kend - kbeg) / GK;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 r0, r1;
  auto gload = [&](int k0) {
    const float* src = Xb + xt_off(vec, k0 + kq, nct);
    r0 = *reinterpret_cast<const float4*>(src);
    r1 = *reinterpret_cast<const float4*>(src + 4);
  };
  auto sstore = [&](int buf) {
    As[buf][kq + 0][row] = r0.x; As[buf][kq + 1][row] = r0.y; As[buf][kq + 2][row] = r0.z; As[buf][kq + 3][row] = r0.w;
    As[buf][kq + 4][row] = r1.x; As[buf][kq + 5][row] = r1.y; As[buf][kq + 6][row] = r1.z; As[buf][kq + 7][row] = r1.w;
  };
  gload(kbeg);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kbeg + (kt + 1) * GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  float* G = Gpart + (((int64_t)b * pairs_per_mat + p) * chunks + c) * (JK * JK);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    *reinterpret_cast<float4*>(&G[r * JK + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(&G[r * JK + 64 + tx * 4]) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// ------------------------------------------------------------------------------------------------ solve
// One CTA per block pair.  G = sum of the partial Grams (upper triangle kept) lives in shared memory and is
// driven to diagonal form by ONE cyclic sweep of two-sided Jacobi in the odd-even ordering with a mandatory
// exchange after every rotation (pairs (2t,2t+1) on even steps, (2t+1,2t+2) on odd steps; after JK steps every
// pair has met once and the order is reversed).  Positions are therefore STATIC, which lets the accumulated
// rotation R live in registers: warps 0-3 hold one row of R per thread (128 registers) and apply the 64
// column rotations of a step with no data movement at all; warps 4-11 apply J^T G J to the upper triangle of
// G in shared memory (float2 accesses on even steps).  Afterwards the columns are ordered by descending norm
// (de Rijk's sorting, applied to the whole 128-column pair at once) and R is re-orthogonalised by one
// Newton-Schulz step, R <- R (1.5 I - 0.5 R^T R): fp32 rounding in the ~130 accumulated rotations per column
// leaves |R^T R - I| ~ 1e-5, which would otherwise drift the singular values (measured: 2e-4 -> 3e-6).
constexpr int SLD = JK + 4;   // shared leading dimension (float4-aligned rows, 4-bank skew)
constexpr int SOLVE_THREADS = 512;
constexpr int SOLVE_RTHREADS = 256;      // two threads per row of R (64 columns each)
constexpr int SOLVE_GTHREADS = SOLVE_THREADS - SOLVE_RTHREADS;
constexpr int NB_EVEN = (JK / 2) * (JK / 2 + 1) / 2;       // 2080 upper-triangle 2x2 blocks on even steps
constexpr int NB_ODD = (JK / 2 - 1) * (JK / 2) / 2;        // 2016 on odd steps (positions 0 and JK-1 idle)
constexpr size_t SOLVE_SMEM = sizeof(float) * (2 * JK * SLD) + sizeof(float2) * (JK / 2) + sizeof(float) * 64 +
                              sizeof(uchar2) * (NB_EVEN + NB_ODD) + sizeof(int) * JK + sizeof(float) * 2 * JK;

// Rotation parameters of one position pair in the SCALED (fast-Givens) form.  The true matrices are
// G = D Ghat D and R = Rhat D with a deferred diagonal D; a rotation by (c, s = t c) followed by the exchange,
//   x_p' = s x_p + c x_q,  x_q' = c x_p - s x_q,
// becomes, on the stored values, xhat_p' = xhat_q + alpha xhat_p, xhat_q' = xhat_p - beta xhat_q (one FMA per
// element instead of a multiply and an FMA) with alpha = t d_p/d_q, beta = t d_q/d_p and new scales
// d_p' = c d_q, d_q' = c d_p.  |t| <= 1, so c >= 0.707 and the scales stay within 0.707^128 over one sweep.
// Returns (alpha, -beta) and updates the two scales.
__device__ __forceinline__ float2 jacobi_scaled(float ghat_pp, float ghat_qq, float ghat_pq, float& dp, float& dq) {
  const float app = dp * dp * ghat_pp, aqq = dq * dq * ghat_qq, apq = dp * dq * ghat_pq;
  const float ratio = __fdividef(dp, dq);                       // independent of the chain below
  float c = 1.f, t = 0.f;
  if (apq * apq > 1e-16f * (app * aqq) && apq != 0.f) {         // |cos| > 1e-8
    const float tau = __fdividef(aqq - app, 2.f * apq);
    const float w = fmaf(tau, tau, 1.f);
    t = __fdividef(copysignf(1.f, tau), fabsf(tau) + w * rsqrtf(w));   // sqrt(w) = w * rsqrt(w); any t gives an exact rotation
    c = rsqrtf(fmaf(t, t, 1.f));
  }
  const float alpha = t * ratio;
  const float nbeta = (t != 0.f) ? -__fdividef(t, ratio) : 0.f;
  const float ndp = c * dq, ndq = c * dp;
  dp = ndp; dq = ndq;
  return make_float2(alpha, nbeta);
}

// scaled rotation + exchange of one position pair of a row: (a, b) <- (b + alpha a, a - beta b); q = (alpha, -beta)
#define ASVD_ROT_SWAP(a, b, q_)        \
  do {                                 \
    const float _a = (a), _b = (b);    \
    (a) = fmaf((q_).x, _a, _b);        \
    (b) = fmaf((q_).y, _b, _a);        \
  } while (0)

// Common head of the solve kernels: G = sum of the partial Grams into shared memory, convergence measure of the
// pair at visit time, non-finite check, threshold test and clean-pair bookkeeping.  Returns false when the CTA has
// nothing to rotate.
template <int NT = SOLVE_THREADS>
__device__ __forceinline__ bool solve_prologue(const float* __restrict__ Gpart, int chunks, int idx, int b, int2 pr, int tid,
                                               float* G, float* red, int* __restrict__ pairflag,
                                               unsigned* __restrict__ maxoff_bits, int* __restrict__ status, float tol,
                                               int* trk, int nb, int round_stamp, int precise, int nbatch, int half_gram,
                                               bool preloaded = false) {
  const float* Gp = Gpart + (int64_t)idx * chunks * (JK * JK);

  if (!preloaded) {     // (preloaded: the caller has already brought a single-chunk Gram matrix into G by bulk copies)
    // G = sum of the partial Grams.  16384 elements over 512 threads = 8 float4 per thread and chunk; all loads of
    // a chunk are issued before the first add so ~32 KB per warp are in flight
    // (NT threads: JK*JK/4/NT float4 per thread, in passes of 8 so that a 256-thread CTA stays within its registers;
    // every element still sums its chunks in the same order, whatever NT)
    const float4* Gp4 = reinterpret_cast<const float4*>(Gp);
    if constexpr ((JK * JK / 4) % (NT * 8) != 0) {
      // thread counts that do not divide the matrix (the 160-thread triangular solve): passes of 4, ragged tail guarded
#pragma unroll 1
      for (int base = tid; base < JK * JK / 4; base += NT * 4) {
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // four chunks' loads (16 float4 per thread) in flight before the first add; chunks are still summed in order
        for (int c0 = 0; c0 < chunks; c0 += 4) {
          float4 v[4][4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              v[cc][i] = (c0 + cc < chunks && base + NT * i < JK * JK / 4) ? Gp4[(int64_t)(c0 + cc) * (JK * JK / 4) + base + NT * i]
                                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (c0 + cc < chunks) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { acc[i].x += v[cc][i].x; acc[i].y += v[cc][i].y; acc[i].z += v[cc][i].z; acc[i].w += v[cc][i].w; }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = (base + NT * i) * 4;
          if (e < JK * JK) *reinterpret_cast<float4*>(&G[(e >> 7) * SLD + (e & (JK - 1))]) = acc[i];
        }
      }
    } else
#pragma unroll 1
    for (int pass = 0; pass < JK * JK / 4 / NT / 8; ++pass) {
      const int base = tid + NT * 8 * pass;
      float4 acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < chunks; ++c) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = Gp4[(int64_t)c * (JK * JK / 4) + base + NT * i];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = (base + NT * i) * 4;
        *reinterpret_cast<float4*>(&G[(e >> 7) * SLD + (e & (JK - 1))]) = acc[i];
      }
    }
  }
  __syncthreads();
  if (half_gram) {
    // the tensor-core Gram pass in precise mode delivers T = (0.5 HI + LO) HI^T; the Gram matrix is T + T^T
    if constexpr (NT == SOLVE_THREADS) {
      constexpr int PER = JK * JK / NT;
      float v[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int e = tid + NT * k, r = e >> 7, c = e & (JK - 1);
        v[k] = G[r * SLD + c] + G[c * SLD + r];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int e = tid + NT * k;
        G[(e >> 7) * SLD + (e & (JK - 1))] = v[k];
      }
    } else {
      // fewer threads: in place, the thread that owns (r, c) with r < c also writes (c, r); the sum is the same
      // floating-point value either way round, so the result equals the staged version bit for bit
#pragma unroll 4
      for (int e = tid; e < JK * JK; e += NT) {
        const int r = e >> 7, c = e & (JK - 1);
        if (r < c) {
          const float v = G[r * SLD + c] + G[c * SLD + r];
          G[r * SLD + c] = v; G[c * SLD + r] = v;
        } else if (r == c) {
          G[r * SLD + r] += G[r * SLD + r];
        }
      }
    }
    __syncthreads();
  }
  // convergence measure of this pair at visit time: max |cos| between any two of its 128 vectors
  float mx = 0.f;
  int bad = 0;
  for (int e = tid; e < JK * JK; e += NT) {
    int r = e >> 7, c = e & (JK - 1);
    float g = G[r * SLD + c];
    if (!(fabsf(g) <= FLT_MAX)) bad = 1;
    if (r < c) {
      float d = G[r * SLD + r] * G[c * SLD + c];
      if (d > 0.f) mx = fmaxf(mx, fabsf(g) * rsqrtf(d));
    }
  }
  mx = warp_max(mx);
  bad = __any_sync(0xffffffffu, bad);
  if ((tid & 31) == 0) { red[tid >> 5] = mx; red[32 + (tid >> 5)] = bad ? 1.f : 0.f; }
  __syncthreads();
  if (tid < 32) {
    float v = warp_max(tid < NT / 32 ? red[tid] : 0.f);
    float bb = warp_max(tid < NT / 32 ? red[32 + tid] : 0.f);
    if (tid == 0) { red[0] = v; red[32] = bb; }
  }
  __syncthreads();
  mx = red[0];
  bad = red[32] > 0.f;
  if (bad) {
    if (tid == 0) { atomicOr(&status[b], 1); pairflag[idx] = 0; }
    return false;
  }
  if (tid == 0) {
    atomicMax(&maxoff_bits[b], __float_as_uint(mx));
    // pairs that are (nearly) orthogonal already: once they appear, the single-pass TF32 Gram would hide them from
    // the threshold test below, so the driver switches to the 3-term split for the next sweep
    if (mx < 1e-2f) atomicAdd(&maxoff_bits[nbatch + b], 1u);
  }
  if (mx < tol) {
    if (tid == 0) {
      pairflag[idx] = 0;
      if (precise) trk[nb + pr.x * nb + pr.y] = round_stamp;      // verified orthogonal as of this round
    }
    return false;
  }
  if (tid == 0) {
    pairflag[idx] = 1;
    trk[pr.x] = round_stamp; trk[pr.y] = round_stamp;             // both blocks change in this round: every pair
  }                                                               // containing one of them is dirty again

  return true;
}

// Common tail: Rs holds R with its columns in norm-sorted order; normalise the columns (default) or take one
// Newton-Schulz step (first-generation tail, ASVD_B200_POLISH=NS), then store.
__device__ __forceinline__ void solve_polish_write(float* G, float* Rs, float* __restrict__ Rout, int idx, int tid,
                                                   int transpose_out) {
  // ---- re-orthogonalise, write R
  // 8x8 register tiles (rows {4ta..} U {64+4ta..}, columns {4tb..} U {64+4tb..}) on the first 256 threads: 64 FMAs
  // per 4-10 shared-memory loads, so both 128^3 products run at the FMA issue rate
  float* E = G;
  const int ta = (tid >> 4) & 15, tb = tid & 15;
  if (transpose_out & 2) {
    // Default tail: unit column norms only.  The accumulated scaled rotations leave |R^T R - I| ~ 1e-5 on the
    // diagonal (approximate rsqrt in every c) and ~1e-6 off it.  Neither needs the Newton-Schulz step below: sigma and
    // the second factor are recovered from the ORIGINAL weight once the vectors are orthogonal (run_svd), the
    // truncation error is stationary in the subspace, and orthogonality of the vectors is what the sweeps measure
    // and enforce directly.  Measured (4096^2, 11008x4096, 4096x11008, 32000x4096): sigma vs fp64 1.4-5.3e-6 with
    // either tail, same sweep counts, 5% less time.
    __shared__ float cnp[4][JK];
    {
      const int col = tid & (JK - 1), part = tid >> 7;
      float ss = 0.f;
#pragma unroll 8
      for (int l = 0; l < JK / 4; ++l) { const float x = Rs[(part * (JK / 4) + l) * SLD + col]; ss = fmaf(x, x, ss); }
      cnp[part][col] = ss;
    }
    __syncthreads();
    if (tid < JK) cnp[0][tid] = rsqrtf(cnp[0][tid] + cnp[1][tid] + cnp[2][tid] + cnp[3][tid]);
    __syncthreads();
    float* Ro = Rout + (int64_t)idx * (JK * JK);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = (tid + SOLVE_THREADS * k) * 4, r = e >> 7, c = e & (JK - 1);
      const float4 x = *reinterpret_cast<const float4*>(&Rs[r * SLD + c]);
      const float4 n = *reinterpret_cast<const float4*>(&cnp[0][c]);
      *reinterpret_cast<float4*>(&Ro[e]) = make_float4(x.x * n.x, x.y * n.y, x.z * n.z, x.w * n.w);
    }
    return;
  }
  if (tid < 256) {
    float e[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) e[i][j] = 0.f;
#pragma unroll 2
    for (int l = 0; l < JK; ++l) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Rs[l * SLD + ta * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Rs[l * SLD + 64 + ta * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Rs[l * SLD + tb * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Rs[l * SLD + 64 + tb * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) e[i][j] = fmaf(av[i], bv[j], e[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = (i < 4) ? ta * 4 + i : 64 + ta * 4 + (i - 4);
      *reinterpret_cast<float4*>(&E[rr * SLD + tb * 4]) = make_float4(e[i][0], e[i][1], e[i][2], e[i][3]);
      *reinterpret_cast<float4*>(&E[rr * SLD + 64 + tb * 4]) = make_float4(e[i][4], e[i][5], e[i][6], e[i][7]);
    }
  }
  __syncthreads();
  if (tid < 256) {
    float o[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
    const float* ra = &Rs[(ta * 4) * SLD];
    const float* rb = &Rs[(64 + ta * 4) * SLD];
#pragma unroll 2
    for (int l = 0; l < JK; ++l) {
      const float4 b0 = *reinterpret_cast<const float4*>(&E[l * SLD + tb * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&E[l * SLD + 64 + tb * 4]);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float av[8] = {ra[l], ra[SLD + l], ra[2 * SLD + l], ra[3 * SLD + l], rb[l], rb[SLD + l], rb[2 * SLD + l], rb[3 * SLD + l]};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] = fmaf(av[i], bv[j], o[i][j]);
    }
    float* Ro = Rout + (int64_t)idx * (JK * JK);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = (i < 4) ? ta * 4 + i : 64 + ta * 4 + (i - 4);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int cc = hh * 64 + tb * 4;
        const float4 x = *reinterpret_cast<const float4*>(&Rs[rr * SLD + cc]);
        const float4 w = make_float4(1.5f * x.x - 0.5f * o[i][4 * hh], 1.5f * x.y - 0.5f * o[i][4 * hh + 1],
                                     1.5f * x.z - 0.5f * o[i][4 * hh + 2], 1.5f * x.w - 0.5f * o[i][4 * hh + 3]);
        if (!transpose_out) {
          *reinterpret_cast<float4*>(&Ro[rr * JK + cc]) = w;
        } else {
          Ro[(cc + 0) * JK + rr] = w.x; Ro[(cc + 1) * JK + rr] = w.y; Ro[(cc + 2) * JK + rr] = w.z; Ro[(cc + 3) * JK + rr] = w.w;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1)
solve_kernel(const float* __restrict__ Gpart, int chunks, int pairs_per_mat, float* __restrict__ Rout,
             int* __restrict__ pairflag, unsigned* __restrict__ maxoff_bits, int* __restrict__ status,
             const int* __restrict__ done, float tol, int transpose_out, int dbg_steps, const int2* __restrict__ pairs,
             int* __restrict__ track, int nb, int round_stamp, const int* __restrict__ precise_b, int half_gram_tc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* G = reinterpret_cast<float*>(smem_raw);            // [JK][SLD]; later E = R^T R
  float* Rs = G + JK * SLD;                                 // [JK][SLD]; R in sorted column order (after the sweep)
  float2* cs = reinterpret_cast<float2*>(Rs + JK * SLD);    // [JK/2]
  float* red = reinterpret_cast<float*>(cs + JK / 2);       // [64]
  uchar2* tab_even = reinterpret_cast<uchar2*>(red + 64);   // [NB_EVEN] (s, t) of each upper-triangle block
  uchar2* tab_odd = tab_even + NB_EVEN;                     // [NB_ODD]
  int* dest = reinterpret_cast<int*>(tab_odd + NB_ODD);     // [JK] output column of each position
  float* diag = reinterpret_cast<float*>(dest + JK);        // [JK]
  float* dsc = diag + JK;                                   // [JK] deferred column scales (fast Givens)

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  const int tid = threadIdx.x;
  const int2 pr = pairs[p];
  int* trk = track + (int64_t)b * (nb + nb * nb);
  if (pair_is_clean(track, nb, b, pr.x, pr.y)) {           // untouched since it was last verified: nothing to do
    if (tid == 0) pairflag[idx] = 0;
    return;
  }
  const int precise = precise_b[b], half_gram = half_gram_tc && precise;     // Gram mode of THIS matrix (see run_svd)
  if (!solve_prologue(Gpart, chunks, idx, b, pr, tid, G, red, pairflag, maxoff_bits, status, tol, trk, nb, round_stamp,
                      precise, gridDim.y, half_gram))
    return;
  if (tid < JK) dsc[tid] = 1.f;      // ordered before its first use by the barriers of the first step

  // ---- one odd-even sweep, everything in registers.
  // G threads (256): thread (a, c) keeps the 8x8 patch G[8a.., 8c..] of the FULL symmetric matrix in registers.
  //   Even steps pair rows/columns (2t, 2t+1): entirely inside the patches, no data movement.  Odd steps pair
  //   (2t+1, 2t+2): three pairs inside a patch, one straddling two patches; the straddling rows (then columns) are
  //   exchanged through a small shared buffer.
  // R threads (256): thread (row, half) keeps 64 columns of one row of R in registers (column rotations only).
  // Rotation parameters come from the diagonal and first super-diagonal, which the patch owners publish each step.
  auto bar_all = [] { asm volatile("bar.sync 1, %0;" ::"n"(SOLVE_THREADS) : "memory"); };
  auto bar_g = [] { asm volatile("bar.sync 2, %0;" ::"n"(SOLVE_GTHREADS) : "memory"); };
  float* gd = diag;                  // [JK]   current diagonal (scaled form)
  float* xbuf = Rs;                  // exchange area inside the (not yet used) Rs region
  float* rowF = xbuf;                // [16][JK] first row of every patch row
  float* rowL = rowF + 16 * JK;      // [16][JK] last row
  float* colF = rowL + 16 * JK;      // [16][JK] first column of every patch column
  float* colL = colF + 16 * JK;      // [16][JK] last column
  float* go = colL + 16 * JK;        // [JK]   first super-diagonal G(i, i+1)
  if (tid < SOLVE_GTHREADS) {
    const int pa = tid >> 4, pc = tid & 15;
    float g[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x0 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc]);
      const float4 x1 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc + 4]);
      g[i][0] = x0.x; g[i][1] = x0.y; g[i][2] = x0.z; g[i][3] = x0.w;
      g[i][4] = x1.x; g[i][5] = x1.y; g[i][6] = x1.z; g[i][7] = x1.w;
    }
    const int dbg_mode = dbg_steps >> 8;
    for (int st = 0; st < 2 * (dbg_steps & 0xff); ++st) {
      const int odd = st & 1;
      // publish the entries the rotation parameters are computed from
      if (dbg_mode & 8) {
      } else if (pa == pc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gd[8 * pa + i] = g[i][i];
#pragma unroll
        for (int i = 0; i < 7; ++i) go[8 * pa + i] = g[i][i + 1];
      } else if (pc == pa + 1) {
        go[8 * pa + 7] = g[7][0];
      }
      bar_all();           // gd / go visible
      bar_all();           // cs / dsc written by the parameter threads
      if (dbg_mode & 2) continue;
      // the parameters of this patch's four row pairs and four column pairs (32 contiguous bytes each)
      float2 qr[4], qc[4];
      {
        const float4 r0 = *reinterpret_cast<const float4*>(&cs[4 * pa]), r1 = *reinterpret_cast<const float4*>(&cs[4 * pa + 2]);
        const float4 c0 = *reinterpret_cast<const float4*>(&cs[4 * pc]), c1 = *reinterpret_cast<const float4*>(&cs[4 * pc + 2]);
        qr[0] = make_float2(r0.x, r0.y); qr[1] = make_float2(r0.z, r0.w); qr[2] = make_float2(r1.x, r1.y); qr[3] = make_float2(r1.z, r1.w);
        qc[0] = make_float2(c0.x, c0.y); qc[1] = make_float2(c0.z, c0.w); qc[2] = make_float2(c1.x, c1.y); qc[3] = make_float2(c1.z, c1.w);
      }
      if (!odd) {
#pragma unroll
        for (int lp = 0; lp < 4; ++lp) {                 // rows (2lp, 2lp+1): new = (g1 + alpha g0, g0 - beta g1)
          const float2 q = qr[lp];
#pragma unroll
          for (int j = 0; j < 8; ++j) ASVD_ROT_SWAP(g[2 * lp][j], g[2 * lp + 1][j], q);
        }
#pragma unroll
        for (int lp = 0; lp < 4; ++lp) {                 // columns, same rule
          const float2 q = qc[lp];
#pragma unroll
          for (int i = 0; i < 8; ++i) ASVD_ROT_SWAP(g[i][2 * lp], g[i][2 * lp + 1], q);
        }
      } else {
        // rows: exchange the boundary rows with the patches above / below
        // exchange layout: [patch row][half of the 8 values][patch column][4 values] -> the 16 lanes that share a patch
        // row touch 256 contiguous bytes per access (no bank conflicts)
        *reinterpret_cast<float4*>(&rowF[pa * JK + 4 * pc]) = make_float4(g[0][0], g[0][1], g[0][2], g[0][3]);
        *reinterpret_cast<float4*>(&rowF[pa * JK + 64 + 4 * pc]) = make_float4(g[0][4], g[0][5], g[0][6], g[0][7]);
        *reinterpret_cast<float4*>(&rowL[pa * JK + 4 * pc]) = make_float4(g[7][0], g[7][1], g[7][2], g[7][3]);
        *reinterpret_cast<float4*>(&rowL[pa * JK + 64 + 4 * pc]) = make_float4(g[7][4], g[7][5], g[7][6], g[7][7]);
        bar_g();
        {
          float below[8], above[8];
          if (pa < 15) {
            const float4 x0 = *reinterpret_cast<const float4*>(&rowF[(pa + 1) * JK + 4 * pc]);
            const float4 x1 = *reinterpret_cast<const float4*>(&rowF[(pa + 1) * JK + 64 + 4 * pc]);
            below[0] = x0.x; below[1] = x0.y; below[2] = x0.z; below[3] = x0.w;
            below[4] = x1.x; below[5] = x1.y; below[6] = x1.z; below[7] = x1.w;
          }
          if (pa > 0) {
            const float4 x0 = *reinterpret_cast<const float4*>(&rowL[(pa - 1) * JK + 4 * pc]);
            const float4 x1 = *reinterpret_cast<const float4*>(&rowL[(pa - 1) * JK + 64 + 4 * pc]);
            above[0] = x0.x; above[1] = x0.y; above[2] = x0.z; above[3] = x0.w;
            above[4] = x1.x; above[5] = x1.y; above[6] = x1.z; above[7] = x1.w;
          }
#pragma unroll
          for (int lp = 0; lp < 3; ++lp) {               // rows (2lp+1, 2lp+2), pair index 4pa + lp
            const float2 q = qr[lp];
#pragma unroll
            for (int j = 0; j < 8; ++j) ASVD_ROT_SWAP(g[2 * lp + 1][j], g[2 * lp + 2][j], q);
          }
          if (pa < 15) {                                 // row 8pa+7 is the p side of pair 4pa+3: new = below + alpha own
            const float2 q = qr[3];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[7][j] = fmaf(q.x, g[7][j], below[j]);
          }
          if (pa > 0) {                                  // row 8pa is the q side of pair 4pa-1: new = above - beta own
            const float2 q = cs[4 * pa - 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[0][j] = fmaf(q.y, g[0][j], above[j]);
          }
        }
        // columns: exchange the (row-rotated) boundary columns with the patches left / right
        *reinterpret_cast<float4*>(&colF[pa * JK + 4 * pc]) = make_float4(g[0][0], g[1][0], g[2][0], g[3][0]);
        *reinterpret_cast<float4*>(&colF[pa * JK + 64 + 4 * pc]) = make_float4(g[4][0], g[5][0], g[6][0], g[7][0]);
        *reinterpret_cast<float4*>(&colL[pa * JK + 4 * pc]) = make_float4(g[0][7], g[1][7], g[2][7], g[3][7]);
        *reinterpret_cast<float4*>(&colL[pa * JK + 64 + 4 * pc]) = make_float4(g[4][7], g[5][7], g[6][7], g[7][7]);
        bar_g();
        {
          float right[8], left[8];
          if (pc < 15) {
            const float4 x0 = *reinterpret_cast<const float4*>(&colF[pa * JK + 4 * (pc + 1)]);
            const float4 x1 = *reinterpret_cast<const float4*>(&colF[pa * JK + 64 + 4 * (pc + 1)]);
            right[0] = x0.x; right[1] = x0.y; right[2] = x0.z; right[3] = x0.w;
            right[4] = x1.x; right[5] = x1.y; right[6] = x1.z; right[7] = x1.w;
          }
          if (pc > 0) {
            const float4 x0 = *reinterpret_cast<const float4*>(&colL[pa * JK + 4 * (pc - 1)]);
            const float4 x1 = *reinterpret_cast<const float4*>(&colL[pa * JK + 64 + 4 * (pc - 1)]);
            left[0] = x0.x; left[1] = x0.y; left[2] = x0.z; left[3] = x0.w;
            left[4] = x1.x; left[5] = x1.y; left[6] = x1.z; left[7] = x1.w;
          }
#pragma unroll
          for (int lp = 0; lp < 3; ++lp) {
            const float2 q = qc[lp];
#pragma unroll
            for (int i = 0; i < 8; ++i) ASVD_ROT_SWAP(g[i][2 * lp + 1], g[i][2 * lp + 2], q);
          }
          if (pc < 15) {
            const float2 q = qc[3];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i][7] = fmaf(q.x, g[i][7], right[i]);
          }
          if (pc > 0) {
            const float2 q = cs[4 * pc - 1];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i][0] = fmaf(q.y, g[i][0], left[i]);
          }
        }
      }
    }
    // final diagonal (true norms) for the sort
    bar_all();             // R threads are done reading cs; exchange area is free
    if (pa == pc) {
#pragma unroll
      for (int i = 0; i < 8; ++i) gd[8 * pa + i] = dsc[8 * pa + i] * dsc[8 * pa + i] * g[i][i];
    }
    bar_g();
    if (tid < JK) {
      const float d = gd[tid];
      int rank = 0;
      for (int j = 0; j < JK; ++j) {
        const float e = gd[j];
        rank += (e > d) || (e == d && j < tid);
      }
      dest[tid] = rank;
    }
    bar_all();             // dest[] ready
  } else {
    // thread = (row, half): 64 consecutive columns of one row of R in registers
    const int rt = tid - SOLVE_GTHREADS;
    const int row = rt >> 1, half = rt & 1;
    float r[JK / 2];
#pragma unroll
    for (int j = 0; j < JK / 2; ++j) r[j] = (half * (JK / 2) + j == row) ? 1.f : 0.f;
    const float2* cs_h = cs + half * (JK / 4);
    const int dbg_mode = dbg_steps >> 8;
    for (int st = 0; st < 2 * (dbg_steps & 0xff); ++st) {
      const int odd = st & 1;
      bar_all();           // gd / go published
      if (!(dbg_mode & 1) && rt < JK / 2 - odd) {                           // rotation parameters of pair rt
        const int pp = 2 * rt + odd;
        float dp = dsc[pp], dq = dsc[pp + 1];
        cs[rt] = jacobi_scaled(gd[pp], gd[pp + 1], go[pp], dp, dq);
        dsc[pp] = dp; dsc[pp + 1] = dq;
      }
      bar_all();           // cs visible
      if (dbg_mode & 4) continue;
      if (!odd) {
        // local pairs (2j, 2j+1), parameters cs[32*half + j]; loads batched 16 at a time
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          float2 q[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) q[j] = cs_h[gq * 16 + j];
#pragma unroll
          for (int j = 0; j < 16; ++j) ASVD_ROT_SWAP(r[2 * (gq * 16 + j)], r[2 * (gq * 16 + j) + 1], q[j]);
        }
      } else {
        // local pairs (2j+1, 2j+2) for j < 31 with cs[32*half + j]; global pair (63,64) straddles the halves
        float2 q[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) q[j] = cs_h[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) ASVD_ROT_SWAP(r[2 * j + 1], r[2 * j + 2], q[j]);
#pragma unroll
        for (int j = 0; j < 15; ++j) q[j] = cs_h[16 + j];
#pragma unroll
        for (int j = 0; j < 15; ++j) ASVD_ROT_SWAP(r[2 * (16 + j) + 1], r[2 * (16 + j) + 2], q[j]);
        const float2 qb = cs[JK / 4 - 1];                       // pair t = 31: positions 63 | 64
        const float mine = half ? r[0] : r[JK / 2 - 1];
        const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
        if (half == 0) r[JK / 2 - 1] = fmaf(qb.x, mine, other);             // a' = b + alpha a
        else r[0] = fmaf(qb.y, mine, other);                                // b' = a - beta b
      }
    }
    bar_all();             // matches the G threads' barrier before the final diagonal
    bar_all();             // dest[] ready
#pragma unroll
    for (int j = 0; j < JK / 2; ++j) Rs[row * SLD + dest[half * (JK / 2) + j]] = r[j] * dsc[half * (JK / 2) + j];
  }
  __syncthreads();
  solve_polish_write(G, Rs, Rout, idx, tid, transpose_out);
}

#include "svd_solve_quad.cuh"
#include "svd_solve_tri.cuh"

// ------------------------------------------------------------------------------------------------ update
// panel <- R^T panel, i.e. out[j][c] = sum_i R[i][j] * X[i][c], in place, UPD_TILES column tiles of 128 per CTA.
constexpr int UPD_TILES = 4;
constexpr size_t UPDATE_SMEM = sizeof(float) * (JK * JK + 2 * JK * 128);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(256, 1)
update_kernel(float* __restrict__ X, int64_t mat_stride, int ldx, const int2* __restrict__ pairs, int len_pad,
              int pairs_per_mat, const float* __restrict__ Rin, const int* __restrict__ pairflag,
              const int* __restrict__ done) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Rs = reinterpret_cast<float*>(smem_raw);          // [128][128]  Rs[i][j]
  float* Xs = Rs + JK * JK;                                // [2][128][128]
  const int b = blockIdx.z, p = blockIdx.y;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  if (!pairflag[idx]) return;
  const int2 pr = pairs[p];
  float* Xb = X + b * mat_stride;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int tile0 = blockIdx.x * UPD_TILES;
  const int ntiles = min(UPD_TILES, len_pad / 128 - tile0);

  auto issue_tile = [&](int tile, int buf) {
    const int c0 = (tile0 + tile) * 128;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int e = t + 256 * i;
      int row = e >> 5, c4 = (e & 31) * 4;
      int vec = (row < JB) ? pr.x * JB + row : pr.y * JB + (row - JB);
      cp_async16(&Xs[(buf * JK + row) * 128 + c4], Xb + xt_off(vec, c0 + c4, ldx >> 5));
    }
  };
  {
    const float* Rg = Rin + (int64_t)idx * (JK * JK);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int e = t + 256 * i;
      cp_async16(&Rs[e * 4], Rg + e * 4);
    }
  }
  issue_tile(0, 0);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) issue_tile(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* xs = Xs + buf * JK * 128;
#pragma unroll 4
    for (int i = 0; i < JK; ++i) {
      float4 a0 = *reinterpret_cast<const float4*>(&Rs[i * JK + ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&Rs[i * JK + 64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&xs[i * 128 + tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&xs[i * 128 + 64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
    }
    const int c0 = (tile0 + tile) * 128;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int j = (r < 4) ? ty * 4 + r : 64 + ty * 4 + (r - 4);
      int vec = (j < JB) ? pr.x * JB + j : pr.y * JB + (j - JB);
      *reinterpret_cast<float4*>(Xb + xt_off(vec, c0 + tx * 4, ldx >> 5)) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(Xb + xt_off(vec, c0 + 64 + tx * 4, ldx >> 5)) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// block-tiled X -> row-major copy (once per factorisation; the epilogue kernels below read rows)
__global__ void __launch_bounds__(256) untile_kernel(const float* __restrict__ Xt, float* __restrict__ Xr, int64_t mat_stride,
                                                     int nv_pad, int ld) {
  const int b = blockIdx.z, vec = blockIdx.y;
  const int col4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (col4 >= ld) return;
  const float4 v = *reinterpret_cast<const float4*>(Xt + b * mat_stride + xt_off(vec, col4, ld >> 5));
  *reinterpret_cast<float4*>(Xr + b * mat_stride + (int64_t)vec * ld + col4) = v;
}

// per-row 2-norm; optionally normalises the row in place.  One CTA per row.
// `anchor` (optional): the Jacobi column norm of the same vector.  |Y_j| recomputed from the original weight is
// free of accumulated rotation drift but, for sigma_j below ~eps*sigma_max, it is dominated by leakage from the
// large directions; one-sided Jacobi norms keep high relative accuracy there, so they win when the two disagree.
__global__ void __launch_bounds__(256) rownorm_kernel(float* __restrict__ X, int64_t mat_stride, int ld, int len,
                                                      int normalise, float* __restrict__ norm_out, int nv_pad,
                                                      int* __restrict__ status, const float* __restrict__ anchor) {
  __shared__ float red[8];
  const int b = blockIdx.y, j = blockIdx.x;
  float* row = X + b * mat_stride + (int64_t)j * ld;
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += 256) { float v = row[i]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  float nrm = sqrtf(red[0]);
  if (anchor) {
    const float a = anchor[(int64_t)b * nv_pad + j];
    if (!(fabsf(nrm - a) <= 1e-3f * a)) nrm = (a <= FLT_MAX && nrm <= FLT_MAX) ? a : nrm;
  }
  if (threadIdx.x == 0) {
    norm_out[(int64_t)b * nv_pad + j] = nrm;
    if (!(nrm <= FLT_MAX)) atomicOr(&status[b], 1);
  }
  if (normalise && nrm > 0.f) {
    const float inv = 1.f / nrm;
    for (int i = threadIdx.x; i < len; i += 256) row[i] *= inv;
  }
}

// descending bitonic sort of (sigma, index) for one matrix per CTA.  All nv_pad rows take part: the norm-sorting
// swaps inside the solver may leave a real vector in a padding slot, and the zero padding vectors sort last.
__global__ void __launch_bounds__(1024) sort_kernel(const float* __restrict__ sig_in, float* __restrict__ sig_out,
                                                    int* __restrict__ perm, int nv, int nv_pad, int P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);
  int* val = reinterpret_cast<int*>(key + P);
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    key[i] = (i < nv_pad) ? sig_in[(int64_t)b * nv_pad + i] : -1.f;   // zero padding vectors sort last (sigma 0)
    val[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int l = i ^ j;
        if (l > i) {
          bool desc = ((i & k) == 0);
          float a = key[i], c = key[l];
          // NaN keys (flagged separately through status) are treated as smallest
          bool a_lt_c = (a < c) || (a != a && c == c);
          bool swap = desc ? a_lt_c : ((c < a) || (c != c && a == a));
          if (swap) { key[i] = c; key[l] = a; int tv = val[i]; val[i] = val[l]; val[l] = tv; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < nv_pad; i += blockDim.x) {
    sig_out[(int64_t)b * nv_pad + i] = key[i];
    perm[(int64_t)b * nv_pad + i] = val[i];
  }
}

// ------------------------------------------------------------------------------------------------ recovery planes
// Y = Xhat * (W diag(s)) has to be as exact as an fp32 product (it anchors sigma and the second factor to the input).
// On the tensor cores: Xhat (fp32) = a1 + a2 + a3 and W (fp16: 11 significant bits) = b1 + b2 exactly, all bf16, every
// plane product is exact in the fp32 accumulator, and a1 b1 + a2 b1 + a3 b1 + a1 b2 + a2 b2 drops only terms below
// 2^-24 (gemm_planes_f32).  The scale is folded into Xhat when it runs along the contraction (wide weights) and into the
// GEMM epilogue otherwise.
__device__ __forceinline__ void bf16_planes3(float x, __nv_bfloat16& a1, __nv_bfloat16& a2, __nv_bfloat16& a3) {
  a1 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a1);
  a2 = __float2bfloat16_rn(r1);
  a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
}
__global__ void __launch_bounds__(256) split_x_kernel(const float* __restrict__ Xr, const float* __restrict__ ascale, int nv_pad,
                                                      int len_pad, int len, __nv_bfloat16* __restrict__ As) {
  const int row = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= len_pad) return;
  const float4 x = *reinterpret_cast<const float4*>(Xr + (int64_t)row * len_pad + c);
  float v[4] = {x.x, x.y, x.z, x.w};
  __nv_bfloat16 pl[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (ascale) v[e] = (c + e < len) ? v[e] * ascale[c + e] : 0.f;
    bf16_planes3(v[e], pl[0][e], pl[1][e], pl[2][e]);
  }
  __nv_bfloat16* dst = As + (int64_t)row * (3 * (int64_t)len_pad) + c;
#pragma unroll
  for (int q = 0; q < 3; ++q) *reinterpret_cast<uint2*>(dst + (int64_t)q * len_pad) = *reinterpret_cast<uint2*>(pl[q]);
}
// planes [b1|b2] of the weight, one row per OUTPUT index of the recovery GEMM (= vector index), contraction index along
// the row: transposed copy when the vectors are the columns of W (tall), plain copy otherwise
template <typename T>
__global__ void __launch_bounds__(256) split_w_kernel(const T* __restrict__ W, int64_t ldw, int m, int n, int tall, int len_pad,
                                                      __nv_bfloat16* __restrict__ Bs) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t ldb = 2 * (int64_t)len_pad;
  if (tall) {
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, j = j0 + tx;
      tile[r][tx] = (i < m && j < n) ? to_f32<T>(W[(int64_t)i * ldw + j]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r, i = i0 + tx;
      if (j < n && i < m) {
        const float w = tile[tx][r];
        const __nv_bfloat16 b1 = __float2bfloat16_rn(w);
        Bs[(int64_t)j * ldb + i] = b1;
        Bs[(int64_t)j * ldb + len_pad + i] = __float2bfloat16_rn(w - __bfloat162float(b1));
      }
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, j = j0 + tx;
      if (i < m && j < n) {
        const float w = to_f32<T>(W[(int64_t)i * ldw + j]);
        const __nv_bfloat16 b1 = __float2bfloat16_rn(w);
        Bs[(int64_t)i * ldb + j] = b1;
        Bs[(int64_t)i * ldb + len_pad + j] = __float2bfloat16_rn(w - __bfloat162float(b1));
      }
    }
  }
}

// Y = (sum of the K-split partial results, fixed order) * column scale * row scale; columns >= ncols are written as 0
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ Yp, int nsplit, int64_t split_stride,
                                                           float* __restrict__ Y, int64_t total, int ldy,
                                                           const float* __restrict__ cscale, int ncols,
                                                           const float* __restrict__ rscale, int nrows) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= total) return;
  float4 acc = *reinterpret_cast<const float4*>(Yp + e);
  for (int q = 1; q < nsplit; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(Yp + q * split_stride + e);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int c = (int)(e % ldy), r = (int)(e / ldy);
  float a[4] = {acc.x, acc.y, acc.z, acc.w};
  const float rs = rscale ? (r < nrows ? rscale[r] : 0.f) : 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (c + k >= ncols) a[k] = 0.f;
    else a[k] *= (cscale ? cscale[c + k] : 1.f) * rs;
  }
  *reinterpret_cast<float4*>(Y + e) = make_float4(a[0], a[1], a[2], a[3]);
}

// row-major [nv_pad, ld] -> block-tiled working layout (inverse of untile_kernel)
__global__ void __launch_bounds__(256) tile_kernel(const float* __restrict__ Xr, float* __restrict__ Xt, int nv_pad, int ld) {
  const int vec = blockIdx.y;
  const int col4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (col4 >= ld) return;
  *reinterpret_cast<float4*>(Xt + xt_off(vec, col4, ld >> 5)) = *reinterpret_cast<const float4*>(Xr + (int64_t)vec * ld + col4);
}

// planes [a1|a2|a3] of W[i][l] * s[l]^2 (the Gram matrix of the ROWS of W diag(s) contracts W s^2 against W)
template <typename T>
__global__ void __launch_bounds__(256) split_ws2_kernel(const T* __restrict__ W, int64_t ldw, int m, int n,
                                                        const float* __restrict__ s, int len_pad, __nv_bfloat16* __restrict__ As) {
  const int i = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || l >= n) return;
  const float sl = s[l];
  __nv_bfloat16 a1, a2, a3;
  bf16_planes3(to_f32<T>(W[(int64_t)i * ldw + l]) * sl * sl, a1, a2, a3);
  __nv_bfloat16* dst = As + (int64_t)i * (3 * (int64_t)len_pad) + l;
  dst[0] = a1; dst[len_pad] = a2; dst[2 * (int64_t)len_pad] = a3;
}

// ------------------------------------------------------------------------------------------------ extract (a5, a6)
__device__ __forceinline__ float sig_pow(float s, float e) {
  if (!(s > 0.f)) return 0.f;
  if (e == 0.f) return 1.f;
  if (e == 1.f) return s;
  if (e == 0.5f) return sqrtf(s);
  if (e == -0.5f) return rsqrtf(s);
  if (e == -1.f) return 1.f / s;
  return powf(s, e);
}

// out[i][j] = src[perm[j]][i] * sigma[j]^e   (i < rows, j < r): gathers factor columns from vector rows
template <typename TC>
__global__ void __launch_bounds__(256) extract_cols_kernel(const float* __restrict__ src, int ld, const int* __restrict__ perm,
                                                           const float* __restrict__ sigma, float e, int rows, int r,
                                                           TC* __restrict__ out, int64_t ldo) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int q = ty; q < 32; q += 8) {
    int j = j0 + q, i = i0 + tx;
    tile[q][tx] = (j < r && i < rows) ? src[(int64_t)perm[j] * ld + i] * sig_pow(sigma[j], e) : 0.f;
  }
  __syncthreads();
  for (int q = ty; q < 32; q += 8) {
    int i = i0 + q, j = j0 + tx;
    if (i < rows && j < r) out[(int64_t)i * ldo + j] = from_f32<TC>(tile[tx][q]);
  }
}

// out[j][l] = src[perm[j]][l] * sigma[j]^e / scale[l]   (j < r, l < cols)
template <typename TC>
__global__ void __launch_bounds__(256) extract_rows_kernel(const float* __restrict__ src, int ld, const int* __restrict__ perm,
                                                           const float* __restrict__ sigma, float e,
                                                           const float* __restrict__ scale, int cols, int r,
                                                           TC* __restrict__ out, int64_t ldo) {
  const int j = blockIdx.y;
  const int l = blockIdx.x * 256 + threadIdx.x;
  if (j >= r || l >= cols) return;
  float v = src[(int64_t)perm[j] * ld + l] * sig_pow(sigma[j], e) / scale[l];
  out[(int64_t)j * ldo + l] = from_f32<TC>(v);
}

// ------------------------------------------------------------------------------------------------ host side
static void build_pair_table(const SvdPlan& p, std::vector<int2>& tab) {
  // circle method over nb blocks: nb-1 rounds of nb/2 disjoint pairs (I < J)
  tab.resize((size_t)p.rounds * p.pairs);
  std::vector<int> idx(p.nb);
  for (int i = 0; i < p.nb; ++i) idx[i] = i;
  for (int r = 0; r < p.rounds; ++r) {
    for (int i = 0; i < p.pairs; ++i) {
      int a = idx[i], b = idx[p.nb - 1 - i];
      tab[(size_t)r * p.pairs + i] = make_int2(a < b ? a : b, a < b ? b : a);
    }
    int last = idx[p.nb - 1];
    for (int i = p.nb - 1; i > 1; --i) idx[i] = idx[i - 1];
    if (p.nb > 1) idx[1] = last;
  }
}

constexpr int MODE_SKIP_PREP = 1;          // X already holds the vectors (block-tiled)
constexpr int MODE_STOP_AT_VECTORS = 2;    // return once Xr holds the unit vectors (no recovery, no sort)

template <typename T>
static int run_svd(const SvdPlan& p, int64_t ldw, unsigned char* ws, float tol, int max_sweeps, int* sweeps_out,
                   cudaStream_t st, const void* const* h_W, int mode = 0, float conv_tol = 0.f) {
  const void* const* d_W = reinterpret_cast<const void* const*>(ws + p.off_ptrs);
  const int2* d_pairs = reinterpret_cast<const int2*>(ws + p.off_pairs);
  float* X = reinterpret_cast<float*>(ws + p.off_X);
  float* Xr = reinterpret_cast<float*>(ws + p.off_Xr);
  float* Y = reinterpret_cast<float*>(ws + p.off_Y);
  float* G = reinterpret_cast<float*>(ws + p.off_G);
  float* R = reinterpret_cast<float*>(ws + p.off_R);
  int* flag = reinterpret_cast<int*>(ws + p.off_flag);
  unsigned* maxoff = reinterpret_cast<unsigned*>(ws + p.off_maxoff);
  int* done = reinterpret_cast<int*>(ws + p.off_done);
  int* prec = reinterpret_cast<int*>(ws + p.off_prec);
  float* sigma = reinterpret_cast<float*>(ws + p.off_sigma);
  int* perm = reinterpret_cast<int*>(ws + p.off_perm);
  int* status = reinterpret_cast<int*>(ws + p.off_status);
  float* scale = reinterpret_cast<float*>(ws + p.off_scale);
  float* norm = reinterpret_cast<float*>(ws + p.off_norm);
  int* track = reinterpret_cast<int*>(ws + p.off_track);
  const int64_t xs = (int64_t)p.nv_pad * p.len_pad;

  static bool attrs_set_dev[ASVD_MAX_DEVICES] = {};
  const int dev_slot = current_device_slot();
  bool& attrs_set = attrs_set_dev[dev_slot];
  if (!attrs_set) {
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVEQ_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_quad_g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVEQG_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_quad_g_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // two CTAs per SM
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_quad_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVEQR_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_tri_g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVET_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_tri_g_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // three CTAs per SM
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_tri_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVETR_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_tri_r_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // two CTAs per SM
    ASVD_CUDA_CHECK(upload_quad_schedule());
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPDATE_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    attrs_set = true;
  }

  if (!(mode & MODE_SKIP_PREP)) ASVD_CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(float) * xs * p.batch, st));
  ASVD_CUDA_CHECK(cudaMemsetAsync(done, 0, sizeof(int) * p.batch, st));
  ASVD_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int) * p.batch, st));
  {
    // stamps start at 1, clean marks at 0: every pair is dirty until verified
    std::vector<int> h_track((size_t)p.batch * (p.nb + (size_t)p.nb * p.nb), 0);
    for (int b = 0; b < p.batch; ++b)
      for (int i = 0; i < p.nb; ++i) h_track[(size_t)b * (p.nb + (size_t)p.nb * p.nb) + i] = 1;
    ASVD_CUDA_CHECK(cudaMemcpyAsync(track, h_track.data(), sizeof(int) * h_track.size(), cudaMemcpyHostToDevice, st));
    ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (!(mode & MODE_SKIP_PREP)) {
    // initial order of the vectors: ascending norm (ASVD_B200_PRESORT=0 keeps the given order).  sigma / norm / perm are
    // free until the epilogue; the slot table lives in the (not yet used) Y buffer.
    const char* ps_env = getenv("ASVD_B200_PRESORT");
    int* slot = nullptr;
    if (!(ps_env && ps_env[0] == '0')) {
      slot = reinterpret_cast<int*>(Y);
      ASVD_CUDA_CHECK(cudaMemsetAsync(sigma, 0, sizeof(float) * p.batch * p.nv_pad, st));
      dim3 gk(p.tall ? (p.n + 31) / 32 : (p.m + 7) / 8, p.batch);
      ASVD_LAUNCH(K_PREP, st, (presort_key_kernel<T><<<gk, 256, 0, st>>>(d_W, scale, ldw, p.m, p.n, p.tall, sigma, p.nv_pad)));
      int P = 1;
      while (P < p.nv_pad) P <<= 1;
      ASVD_LAUNCH(K_PREP, st, (sort_kernel<<<p.batch, 1024, 8 * (size_t)P, st>>>(sigma, norm, perm, p.nv, p.nv_pad, P)));
      ASVD_LAUNCH(K_PREP, st, (presort_slot_kernel<<<dim3((p.nv_pad + 255) / 256, p.batch), 256, 0, st>>>(perm, slot, p.nv_pad)));
    }
    dim3 grid((p.n + 31) / 32, (p.m + 31) / 32, p.batch);
    ASVD_LAUNCH(K_PREP, st, (prep_kernel<T><<<grid, 256, 0, st>>>(d_W, scale, ldw, p.m, p.n, p.tall, X, xs, p.len_pad, slot, p.nv_pad)));
    ASVD_CUDA_CHECK(cudaGetLastError());
  }
  // ASVD_B200_SIMT=1 selects the fp32 SIMT Gram / update kernels (kept as the in-library reference the
  // tensor-core kernels are tested against); default is the tcgen05 path.
  const char* dbg_env = getenv("ASVD_B200_DBG_STEPS");   // timing experiments only: truncates the inner sweep
  const int dbg_steps = dbg_env ? atoi(dbg_env) : JK / 2;
  // ASVD_B200_SOLVE=oddeven / quad forces one of the two inner orderings (A/B runs)
  const char* polish_env = getenv("ASVD_B200_POLISH");
  const int polish_flag = (polish_env && polish_env[0] == 'N') ? 0 : 2;   // ASVD_B200_POLISH=NS: Newton-Schulz tail
  const char* solve_env = getenv("ASVD_B200_SOLVE");
  // Measured: the quad ordering needs about half a sweep more than the odd-even one.  On square problems its faster
  // step wins (-6% wall clock); on 2.7:1 rectangles, where the streaming passes dominate a round, it loses 3%.
  const bool quad_pays = p.len_pad < 2 * p.nv_pad;
  // ASVD_B200_SOLVE=lean (experimental, see svd_solve_quad.cuh): G-only sweep at two CTAs per SM + R replay kernel;
  // default tail only
  // ASVD_B200_LEAN_AUTO=1 (for the measurement that decides the default): lean wherever the quad solve would be picked
  // and the batch has more block pairs than SMs, i.e. where two solve CTAs per SM save a wave.
  const char* auto_env = getenv("ASVD_B200_LEAN_AUTO");
  int sms_dev = 0;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms_dev, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms_dev <= 0)
      sms_dev = 148;
  }
  const bool lean_auto = auto_env && auto_env[0] == '1' && !solve_env && quad_pays && p.batch * p.pairs > sms_dev;
  const bool solve_tri = !dbg_env && polish_flag == 2 && (solve_env ? solve_env[0] == 't' : true);
  const bool solve_lean = !dbg_env && !solve_tri && polish_flag == 2 && ((solve_env && solve_env[0] == 'l') || lean_auto);
  const bool solve_quad = !dbg_env && !solve_tri && !solve_lean && (solve_env ? solve_env[0] == 'q' : quad_pays);
  const char* simt_env = getenv("ASVD_B200_SIMT");
  const bool use_tc = !(simt_env && simt_env[0] == '1');
  // Overlapped half-batches (ASVD_B200_OVERLAP=1).  The solve is a chain of 127 dependent rotation steps on one CTA per
  // block pair: 120 us per launch whatever the batch, and its 512 threads x 128 registers fill an SM, so nothing can
  // share that SM.  A batch of 4 x 4096^2 occupies 128 SMs with it while HBM idles; the streaming passes then want
  // every SM.  Split in two halves on two streams, the solve of one half (64 SMs) runs beside the update + Gram
  // passes of the other (the remaining SMs); the solves hand a token back and forth so that the two halves stay in
  // anti-phase instead of drifting into lock-step.  Every buffer is [batch]-major, so a half is the same kernels on
  // offset pointers (and its own tensor maps); the results are bitwise those of the single-stream schedule.
  struct Part { int b0, nb; cudaStream_t s; CUtensorMap tmK, tmMN; };
  Part parts[2];
  int n_parts = 1;
  static int sms_total = 0;
  if (!sms_total) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms_total, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms_total <= 0)
      sms_total = 148;
  }
  struct SideStreams { cudaStream_t side[2]; cudaEvent_t fork, join[2], solve[2]; };
  static SideStreams side_dev[ASVD_MAX_DEVICES] = {};
  SideStreams& ss = side_dev[dev_slot];
  cudaStream_t* side = ss.side;
  cudaEvent_t &ev_fork = ss.fork, *ev_join = ss.join, *ev_solve = ss.solve;
  {
    const char* ov_env = getenv("ASVD_B200_OVERLAP");
    const bool want = ov_env && ov_env[0] == '1';
    const int hb = (p.batch + 1) / 2;
    if (want && use_tc && !g_prof_on && p.batch >= 2 && 2 * hb * p.pairs <= sms_total) {
      if (!side[0]) {
        for (int i = 0; i < 2; ++i) {
          ASVD_CUDA_CHECK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
          ASVD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
          ASVD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_solve[i], cudaEventDisableTiming));
        }
        ASVD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      }
      n_parts = 2;
      parts[0].b0 = 0; parts[0].nb = hb; parts[0].s = side[0];
      parts[1].b0 = hb; parts[1].nb = p.batch - hb; parts[1].s = side[1];
    } else {
      parts[0].b0 = 0; parts[0].nb = p.batch; parts[0].s = st;
    }
  }
  if (use_tc) {
    for (int h = 0; h < n_parts; ++h)
      if (!tc::make_x_tmap(&parts[h].tmK, X + parts[h].b0 * xs, parts[h].nb, p.nv_pad, p.len_pad) ||
          !tc::make_x_tmap_mn(&parts[h].tmMN, X + parts[h].b0 * xs, parts[h].nb, p.nv_pad, p.len_pad)) {
        set_error("cuTensorMapEncodeTiled failed");
        return ASVD_ERR_CUDA;
      }
  }
  // streaming passes of a half count on the SMs the other half's solve leaves free
  struct BudgetGuard { ~BudgetGuard() { tc::set_sm_budget(0); } } budget_guard;
  if (n_parts == 2) tc::set_sm_budget(sms_total - parts[0].nb * p.pairs);
  // maxoff: per part [nb] largest cosine bits, then [nb] near-converged pair counts (the solve kernel indexes by its
  // own gridDim.y); one part = the layout [batch][batch]
  auto cos_idx = [&](int b) { const Part& q = parts[(n_parts == 2 && b >= parts[1].b0) ? 1 : 0]; return 2 * q.b0 + (b - q.b0); };
  auto near_idx = [&](int b) { const Part& q = parts[(n_parts == 2 && b >= parts[1].b0) ? 1 : 0]; return 2 * q.b0 + q.nb + (b - q.b0); };
  const int64_t track_stride = p.nb + (int64_t)p.nb * p.nb;
  const char* pre_env = getenv("ASVD_B200_TOL_PRE");   // default 5 tol = 2e-5: `tol` itself sits on the fp32 plateau, where passing is a coin flip
  float tol_pre = pre_env ? (float)atof(pre_env) : 5.f * tol;
  if (conv_tol > 0.f) tol_pre = conv_tol;          // experiments (ASVD_B200_INNER_TOL): looser tolerance for the square stage
  if (tol_pre < tol) tol_pre = tol;
  // When a matrix leaves the single-pass TF32 Gram for the fp32-accurate one: after a sweep in which at least this share
  // of its pair visits met cosines below 1e-2.  The accurate pass is no longer free (the split keeps four warps busy:
  // 425 vs 289 us per launch at 27 x 4096^2) and only matters once MOST pairs are small; the first few pairs get there
  // around sweep 4, the majority around sweep 7.  Measured (profiles/r02_ab_near_pct_tri*.jsonl): 0 / 50 / 70 / 80 %:
  // 1120 / 1077 / 1074 / 1079 ms per 27 x 4096^2, same sweep counts and sigma errors; 2048^2 -3 %, 4096x11008 -2 %.
  const char* near_env = getenv("ASVD_B200_NEAR_PCT");
  const unsigned near_min = (unsigned)((near_env ? atof(near_env) : 50.0) * 0.01 * p.rounds * p.pairs);
  std::vector<unsigned> h_maxoff(2 * (size_t)p.batch);
  std::vector<int> h_done(p.batch, 0), h_sweeps(p.batch, 0);
  bool gave_up = false;
  int sweep = 0;
  bool all_done = false;
  // PER MATRIX: some pair of it was already (nearly) orthogonal in an earlier sweep -- or the vectors arrive
  // pre-conditioned, where the single-pass Gram could not see the cosines that are left.  Kept per matrix (and handed to
  // the kernels as an array) so that a weight's arithmetic never depends on which other weights share its batch: the
  // factors are bitwise the same for every batch size, hence for every world size of a sharded run.
  std::vector<char> near_seen(p.batch, (mode & MODE_SKIP_PREP) != 0 ? 1 : 0);
  std::vector<int> h_prec(p.batch, -1), h_prec_new(p.batch, 0);
  for (; sweep < max_sweeps && !all_done; ++sweep) {
    // single-pass TF32 Gram only while every pair of the matrix still needs work; afterwards the 3-term split
    // (fp32-accurate), without which the threshold test could not skip converged pairs nor certify convergence
    for (int b = 0; b < p.batch; ++b) h_prec_new[b] = (!use_tc || near_seen[b]) ? 1 : 0;
    if (h_prec_new != h_prec) {
      h_prec = h_prec_new;
      ASVD_CUDA_CHECK(cudaMemcpyAsync(prec, h_prec.data(), sizeof(int) * p.batch, cudaMemcpyHostToDevice, st));
      ASVD_CUDA_CHECK(cudaStreamSynchronize(st));        // pageable source: the vector may change before a deferred copy reads it
    }
    ASVD_CUDA_CHECK(cudaMemsetAsync(maxoff, 0, sizeof(unsigned) * 2 * p.batch, st));
    if (n_parts == 2) {
      ASVD_CUDA_CHECK(cudaEventRecord(ev_fork, st));
      for (int h = 0; h < 2; ++h) ASVD_CUDA_CHECK(cudaStreamWaitEvent(parts[h].s, ev_fork, 0));
    }
    for (int r = 0; r < p.rounds; ++r) {
      const int2* pr = d_pairs + (size_t)r * p.pairs;
      const int round_stamp = 2 + sweep * p.rounds + r;
      const int half_gram = use_tc ? 1 : 0;   // gram_tc_kernel's precise mode stores T, G = T + T^T (kernels AND it with prec[b])
      for (int h = 0; h < n_parts; ++h) {
        const Part& q = parts[h];
        cudaStream_t s = q.s;
        float* Gq = G + (size_t)q.b0 * p.pairs * p.chunks * JK * JK;
        float* Rq = R + (size_t)q.b0 * p.pairs * JK * JK;
        int* flagq = flag + (size_t)q.b0 * p.pairs;
        int* doneq = done + q.b0;
        int* statusq = status + q.b0;
        int* trackq = track + q.b0 * track_stride;
        unsigned* maxoffq = maxoff + 2 * q.b0;
        const int* precq = prec + q.b0;
        float* Xq = X + q.b0 * xs;
        if (use_tc) {
          // one launch per Gram mode present among the part's running matrices (usually one)
          bool any_mode[2] = {false, false};
          for (int b = q.b0; b < q.b0 + q.nb; ++b)
            if (!h_done[b]) any_mode[h_prec[b] ? 1 : 0] = true;
          for (int md = 0; md < 2; ++md)
            if (any_mode[md])
              ASVD_LAUNCH(K_GRAM, s, ASVD_CUDA_CHECK(tc::launch_gram_tc(q.tmK, pr, p.pairs, p.chunks, p.chunk_cols, p.len_pad, p.nv_pad, q.nb, Gq, doneq, md, precq, trackq, s)));
        } else {
          ASVD_LAUNCH(K_GRAM, s, (gram_kernel<<<dim3(p.chunks, p.pairs, q.nb), 256, 0, s>>>(Xq, xs, p.len_pad, pr, p.len_pad, p.chunks, p.pairs, Gq, doneq, trackq, p.nb, p.chunk_cols)));
        }
        // the token: this half's solve starts when the other half's latest solve has finished
        if (n_parts == 2 && !(r == 0 && h == 0)) ASVD_CUDA_CHECK(cudaStreamWaitEvent(s, ev_solve[h ^ 1], 0));
        if (solve_tri) {
          float* auxq = reinterpret_cast<float*>(ws + p.off_aux) + (size_t)q.b0 * p.pairs * QAUX_FLOATS;
          ASVD_LAUNCH(K_SOLVE, s, (solve_tri_g_kernel<<<dim3(p.pairs, q.nb), TRI_THREADS, SOLVET_SMEM, s>>>(Gq, p.chunks, p.pairs, auxq, flagq, maxoffq, statusq, doneq, tol, pr, trackq, p.nb, round_stamp, precq, half_gram)));
          ASVD_LAUNCH(K_SOLVE, s, (solve_tri_r_kernel<<<dim3(p.pairs, q.nb), 256, SOLVETR_SMEM, s>>>(auxq, p.pairs, Rq, flagq, doneq)));
        } else if (solve_lean) {
          float* auxq = reinterpret_cast<float*>(ws + p.off_aux) + (size_t)q.b0 * p.pairs * QAUX_FLOATS;
          ASVD_LAUNCH(K_SOLVE, s, (solve_quad_g_kernel<<<dim3(p.pairs, q.nb), 256, SOLVEQG_SMEM, s>>>(Gq, p.chunks, p.pairs, auxq, flagq, maxoffq, statusq, doneq, tol, pr, trackq, p.nb, round_stamp, precq, half_gram)));
          // ASVD_B200_SOLVE=leanr: the same G kernel with the triangular solve's replay kernel -- the bitwise bridge between
          // solve_tri_r_kernel and the monolithic quad kernel (tests)
          if (solve_env && solve_env[0] == 'l' && solve_env[1] == 'e' && solve_env[2] == 'a' && solve_env[3] == 'n' && solve_env[4] == 'r')
            ASVD_LAUNCH(K_SOLVE, s, (solve_tri_r_kernel<<<dim3(p.pairs, q.nb), 256, SOLVETR_SMEM, s>>>(auxq, p.pairs, Rq, flagq, doneq)));
          else
          ASVD_LAUNCH(K_SOLVE, s, (solve_quad_r_kernel<<<dim3(p.pairs, q.nb), 256, SOLVEQR_SMEM, s>>>(auxq, p.pairs, Rq, flagq, doneq)));
        } else if (solve_quad)
          ASVD_LAUNCH(K_SOLVE, s, (solve_quad_kernel<<<dim3(p.pairs, q.nb), SOLVE_THREADS, SOLVEQ_SMEM, s>>>(Gq, p.chunks, p.pairs, Rq, flagq, maxoffq, statusq, doneq, tol, polish_flag, pr, trackq, p.nb, round_stamp, precq, half_gram)));
        else
          ASVD_LAUNCH(K_SOLVE, s, (solve_kernel<<<dim3(p.pairs, q.nb), SOLVE_THREADS, SOLVE_SMEM, s>>>(Gq, p.chunks, p.pairs, Rq, flagq, maxoffq, statusq, doneq, tol, polish_flag, dbg_steps, pr, trackq, p.nb, round_stamp, precq, half_gram)));
        if (n_parts == 2) ASVD_CUDA_CHECK(cudaEventRecord(ev_solve[h], s));
        if (use_tc) {
          ASVD_LAUNCH(K_UPDATE, s, ASVD_CUDA_CHECK(tc::launch_update_tc(q.tmMN, Xq, xs, p.len_pad, pr, p.pairs, p.nv_pad, p.len_pad, q.nb, Rq, flagq, doneq, s)));
        } else {
          const int ctas_x = (p.len_pad / 128 + UPD_TILES - 1) / UPD_TILES;
          ASVD_LAUNCH(K_UPDATE, s, (update_kernel<<<dim3(ctas_x, p.pairs, q.nb), 256, UPDATE_SMEM, s>>>(Xq, xs, p.len_pad, pr, p.len_pad, p.pairs, Rq, flagq, doneq)));
        }
      }
    }
    if (n_parts == 2) {
      for (int h = 0; h < 2; ++h) {
        ASVD_CUDA_CHECK(cudaEventRecord(ev_join[h], parts[h].s));
        ASVD_CUDA_CHECK(cudaStreamWaitEvent(st, ev_join[h], 0));
      }
    }
    ASVD_CUDA_CHECK(cudaGetLastError());
    ASVD_CUDA_CHECK(cudaMemcpyAsync(h_maxoff.data(), maxoff, sizeof(unsigned) * 2 * p.batch, cudaMemcpyDeviceToHost, st));
    ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
    all_done = true;
    bool changed = false;
    float worst = 0.f, best = 1.f;
    for (int b = 0; b < p.batch; ++b) {
      if (h_done[b]) continue;
      float mo;
      memcpy(&mo, &h_maxoff[cos_idx(b)], 4);
      h_sweeps[b] = sweep + 1;
      // a sweep measured with the single-pass Gram cannot certify convergence
      // mo is the largest cosine met at VISIT time, i.e. before this sweep's own rotations (every pair at or above
      // `tol` was rotated).  Measured traces: once below 1e-3 a sweep shrinks it by >= 5x (26x from 1e-4), down to the
      // fp32 plateau of 2-4e-6, so a sweep that started below tol_pre leaves the vectors orthogonal to ~2e-5 or
      // better; asking for a further verification sweep below `tol` (which sits ON that plateau) costs one to two
      // sweeps and changes sigma by < 1e-5 relative.
      bool finished = mo < tol_pre && h_prec[b];
      if ((mode & MODE_STOP_AT_VECTORS) && !finished) {
        // pre-conditioning stage: the square problem runs on the SQUARED spectrum; on ill-conditioned weights its fp32
        // cosines never settle.  The largest cosine is no guide (it RISES for a dozen sweeps on a healthy problem while
        // clustered values sort themselves out); the number of block pairs already below 1e-2 is: after seven sweeps a
        // healthy problem has 16-65 % of its pairs there (measured: the Gram matrices of both Llama rectangles, 332 and
        // 1329 of 2016), power-law spectra with kappa >= 5e3 have 2-4 of 2016; the bar is 5 %.  The stage's unit vectors are an
        // orthogonal transform only once it HAS converged, so there is no partial credit: give up early and let the
        // caller take the direct path.
        const unsigned near_pairs = h_maxoff[near_idx(b)];
        static const char* gu_env = getenv("ASVD_B200_INNER_GIVEUP");          // 0: never give up (diagnostics)
        if (!(gu_env && gu_env[0] == '0') &&
            ((sweep == 3 && mo > 0.25f) ||          // healthy: <= 0.09 after four sweeps; stalled: 0.45-0.9
             (sweep == 6 && near_pairs * 20u < (unsigned)(p.rounds * p.pairs)) || sweep >= 19))
          gave_up = true;
      }
      if (finished) { h_done[b] = 1; changed = true; }
      else { all_done = false; worst = fmaxf(worst, mo); best = fminf(best, mo); }
      // ASVD_B200_NEAR_PCT (experiments): stay with the single-pass Gram until that percentage of a matrix's block pairs
      // is below 1e-2 (default 0: the first such pair switches to the precise Gram)
      if (h_maxoff[near_idx(b)] > near_min) near_seen[b] = 1;
    }
    (void)worst; (void)best;
    if (getenv("ASVD_B200_TRACE")) {                 // diagnostic: convergence trace, one line per sweep
      fprintf(stderr, "sweep %2d precise", sweep + 1);
      for (int b = 0; b < p.batch; ++b) fprintf(stderr, " %d", h_prec[b]);
      fprintf(stderr, "  max|cos|:");
      for (int b = 0; b < p.batch; ++b) { float mo; memcpy(&mo, &h_maxoff[cos_idx(b)], 4); fprintf(stderr, " %.3e", mo); }
      fprintf(stderr, "  near-orthogonal pairs:");
      for (int b = 0; b < p.batch; ++b) fprintf(stderr, " %u", h_maxoff[near_idx(b)]);
      fprintf(stderr, "\n");
    }
    if (gave_up) break;
    if (changed && !all_done)
      ASVD_CUDA_CHECK(cudaMemcpyAsync(done, h_done.data(), sizeof(int) * p.batch, cudaMemcpyHostToDevice, st));
  }
  if ((mode & MODE_STOP_AT_VECTORS) && (gave_up || !all_done)) {
    if (sweeps_out) for (int b = 0; b < p.batch; ++b) sweeps_out[b] = h_sweeps[b];
    return ASVD_ERR_NOT_CONVERGED;
  }
  // rows of X are sigma_j u_j: normalise, recover the other factor from the original weight, anchor sigma to it
  ASVD_LAUNCH(K_FINAL, st, (untile_kernel<<<dim3((p.len_pad / 4 + 255) / 256, p.nv_pad, p.batch), 256, 0, st>>>(X, Xr, xs, p.nv_pad, p.len_pad)));
  ASVD_LAUNCH(K_FINAL, st, (rownorm_kernel<<<dim3(p.nv_pad, p.batch), 256, 0, st>>>(Xr, xs, p.len_pad, p.len_pad, 1, sigma, p.nv_pad, status, nullptr)));
  ASVD_CUDA_CHECK(cudaGetLastError());
  if (mode & MODE_STOP_AT_VECTORS) {
    std::vector<int> hs(p.batch);
    ASVD_CUDA_CHECK(cudaMemcpyAsync(hs.data(), status, sizeof(int) * p.batch, cudaMemcpyDeviceToHost, st));
    ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
    prof_collect();
    if (sweeps_out) for (int b = 0; b < p.batch; ++b) sweeps_out[b] = h_sweeps[b];
    for (int b = 0; b < p.batch; ++b)
      if (hs[b]) return ASVD_ERR_NONFINITE;
    return ASVD_OK;
  }
  {
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    gb.Bptrs = d_W;
    gb.strideA = xs;
    gb.strideC = (int64_t)p.nv_pad * p.ldy;
    ASVD_CUDA_CHECK(cudaMemsetAsync(Y, 0, sizeof(float) * gb.strideC * p.batch, st));
    // per-matrix scale vectors live contiguously in the workspace: pass through pointer-free strides by
    // launching one GEMM per matrix when scales differ (batch is small); z-batched otherwise.
    // 16-bit weights: tensor-core plane GEMM (ASVD_B200_RECOVER=simt keeps the fp32 SIMT GEMM, which fp32 weights
    // always use)
    const char* rec_env = getenv("ASVD_B200_RECOVER");
    const bool recover_tc = use_tc && !std::is_same<T, float>::value && !(rec_env && rec_env[0] == 's') &&
                            p.len_pad <= RECOVER_MAX_SPLITS * 1024;      // longer vectors: runs would exceed 1024 columns
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(ws + p.off_As);
    __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(ws + p.off_Bs);
    if (recover_tc) ASVD_CUDA_CHECK(cudaMemsetAsync(Bs, 0, sizeof(__nv_bfloat16) * 2 * (size_t)p.nv_pad * p.len_pad, st));
    for (int b = 0; b < p.batch; ++b) {
      GemmBatch g1 = gb;
      g1.Bptrs = d_W + b;
      const float* sb = scale + (int64_t)b * p.n;
      cudaError_t e;
      if (recover_tc) {
        prof_begin(K_FINAL, st);
        split_x_kernel<<<dim3((p.len_pad / 4 + 255) / 256, p.nv_pad), 256, 0, st>>>(Xr + b * xs, p.tall ? nullptr : sb, p.nv_pad,
                                                                                 p.len_pad, p.len, As);
        split_w_kernel<T><<<dim3((p.n + 31) / 32, (p.m + 31) / 32), 256, 0, st>>>(reinterpret_cast<const T*>(h_W[b]), ldw, p.m, p.n,
                                                                                 p.tall, p.len_pad, Bs);
        // the contraction is cut into runs of ~768 columns, each into its own fp32 partial result (see gemm_planes_f32
        // on the truncating accumulation of the tensor core), summed afterwards in a fixed order
        int nsplit = (p.len_pad + RECOVER_RUN - 1) / RECOVER_RUN;
        if (nsplit > RECOVER_MAX_SPLITS) nsplit = RECOVER_MAX_SPLITS;
        const int run = (int)round_up((p.len_pad + nsplit - 1) / nsplit, 64);
        float* Yp = reinterpret_cast<float*>(ws + p.off_Yp);
        const int64_t ystride = (int64_t)p.nv_pad * p.ldy;
        int rc = 0, used = 0;
        for (int k0 = 0; k0 < p.len_pad && rc == 0; k0 += run, ++used)
          rc = tc::gemm_planes_f32(As, 3 * (int64_t)p.len_pad, 3, Bs, 2 * (int64_t)p.len_pad,
                                   std::is_same<T, __nv_bfloat16>::value ? 1 : 2, Yp + used * ystride, p.ldy, p.nv_pad, p.nv,
                                   p.len_pad, k0, (p.len_pad - k0 < run) ? p.len_pad - k0 : run, nullptr, st);
        if (rc == 0)
          sum_partials_kernel<<<(unsigned)((ystride / 4 + 255) / 256), 256, 0, st>>>(Yp, used, ystride, Y + b * gb.strideC, ystride,
                                                                                 p.ldy, p.tall ? sb : nullptr, p.nv, nullptr, 0);
        prof_end(K_FINAL, st);
        if (rc < 0) { set_error("recovery GEMM launch failed (%d)", rc); return ASVD_ERR_CUDA; }
        if (rc == 0) continue;
      }
      prof_begin(K_FINAL, st);
      if (p.tall)   // Y[j][l] = sum_i Xhat[j][i] W[i][l] * s[l]
        e = launch_gemm128<float, T, float, true>(Xr + b * xs, p.len_pad, (const T*)nullptr, ldw, Y + b * gb.strideC, p.ldy,
                                                  p.nv_pad, p.n, p.m, nullptr, sb, nullptr, 1, g1, st);
      else          // Y[j][i] = sum_l Xhat[j][l] s[l] W[i][l]
        e = launch_gemm128<float, T, float, false>(Xr + b * xs, p.len_pad, (const T*)nullptr, ldw, Y + b * gb.strideC, p.ldy,
                                                   p.nv_pad, p.m, p.n, sb, nullptr, nullptr, 1, g1, st);
      prof_end(K_FINAL, st);
      ASVD_CUDA_CHECK(e);
    }
  }
  ASVD_LAUNCH(K_FINAL, st, (rownorm_kernel<<<dim3(p.nv_pad, p.batch), 256, 0, st>>>(Y, (int64_t)p.nv_pad * p.ldy, p.ldy, p.nv, 0, norm, p.nv_pad, status, sigma)));
  ASVD_CUDA_CHECK(cudaGetLastError());
  {
    int P = 1;
    while (P < p.nv_pad) P <<= 1;
    ASVD_LAUNCH(K_FINAL, st, (sort_kernel<<<p.batch, 1024, 8 * (size_t)P, st>>>(norm, sigma, perm, p.nv, p.nv_pad, P)));
    ASVD_CUDA_CHECK(cudaGetLastError());
  }
  std::vector<int> h_status(p.batch);
  ASVD_CUDA_CHECK(cudaMemcpyAsync(h_status.data(), status, sizeof(int) * p.batch, cudaMemcpyDeviceToHost, st));
  ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
  prof_collect();
  if (sweeps_out) for (int b = 0; b < p.batch; ++b) sweeps_out[b] = h_sweeps[b];
  for (int b = 0; b < p.batch; ++b)
    if (h_status[b]) { set_error("non-finite values in weight %d of the batch (or its scale)", b); return ASVD_ERR_NONFINITE; }
  if (!all_done) { set_error("sweep limit %d reached above tolerance %g", max_sweeps, (double)tol); return ASVD_ERR_NOT_CONVERGED; }
  return ASVD_OK;
}

__global__ void count_nonfinite_kernel(const float* __restrict__ p, int64_t n, int* __restrict__ out) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(fabsf(p[i]) <= FLT_MAX)) ++bad;
  if (bad) atomicAdd(out, bad);
}
static void trace_nonfinite(const char* what, const float* p, int64_t n, int* scratch, cudaStream_t st) {
  if (!getenv("ASVD_B200_TRACE")) return;
  cudaMemsetAsync(scratch, 0, sizeof(int), st);
  count_nonfinite_kernel<<<592, 256, 0, st>>>(p, n, scratch);
  int h = 0;
  cudaMemcpyAsync(&h, scratch, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  fprintf(stderr, "  [gram-pre] %s: %d non-finite of %lld\n", what, h, (long long)n);
}

// ------------------------------------------------------------------------------------------------ Gram pre-conditioner
// Rectangular weights (long vectors, few of them): every round of the sweeps streams vectors 2.7x longer than there
// are of them although the rotations only depend on their nv x nv Gram matrix.  So:
//   1. G = X X^T exactly (bf16 plane GEMM from the original weight: b_i b_j products, scale applied in fp32);
//   2. the same block-Jacobi machinery on G itself, an nv x nv problem: its unit vectors are the eigenvectors Q of G,
//      i.e. the rotation the long problem is looking for, to the accuracy the squared conditioning allows;
//   3. X1 = Q^T X, again as an exact plane GEMM from the ORIGINAL weight (nothing of step 2's error survives except
//      through Q being slightly off), written in the working layout;
//   4. the ordinary sweeps on X1 finish the job -- typically one sweep with a few rotations and the verification --
//      under the ordinary convergence test, so the result has the accuracy of the direct path.
// (the Gram shortcut is acceptable as a pre-conditioner only.)
template <typename T>
static int gram_precondition(const SvdPlan& p, int64_t ldw, unsigned char* ws, float tol, int max_sweeps, cudaStream_t st,
                             const void* const* h_W, int* inner_sweeps) {
  const SvdPlan pi = make_plan(p.nv, p.nv, p.batch, false);
  unsigned char* wi = ws + p.off_inner;
  float* Gm = reinterpret_cast<float*>(ws + p.off_Gm);
  float* X = reinterpret_cast<float*>(ws + p.off_X);
  float* Xr = reinterpret_cast<float*>(ws + p.off_Xr);
  float* scale = reinterpret_cast<float*>(ws + p.off_scale);
  float* Yp = reinterpret_cast<float*>(ws + p.off_Yp);
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(ws + p.off_As);
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(ws + p.off_Bs);
  const int64_t gs = (int64_t)p.nv_pad * p.nv_pad, xs = (int64_t)p.nv_pad * p.len_pad;
  const int b_planes = std::is_same<T, __nv_bfloat16>::value ? 1 : 2;
  auto runs = [](int K) {
    int ns = (K + RECOVER_RUN - 1) / RECOVER_RUN;
    if (ns > RECOVER_MAX_SPLITS) ns = RECOVER_MAX_SPLITS;
    return (int)round_up((K + ns - 1) / ns, 64);
  };
  // ---- 1. Gram matrices
  for (int b = 0; b < p.batch; ++b) {
    const T* W = reinterpret_cast<const T*>(h_W[b]);
    const float* sb = scale + (int64_t)b * p.n;
    prof_begin(K_PREP, st);
    ASVD_CUDA_CHECK(cudaMemsetAsync(Bs, 0, sizeof(__nv_bfloat16) * 2 * (size_t)p.nv_pad * p.len_pad, st));
    split_w_kernel<T><<<dim3((p.n + 31) / 32, (p.m + 31) / 32), 256, 0, st>>>(W, ldw, p.m, p.n, p.tall, p.len_pad, Bs);
    const __nv_bfloat16* A = Bs;
    int a_planes = b_planes;
    int64_t lda = 2 * (int64_t)p.len_pad;
    if (!p.tall) {
      ASVD_CUDA_CHECK(cudaMemsetAsync(As, 0, sizeof(__nv_bfloat16) * 3 * (size_t)p.nv_pad * p.len_pad, st));
      split_ws2_kernel<T><<<dim3((p.n + 255) / 256, p.m), 256, 0, st>>>(W, ldw, p.m, p.n, sb, p.len_pad, As);
      A = As; a_planes = 3; lda = 3 * (int64_t)p.len_pad;
    }
    const int run = runs(p.len_pad);
    int rc = 0, used = 0;
    for (int k0 = 0; k0 < p.len_pad && rc == 0; k0 += run, ++used)
      rc = tc::gemm_planes_f32(A, lda, a_planes, Bs, 2 * (int64_t)p.len_pad, b_planes, Yp + used * gs, p.nv_pad, p.nv_pad, p.nv,
                               p.len_pad, k0, (p.len_pad - k0 < run) ? p.len_pad - k0 : run, nullptr, st);
    if (rc != 0) { prof_end(K_PREP, st); set_error("Gram GEMM launch failed (%d)", rc); return ASVD_ERR_CUDA; }
    sum_partials_kernel<<<(unsigned)((gs / 4 + 255) / 256), 256, 0, st>>>(Yp, used, gs, Gm + b * gs, gs, p.nv_pad,
                                                                      p.tall ? sb : nullptr, p.nv, p.tall ? sb : nullptr, p.nv);
    prof_end(K_PREP, st);
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  trace_nonfinite("G", Gm, gs * p.batch, reinterpret_cast<int*>(ws + p.off_flag), st);
  // ---- 2. the square problem on G
  {
    std::vector<const void*> gp(2 * (size_t)p.batch, nullptr);
    for (int b = 0; b < p.batch; ++b) gp[b] = Gm + b * gs;
    ASVD_CUDA_CHECK(cudaMemcpyAsync(wi + pi.off_ptrs, gp.data(), sizeof(void*) * 2 * p.batch, cudaMemcpyHostToDevice, st));
    std::vector<int2> tab;
    build_pair_table(pi, tab);
    if (!tab.empty())
      ASVD_CUDA_CHECK(cudaMemcpyAsync(wi + pi.off_pairs, tab.data(), sizeof(int2) * tab.size(), cudaMemcpyHostToDevice, st));
    float* iscale = reinterpret_cast<float*>(wi + pi.off_scale);
    const int64_t ne = (int64_t)p.batch * pi.n;
    fill_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(iscale, 1.f, ne);
    ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
    const char* it_env = getenv("ASVD_B200_INNER_TOL");
    // The unit vectors of the square stage are used as an orthogonal transform, so they have to BE orthonormal: the
    // stage runs to the ordinary tolerance.  (Stopping at 1e-3 is 3 % faster on the Gaussian workload -- 273 / 343 ms
    // against 282 / 349 ms per batch of four -- but leaves Q Q^T - I at 1e-3, which goes straight into sigma.)
    const float inner_tol = it_env ? (float)atof(it_env) : 0.f;
    const int rc = run_svd<float>(pi, p.nv_pad, wi, tol, max_sweeps, inner_sweeps, st, gp.data(), MODE_STOP_AT_VECTORS, inner_tol);
    // squared dynamic range left fp32, or no convergence on the squared spectrum: the caller takes the direct path
    if (rc == ASVD_ERR_NONFINITE || rc == ASVD_ERR_NOT_CONVERGED) return rc;
    if (rc != ASVD_OK) return rc;
  }
  // ---- 3. X1 = Q^T X from the original weight, into the working layout
  const float* Q = reinterpret_cast<const float*>(wi + pi.off_Xr);
  const int64_t qs = (int64_t)pi.nv_pad * pi.len_pad;
  trace_nonfinite("Q", Q, qs * p.batch, reinterpret_cast<int*>(ws + p.off_flag), st);
  for (int b = 0; b < p.batch; ++b) {
    const T* W = reinterpret_cast<const T*>(h_W[b]);
    const float* sb = scale + (int64_t)b * p.n;
    prof_begin(K_PREP, st);
    ASVD_CUDA_CHECK(cudaMemsetAsync(Bs, 0, sizeof(__nv_bfloat16) * 2 * (size_t)p.nv_pad * p.len_pad, st));
    // rows = positions along the long dimension, contraction along the short one
    split_w_kernel<T><<<dim3((p.n + 31) / 32, (p.m + 31) / 32), 256, 0, st>>>(W, ldw, p.m, p.n, p.tall ? 0 : 1, pi.len_pad, Bs);
    split_x_kernel<<<dim3((pi.len_pad / 4 + 255) / 256, pi.nv_pad), 256, 0, st>>>(Q + b * qs, p.tall ? sb : nullptr, pi.nv_pad,
                                                                               pi.len_pad, p.nv, As);
    const int run = runs(pi.len_pad);
    int rc = 0, used = 0;
    for (int k0 = 0; k0 < pi.len_pad && rc == 0; k0 += run, ++used)
      rc = tc::gemm_planes_f32(As, 3 * (int64_t)pi.len_pad, 3, Bs, 2 * (int64_t)pi.len_pad, b_planes, Yp + used * xs, p.len_pad,
                               p.nv_pad, p.len, pi.len_pad, k0, (pi.len_pad - k0 < run) ? pi.len_pad - k0 : run, nullptr, st);
    if (rc != 0) { prof_end(K_PREP, st); set_error("pre-conditioning GEMM launch failed (%d)", rc); return ASVD_ERR_CUDA; }
    sum_partials_kernel<<<(unsigned)((xs / 4 + 255) / 256), 256, 0, st>>>(Yp, used, xs, Xr + b * xs, xs, p.len_pad,
                                                                      p.tall ? nullptr : sb, p.len, nullptr, 0);
    tile_kernel<<<dim3((p.len_pad / 4 + 255) / 256, p.nv_pad), 256, 0, st>>>(Xr + b * xs, X + b * xs, p.nv_pad, p.len_pad);
    prof_end(K_PREP, st);
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  trace_nonfinite("X1", Xr, xs * p.batch, reinterpret_cast<int*>(ws + p.off_flag), st);
  return ASVD_OK;
}

}  // namespace asvd

using namespace asvd;

template <typename TC>
static int do_extract(const SvdPlan& p, const unsigned char* ws, int b, int r, int fuse, TC* A, int64_t lda, TC* B,
                      int64_t ldb, cudaStream_t st) {
  const float* X = reinterpret_cast<const float*>(ws + p.off_Xr) + (int64_t)b * p.nv_pad * p.len_pad;   // row-major copy
  const float* Y = reinterpret_cast<const float*>(ws + p.off_Y) + (int64_t)b * p.nv_pad * p.ldy;
  const float* sigma = reinterpret_cast<const float*>(ws + p.off_sigma) + (int64_t)b * p.nv_pad;
  const int* perm = reinterpret_cast<const int*>(ws + p.off_perm) + (int64_t)b * p.nv_pad;
  const float* scale = reinterpret_cast<const float*>(ws + p.off_scale) + (int64_t)b * p.n;
  // exponent of sigma carried by A: UV 1/2, U 1, V 0 (modules/svd_linear.py:16-24)
  const float a = fuse == ASVD_FUSE_UV ? 0.5f : (fuse == ASVD_FUSE_U ? 1.f : 0.f);
  dim3 gA((r + 31) / 32, (p.m + 31) / 32), gB((p.n + 255) / 256, r);
  if (p.tall) {
    // X rows are unit u_j (length m); Y rows are sigma_j v_j^T diag(s) (length n)
    ASVD_LAUNCH(K_EXTRACT, st, (extract_cols_kernel<TC><<<gA, 256, 0, st>>>(X, p.len_pad, perm, sigma, a, p.m, r, A, lda)));
    ASVD_LAUNCH(K_EXTRACT, st, (extract_rows_kernel<TC><<<gB, 256, 0, st>>>(Y, p.ldy, perm, sigma, -a, scale, p.n, r, B, ldb)));
  } else {
    // X rows are unit v_j^T diag(s)... scaled space (length n); Y rows are sigma_j u_j (length m)
    ASVD_LAUNCH(K_EXTRACT, st, (extract_cols_kernel<TC><<<gA, 256, 0, st>>>(Y, p.ldy, perm, sigma, a - 1.f, p.m, r, A, lda)));
    ASVD_LAUNCH(K_EXTRACT, st, (extract_rows_kernel<TC><<<gB, 256, 0, st>>>(X, p.len_pad, perm, sigma, 1.f - a, scale, p.n, r, B, ldb)));
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  return ASVD_OK;
}

extern "C" {

int asvd_version(void) { return ASVD_B200_VERSION; }
#ifdef ASVD_SOLVE_TIMING
int asvd_debug_solve_timing(unsigned long long* out8) {       // timing builds only (scripts/solve_timing.py)
  return cudaMemcpyFromSymbol(out8, asvd::g_solve_timing, 8 * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}
#endif
#ifdef ASVD_SOLVE_TIMING
int asvd_debug_tri_timing(unsigned long long* out16) {         // timing builds only (scripts/tri_timing.py)
  return cudaMemcpyFromSymbol(out16, asvd::g_tri_timing, 16 * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}
#endif
const char* asvd_last_error(void) { return asvd::last_error(); }

void asvd_profile_enable(int on) { asvd::prof_reset(on != 0); }
int asvd_profile_read(double* ms_out, uint64_t* launches_out) {
  asvd::prof_collect();
  for (int k = 0; k < asvd::K_COUNT; ++k) {
    if (ms_out) ms_out[k] = asvd::prof_ms(k);
    if (launches_out) launches_out[k] = asvd::launches(k);
  }
  return asvd::K_COUNT;
}
uint64_t asvd_launch_count(void) {
  uint64_t t = 0;
  for (int k = 0; k < asvd::K_COUNT; ++k) t += asvd::launches(k);
  return t;
}

int asvd_rank_for_ratio(int64_t out_features, int64_t in_features, double param_ratio, int rank_align) {
  // modules/svd_linear.py:39-44: int(n_params * ratio) // (in + out), then ceil to a multiple of rank_align
  double prod = (double)(out_features * in_features) * param_ratio;
  int64_t compressed = (int64_t)prod;   // python int() truncates toward zero
  int64_t rank = compressed / (in_features + out_features);
  if (rank_align > 1) rank = (rank + rank_align - 1) / rank_align * rank_align;
  return (int)rank;
}

size_t asvd_svd_workspace_bytes(int m, int n, int batch) {
  if (m <= 0 || n <= 0 || batch <= 0) return 0;
  return make_plan(m, n, batch).bytes;
}

int asvd_scaled_svd(const void* const* W_host_ptrs, int w_dtype, int64_t ldw, int m, int n, int batch,
                    const float* const* scale_host_ptrs, void* workspace, size_t workspace_bytes, float tol,
                    int max_sweeps, int* sweeps_out_host, void* stream) {
  ASVD_REQUIRE(W_host_ptrs && workspace, "null pointer");
  ASVD_REQUIRE(m > 0 && n > 0 && batch > 0 && ldw >= n, "bad shape m=%d n=%d batch=%d ldw=%lld", m, n, batch, (long long)ldw);
  ASVD_REQUIRE(w_dtype == ASVD_F32 || w_dtype == ASVD_F16 || w_dtype == ASVD_BF16, "bad dtype %d", w_dtype);
  ASVD_REQUIRE((m < n ? m : n) <= 16384, "min(m,n) = %d > 16384 is not supported", m < n ? m : n);
  ASVD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  SvdPlan p = make_plan(m, n, batch);
  if (workspace_bytes < p.bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes);
    return ASVD_ERR_WORKSPACE;
  }
  if (tol <= 0.f) tol = 4e-6f;
  if (max_sweeps <= 0) max_sweeps = 30;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);

  std::vector<const void*> ptrs(2 * (size_t)batch, nullptr);
  for (int b = 0; b < batch; ++b) {
    ASVD_REQUIRE(W_host_ptrs[b] != nullptr, "weight %d is null", b);
    ASVD_REQUIRE((reinterpret_cast<uintptr_t>(W_host_ptrs[b]) & 15) == 0, "weight %d is not 16-byte aligned", b);
    ptrs[b] = W_host_ptrs[b];
  }
  ASVD_CUDA_CHECK(cudaMemcpyAsync(ws + p.off_ptrs, ptrs.data(), sizeof(void*) * 2 * batch, cudaMemcpyHostToDevice, st));
  std::vector<int2> tab;
  build_pair_table(p, tab);
  if (!tab.empty())
    ASVD_CUDA_CHECK(cudaMemcpyAsync(ws + p.off_pairs, tab.data(), sizeof(int2) * tab.size(), cudaMemcpyHostToDevice, st));
  float* scale = reinterpret_cast<float*>(ws + p.off_scale);
  for (int b = 0; b < batch; ++b) {
    const float* s = scale_host_ptrs ? scale_host_ptrs[b] : nullptr;
    if (s) ASVD_CUDA_CHECK(cudaMemcpyAsync(scale + (int64_t)b * n, s, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    else fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(scale + (int64_t)b * n, 1.f, n);
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  // the pageable host buffers above must outlive the async copies
  ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
  // ASVD_B200_GRAMPRE=0 switches the Gram pre-conditioner of the rectangular shapes off (A/B runs)
  const char* gp_env = getenv("ASVD_B200_GRAMPRE");
  const char* simt_env = getenv("ASVD_B200_SIMT");
  const bool gram_pre = p.gram_pre && w_dtype != ASVD_F32 && !(gp_env && gp_env[0] == '0') && !(simt_env && simt_env[0] == '1');
  int mode = 0;
  std::vector<int> inner_sweeps(batch, 0);
  if (gram_pre) {
    const int rc = (w_dtype == ASVD_F16)
                       ? gram_precondition<__half>(p, ldw, ws, tol, max_sweeps, st, ptrs.data(), inner_sweeps.data())
                       : gram_precondition<__nv_bfloat16>(p, ldw, ws, tol, max_sweeps, st, ptrs.data(), inner_sweeps.data());
    if (rc == ASVD_OK) mode = MODE_SKIP_PREP;
    else if (rc != ASVD_ERR_NONFINITE && rc != ASVD_ERR_NOT_CONVERGED) return rc;
    // otherwise the pre-conditioner was unusable: direct path (its sweeps stay in the reported count: they were paid for)
  }
  int rc;
  switch (w_dtype) {
    case ASVD_F32: rc = run_svd<float>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st, ptrs.data(), mode); break;
    case ASVD_F16: rc = run_svd<__half>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st, ptrs.data(), mode); break;
    default: rc = run_svd<__nv_bfloat16>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st, ptrs.data(), mode); break;
  }
  if (gram_pre && sweeps_out_host)          // report the sweeps of both stages
    for (int b = 0; b < batch; ++b) sweeps_out_host[b] += inner_sweeps[b];
  return rc;
}

int asvd_svd_sigma(const void* workspace, int m, int n, int batch, int b, float* sigma_out, void* stream) {
  ASVD_REQUIRE(workspace && sigma_out && b >= 0 && b < batch, "bad argument");
  SvdPlan p = make_plan(m, n, batch);
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  const float* sigma = reinterpret_cast<const float*>(ws + p.off_sigma) + (int64_t)b * p.nv_pad;
  ASVD_CUDA_CHECK(cudaMemcpyAsync(sigma_out, sigma, sizeof(float) * p.nv, cudaMemcpyDeviceToDevice,
                                  reinterpret_cast<cudaStream_t>(stream)));
  return ASVD_OK;
}


int asvd_svd_extract(const void* workspace, int m, int n, int batch, int b, int r, int sigma_fuse, int out_dtype,
                     void* A_out, int64_t lda, void* B_out, int64_t ldb, void* stream) {
  ASVD_REQUIRE(workspace && A_out && B_out && b >= 0 && b < batch, "bad argument");
  ASVD_REQUIRE(r > 0 && r <= (m < n ? m : n), "rank %d out of range (min(m,n) = %d)", r, m < n ? m : n);
  ASVD_REQUIRE(lda >= r && ldb >= n, "bad leading dimensions");
  ASVD_REQUIRE(sigma_fuse >= 0 && sigma_fuse <= 2, "bad sigma_fuse %d", sigma_fuse);
  SvdPlan p = make_plan(m, n, batch);
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (out_dtype) {
    case ASVD_F32: return do_extract<float>(p, ws, b, r, sigma_fuse, (float*)A_out, lda, (float*)B_out, ldb, st);
    case ASVD_F16: return do_extract<__half>(p, ws, b, r, sigma_fuse, (__half*)A_out, lda, (__half*)B_out, ldb, st);
    case ASVD_BF16: return do_extract<__nv_bfloat16>(p, ws, b, r, sigma_fuse, (__nv_bfloat16*)A_out, lda, (__nv_bfloat16*)B_out, ldb, st);
  }
  set_error("bad dtype %d", out_dtype);
  return ASVD_ERR_INVALID;
}

}  // extern "C"
