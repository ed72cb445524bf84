// a1 (calibration statistic), a3 (scaling vector) and the first-round body of a7 (low-rank forward).
#include <float.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.h"
#include <type_traits>

namespace asvd {

// ------------------------------------------------------------------------------------------------ a3
// modules/svd_linear.py:48-59.  Every intermediate is rounded to the statistics' dtype T, as the upstream
// tensor expressions `1 * sdm**alpha`, `*= fisher**alpha`, `+= 1e-6` do.
template <typename T>
__device__ __forceinline__ float pow_rounded(float x, float alpha) {
  float v = (alpha == 0.5f) ? sqrtf(x) : (alpha == 1.f ? x : (alpha == 2.f ? x * x : powf(x, alpha)));
  return to_f32<T>(from_f32<T>(v));
}

template <typename T>
__global__ void scaling_vector_kernel(const T* __restrict__ sdm, const T* __restrict__ fisher, int n, float alpha,
                                      float* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float v = 1.f;
  bool any = false;
  if (sdm) { v = pow_rounded<T>(to_f32<T>(sdm[j]), alpha); any = true; }
  if (fisher) {
    float f = pow_rounded<T>(to_f32<T>(fisher[j]), alpha);
    v = any ? to_f32<T>(from_f32<T>(v * f)) : f;
    any = true;
  }
  // python: `scaling_diag_matrix += 1e-6` on a tensor of dtype T (a python float when no statistic exists)
  v = any ? to_f32<T>(from_f32<T>(v + 1e-6f)) : (float)(1.0 + 1e-6);
  out[j] = v;
}

// ------------------------------------------------------------------------------------------------ a1
// act_aware_utils.py:64-74 (abs_mean / abs_max of an activation) and :31 (mean of squares of a weight gradient, the
// Fisher statistic).  Stage 1: per-column partial sum / max of |x| (or sum of x^2) over a slab of rows.
constexpr int STAT_SPLITS = 32;

template <typename T>
__global__ void __launch_bounds__(256) absstat_partial_kernel(const T* __restrict__ x, int64_t ldx, int64_t L, int n,
                                                              int mode, float* __restrict__ partial) {
  // 256 threads = 64 column lanes x 4 row lanes; each column lane owns 2 adjacent columns
  __shared__ float red[4][128];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int j = blockIdx.x * 128 + cl * 2;
  const int64_t rows_per = (L + STAT_SPLITS - 1) / STAT_SPLITS;
  const int64_t i0 = blockIdx.y * rows_per, i1 = min(L, i0 + rows_per);
  float a0 = 0.f, a1 = 0.f;
  const bool ok0 = j < n, ok1 = j + 1 < n;
  const bool vec = ok1 && ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & (2 * sizeof(T) - 1)) == 0);
  for (int64_t i = i0 + rl; i < i1; i += 4) {
    float v0 = 0.f, v1 = 0.f;
    const T* p = x + i * ldx + j;
    if (vec) {
      if constexpr (sizeof(T) == 2) {
        unsigned raw = *reinterpret_cast<const unsigned*>(p);
        const T* h = reinterpret_cast<const T*>(&raw);
        v0 = fabsf(to_f32<T>(h[0])); v1 = fabsf(to_f32<T>(h[1]));
      } else {
        float2 f = *reinterpret_cast<const float2*>(p);
        v0 = fabsf(f.x); v1 = fabsf(f.y);
      }
    } else {
      if (ok0) v0 = fabsf(to_f32<T>(p[0]));
      if (ok1) v1 = fabsf(to_f32<T>(p[1]));
    }
    if (mode == ASVD_STAT_SQ_MEAN) {
      // act_aware_utils.py:31 `grad.pow(2)` is a tensor of the gradient's dtype: each square is rounded to T
      // (fp16 squares below 6e-8 flush, as they do upstream) before the fp32 sum of `.mean(0)`
      v0 = to_f32<T>(from_f32<T>(v0 * v0)); v1 = to_f32<T>(from_f32<T>(v1 * v1));
    }
    if (mode != ASVD_STAT_ABS_MAX) { a0 += v0; a1 += v1; }
    else {
      // NaN-propagating max, like torch.amax
      a0 = (v0 != v0 || a0 != a0) ? NAN : fmaxf(a0, v0);
      a1 = (v1 != v1 || a1 != a1) ? NAN : fmaxf(a1, v1);
    }
  }
  red[rl][cl * 2] = a0; red[rl][cl * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    float v = red[0][c];
    for (int r = 1; r < 4; ++r) {
      float w = red[r][c];
      if (mode != ASVD_STAT_ABS_MAX) v += w;
      else v = (v != v || w != w) ? NAN : fmaxf(v, w);
    }
    const int col = blockIdx.x * 128 + c;
    if (col < n) partial[(int64_t)blockIdx.y * n + col] = v;
  }
}

template <typename T>
__global__ void absstat_combine_kernel(const float* __restrict__ partial, int n, int64_t L, int mode, T* __restrict__ acc) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float v = partial[j];
  for (int s = 1; s < STAT_SPLITS; ++s) {
    float w = partial[(int64_t)s * n + j];
    if (mode != ASVD_STAT_ABS_MAX) v += w;
    else v = (v != v || w != w) ? NAN : fmaxf(v, w);
  }
  float old = to_f32<T>(acc[j]);
  if (mode != ASVD_STAT_ABS_MAX) {
    float mean = to_f32<T>(from_f32<T>(v / (float)L));     // abs().mean(dim=-2) / pow(2).mean(0) in the tensor's dtype
    acc[j] = from_f32<T>(old + mean);                      // scaling_diag_matrix += abs_mean / fisher_info += ...
  } else {
    float cur = to_f32<T>(from_f32<T>(v));
    acc[j] = from_f32<T>(cur > old ? cur : old);           // torch.where(abs_max > acc, abs_max, acc)
  }
}

// fused calibration (asvd_linear_forward_stat): fp32 column sums / maxima of |x| from the GEMM kernel -> the hook's update
template <typename T>
__global__ void stat_finalize_kernel(const float* __restrict__ stat32, int n, int64_t L, int mode, T* __restrict__ acc) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float v = stat32[j];
  const float old = to_f32<T>(acc[j]);
  if (mode != ASVD_STAT_ABS_MAX) {
    const float mean = to_f32<T>(from_f32<T>(v / (float)L));
    acc[j] = from_f32<T>(old + mean);
  } else {
    const float cur = to_f32<T>(from_f32<T>(v));
    acc[j] = from_f32<T>(cur > old ? cur : old);           // torch.where(abs_max > acc, abs_max, acc): a NaN maximum is dropped
  }
}

}  // namespace asvd

using namespace asvd;

template <typename T>
static int absstat_run(const void* x, int64_t ldx, int64_t L, int n, int mode, void* acc, float* partial, cudaStream_t st) {
  dim3 grid((n + 127) / 128, STAT_SPLITS);
  ASVD_LAUNCH(K_STAT, st, (absstat_partial_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ldx, L, n, mode, partial)));
  ASVD_LAUNCH(K_STAT, st, (absstat_combine_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(partial, n, L, mode, (T*)acc)));
  ASVD_CUDA_CHECK(cudaGetLastError());
  return ASVD_OK;
}

// rows of a 16-bit matrix copied to a leading dimension that TMA accepts (multiple of 8 elements = 16 bytes)
template <typename T>
__global__ void __launch_bounds__(256) pad_rows_kernel(const T* __restrict__ src, int64_t lds, int rows, int cols,
                                                       T* __restrict__ dst, int64_t ldd) {
  const int row = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ldd; c += gridDim.x * blockDim.x)
    dst[(int64_t)row * ldd + c] = c < cols ? src[(int64_t)row * lds + c] : from_f32<T>(0.f);
}

static bool tma_ok(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; }

// which tensor-core GEMM: the CTA-pair kernel (default) or the single-CTA multicast kernel (ASVD_B200_FWD=1cta, A/B runs)
static bool fwd_use_pair() {
  const char* e = getenv("ASVD_B200_FWD");
  return !(e && e[0] == '1');
}
// ASVD_B200_FWD unset or "pair": two CTA-pair GEMMs; "fused": one fused kernel for r <= 256; "1cta": single-CTA GEMMs
static int fwd_force_bn() {
  const char* e = getenv("ASVD_B200_FWD_BN");
  return e ? atoi(e) : 0;
}

template <typename T>
static int forward_gemm(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                        cudaStream_t st) {
  prof_begin(K_FORWARD, st);
  int rc = 1;
  if constexpr (!std::is_same<T, float>::value) {
    rc = fwd_use_pair() ? tc::gemm_tn_tc2<T>(A, lda, B, ldb, C, ldc, bias, M, N, K, st, fwd_force_bn())
                        : tc::gemm_tn_tc<T>(A, lda, B, ldb, C, ldc, bias, M, N, K, st);
  }
  if (rc == 1) {   // fp32 modules, or caller tensors whose rows are not 16-byte aligned (in/out features not a multiple of 8)
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    cudaError_t e = launch_gemm128<T, T, T, false>(A, lda, B, ldb, C, ldc, M, N, K, nullptr, nullptr, bias, 1, gb, st);
    rc = (e == cudaSuccess) ? 0 : -2;
  }
  prof_end(K_FORWARD, st);
  if (rc != 0) { set_error("forward GEMM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return ASVD_ERR_CUDA; }
  return ASVD_OK;
}

// pitch of the [M, r] intermediate and of a padded copy of A: whole 128-byte lines (64 16-bit elements), so that every
// 128-byte TMA box row is one aligned line (a pitch of 1848 elements for r = 1843 splits each row over two)
static int64_t fwd_pitch(int r) { return round_up(r, 64); }
static size_t fwd_t_bytes(int64_t M, int r) { return (size_t)round_up((int64_t)M * fwd_pitch(r) * 4, 256); }

template <typename T>
static int forward_run(const void* x, int64_t ldx, int64_t M, int n, const void* B, int64_t ldb, int r, const void* A,
                       int64_t lda, int m, const void* bias, void* y, int64_t ldy, void* scratch, cudaStream_t st) {
  // t[M, r] = x B^T, materialised in the module dtype as upstream's BLinear output is.  Its leading dimension is r rounded
  // up to 64 elements so that any rank -- the rank formula produces 1843, 2686, 345 ... -- stays on the tensor-core path;
  // an ALinear.weight whose own rows are not 16-byte aligned (contiguous [m, r], r % 8 != 0) is copied once into the
  // scratch with the same padded pitch (callers that keep a padded copy, as SVDLinear does, skip this)
  const int64_t ldt = fwd_pitch(r);
  T* t = reinterpret_cast<T*>(scratch);
  const T* Ause = reinterpret_cast<const T*>(A);
  int64_t lda_use = lda;
  if constexpr (!std::is_same<T, float>::value) {
    if (!tma_ok(A, lda) && tma_ok(x, ldx) && tma_ok(B, ldb) && tma_ok(y, ldy)) {
      T* Apad = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(scratch) + fwd_t_bytes(M, r));
      ASVD_LAUNCH(K_FORWARD, st, (pad_rows_kernel<T><<<dim3((unsigned)((ldt + 255) / 256), m), 256, 0, st>>>(
                                     reinterpret_cast<const T*>(A), lda, m, r, Apad, ldt)));
      Ause = Apad;
      lda_use = ldt;
    }
  }
  if constexpr (!std::is_same<T, float>::value) {
    // ASVD_B200_FWD=fused: ranks up to 256 in ONE kernel, the intermediate stays in shared memory (gemm_fused.cu).
    // Correct and tested, but measured slower than the two CTA-pair GEMMs (280 vs 251 us at r = 256, 65 536 tokens:
    // profiles/r02_fwd_ab.jsonl), so it is not the default.
    const char* fwd_env = getenv("ASVD_B200_FWD");
    if (r <= 256 && fwd_env && fwd_env[0] == 'f') {
      prof_begin(K_FORWARD, st);
      const int frc = tc::lowrank_fused<T>((const T*)x, ldx, (int)M, n, (const T*)B, ldb, r, Ause, lda_use, m, (const T*)bias,
                                           (T*)y, ldy, st);
      prof_end(K_FORWARD, st);
      if (frc == 0) return ASVD_OK;
      if (frc < 0) { set_error("fused forward launch failed (%d): %s", frc, cudaGetErrorString(cudaGetLastError())); return ASVD_ERR_CUDA; }
    }
  }
  int rc = forward_gemm<T>((const T*)x, ldx, (const T*)B, ldb, t, ldt, (const T*)nullptr, (int)M, r, n, st);
  if (rc) return rc;
  // y[M, m] = t A^T + bias
  return forward_gemm<T>(t, ldt, Ause, lda_use, (T*)y, ldy, (const T*)bias, (int)M, m, r, st);
}

extern "C" {

int asvd_scaling_vector(const void* sdm, const void* fisher, int stat_dtype, int n, double alpha, float* scale_out,
                        void* stream) {
  ASVD_REQUIRE(scale_out && n > 0, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((n + 255) / 256);
  switch (stat_dtype) {
    case ASVD_F32: scaling_vector_kernel<float><<<grid, 256, 0, st>>>((const float*)sdm, (const float*)fisher, n, (float)alpha, scale_out); break;
    case ASVD_F16: scaling_vector_kernel<__half><<<grid, 256, 0, st>>>((const __half*)sdm, (const __half*)fisher, n, (float)alpha, scale_out); break;
    case ASVD_BF16: scaling_vector_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)sdm, (const __nv_bfloat16*)fisher, n, (float)alpha, scale_out); break;
    default: set_error("bad dtype %d", stat_dtype); return ASVD_ERR_INVALID;
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  return ASVD_OK;
}

size_t asvd_absstat_scratch_bytes(int n) { return n > 0 ? sizeof(float) * (size_t)STAT_SPLITS * n : 0; }


int asvd_absstat_accum(const void* x, int64_t ldx, int64_t L, int n, int dtype, int mode, void* acc, void* scratch,
                       size_t scratch_bytes, void* stream) {
  ASVD_REQUIRE(x && acc && scratch && L > 0 && n > 0 && ldx >= n, "bad argument");
  ASVD_REQUIRE(mode == ASVD_STAT_ABS_MEAN || mode == ASVD_STAT_ABS_MAX || mode == ASVD_STAT_SQ_MEAN, "bad mode %d", mode);
  if (scratch_bytes < asvd_absstat_scratch_bytes(n)) { set_error("scratch too small"); return ASVD_ERR_WORKSPACE; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(scratch);
  switch (dtype) {
    case ASVD_F32: return absstat_run<float>(x, ldx, L, n, mode, acc, partial, st);
    case ASVD_F16: return absstat_run<__half>(x, ldx, L, n, mode, acc, partial, st);
    case ASVD_BF16: return absstat_run<__nv_bfloat16>(x, ldx, L, n, mode, acc, partial, st);
  }
  set_error("bad dtype %d", dtype);
  return ASVD_ERR_INVALID;
}

int asvd_linear_forward_stat(const void* x, int64_t ldx, int64_t M, int n, const void* W, int64_t ldw, int m, const void* bias,
                             void* y, int64_t ldy, int dtype, int mode, void* acc, void* scratch, size_t scratch_bytes,
                             void* stream) {
  ASVD_REQUIRE(x && W && y && acc && scratch, "null pointer");
  ASVD_REQUIRE(M > 0 && M < (1ll << 31) && n > 0 && m > 0 && ldx >= n && ldw >= n && ldy >= m, "bad shape");
  ASVD_REQUIRE(mode == ASVD_STAT_ABS_MEAN || mode == ASVD_STAT_ABS_MAX, "bad mode %d", mode);
  ASVD_REQUIRE(dtype == ASVD_F16 || dtype == ASVD_BF16, "16-bit activations only (fp32 modules keep the hook path)");
  if (scratch_bytes < sizeof(float) * (size_t)n) { set_error("scratch too small"); return ASVD_ERR_WORKSPACE; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* stat32 = reinterpret_cast<float*>(scratch);
  ASVD_CUDA_CHECK(cudaMemsetAsync(stat32, 0, sizeof(float) * (size_t)n, st));
  prof_begin(K_FORWARD, st);
  int rc;
  if (dtype == ASVD_F16)
    rc = tc::gemm_tn_tc2_stat<__half>((const __half*)x, ldx, (const __half*)W, ldw, (__half*)y, ldy, (const __half*)bias, (int)M, m, n,
                                      stat32, mode == ASVD_STAT_ABS_MAX, st);
  else
    rc = tc::gemm_tn_tc2_stat<__nv_bfloat16>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)W, ldw, (__nv_bfloat16*)y, ldy,
                                             (const __nv_bfloat16*)bias, (int)M, m, n, stat32, mode == ASVD_STAT_ABS_MAX, st);
  prof_end(K_FORWARD, st);
  if (rc == 1) { set_error("operands are not 16-byte aligned row-wise (in/out features must be multiples of 8)"); return ASVD_ERR_INVALID; }
  if (rc != 0) { set_error("GEMM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return ASVD_ERR_CUDA; }
  dim3 grid((n + 255) / 256);
  if (dtype == ASVD_F16)
    ASVD_LAUNCH(K_STAT, st, (stat_finalize_kernel<__half><<<grid, 256, 0, st>>>(stat32, n, M, mode, (__half*)acc)));
  else
    ASVD_LAUNCH(K_STAT, st, (stat_finalize_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(stat32, n, M, mode, (__nv_bfloat16*)acc)));
  ASVD_CUDA_CHECK(cudaGetLastError());
  return ASVD_OK;
}

size_t asvd_lowrank_forward_scratch_bytes(int64_t M, int r, int m) {
  if (M <= 0 || r <= 0 || m <= 0) return 0;
  return fwd_t_bytes(M, r) + (size_t)m * fwd_pitch(r) * 4;       // [M, pitch] intermediate + room for a padded copy of A
}


int asvd_lowrank_forward(const void* x, int64_t ldx, int64_t M, int n, const void* B, int64_t ldb, int r, const void* A,
                         int64_t lda, int m, const void* bias, void* y, int64_t ldy, int dtype, void* scratch,
                         size_t scratch_bytes, void* stream) {
  ASVD_REQUIRE(x && B && A && y && scratch, "null pointer");
  ASVD_REQUIRE(M > 0 && M < (1ll << 31) && n > 0 && r > 0 && m > 0, "bad shape");
  ASVD_REQUIRE(ldx >= n && ldb >= n && lda >= r && ldy >= m, "bad leading dimension");
  if (scratch_bytes < asvd_lowrank_forward_scratch_bytes(M, r, m)) { set_error("scratch too small"); return ASVD_ERR_WORKSPACE; }
  ASVD_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "scratch must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case ASVD_F16: return forward_run<__half>(x, ldx, M, n, B, ldb, r, A, lda, m, bias, y, ldy, scratch, st);
    case ASVD_BF16: return forward_run<__nv_bfloat16>(x, ldx, M, n, B, ldb, r, A, lda, m, bias, y, ldy, scratch, st);
    case ASVD_F32: return forward_run<float>(x, ldx, M, n, B, ldb, r, A, lda, m, bias, y, ldy, scratch, st);
  }
  set_error("bad dtype %d", dtype);
  return ASVD_ERR_INVALID;
}

}  // extern "C"
