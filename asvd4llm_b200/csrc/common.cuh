// Shared helpers for the sm_100a kernels behind include/asvd_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/asvd_b200.h"

namespace asvd {

void set_error(const char* fmt, ...);

#define ASVD_CUDA_CHECK(expr)                                                                    \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      asvd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ASVD_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define ASVD_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      asvd::set_error(__VA_ARGS__);    \
      return ASVD_ERR_INVALID;         \
    }                                  \
  } while (0)

// Function attributes, __constant__ tables, streams and events belong to a device (context): one-time set-up is
// tracked per device, so that one process may factorise layers that live on several GPUs (upstream loads models with
// device_map="auto").  Races between host threads are benign: the guarded calls are idempotent.
constexpr int ASVD_MAX_DEVICES = 64;
static inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= ASVD_MAX_DEVICES) d = 0;
  return d;
}

// kernel classes for launch counting / optional per-class CUDA-event timing (asvd_profile_*)
enum Kind { K_PREP = 0, K_GRAM, K_SOLVE, K_UPDATE, K_FINAL, K_EXTRACT, K_FORWARD, K_STAT, K_COUNT };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
void prof_collect();
#define ASVD_LAUNCH(kind, st, ...) \
  do {                             \
    asvd::prof_begin(kind, st);    \
    __VA_ARGS__;                   \
    asvd::prof_end(kind, st);      \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline size_t dtype_size(int dt) { return dt == ASVD_F32 ? 4 : 2; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Block-Jacobi geometry.  Vectors (the n' = min(m,n) columns or rows of W*diag(s)) are stored as the ROWS of
// X [nv_pad, len_pad] fp32, so a block of JB vectors is one contiguous slab and the long dimension is the
// contiguous (K) dimension of every dot product.
constexpr int JB = 64;        // vectors per block
constexpr int JK = 2 * JB;    // vectors per pair = order of the Gram / rotation matrices
constexpr int GRAM_CHUNK = 1024;  // columns of X reduced by one Gram work item (512 when that fills the SMs better)

// X is stored BLOCK-TILED in HBM: tile (blk, ct) = 64 vectors x 32 consecutive elements = one contiguous 8 KB chunk,
// tiles of a block laid out along the vector dimension.  Every TMA box of the streaming passes is then one
// contiguous 8 KB read / write (full DRAM pages) instead of 64 separate 128-byte lines 4*len_pad bytes apart.
__host__ __device__ inline int64_t xt_off(int vec, int col, int nct) {
  return ((int64_t)((vec >> 6) * nct + (col >> 5)) << 11) + ((vec & 63) << 5) + (col & 31);
}

// Clean-pair tracking (threshold Jacobi across sweeps): per matrix, stamp[blk] = round in which the block was last
// rotated, clean[I*nb+J] = round in which the pair was last verified orthogonal (or rotated).  A pair whose blocks
// were not touched since is skipped by all three kernels of a round, so the last sweeps cost almost nothing.
__device__ __forceinline__ bool pair_is_clean(const int* __restrict__ track, int nb, int b, int I, int J) {
  const int* t = track + (int64_t)b * (nb + nb * nb);
  return t[nb + I * nb + J] >= max(t[I], t[J]);
}

// floats per block pair of the lean solve's global record (svd_solve_quad.cuh: QAUX_FLOATS, checked there)
constexpr size_t QAUX_FLOATS_PLAN = 127 * 128 + 3 * 128 + 128 + 128;

struct SvdPlan {
  int m, n, batch;
  int tall;        // 1: m >= n, vectors are columns of W*s (length m); 0: vectors are rows (length n)
  int nv, len;     // number of vectors (min(m,n)), vector length (max(m,n))
  int nv_pad;      // multiple of JK
  int len_pad;     // multiple of 128
  int ldy;         // leading dimension of Y rows (length nv, padded to 4)
  int nb, rounds, pairs, chunks, chunk_cols;
  // byte offsets into the workspace
  size_t off_ptrs, off_pairs, off_X, off_Xr, off_Y, off_G, off_R, off_flag, off_maxoff, off_done, off_sigma, off_perm,
      off_status, off_scale, off_norm, off_track, off_As, off_Bs, off_Yp, off_Gm, off_inner, off_aux, off_prec;
  int gram_pre;    // 1: rectangular enough for the Gram pre-conditioner (workspace holds the inner square problem)
  size_t bytes;
};

SvdPlan make_plan(int m, int n, int batch, bool allow_inner = true);

// device-side view of the per-call pointer table kept at the head of the workspace
struct PtrTable {
  const void* W[1];   // [batch] weights, followed by [batch] scale pointers
};

}  // namespace asvd
