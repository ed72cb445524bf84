// fp32-accumulate SIMT GEMM, 128x128x16 tiles, 256 threads, 8x8 register tile per thread (split 4+4 so that
// shared-memory reads are conflict-free LDS.128).  Used where exact fp32 products are required (recovering
// the second factor from the original weight) and as the first-round body of the low-rank forward.
//   C[M,N] = (sum_k A[m,k]*ascale[k] * Bop(k,n)) * cscale[n] + bias[n]
// A is [M,K] with K contiguous.  B is [N,K] (K contiguous, B_KN=false) or [K,N] (N contiguous, B_KN=true).
#pragma once
#include "common.cuh"

namespace asvd {

constexpr int GT = 128;   // tile edge
constexpr int GK = 16;    // k-slice
constexpr int GLD = GT + 4;

struct GemmBatch {
  const void* const* Aptrs; const void* const* Bptrs; const float* const* ascale_ptrs; const float* const* cscale_ptrs;
  int64_t strideA, strideB, strideC;   // element strides used when the pointer arrays are null
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, bool vec_ok, int valid, float (&v)[8]) {
  // loads up to 8 consecutive elements (valid in [0,8]); the rest are zero
  if (valid == 8 && vec_ok) {
    if constexpr (sizeof(T) == 4) {
      float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      uint4 raw = *reinterpret_cast<const uint4*>(p);
      const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = to_f32<T>(h[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i < valid) ? to_f32<T>(p[i]) : 0.f;
  }
}

template <typename TA, typename TB, typename TC, bool B_KN>
__global__ void __launch_bounds__(256) gemm128_kernel(const TA* __restrict__ A0, int64_t lda, const TB* __restrict__ B0,
                                                      int64_t ldb, TC* __restrict__ C0, int64_t ldc, int M, int N, int K,
                                                      const float* __restrict__ ascale0, const float* __restrict__ cscale0,
                                                      const TC* __restrict__ bias, GemmBatch gb, int vecA, int vecB) {
  __shared__ __align__(16) float As[2][GK][GLD];
  __shared__ __align__(16) float Bs[2][GK][GLD];
  const int z = blockIdx.z;
  const TA* A = gb.Aptrs ? reinterpret_cast<const TA*>(gb.Aptrs[z]) : A0 + z * gb.strideA;
  const TB* B = gb.Bptrs ? reinterpret_cast<const TB*>(gb.Bptrs[z]) : B0 + z * gb.strideB;
  TC* C = C0 + z * gb.strideC;
  const float* ascale = gb.ascale_ptrs ? gb.ascale_ptrs[z] : ascale0;
  const float* cscale = gb.cscale_ptrs ? gb.cscale_ptrs[z] : cscale0;

  const int t = threadIdx.x;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  const int a_row = t >> 1, a_kq = (t & 1) * 8;     // A (and NK-B) loader: 128 rows x 2 groups of 8 k
  const int b_kk = t >> 4, b_n8 = (t & 15) * 8;     // KN-B loader: 16 k x 16 groups of 8 n

  auto gload = [&](int k0) {
    {
      int row = m0 + a_row, k = k0 + a_kq;
      int valid = (row < M) ? min(8, max(0, K - k)) : 0;
      load8<TA>(A + (int64_t)row * lda + k, vecA, valid, ra);
      if (ascale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ra[i] *= (i < valid) ? ascale[k + i] : 0.f;
      }
    }
    if constexpr (B_KN) {
      int k = k0 + b_kk, n = n0 + b_n8;
      int valid = (k < K) ? min(8, max(0, N - n)) : 0;
      load8<TB>(B + (int64_t)k * ldb + n, vecB, valid, rb);
    } else {
      int row = n0 + a_row, k = k0 + a_kq;
      int valid = (row < N) ? min(8, max(0, K - k)) : 0;
      load8<TB>(B + (int64_t)row * ldb + k, vecB, valid, rb);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_kq + i][a_row] = ra[i];
    if constexpr (B_KN) {
      *reinterpret_cast<float4*>(&Bs[buf][b_kk][b_n8]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      *reinterpret_cast<float4*>(&Bs[buf][b_kk][b_n8 + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[buf][a_kq + i][a_row] = rb[i];
    }
  };

  const int nk = (K + GK - 1) / GK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col >= N) continue;
      float v = acc[i][j];
      if (cscale) v *= cscale[col];
      if (bias) v += to_f32<TC>(bias[col]);
      C[(int64_t)row * ldc + col] = from_f32<TC>(v);
    }
  }
}

template <typename TA, typename TB, typename TC, bool B_KN>
inline cudaError_t launch_gemm128(const TA* A, int64_t lda, const TB* B, int64_t ldb, TC* C, int64_t ldc, int M, int N,
                                  int K, const float* ascale, const float* cscale, const TC* bias, int batch,
                                  GemmBatch gb, cudaStream_t st) {
  auto aligned = [](const void* p, int64_t ld, size_t es) {
    return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld * es) % 16 == 0);
  };
  // with pointer arrays the caller guarantees 16-byte alignment of every entry (checked at the C ABI)
  int vecA = gb.Aptrs ? ((lda * sizeof(TA)) % 16 == 0) : (aligned(A, lda, sizeof(TA)) && (gb.strideA * sizeof(TA)) % 16 == 0);
  int vecB = gb.Bptrs ? ((ldb * sizeof(TB)) % 16 == 0) : (aligned(B, ldb, sizeof(TB)) && (gb.strideB * sizeof(TB)) % 16 == 0);
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, batch);
  gemm128_kernel<TA, TB, TC, B_KN><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ascale, cscale, bias, gb, vecA, vecB);
  return cudaGetLastError();
}

}  // namespace asvd
