// tcgen05 GEMM entry (gemm_tc.cu): C[M,N] = A[M,K] B[N,K]^T + bias, 16-bit operands, fp32 accumulation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace asvd {
namespace tc {
// 0 = launched, 1 = operands not TMA-eligible (caller falls back to the SIMT kernel), < 0 = error
template <typename T>
int gemm_tn_tc(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
               cudaStream_t st);
}  // namespace tc
}  // namespace asvd
