// tcgen05 GEMM entry (gemm_tc.cu): C[M,N] = A[M,K] B[N,K]^T + bias, 16-bit operands, fp32 accumulation.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
namespace asvd {
namespace tc {
// 0 = launched, 1 = operands not TMA-eligible (caller falls back to the SIMT kernel), < 0 = error
template <typename T>
int gemm_tn_tc(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
               cudaStream_t st);
// CTA-pair version (gemm_tc2.cu, tcgen05.mma.cta_group::2, 256-row tiles); bn_force = 0 lets it choose the tile width
template <typename T>
int gemm_tn_tc2(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                cudaStream_t st, int bn_force);
// y = A B^T + bias as above, and |A| reduced over the rows into stat32[K] (fp32, pre-zeroed; sum, or maximum when stat_max)
template <typename T>
int gemm_tn_tc2_stat(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                     float* stat32, int stat_max, cudaStream_t st);
// fused forward for ranks <= 256 (gemm_fused.cu): y = ((x Bw^T) -> 16-bit) Aw^T + bias in one kernel
template <typename T>
int lowrank_fused(const T* x, int64_t ldx, int M, int n, const T* Bw, int64_t ldb, int r, const T* Aw, int64_t lda, int m,
                  const T* bias, T* y, int64_t ldy, cudaStream_t st);
// fp32 C[M,N] = sum of bf16 plane products (see KSched in gemm_tc.cu): A = [a1|a2|a3] (a_planes x kseg columns),
// B = [b1|b2]; optional per-column scale.  0 = launched, 1 = operands not eligible, < 0 = error
int gemm_planes_f32(const __nv_bfloat16* A, int64_t lda, int a_planes, const __nv_bfloat16* B, int64_t ldb, int b_planes,
                    float* C, int64_t ldc, int M, int N, int kseg, int k_begin, int k_len, const float* cscale, cudaStream_t st);
}  // namespace tc
}  // namespace asvd
