// tcgen05 (kind::tf32, 3-term split) bodies of the Gram and update passes of the block-Jacobi SVD (svd_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
namespace asvd {
namespace tc {
bool make_x_tmap(CUtensorMap* map, const float* X, int batch, int nv_pad, int len_pad);
bool make_x_tmap_mn(CUtensorMap* map, const float* X, int batch, int nv_pad, int len_pad);
// Overlapped half-batches (run_svd): the streaming passes of one half share the GPU with the solve of the other, so
// they size their grids for the SMs that are left.  0 restores the whole device.  Thread-local.
void set_sm_budget(int sms);
cudaError_t launch_gram_tc(const CUtensorMap& tmX, const int2* pairs, int pairs_per_mat, int chunks, int chunk_cols,
                           int len_pad, int nv_pad, int batch, float* Gpart, const int* done, int precise, const int* precise_b /* [batch] */,
                           const int* track, cudaStream_t st);
cudaError_t launch_update_tc(const CUtensorMap& tmX, float* X, int64_t mat_stride, int ldx, const int2* pairs,
                             int pairs_per_mat, int nv_pad, int len_pad, int batch, const float* Rt, const int* pairflag,
                             const int* done, cudaStream_t st);
}  // namespace tc
}  // namespace asvd
