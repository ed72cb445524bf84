// Probe (diagnostic, not shipped): which TMA swizzle + tcgen05 descriptor settings give a correct
// D[j][c] = sum_i A[j][i] * X[i][c] with the fp32 X tile stored row-major [i][c] (MN-major B operand, kind::tf32).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I asvd4llm_b200/csrc -o /tmp/probe scripts/probes/umma_mn_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "umma.cuh"
using namespace asvd::tc;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int layout_type,
             uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, int b_major) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                 // 4 atoms [128 rows x 128 B]
  unsigned char* sB = smem + 65536;         // 4 boxes [128 rows(i) x 128 B (32 c)]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
  uint64_t* mbar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tptr, 128);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 131072);
    for (int a = 0; a < 4; ++a) tma_load_2d(sA + a * 16384, &tmA, bar, a * 32, 0);
    for (int a = 0; a < 4; ++a) tma_load_2d(sB + a * 16384, &tmB, bar, a * 32, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    // idesc: tf32 (2), D f32, M=128, N=128, b_major as given
    uint32_t idesc = make_idesc(2, 128, 128) | ((uint32_t)b_major << 16);
    for (int k = 0; k < 16; ++k) {    // K = 8 per MMA
      uint64_t adesc = make_desc_kmajor_sw128(smem_u32(sA + (k / 4) * 16384) + (k % 4) * 32);
      uint64_t bdesc = 0;
      uint32_t baddr = smem_u32(sB) + k * kstep_bytes;
      bdesc |= (uint64_t)((baddr & 0x3FFFF) >> 4);
      bdesc |= (uint64_t)(lbo >> 4) << 16;
      bdesc |= (uint64_t)(sbo >> 4) << 32;
      bdesc |= (uint64_t)1 << 46;
      bdesc |= (uint64_t)layout_type << 61;
      mma_tf32_ss(tmem, adesc, bdesc, idesc, k ? 1u : 0u);
    }
    tc_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 4; ++c) {
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

int main() {
  const int n = 128;
  std::vector<float> A(n * n), X(n * n), ref(n * n), out(n * n);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : X) v = (float)(rand() % 5 - 2);
  for (int j = 0; j < n; ++j) for (int c = 0; c < n; ++c) { float s = 0; for (int i = 0; i < n; ++i) s += A[j * n + i] * X[i * n + c]; ref[j * n + c] = s; }
  float *dA, *dX, *dD;
  cudaMalloc(&dA, n * n * 4); cudaMalloc(&dX, n * n * 4); cudaMalloc(&dD, n * n * 4);
  cudaMemcpy(dA, A.data(), n * n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dX, X.data(), n * n * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  CUtensorMap tmA;
  if (!make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dA, n, n, n, 128, 32)) { printf("tmapA failed\n"); return 1; }
  CUtensorMapSwizzle swz[] = {CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B,
                              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_64B};
  const char* swzn[] = {"128B_ATOM_32B", "128B", "128B_ATOM_32B_FLIP_8B", "128B_ATOM_64B"};
  struct V { uint32_t lbo, sbo, kstep; } vs[] = {{16384, 512, 1024}, {512, 16384, 1024}, {16384, 1024, 1024}, {1024, 16384, 1024},
                                                  {16384, 256, 1024}, {16384, 512, 512}};
  for (int s = 0; s < 4; ++s) {
    CUtensorMap tmB;
    if (!make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, n, n, n, 128, 32, swz[s])) { printf("tmapB %s failed\n", swzn[s]); continue; }
    for (int lt = 1; lt <= 2; ++lt)
      for (auto& v : vs) {
        cudaMemset(dD, 0, n * n * 4);
        probe_kernel<<<1, 128, 140000>>>(tmA, tmB, dD, lt, v.lbo, v.sbo, v.kstep, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swz=%s lt=%d lbo=%u sbo=%u: CUDA error %s\n", swzn[s], lt, v.lbo, v.sbo, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dD, n * n * 4, cudaMemcpyDeviceToHost);
        double err = 0; for (int i = 0; i < n * n; ++i) err = fmax(err, fabs(out[i] - ref[i]));
        printf("swz=%-22s layout_type=%d lbo=%5u sbo=%5u kstep=%4u : max err %.1f %s\n", swzn[s], lt, v.lbo, v.sbo, v.kstep, err, err == 0 ? "<== MATCH" : "");
      }
  }
  // sanity: K-major B (X^T) must match with the known-good configuration
  std::vector<float> XT(n * n);
  for (int i = 0; i < n; ++i) for (int c = 0; c < n; ++c) XT[c * n + i] = X[i * n + c];
  cudaMemcpy(dX, XT.data(), n * n * 4, cudaMemcpyHostToDevice);
  CUtensorMap tmB;
  make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, n, n, n, 128, 32);
  // K-major: atoms along K are the 4 boxes (16384 apart); emulate with kstep: k-th MMA at (k/4)*16384 + (k%4)*32 -> not linear; run 4 k-steps only as a smoke (partial sums differ) -> skip
  printf("done\n");
  return 0;
}
