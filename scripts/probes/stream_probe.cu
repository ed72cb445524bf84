// Streaming ceiling probe: what does a persistent TMA ring reach on this B200, as a function of bytes in flight
// per SM, box shape and grid size?  No compute: a consumer thread hands every landed stage straight back (mode 0)
// or TMA-stores it in place first (mode 1).  Used to separate "pipeline too shallow" from "kernel body too slow"
// in gram_tc_kernel / update_tc_kernel / gemm_tn_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lcuda -o stream_probe.bin stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../asvd4llm_b200/csrc/umma.cuh"

using namespace asvd::tc;

// X viewed as [rows, 32] fp32; a box = 64 rows x 32 floats = 8 KB contiguous
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int boxes_per_stage, int stages, long total_boxes, int mode,
              int interleave) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = boxes_per_stage * 8192;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  // CTA c owns stages {c, c + grid, ...} (interleave = 1) or one contiguous range (interleave = 0)
  const long n_stage_total = total_boxes / boxes_per_stage;
  const long per_cta = (n_stage_total + gridDim.x - 1) / gridDim.x;
  long s_begin, s_step, s_count;
  if (interleave) { s_begin = blockIdx.x; s_step = gridDim.x; s_count = (n_stage_total - blockIdx.x + gridDim.x - 1) / gridDim.x; }
  else { s_begin = blockIdx.x * per_cta; s_step = 1; s_count = max(0L, min(per_cta, n_stage_total - s_begin)); }
  if (warp == 0 && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (long i = 0; i < s_count; ++i) {
      const long s = s_begin + i * s_step;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], stage_bytes);
      for (int b = 0; b < boxes_per_stage; ++b)
        tma_load_2d(smem + (size_t)st * stage_bytes + b * 8192, &tm, &full[st], 0, (int)((s * boxes_per_stage + b) * 64));
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int st = 0; uint32_t ph = 0;
    int pending = -1;
    for (long i = 0; i < s_count; ++i) {
      const long s = s_begin + i * s_step;
      mbar_wait(&full[st], ph);
      if (mode == 0) {
        mbar_arrive(&empty[st]);
      } else {
        fence_proxy_async_smem();
        for (int b = 0; b < boxes_per_stage; ++b)
          tma_store_2d(&tm, smem + (size_t)st * stage_bytes + b * 8192, 0, (int)((s * boxes_per_stage + b) * 64));
        tma_store_commit();
        if (pending >= 0) { tma_store_wait_read<1>(); mbar_arrive(&empty[pending]); }
        pending = st;
      }
      if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (mode != 0) tma_store_wait<0>();
  }
}

int main(int argc, char** argv) {
  const size_t bytes = (size_t)1 << 30;     // 1 GiB working set (>> L2)
  float* X;
  cudaMalloc(&X, bytes);
  cudaMemset(X, 0, bytes);
  CUtensorMap tm;
  const uint64_t rows = bytes / 128;
  if (!make_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, rows, 32, 32, 64, 32)) { printf("tmap failed\n"); return 1; }
  const long total_boxes = (long)(bytes / 8192);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("mode grid boxes/stage stages KB_in_flight interleave  ms  GB/s(algorithmic: read, or read+write)\n");
  const int grids[] = {148, 128, 296};
  for (int mode = 0; mode < 2; ++mode)
    for (int gi = 0; gi < 3; ++gi)
      for (int bps : {1, 2, 4})
        for (int stages : {2, 4, 6, 8, 12, 16, 24})
          for (int il = 0; il < 2; ++il) {
            const int grid = grids[gi];
            const size_t smem = (size_t)stages * bps * 8192 + 2 * stages * 8 + 1024 + 64;
            const size_t cap = grid > 148 ? 110 * 1024 : 226 * 1024;
            if (smem > cap) continue;
            if (il == 1 && !(stages == 8 || stages == 4)) continue;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
              cudaEventRecord(e0);
              stream_kernel<<<grid, 128, smem>>>(tm, bps, stages, total_boxes, mode, il);
              cudaEventRecord(e1);
              cudaEventSynchronize(e1);
              float ms;
              cudaEventElapsedTime(&ms, e0, e1);
              if (ms < best) best = ms;
            }
            cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
            printf("%d %4d %2d %3d %5d %d  %.3f  %.0f\n", mode, grid, bps, stages, stages * bps * 8, il, best,
                   (mode ? 2.0 : 1.0) * bytes / best * 1e-6);
          }
  return 0;
}
