// Throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a: 16 independent accumulators per thread, 1024 threads per SM.
// Variants: 0 = scalar FFMA, 1 = FFMA2 aligned pairs, 2 = FFMA2 with the b operand's halves swapped, 3 = FFMA2 with a
// broadcast multiplier built from one scalar.  Prints lane-FMAs per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <int V> __global__ void __launch_bounds__(1024, 1) k(float2* p, float qx, float qy, int iters, long long* cyc) {
  float2 g[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) g[i] = p[(threadIdx.x * 16 + i) % 4096];
  const float2 q = make_float2(qx, qy);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float2 a = g[i], b = g[i + 8];
      if (V == 0) {
        g[i].x = fmaf(qx, b.x, a.x); g[i].y = fmaf(qx, b.y, a.y);
        g[i + 8].x = fmaf(qy, a.x, b.x); g[i + 8].y = fmaf(qy, a.y, b.y);
      } else if (V == 1) {
        g[i] = ffma2(q, b, a); g[i + 8] = ffma2(q, a, b);
      } else if (V == 2) {
        g[i] = ffma2(q, make_float2(b.y, b.x), a); g[i + 8] = ffma2(q, make_float2(a.y, a.x), b);
      } else {
        g[i] = ffma2(make_float2(qx, qx), b, a); g[i + 8] = ffma2(make_float2(qy, qy), a, b);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
#pragma unroll
  for (int i = 0; i < 16; ++i) p[(threadIdx.x * 16 + i) % 4096] = g[i];
}
int main() {
  float2* p; long long* c;
  cudaMalloc(&p, 4096 * sizeof(float2)); cudaMemset(p, 0, 4096 * sizeof(float2)); cudaMalloc(&c, 148 * 8);
  const int iters = 4096;
  for (int v = 0; v < 4; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) k<0><<<148, 1024>>>(p, 1e-3f, -1e-3f, iters, c);
      if (v == 1) k<1><<<148, 1024>>>(p, 1e-3f, -1e-3f, iters, c);
      if (v == 2) k<2><<<148, 1024>>>(p, 1e-3f, -1e-3f, iters, c);
      if (v == 3) k<3><<<148, 1024>>>(p, 1e-3f, -1e-3f, iters, c);
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double fmas = 1024.0 * 32 * iters;          // lane-FMAs per SM (16 accumulators x 2 lanes per iteration per thread)
    printf("variant %d: %lld cycles, %.1f lane-FMAs per clock per SM\n", v, h, fmas / h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
