// Would packed FMAs (fma.rn.f32x2) pay in the replay kernel?  The replay's inner loop -- per step two LDS.128 of rotation
// parameters and 64 FMAs on an 8x8 register patch (column rotations) -- with scalar FFMA on r[i][j] against FFMA2 on the
// transposed patch (pairs of rows), 256 threads, two CTAs per SM, 127 steps x 64 repetitions.  Prints clocks per step.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int V, int MOVES> __global__ void __launch_bounds__(256, 2) k(float* out, const float4* par, int reps, long long* cyc) {
  __shared__ float4 q[64 * 32];
  for (int i = threadIdx.x; i < 64 * 32; i += 256) q[i] = par[i];
  __syncthreads();
  const int pc = threadIdx.x & 15;
  float r[8][8];
  unsigned long long rp[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[i][j] = (i == j) ? 1.f : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) rp[j][k2] = pk(j == 2 * k2 ? 1.f : 0.f, j == 2 * k2 + 1 ? 1.f : 0.f);
  long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
    for (int s4 = 0; s4 < 128; s4 += 4) {
#pragma unroll
    for (int sj = 0; sj < 4; ++sj) {
      const int s = s4 + sj;
      const float4 a = q[(s & 63) * 32 + pc], b = q[(s & 63) * 32 + 16 + pc];
      const float qx[4] = {a.x, a.z, b.x, b.z}, qy[4] = {a.y, a.w, b.y, b.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int p = kk, qq = 4 + ((kk + sj) & 3);
        if (V == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float u = r[i][p], v = r[i][qq];
            r[i][p] = fmaf(qx[kk], v, u);
            r[i][qq] = fmaf(qy[kk], u, v);
          }
        } else {
          const unsigned long long X = pk(qx[kk], qx[kk]), Y = pk(qy[kk], qy[kk]);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const unsigned long long u = rp[p][k2], v = rp[qq][k2];
            rp[p][k2] = ffma2(X, v, u);
            rp[qq][k2] = ffma2(Y, u, v);
          }
        }
      }
    }
    if (MOVES) {
      const int src = (threadIdx.x & 16) | ((pc + 1) & 15);
      if (V == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 4; j < 8; ++j) r[i][j] = __shfl_sync(0xffffffffu, r[i][j], src);
      } else {
#pragma unroll
        for (int j = 4; j < 8; ++j)
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(rp[j][k2]));
            rp[j][k2] = pk(__shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src));
          }
      }
    }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += r[i][j];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) { float2 f = *reinterpret_cast<float2*>(&rp[j][k2]); acc += f.x + f.y; }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
int main() {
  float* out; float4* par; long long* c;
  cudaMalloc(&out, 296 * 256 * 4); cudaMalloc(&par, 128 * 32 * 16); cudaMalloc(&c, 296 * 8);
  float4* h = new float4[128 * 32];
  for (int i = 0; i < 128 * 32; ++i) h[i] = make_float4(1e-3f * (i % 7), -1e-3f * (i % 5), 2e-3f, -2e-3f);
  cudaMemcpy(par, h, 128 * 32 * 16, cudaMemcpyHostToDevice);
  const int reps = 64;
  for (int mv = 0; mv < 2; ++mv)
  for (int v = 0; v < 2; ++v)
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0 && mv == 0) k<0, 0><<<296, 256>>>(out, par, reps, c);
      if (v == 1 && mv == 0) k<1, 0><<<296, 256>>>(out, par, reps, c);
      if (v == 0 && mv == 1) k<0, 1><<<296, 256>>>(out, par, reps, c);
      if (v == 1 && mv == 1) k<1, 1><<<296, 256>>>(out, par, reps, c);
      cudaDeviceSynchronize();
      long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
      printf("variant %d (%s)%s: %.1f clocks per step (two CTAs per SM)\n", v, v ? "FFMA2, transposed patch" : "scalar FFMA", mv ? " + 32 shuffles per 4 steps" : "", (double)hc / (reps * 128.0));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
