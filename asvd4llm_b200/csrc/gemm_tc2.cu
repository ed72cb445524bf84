// CTA-pair tcgen05 GEMM for the low-rank forward (a7): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]), 16-bit operands, both
// K-major (activations [tokens, features], nn.Linear weights [out, in]), fp32 accumulation in tensor memory.
//
// One MMA spans TWO SMs (tcgen05.mma.cta_group::2, M = 256): a cluster of two CTAs owns a 256 x BN output tile; each CTA
// stages its own 128 rows of A and its own BN/2 rows of B, the tensor cores of the pair exchange the B halves, so the
// weight operand crosses L2 -> SM once per 256 output rows and each SM reads half of it from its own shared memory
// (the single-CTA kernel in gemm_tc.cu multicasts full B tiles instead and is bound by exactly that traffic at r = 256).
//   warp 0     TMA producer (both CTAs): A tile [128 x 64] + B half [BN/2 x 64] per stage, 128-byte swizzle, bytes
//                             counted on the LEADER's full barrier; a stage is refilled when the pair's MMAs retired it
//   warp 1     MMA issuer (leader CTA only): 4 x tcgen05.mma (M=256, N=BN, K=16) per stage into one of two TMEM
//                             accumulators; tcgen05.commit (multicast) frees the stage / publishes the accumulator in
//                             both CTAs
//   warp 2     TMEM allocator (cta_group::2, 512 columns)
//   warps 4-11 epilogue (both CTAs, own 128 rows): tcgen05.ld -> + bias -> 16-bit -> swizzled staging -> TMA store, 64
//                             columns at a time, overlapped with the next tile's MMAs through the second accumulator;
//                             "drained" arrives on the leader's barrier from both CTAs
// BN is a launch-time choice among 64 / 128 / 192 / 256 (256 unless N is smaller): ranks such as 1843 are not
// multiples of anything.  Ragged edges are TMA's: out-of-range rows / columns / K read as zero
// and are clipped on store.
#include "common.cuh"
#include "umma.cuh"
#include "gemm_tc.h"
#include <type_traits>
#include <stdlib.h>

namespace asvd {
namespace tc {

namespace g2 {
constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BK = 64;                  // 16-bit elements per stage row = 128 bytes = one swizzle atom
constexpr int THREADS = 384;
constexpr int STAGES = 5;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int B_BYTES = 128 * BK * 2;   // 16 KB: up to 128 rows of B per CTA
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NBUF = 2;                 // staging buffers per 128-column half
constexpr int STG_BYTES = BM * 128;     // [128 rows x 64 columns] 16-bit
constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
constexpr int BAR_OFFSET = STAGING_OFFSET + 2 * NBUF * STG_BYTES;
constexpr int SMEM_TOTAL = BAR_OFFSET + 256 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
}  // namespace g2

template <typename T> __device__ __forceinline__ uint32_t pack2_(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2_<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2_<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// STAT (calibration, SURVEY 8f N3 / act_aware_utils.py:64-74): the kernel also reduces |A| over the rows -- per column of
// the activation the sum (abs_mean) or the maximum (abs_max) -- into stat32[K] (fp32, pre-zeroed): the spare warp of every
// CTA walks 128 x 64 blocks of A with plain coalesced loads while the tensor pipe works (the blocks are the ones the TMA
// just pulled through L2: no second HBM pass), so the statistic is a side output of the GEMM that consumes the
// activation instead of a separate pass over it.
template <typename T, bool STAT>
__global__ void __launch_bounds__(g2::THREADS, 1)
gemm_tn2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const T* __restrict__ bias, int M, int N, int K, int BN,
                const T* __restrict__ Araw, int64_t lda, float* __restrict__ stat32, int stat_max) {
  using namespace g2;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);   // used in the leader only
  uint64_t* empty = full + STAGES;        // per CTA: stage free (multicast commit)
  uint64_t* tfull = empty + STAGES;       // [2] per CTA: accumulator ready (multicast commit)
  uint64_t* tempty = tfull + 2;           // [2] leader only: accumulator drained by both CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = ((M + 2 * BM - 1) / (2 * BM)) * num_n;
  const int num_k = (K + BK - 1) / BK;
  const int bnh = BN >> 1;                       // rows of B per CTA
  const int nhalves = (BN + 127) >> 7;           // 128-column halves with epilogue work

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * 4 * nhalves); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's barriers and tensor memory exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cid; t < num_tiles; t += nclusters) {
        const int m0 = (t / num_n) * (2 * BM) + crank * BM, n0 = (t % num_n) * BN + crank * bnh;
        for (int k = 0; k < num_k; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* a = smem + stage * STAGE_BYTES;
          if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (uint32_t)(A_BYTES + bnh * 128));   // both CTAs' tiles
          const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
          tma_load_2d_2sm(a, &tmA, bar, k * BK, m0);
          tma_load_2d_2sm(a + A_BYTES, &tmB, bar, k * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // whole warp, uniform control flow; one elected lane issues the MMAs and commits
      const uint32_t idesc = make_idesc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 2 * BM, BN);
      const uint64_t adesc0 = make_desc_kmajor_sw128(smem_u32(smem));
      const uint64_t bdesc0 = make_desc_kmajor_sw128(smem_u32(smem + A_BYTES));
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = cid; t < num_tiles; t += nclusters, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&tempty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int k = 0; k < num_k; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = adesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
            const uint32_t acc0 = k ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)          // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
              mma_f16_ss_2sm(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk ? 1u : acc0);
            tc_commit_2sm(&empty[stage], (uint16_t)3);     // the pair is done with the stage: both producers may refill
            if (k == num_k - 1) tc_commit_2sm(&tfull[buf], (uint16_t)3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (STAT && warp == 3) {
    using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
    const int kb = (K + 63) >> 6, mb = (M + 127) >> 7;
    for (int item = blockIdx.x; item < mb * kb; item += gridDim.x) {
      const int r0 = (item / kb) << 7, c = ((item % kb) << 6) + 2 * lane;       // K is even (16-byte rows): c + 1 < K
      if (c >= K) continue;
      const int r1 = min(M, r0 + 128);
      float s0 = 0.f, s1 = 0.f;
      unsigned m0u = 0u, m1u = 0u;                  // |x| >= 0: its bit pattern orders like the value, NaN above everything
      const T* p = Araw + (int64_t)r0 * lda + c;
#pragma unroll 4
      for (int r = r0; r < r1; ++r, p += lda) {
        const T2 v = *reinterpret_cast<const T2*>(p);
        const float a = fabsf(to_f32<T>(v.x)), b = fabsf(to_f32<T>(v.y));
        if (stat_max) { m0u = max(m0u, __float_as_uint(a)); m1u = max(m1u, __float_as_uint(b)); }
        else { s0 += a; s1 += b; }
      }
      if (stat_max) {
        atomicMax(reinterpret_cast<unsigned*>(stat32) + c, m0u);
        atomicMax(reinterpret_cast<unsigned*>(stat32) + c + 1, m1u);
      } else {
        atomicAdd(stat32 + c, s0);
        atomicAdd(stat32 + c + 1, s1);
      }
    }
  } else if (warp >= 4) {
    // epilogue: warp = (TMEM lane quadrant q, 128-column half h).  Accumulator -> registers -> + bias -> 16-bit ->
    // staging tile in shared memory (rows of 128 B, 16-byte chunks XOR row&7 = the 128-byte TMA swizzle) -> TMA store.
    const int q = (warp - 4) & 3, h = (warp - 4) >> 2;
    const bool active = h < nhalves;
    const int nsub = active ? ((BN - h * 128) >= 128 ? 2 : 1) : 0;     // 64-column groups of this half
    const int r = q * 32 + lane;                                      // row of the CTA's tile = TMEM lane
    unsigned char* stg_base = smem + STAGING_OFFSET + h * NBUF * STG_BYTES;
    int it = 0, nstore = 0;
    for (int t = cid; t < num_tiles; t += nclusters, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int m0 = (t / num_n) * (2 * BM) + crank * BM, n0 = (t % num_n) * BN;
      if (!active) continue;
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < nsub; ++sub) {
        unsigned char* stg = stg_base + (nstore % NBUF) * STG_BYTES;
        ++nstore;
        if (q == 0 && lane == 0) tma_store_wait_read<NBUF - 1>();   // the store that last used this buffer has read it
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * sub + cc;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + h * 128 + c * 32), v);
          tmem_ld_wait();
          const int col0 = n0 + h * 128 + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < N) f[j] += to_f32<T>(bias[col0 + j]);
          }
          unsigned char* rowp = stg + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (cc * 4 + j) ^ (r & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) =
                make_uint4(pack2_<T>(f[8 * j], f[8 * j + 1]), pack2_<T>(f[8 * j + 2], f[8 * j + 3]),
                           pack2_<T>(f[8 * j + 4], f[8 * j + 5]), pack2_<T>(f[8 * j + 6], f[8 * j + 7]));
          }
        }
        if (sub == nsub - 1) {                                       // accumulator fully read: hand it back to the leader
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
        if (q == 0 && lane == 0) {
          tma_store_2d(&tmC, stg, n0 + h * 128 + sub * 64, m0);
          tma_store_commit();
        }
      }
    }
    if (active && q == 0 && lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA leaves while its peer may still target its barriers / tensor memory
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// BN in {64, 128, 192, 256}.  Measured at M = 65 536 (profiles/r02_fwd_ab.jsonl): the widest tile always wins -- 256 / 192 /
// 128 columns ran r = 256 in 290 / 347 / 384 us, r = 1843 in 1829 / 1917 / 2425 us -- because a narrower tile halves
// the operand reuse of the MMA (the A tile is re-read per N tile) long before the padded columns of the last tile or
// the last wave of clusters cost as much.  So: 256, or the smallest multiple of 64 that covers N.
static int choose_bn(int M, int N, int nclusters_max) {
  (void)M; (void)nclusters_max;
  if (N >= 256) return 256;
  return (N + 63) / 64 * 64;
}

// returns 0 on launch, 1 if the operands do not meet TMA's alignment rules, <0 on error.  bn_force: 0 = choose.
// stat32 != nullptr: also reduce |A| over the rows into stat32[K] (sum, or maximum when stat_max), see the kernel.
template <typename T, bool STAT>
static int gemm_tn_tc2_impl(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                            cudaStream_t st, int bn_force, float* stat32, int stat_max) {
  using namespace g2;
  auto ok = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; };
  if (!ok(A, lda) || !ok(B, ldb) || !ok(C, ldc) || K < 1) return 1;
  static int sms_dev[ASVD_MAX_DEVICES] = {};
  static bool attr[ASVD_MAX_DEVICES] = {};
  const int slot = current_device_slot();
  if (!sms_dev[slot]) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_dev[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  if (!attr[slot]) {
    if (cudaFuncSetAttribute(gemm_tn2_kernel<T, STAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) return -2;
    attr[slot] = true;
  }
  const int ncl_max = sms_dev[slot] / 2;
  const int BN = bn_force ? bn_force : choose_bn(M, N, ncl_max);
  if (BN < 64 || BN > 256 || BN % 64) return -1;
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB, tmC;
  if (!make_tmap_2d(&tmA, dt, 2, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return -1;
  if (!make_tmap_2d(&tmB, dt, 2, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)BN / 2, BK)) return -1;
  if (!make_tmap_2d(&tmC, dt, 2, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, BM, 64)) return -1;
  const int64_t tiles = (int64_t)((M + 255) / 256) * ((N + BN - 1) / BN);
  int nclusters = (int)(tiles < ncl_max ? tiles : ncl_max);
  if (const char* mc = getenv("ASVD_B200_FWD_MAXCL")) {        // experiments: leave SMs to a kernel on another stream
    const int lim = atoi(mc);
    if (lim > 0 && lim < nclusters) nclusters = lim;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * nclusters);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = 2; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, gemm_tn2_kernel<T, STAT>, tmA, tmB, tmC, bias, M, N, K, BN, A, lda, stat32, stat_max) != cudaSuccess) return -2;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <typename T>
int gemm_tn_tc2(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                cudaStream_t st, int bn_force) {
  return gemm_tn_tc2_impl<T, false>(A, lda, B, ldb, C, ldc, bias, M, N, K, st, bn_force, nullptr, 0);
}
template <typename T>
int gemm_tn_tc2_stat(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
                     float* stat32, int stat_max, cudaStream_t st) {
  return gemm_tn_tc2_impl<T, true>(A, lda, B, ldb, C, ldc, bias, M, N, K, st, 0, stat32, stat_max);
}

template int gemm_tn_tc2_stat<__half>(const __half*, int64_t, const __half*, int64_t, __half*, int64_t, const __half*, int, int, int,
                                      float*, int, cudaStream_t);
template int gemm_tn_tc2_stat<__nv_bfloat16>(const __nv_bfloat16*, int64_t, const __nv_bfloat16*, int64_t, __nv_bfloat16*, int64_t,
                                             const __nv_bfloat16*, int, int, int, float*, int, cudaStream_t);
template int gemm_tn_tc2<__half>(const __half*, int64_t, const __half*, int64_t, __half*, int64_t, const __half*, int, int, int,
                                 cudaStream_t, int);
template int gemm_tn_tc2<__nv_bfloat16>(const __nv_bfloat16*, int64_t, const __nv_bfloat16*, int64_t, __nv_bfloat16*, int64_t,
                                        const __nv_bfloat16*, int, int, int, cudaStream_t, int);

}  // namespace tc
}  // namespace asvd
