// tcgen05 GEMM for the low-rank forward (a7): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]), 16-bit operands (both
// K-major, as activations [tokens, features] and nn.Linear weights [out, in] are), fp32 accumulation in TMEM.
//
// Persistent, warp-specialised, one CTA per SM, launched as CLUSTERS OF TWO CTAs that work on vertically adjacent
// output tiles (same columns): each CTA fetches half of the B tile and TMA-multicasts it to both, so the weight
// operand -- re-read from L2 once per 128 output rows -- crosses the L2 -> SM fabric once per 256 rows.  (Measured before
// this: x B^T at r = 256 moved 1.07 GB of weights through L2 next to 0.54 GB of activations and ran at the L2 cap.)
//   warp 0     TMA producer : 128 x 64 tile of A and half of the BN x 64 tile of B per stage (128-byte swizzle), 3 stages;
//                             a stage is refilled when the MMAs of BOTH CTAs have retired it
//   warp 1     MMA issuer   : one elected thread issues 4 x tcgen05.mma (M=128, N=BN, K=16) per stage into one of
//                             two TMEM accumulators; tcgen05.commit releases the stage / publishes the accumulator
//   warp 2     TMEM allocator (2 x BN columns)
//   warps 4-11 epilogue     : tcgen05.ld (32 lanes x 32 columns per warp and step) -> + bias -> 16-bit -> swizzled
//                             16 KB staging buffer per 128-column half, 64 columns at a time -> TMA store (full 128-byte
//                             lines), overlapped with the next tile's MMAs through the second accumulator
// SVDLinear.forward = two launches: t = x B^T, y = t A^T + b (the [tokens, r] intermediate stays L2-resident for
// the sizes of BASELINE config 4: 64 Ki x 256 x 2 B = 32 MiB per 128 MiB L2).
#include "common.cuh"
#include "umma.cuh"
#include <type_traits>

namespace asvd {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;            // 16-bit elements per stage row = 128 bytes = one swizzle atom
constexpr int GEMM_THREADS = 384;      // warps 0-3: producer / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int STAGING_HALF = BM * 128;       // 64 output columns of one 128-column half: [128 rows x 128 B], used twice per tile

// NBUF = staging buffers per 128-column half.  Long-K products (x B^T) are bound by activation bytes in flight: 4 stages
// (3 in flight = 48 KB per SM; scripts/probes/stream_probe.cu: 32 KB caps at 5.2 TB/s) and one staging buffer.  Short-K
// products (t A^T) are bound by output stores in flight: 3 stages and two staging buffers per half, so two 16 KB TMA
// stores per half are outstanding.
template <int BN, int NBUF> struct GemmSmem {
  static constexpr int STAGES = NBUF == 1 ? 4 : 3;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGING_OFFSET + (BN / 128) * NBUF * STAGING_HALF;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + tmem pointer + alignment slack
};

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// K schedule: the contraction runs over nseg segments of nk k-steps; segment s reads A columns a_off[s] + k and B
// columns b_off[s] + k.  A plain GEMM is one segment at offset 0.  The recovery GEMM of the SVD (fp32 x 16-bit product
// evaluated exactly with bf16 planes: A = [a1|a2|a3], B = [b1|b2]) runs a1 b1 + a2 b1 + a3 b1 + a1 b2 + a2 b2 as five
// segments accumulating into the same TMEM tile.
struct KSched {
  int nseg, nk;
  int a_off[6], b_off[6];
};

template <typename T, int BN, int NBUF, bool F32OUT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const T* __restrict__ bias, int M, int N, int K,
               float* __restrict__ Cf, int64_t ldcf, const float* __restrict__ cscale, const KSched ks) {
  using S = GemmSmem<BN, NBUF>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // [2] accumulator ready
  uint64_t* tempty = tfull + 2;          // [2] accumulator drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = ((num_m + 1) / 2) * num_n;          // pairs of vertically adjacent tiles, one pair per cluster
  const int num_k = ks.nseg * ks.nk;
  const int crank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!F32OUT) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * (BN / 128)); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cid; t < num_tiles; t += nclusters) {
        const int m0 = ((t / num_n) * 2 + crank) * BM, n0 = (t % num_n) * BN;
        for (int k = 0; k < num_k; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* a = smem + stage * S::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);     // own A tile + both halves of the B tile
          const int seg = k / ks.nk, kk = (k - seg * ks.nk) * BK;
          tma_load_2d(a, &tmA, &full[stage], ks.a_off[seg] + kk, m0);
          tma_load_2d_multicast(a + S::A_BYTES + crank * (S::B_BYTES / 2), &tmB, &full[stage], ks.b_off[seg] + kk,
                                n0 + crank * (BN / 2), (uint16_t)3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp, uniform control flow; one elected lane issues the MMAs and commits (see elect_one)
    constexpr uint32_t idesc = make_idesc(sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1 : 0, BM, BN);
    const uint64_t adesc0 = make_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t bdesc0 = make_desc_kmajor_sw128(smem_u32(smem + S::A_BYTES));
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = cid; t < num_tiles; t += nclusters, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&tempty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
      for (int k = 0; k < num_k; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = adesc0 + (uint64_t)(stage * (S::STAGE_BYTES >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (S::STAGE_BYTES >> 4));
          const uint32_t acc0 = k ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)          // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
            mma_f16_ss(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk ? 1u : acc0);
          tc_commit_multicast(&empty[stage], (uint16_t)3);     // this CTA is done with the stage: tell both producers
          if (k == num_k - 1) tc_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // epilogue: warp = (TMEM lane quadrant q, 128-column half h).  Accumulator -> registers -> + bias -> 16-bit ->
    // staging tile in shared memory (rows of 128 B, 16-byte chunks XOR row&7 = the 128-byte TMA swizzle) -> TMA store.
    const int q = (warp - 4) & 3, h = (warp - 4) >> 2;
    const bool active = h < BN / 128;
    const int r = q * 32 + lane;                                  // row of the tile = TMEM lane
    unsigned char* stg_base = smem + S::STAGING_OFFSET + h * NBUF * STAGING_HALF;
    int it = 0, nstore = 0;
    for (int t = cid; t < num_tiles; t += nclusters, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int m0 = ((t / num_n) * 2 + crank) * BM, n0 = (t % num_n) * BN;
      if (!active) continue;
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      if constexpr (F32OUT) {
        // fp32 result straight from the accumulator to global memory (each thread owns a row: 128 contiguous bytes per
        // 32-column step), optionally scaled per column
        const int row = m0 + r;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * 128 + c * 32), v);
          tmem_ld_wait();
          const int col0 = n0 + h * 128 + c * 32;
          if (row < M) {
            float* dst = Cf + (int64_t)row * ldcf + col0;
            if (col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
                if (cscale) {
                  const float4 sc = *reinterpret_cast<const float4*>(cscale + col0 + 4 * j);
                  o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
                }
                reinterpret_cast<float4*>(dst)[j] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) dst[j] = __uint_as_float(v[j]) * (cscale ? cscale[col0 + j] : 1.f);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
      } else {
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {                          // 64 columns at a time through a 16 KB staging buffer
        // the store that last used this staging buffer must have finished reading it
        unsigned char* stg = stg_base + (nstore % NBUF) * STAGING_HALF;
        ++nstore;
        if (q == 0 && lane == 0) tma_store_wait_read<NBUF - 1>();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * sub + cc;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * 128 + c * 32), v);
          tmem_ld_wait();
          const int col0 = n0 + h * 128 + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < N) f[j] += to_f32<T>(bias[col0 + j]);
          }
          // 32 columns = 64 bytes = chunks 4*cc .. +3 of the 128-byte row
          unsigned char* rowp = stg + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (cc * 4 + j) ^ (r & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) =
                make_uint4(pack2<T>(f[8 * j], f[8 * j + 1]), pack2<T>(f[8 * j + 2], f[8 * j + 3]),
                           pack2<T>(f[8 * j + 4], f[8 * j + 5]), pack2<T>(f[8 * j + 6], f[8 * j + 7]));
          }
        }
        if (sub == 1) {                                            // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
        if (q == 0 && lane == 0) {
          tma_store_2d(&tmC, stg, n0 + h * 128 + sub * 64, m0);
          tma_store_commit();
        }
      }
      }
    }
    if (!F32OUT && active && q == 0 && lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

template <typename T, int BN, int NBUF, bool F32OUT = false>
static cudaError_t launch_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const T* bias, int M, int N,
                             int K, int sms, cudaStream_t st, float* Cf = nullptr, int64_t ldcf = 0, const float* cscale = nullptr,
                             const KSched* sched = nullptr) {
  using S = GemmSmem<BN, NBUF>;
  static bool attr[ASVD_MAX_DEVICES] = {};
  const int dev = current_device_slot();
  if (!attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<T, BN, NBUF, F32OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const int pairs = (((M + BM - 1) / BM + 1) / 2) * ((N + BN - 1) / BN);
  const int nclusters = pairs < sms / 2 ? pairs : sms / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * nclusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = 2; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 1;
  KSched ks;
  if (sched) ks = *sched;
  else { memset(&ks, 0, sizeof(ks)); ks.nseg = 1; ks.nk = (K + BK - 1) / BK; }
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tn_kernel<T, BN, NBUF, F32OUT>, tmA, tmB, tmC, bias, M, N, K, Cf, ldcf, cscale, ks);
  if (le != cudaSuccess) return le;
  return cudaGetLastError();
}

// returns 0 on launch, 1 if the operands do not meet TMA's alignment rules (caller uses the SIMT kernel), <0 on error
template <typename T>
int gemm_tn_tc(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, const T* bias, int M, int N, int K,
               cudaStream_t st) {
  auto ok = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; };
  if (!ok(A, lda) || !ok(B, ldb) || !ok(C, ldc) || K < 8) return 1;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB, tmC;
  const int BN = (N > 128) ? 256 : 128;
  if (!make_tmap_2d(&tmA, dt, 2, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return -1;
  if (!make_tmap_2d(&tmB, dt, 2, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)BN / 2, BK)) return -1;   // half tile per CTA
  if (!make_tmap_2d(&tmC, dt, 2, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, BM, 64)) return -1;
  const bool short_k = K <= 512;
  cudaError_t e = (BN == 256) ? (short_k ? launch_tn<T, 256, 2>(tmA, tmB, tmC, bias, M, N, K, sms, st)
                                         : launch_tn<T, 256, 1>(tmA, tmB, tmC, bias, M, N, K, sms, st))
                              : (short_k ? launch_tn<T, 128, 2>(tmA, tmB, tmC, bias, M, N, K, sms, st)
                                         : launch_tn<T, 128, 1>(tmA, tmB, tmC, bias, M, N, K, sms, st));
  return e == cudaSuccess ? 0 : -2;
}

// C[M,N] (fp32, overwritten) = sum over plane pairs of A_i[:, k] * B_j[:, k]^T for k in [k_begin, k_begin + k_len), bf16 planes
int gemm_planes_f32(const __nv_bfloat16* A, int64_t lda, int a_planes, const __nv_bfloat16* B, int64_t ldb, int b_planes,
                    float* C, int64_t ldc, int M, int N, int kseg, int k_begin, int k_len, const float* cscale,
                    cudaStream_t st) {
  auto ok = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; };
  if (!ok(A, lda) || !ok(B, ldb) || kseg % BK != 0 || k_begin % BK != 0 || k_len % BK != 0 || k_len <= 0 || k_begin + k_len > kseg ||
      a_planes < 1 || a_planes > 3 || b_planes < 1 || b_planes > 2)
    return 1;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  CUtensorMap tmA, tmB, tmC;
  constexpr int BN = 256;
  if (!make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, (uint64_t)M, (uint64_t)a_planes * kseg, (uint64_t)lda, BM, BK)) return -1;
  if (!make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, (uint64_t)N, (uint64_t)b_planes * kseg, (uint64_t)ldb, BN / 2, BK)) return -1;
  tmC = tmA;                                   // unused by the fp32 epilogue
  // Plane products a_i b_j with i + j <= 2 (the others are below fp32 resolution: 2^-8(i+j) relative), SMALLEST FIRST.
  // The tensor core adds each K = 16 group of products into the fp32 accumulator with truncation, about half an ulp of
  // the accumulator per instruction and always the same way: with the leading a1 b1 segment first, the 1280
  // instructions of a 4096-long contraction biased sigma by 4e-5..5e-4 relative (measured against fp64).  Accumulating
  // the correction planes while the accumulator is still small leaves only the 256 instructions of a1 b1 at full
  // magnitude -- and the caller splits long contractions into runs of at most ~1024 (k_begin, k_len), one fp32 partial
  // result each, summed with ordinary rounding afterwards.
  KSched ks;
  memset(&ks, 0, sizeof(ks));
  ks.nk = k_len / BK;
  for (int sum = 2; sum >= 0; --sum)
    for (int j = b_planes - 1; j >= 0; --j) {
      const int i = sum - j;
      if (i < 0 || i >= a_planes || ks.nseg >= 6) continue;
      ks.a_off[ks.nseg] = i * kseg + k_begin; ks.b_off[ks.nseg] = j * kseg + k_begin; ++ks.nseg;
    }
  cudaError_t e = launch_tn<__nv_bfloat16, BN, 1, true>(tmA, tmB, tmC, nullptr, M, N, kseg, sms, st, C, ldc, cscale, &ks);
  return e == cudaSuccess ? 0 : -2;
}

template int gemm_tn_tc<__half>(const __half*, int64_t, const __half*, int64_t, __half*, int64_t, const __half*, int, int, int, cudaStream_t);
template int gemm_tn_tc<__nv_bfloat16>(const __nv_bfloat16*, int64_t, const __nv_bfloat16*, int64_t, __nv_bfloat16*, int64_t,
                                       const __nv_bfloat16*, int, int, int, cudaStream_t);

}  // namespace tc
}  // namespace asvd
