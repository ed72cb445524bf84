// Fused SVDLinear.forward (a7) for ranks up to 256: y[M, m] = ((x[M, n] Bw[r, n]^T) -> module dtype) Aw[m, r]^T + bias in ONE
// kernel; the [tokens, r] intermediate never leaves the SM (upstream: two nn.Linear calls, modules/svd_linear.py:105-109).
// Selected with ASVD_B200_FWD=fused (the default is the pair of CTA-pair GEMMs in gemm_tc2.cu; see ops_misc.cu).
//
// One cluster of two CTAs (tcgen05.mma.cta_group::2, M = 256) owns a 256-row tile of x:
//   phase 1  T[256, R] = x_tile Bw^T   K loop over n; per 64-wide step each CTA loads its 128 rows of x and its R/2 rows of
//            Bw (R = r rounded up to 64); fp32 accumulator = TMEM columns [0, R)
//   drain    T -> 16-bit, packed two per column, written BACK INTO TENSOR MEMORY (columns [384, 384 + R/2), tcgen05.st):
//            the second product takes its A operand from tensor memory, so the intermediate costs no shared-memory
//            space or bandwidth.  (The first version parked T in shared memory; its phase 2 moved 115 B/clk through a
//            128 B/clk shared memory -- operand reads of both products' tiles, TMA landing, store staging -- and the
//            tensor pipe sat at 57 %, profiles/r02_ncu_fwd_fused256b.)
//   phase 2  for each 192-column chunk of y: Y = T Aw_chunk^T, K = R, A from tensor memory, B = Aw rows streamed from L2;
//            accumulators alternate between TMEM columns [0, 192) and [192, 384); the epilogue (+ bias -> 16-bit ->
//            swizzled staging -> TMA store, three teams of four warps, 64 columns each) overlaps the next chunk's MMAs
// Shared memory: ring of eleven 16 KB units (phase 1 uses two per step: x, Bw; phase 2 one: Aw) + 3 x 16 KB store staging.
// Work items are (tile, chunk range): the tiles of the last, partial wave of clusters are split between several
// clusters, each repeating phase 1 and taking a share of the chunks (65 536 tokens = 256 tiles on 74 clusters: 3 full
// waves + 34 tiles on 68 clusters at 3/4 of a wave instead of a 4th full wave).
// Why only r <= 256: TMEM holds 512 columns = fp32 T (R) during phase 1, then 16-bit T (R/2) + two output accumulators.
#include "common.cuh"
#include "umma.cuh"
#include "gemm_tc.h"
#include <type_traits>

namespace asvd {
namespace tc {

namespace fz {
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int TEAMS = 3;                // epilogue teams of four warps (one per TMEM lane quadrant), 64 output columns each
constexpr int CHUNK = 64 * TEAMS;       // columns of y per phase-2 accumulator
constexpr int THREADS = 32 * (4 + 4 * TEAMS);   // warps 0-3: producer / MMA / TMEM alloc / spare; 12 drain + epilogue warps
constexpr int UNIT = 16384;             // [128 rows x 128 B]
constexpr int NU = 11;                  // ring units
constexpr int RMAX = 256;
constexpr int T16_COL = 2 * CHUNK;      // TMEM column of the packed 16-bit intermediate (128 columns at most)
constexpr int STAGING_OFFSET = NU * UNIT;
constexpr int BAR_OFFSET = STAGING_OFFSET + TEAMS * UNIT;
constexpr int SMEM_TOTAL = BAR_OFFSET + 512 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
static_assert(T16_COL + RMAX / 2 <= 512, "tensor memory budget");
}  // namespace fz

template <typename T> __device__ __forceinline__ uint32_t pack2f(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2f<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2f<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct FusedSched {
  int num_tiles;     // 256-row tiles
  int nch;           // 192-column chunks of y
  int full;          // tiles processed whole (a multiple of the cluster count)
  int split;         // work items per tile of the tail
  int items;         // full + (num_tiles - full) * split
};

__device__ __forceinline__ void fused_item(const FusedSched& s, int w, int& tile, int& c0, int& c1) {
  if (w < s.full) { tile = w; c0 = 0; c1 = s.nch; return; }
  const int tw = w - s.full;
  tile = s.full + tw / s.split;
  const int part = tw % s.split;
  c0 = part * s.nch / s.split;
  c1 = (part + 1) * s.nch / s.split;
}

// columns of chunk c that exist, rounded up to the 64 an epilogue team handles (the MMA's N)
__device__ __forceinline__ int chunk_cols(int c, int m) {
  const int rem = m - c * fz::CHUNK;
  return rem >= fz::CHUNK ? fz::CHUNK : ((rem + 63) & ~63);
}

template <typename T>
__global__ void __launch_bounds__(fz::THREADS, 1)
lowrank_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY,
                     const T* __restrict__ bias, int M, int n, int R, int m, const FusedSched sched) {
  using namespace fz;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);   // [NU] leader only: unit landed (both CTAs' bytes)
  uint64_t* empty = full + NU;            // [NU] per CTA: unit free (multicast commit)
  uint64_t* tfull = empty + NU;           // [2]  per CTA: output accumulator written (multicast commit)
  uint64_t* tempty = tfull + 2;           // [2]  leader only: output accumulator drained by both CTAs
  uint64_t* t32full = tempty + 2;         // [1]  per CTA: phase-1 accumulator complete (multicast commit)
  uint64_t* tready = t32full + 1;         // [1]  leader only: the 16-bit intermediate is in tensor memory in both CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int num_k1 = (n + BK - 1) / BK;
  const int num_k2 = R / BK;
  const int rh = R >> 1;                         // rows of Bw per CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NU; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * 4 * TEAMS); }
    mbar_init(t32full, 1);
    mbar_init(tready, 2 * 4 * TEAMS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int u = 0; uint32_t ph = 0;
      auto take = [&](const void* map, uint32_t bytes_pair, int x, int y) {
        mbar_wait(&empty[u], ph ^ 1);
        if (crank == 0) mbar_arrive_expect_tx(&full[u], bytes_pair);
        tma_load_2d_2sm(ring + u * UNIT, map, mapa_u32(smem_u32(&full[u]), 0), x, y);
        if (++u == NU) { u = 0; ph ^= 1; }
      };
      for (int w = cid; w < sched.items; w += nclusters) {
        int tile, c0, c1;
        fused_item(sched, w, tile, c0, c1);
        const int m0 = tile * (2 * BM) + crank * BM;
        for (int k = 0; k < num_k1; ++k) {
          take(&tmX, 2u * UNIT, k * BK, m0);
          take(&tmB, 2u * (uint32_t)(rh * 128), k * BK, crank * rh);
        }
        for (int c = c0; c < c1; ++c) {
          const int nh = chunk_cols(c, m) >> 1;            // rows of Aw the pair's MMA takes from each CTA
          for (int kb = 0; kb < num_k2; ++kb) take(&tmA, 2u * (uint32_t)((CHUNK / 2) * 128), kb * BK, c * CHUNK + crank * nh);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (crank == 0) {
      const int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      const uint32_t idesc1 = make_idesc(fmt, 2 * BM, R);
      const uint64_t ring_desc = make_desc_kmajor_sw128(smem_u32(ring));
      int u = 0; uint32_t ph = 0;
      uint32_t use[2] = {0, 0};                     // uses of the output accumulators: columns [0,192) and [192,384)
      uint32_t item = 0;
      for (int w = cid; w < sched.items; w += nclusters, ++item) {
        int tile, c0, c1;
        fused_item(sched, w, tile, c0, c1);
        // phase 1 writes columns [0, R), which both output accumulators of the previous item overlap: both drained
        mbar_wait(&tempty[0], (use[0] & 1) ^ 1);
        mbar_wait(&tempty[1], (use[1] & 1) ^ 1);
        tc_fence_after();
        for (int k = 0; k < num_k1; ++k) {
          const int ua = u; const uint32_t pa = ph;
          if (++u == NU) { u = 0; ph ^= 1; }
          const int ub = u; const uint32_t pb = ph;
          if (++u == NU) { u = 0; ph ^= 1; }
          mbar_wait(&full[ua], pa);
          mbar_wait(&full[ub], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = ring_desc + (uint64_t)(ua * (UNIT >> 4));
            const uint64_t bdesc = ring_desc + (uint64_t)(ub * (UNIT >> 4));
            const uint32_t acc0 = k ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              mma_f16_ss_2sm(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc1, kk ? 1u : acc0);
            tc_commit_2sm(&empty[ua], (uint16_t)3);
            tc_commit_2sm(&empty[ub], (uint16_t)3);
            if (k == num_k1 - 1) tc_commit_2sm(t32full, (uint16_t)3);
          }
          __syncwarp();
        }
        // the 16-bit intermediate is in tensor memory (both CTAs); columns [0, R) are free again
        mbar_wait(tready, item & 1);
        tc_fence_after();
        const int nc = c1 - c0;
        for (int c = 0; c < nc; ++c) {
          const int b = c & 1;
          const uint32_t idesc2 = make_idesc(fmt, 2 * BM, chunk_cols(c0 + c, m));
          mbar_wait(&tempty[b], (use[b] & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * CHUNK);
          for (int kb = 0; kb < num_k2; ++kb) {
            const int uu = u; const uint32_t pu = ph;
            if (++u == NU) { u = 0; ph ^= 1; }
            mbar_wait(&full[uu], pu);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_tmem = tmem_base + (uint32_t)(T16_COL + kb * (BK / 2));     // 32 columns per 64-wide K block
              const uint64_t bdesc = ring_desc + (uint64_t)(uu * (UNIT >> 4));
              const uint32_t acc0 = kb ? 1u : 0u;
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)          // K = 16 = 8 packed columns of A, 32 bytes of B
                mma_f16_ts_2sm(d_tmem, a_tmem + (uint32_t)(kk * 8), bdesc + (uint64_t)(kk * 2), idesc2, kk ? 1u : acc0);
              tc_commit_2sm(&empty[uu], (uint16_t)3);
              if (kb == num_k2 - 1) tc_commit_2sm(&tfull[b], (uint16_t)3);
            }
            __syncwarp();
          }
          ++use[b];
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ drain + epilogue (both CTAs, own 128 rows)
    // 12 warps = 3 teams x 4 TMEM lane quadrants.  A team owns 64 columns of every output accumulator and one 16 KB
    // staging buffer: a chunk costs each warp two tcgen05.ld (issued together), one pass of packing, one TMA store per team.
    const int q = (warp - 4) & 3, team = (warp - 4) >> 2;
    const int r = q * 32 + lane;                                      // row of the CTA's tile = TMEM lane
    unsigned char* stg = smem + STAGING_OFFSET + team * UNIT;
    uint32_t use[2] = {0, 0};
    uint32_t item = 0;
    const uint32_t tempty_l[2] = {mapa_u32(smem_u32(&tempty[0]), 0), mapa_u32(smem_u32(&tempty[1]), 0)};
    const uint32_t tready_l = mapa_u32(smem_u32(tready), 0);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int w = cid; w < sched.items; w += nclusters, ++item) {
      int tile, c0, c1;
      fused_item(sched, w, tile, c0, c1);
      const int m0 = tile * (2 * BM) + crank * BM;
      // ---- drain: fp32 columns [64 g, 64 g + 64) -> 32 packed 16-bit columns at T16_COL + 32 g (g = team, team + 3, ...)
      mbar_wait(t32full, item & 1);
      tc_fence_after();
      for (int g = team; g * 64 < R; g += TEAMS) {
        uint32_t v0[32], v1[32], pk[32];
        tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(g * 64), v0);
        tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(g * 64 + 32), v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = pack2f<T>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
          pk[16 + j] = pack2f<T>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        }
        tmem_st_32x32b_x32(tmem_base + lane_base + (uint32_t)(T16_COL + g * 32), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tready_l);
      // ---- chunks of y
      const int nc = c1 - c0;
#pragma unroll 1
      for (int c = 0; c < nc; ++c) {
        const int b = c & 1;
        const int ncols = chunk_cols(c0 + c, m);
        const int n0 = (c0 + c) * CHUNK + team * 64;
        mbar_wait(&tfull[b], use[b] & 1);
        tc_fence_after();
        ++use[b];
        const bool mine = team * 64 < ncols;                          // warp-uniform
        uint32_t v0[32], v1[32];
        if (mine) {
          const uint32_t col = (uint32_t)(b * CHUNK + team * 64);
          tmem_ld_32x32b_x32(tmem_base + lane_base + col, v0);
          tmem_ld_32x32b_x32(tmem_base + lane_base + col + 32u, v1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_l[b]);              // this warp's share of the accumulator is in registers
        if (!mine) continue;
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + j < m) v0[j] = __float_as_uint(__uint_as_float(v0[j]) + to_f32<T>(bias[n0 + j]));
            if (n0 + 32 + j < m) v1[j] = __float_as_uint(__uint_as_float(v1[j]) + to_f32<T>(bias[n0 + 32 + j]));
          }
        }
        if (q == 0 && lane == 0) tma_store_wait_read<0>();            // the team's previous store has read the buffer
        asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
        unsigned char* rowp = stg + r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(rowp + ((j ^ (r & 7)) * 16)) =
              make_uint4(pack2f<T>(__uint_as_float(v0[8 * j]), __uint_as_float(v0[8 * j + 1])),
                         pack2f<T>(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3])),
                         pack2f<T>(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5])),
                         pack2f<T>(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7])));
          *reinterpret_cast<uint4*>(rowp + (((4 + j) ^ (r & 7)) * 16)) =
              make_uint4(pack2f<T>(__uint_as_float(v1[8 * j]), __uint_as_float(v1[8 * j + 1])),
                         pack2f<T>(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3])),
                         pack2f<T>(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5])),
                         pack2f<T>(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7])));
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
        if (q == 0 && lane == 0) {
          tma_store_2d(&tmY, stg, n0, m0);
          tma_store_commit();
        }
      }
    }
    if (q == 0 && lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// 0 = launched, 1 = not eligible (rank above 256 or operands TMA cannot address), < 0 = error
template <typename T>
int lowrank_fused(const T* x, int64_t ldx, int M, int n, const T* Bw, int64_t ldb, int r, const T* Aw, int64_t lda, int m,
                  const T* bias, T* y, int64_t ldy, cudaStream_t st) {
  using namespace fz;
  auto ok = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; };
  if (r < 1 || r > RMAX || !ok(x, ldx) || !ok(Bw, ldb) || !ok(Aw, lda) || !ok(y, ldy)) return 1;
  static int sms_dev[ASVD_MAX_DEVICES] = {};
  static bool attr[ASVD_MAX_DEVICES] = {};
  const int slot = current_device_slot();
  if (!sms_dev[slot]) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_dev[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  if (!attr[slot]) {
    if (cudaFuncSetAttribute(lowrank_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) return -2;
    attr[slot] = true;
  }
  const int R = (r + 63) / 64 * 64;
  const int ncl_max = sms_dev[slot] / 2;
  FusedSched s;
  s.num_tiles = (M + 2 * BM - 1) / (2 * BM);
  s.nch = (m + CHUNK - 1) / CHUNK;
  s.full = (s.num_tiles / ncl_max) * ncl_max;
  const int tail = s.num_tiles - s.full;
  s.split = 1;
  if (tail > 0) {
    s.split = ncl_max / tail;
    if (s.split > s.nch) s.split = s.nch;
    if (s.split < 1) s.split = 1;
  }
  s.items = s.full + tail * s.split;
  const int nclusters = s.items < ncl_max ? s.items : ncl_max;
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmX, tmB, tmA, tmY;
  if (!make_tmap_2d(&tmX, dt, 2, x, (uint64_t)M, (uint64_t)n, (uint64_t)ldx, BM, BK)) return -1;
  if (!make_tmap_2d(&tmB, dt, 2, Bw, (uint64_t)r, (uint64_t)n, (uint64_t)ldb, (uint32_t)R / 2, BK)) return -1;
  if (!make_tmap_2d(&tmA, dt, 2, Aw, (uint64_t)m, (uint64_t)r, (uint64_t)lda, CHUNK / 2, BK)) return -1;
  if (!make_tmap_2d(&tmY, dt, 2, y, (uint64_t)M, (uint64_t)m, (uint64_t)ldy, BM, 64)) return -1;
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
  if (cudaLaunchKernelEx(&cfg, lowrank_fused_kernel<T>, tmX, tmB, tmA, tmY, bias, M, n, R, m, s) != cudaSuccess) return -2;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template int lowrank_fused<__half>(const __half*, int64_t, int, int, const __half*, int64_t, int, const __half*, int64_t, int,
                                   const __half*, __half*, int64_t, cudaStream_t);
template int lowrank_fused<__nv_bfloat16>(const __nv_bfloat16*, int64_t, int, int, const __nv_bfloat16*, int64_t, int,
                                          const __nv_bfloat16*, int64_t, int, const __nv_bfloat16*, __nv_bfloat16*, int64_t,
                                          cudaStream_t);

}  // namespace tc
}  // namespace asvd
