// Tensor-core (tcgen05, kind::tf32) bodies of the two streaming passes of the block-Jacobi SVD.
//
// fp32 accuracy on a TF32 pipe: every fp32 operand x is split as x = hi + lo with hi = x rounded or truncated to tf32 (exactly
// representable, so the MMA's internal truncation is a no-op) and lo = x - hi (|lo| <= 2^-11 |x|); a product is
// accumulated as hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator, dropping only the 2^-22 lo*lo term.  The
// split is done in shared memory by dedicated warps between the TMA landing and the MMA issue.
//
// gram_tc_kernel: G_c = P_c P_c^T for the 128-vector panel of a block pair over a chunk of columns.  Both MMA
// operands are the SAME K-major tile (vectors are rows of X, the contraction runs along the contiguous dimension).
//   warp 0  TMA producer (two 64-row boxes per stage: block I rows, block J rows), 128-byte swizzle
//   warp 1  MMA issuer (whole warp, one elected lane): 4 K-steps of tcgen05.mma M=128 N=128 K=8 per stage
//   warp 2  TMEM allocator
//   warps 4-7  epilogue: tcgen05.ld -> global partial Gram
//   warps 8-15 split (precise mode, two groups of four taking the tiles in turn): thread t owns row t of the landed tile;
//              hi = x with the low mantissa bits cleared (the raw tile stays as it is), the
//              A operand (0.5 hi | lo) goes to TENSOR MEMORY (tcgen05.st)
// Precise mode computes only T = (0.5 HI + LO) HI^T -- two MMAs per K-step, A from tensor memory, B = the hi tile
// in shared memory -- and the solve kernel forms G = T + T^T = HI HI^T + LO HI^T + HI LO^T.  Against three
// shared-memory MMAs per K-step this halves the shared-memory traffic per landed byte (the old bound of this
// kernel: 10 bytes moved per byte landed) and needs no lo tile in shared memory at all.
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace asvd {
namespace tc {

constexpr int GR_NH = 10;                        // landing ring (TMA -> hi tiles): 10 x 16 KB in flight per SM
constexpr int GR_NA = 4;                         // A-operand ring in tensor memory: 64 columns (0.5 hi | lo) per stage
constexpr int GR_HI_BYTES = 128 * 128;           // 128 rows x 32 floats
constexpr int GR_BAR_OFFSET = GR_NH * GR_HI_BYTES;
constexpr uint32_t GR_TMEM_A = 256;              // accumulators at columns 0 and 128, A ring from column 256
constexpr int GR_SMEM = GR_BAR_OFFSET + 512 + 1024;
constexpr int GR_THREADS = 512;               // 16 warps: TMA, MMA, TMEM alloc, (idle), 4 epilogue, 2 x 4 split

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// hi/lo split of a 16 KB tile by 128 threads (layout-agnostic: same offsets in both tiles).  hi = x with the 13 low
// mantissa bits cleared -- what the tensor core makes of the RAW tile when it reads it as a kind::tf32 operand --, so the
// landed tile is left as it is and only lo = x - hi (exact) is written: a quarter less shared-memory traffic in the
// split than rounding hi and writing it back in place (what this function did until round 2's second session), and no
// generic-proxy write to a tile the epilogue later overwrites.  All loads are issued before the first store.
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split_tile(const float4* hi, float4* lo, int t) {
  float4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = hi[t + 128 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    lo[t + 128 * j] = make_float4(v[j].x - trunc_tf32(v[j].x), v[j].y - trunc_tf32(v[j].y), v[j].z - trunc_tf32(v[j].z),
                                  v[j].w - trunc_tf32(v[j].w));
}

// One launch serves the matrices whose mode precise_b[b] equals `precise` (the mode is a per-matrix state, so that a
// weight's arithmetic never depends on its batch-mates; a round of a batch in transition is two launches).
// precise = 1: 3-term split (fp32-accurate Gram).  precise = 0: one TF32 pass on the raw tile (the MMA ignores the low
// mantissa bits); used while the off-diagonal cosines are still >= 1e-2, where 1e-3 relative accuracy of G only
// perturbs the rotation angles (R stays exactly orthogonal, so nothing is lost but a little convergence speed).
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmX, const int2* __restrict__ pairs, int pairs_per_mat, int chunks,
               int chunk_cols, int len_pad, int nv_pad, int n_items, float* __restrict__ Gpart,
               const int* __restrict__ done, int precise, const int* __restrict__ precise_b, const int* __restrict__ track) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GR_BAR_OFFSET);   // [NH] TMA landed
  uint64_t* empty = full + GR_NH;                                        // [NH] MMAs reading the hi tile retired
  uint64_t* a_ready = empty + GR_NH;                                     // [NA] split done (hi rewritten, A in TMEM)
  uint64_t* a_empty = a_ready + GR_NA;                                   // [NA] MMAs reading the A stage retired
  uint64_t* tfull = a_empty + GR_NA;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmX);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GR_NH; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < GR_NA; ++i) { mbar_init(&a_ready[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_mat = pairs_per_mat * chunks;

  if (warp == 0) {
    if (lane == 0) {
      int hs = 0; uint32_t hph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / per_mat, p = (item % per_mat) / chunks, c = item % chunks;
        if (done[b] || precise_b[b] != precise) continue;   // the launch serves the matrices in ITS Gram mode
        const int2 pr = pairs[p];
        if (pair_is_clean(track, nv_pad / JB, b, pr.x, pr.y)) continue;
        const int k0 = c * chunk_cols, k1 = min(len_pad, k0 + chunk_cols);
        for (int k = k0; k < k1; k += 32) {
          mbar_wait(&empty[hs], hph ^ 1);
          unsigned char* hi = smem + hs * GR_HI_BYTES;
          mbar_arrive_expect_tx(&full[hs], GR_HI_BYTES);
          // block-tiled X: tile (block, k/32) is one contiguous 8 KB chunk = rows [tile*64, tile*64+64) of a [.., 32] view
          const int nb = nv_pad / JB, nct = len_pad >> 5;
          tma_load_2d(hi, &tmX, &full[hs], 0, ((b * nb + pr.x) * nct + (k >> 5)) * JB);
          tma_load_2d(hi + GR_HI_BYTES / 2, &tmX, &full[hs], 0, ((b * nb + pr.y) * nct + (k >> 5)) * JB);
          if (++hs == GR_NH) { hs = 0; hph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the (uniform) loop; one elected lane issues the MMAs and their commits
    constexpr uint32_t idesc = make_idesc(2, 128, 128);
    const uint64_t dh0 = make_desc_kmajor_sw128(smem_u32(smem));
    int hs = 0; uint32_t hph = 0;
    int as = 0; uint32_t aph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / per_mat, c = item % chunks;
      if (done[b] || precise_b[b] != precise) continue;   // the launch serves the matrices in ITS Gram mode
      { const int2 pr = pairs[(item % per_mat) / chunks]; if (pair_is_clean(track, nv_pad / JB, b, pr.x, pr.y)) continue; }
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      ++it;
      mbar_wait(&tempty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
      const int k0 = c * chunk_cols, k1 = min(len_pad, k0 + chunk_cols);
      for (int k = k0; k < k1; k += 32) {
        if (precise) mbar_wait(&a_ready[as], aph);
        else mbar_wait(&full[hs], hph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dh = dh0 + (uint64_t)(hs * (GR_HI_BYTES >> 4));
          const uint32_t acc0 = (k != k0) ? 1u : 0u;
          if (precise) {
            const uint32_t a_tmem = tmem_base + GR_TMEM_A + (uint32_t)(as * 64);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {              // K = 8 tf32 = 32 bytes of B, 8 columns of A per step
              const uint64_t o = (uint64_t)(kk * 2);
              mma_tf32_ts(d_tmem, a_tmem + kk * 8, dh + o, idesc, kk ? 1u : acc0);          // 0.5 hi * hi^T
              mma_tf32_ts(d_tmem, a_tmem + 32 + kk * 8, dh + o, idesc, 1u);                 // lo * hi^T
            }
            tc_commit(&a_empty[as]);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t o = (uint64_t)(kk * 2);
              mma_tf32_ss(d_tmem, dh + o, dh + o, idesc, kk ? 1u : acc0);
            }
          }
          tc_commit(&empty[hs]);
          if (k + 32 >= k1) tc_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++hs == GR_NH) { hs = 0; hph ^= 1; }
        if (precise) { if (++as == GR_NA) { as = 0; aph ^= 1; } }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / per_mat;
      if (done[b] || precise_b[b] != precise) continue;   // the launch serves the matrices in ITS Gram mode
      { const int2 pr = pairs[(item % per_mat) / chunks]; if (pair_is_clean(track, nv_pad / JB, b, pr.x, pr.y)) continue; }
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      ++it;
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      float* G = Gpart + (int64_t)item * (JK * JK) + (q * 32 + lane) * JK;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + c * 32), v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(G + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  } else if (warp >= 8 && precise) {
    // Two groups of four split warps take the landed tiles in turn (group = stage parity).  One group alone has to get
    // through load -> round -> store -> tcgen05.st -> wait -> fence -> arrive once per 16 KB tile, ~780 clocks at the
    // HBM rate: measured, the accurate pass ran at 0.65 of the copy bandwidth against 0.9 for the single pass.
    const int q = (warp - 8) & 3, grp = (warp - 8) >> 2;
    const int t = q * 32 + lane;                             // row of the tile = TMEM lane; warps 8+q and 12+q own lanes 32q..
    int hs = 0; uint32_t hph = 0;
    int as = 0; uint32_t aph = 0;
    int stage = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / per_mat, c = item % chunks;
      if (done[b] || precise_b[b] != precise) continue;   // the launch serves the matrices in ITS Gram mode
      { const int2 pr = pairs[(item % per_mat) / chunks]; if (pair_is_clean(track, nv_pad / JB, b, pr.x, pr.y)) continue; }
      const int k0 = c * chunk_cols, k1 = min(len_pad, k0 + chunk_cols);
      for (int k = k0; k < k1; k += 32, ++stage) {
        if ((stage & 1) != grp) {                            // the other group's tile: only the ring positions advance
          if (++hs == GR_NH) { hs = 0; hph ^= 1; }
          if (++as == GR_NA) { as = 0; aph ^= 1; }
          continue;
        }
        mbar_wait(&full[hs], hph);
        // row t of the 128-byte-swizzled tile: 16-byte chunk j lives at chunk j ^ (t & 7)
        const uint32_t row = smem_u32(smem + hs * GR_HI_BYTES + t * 128);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)      // explicit shared-space loads (the 1 KB-aligned base is computed through an integer: generic otherwise)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "r"(row + ((j ^ (t & 7)) << 4)));
        uint32_t hh[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          float h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // hi = x with the 13 low mantissa bits cleared: exactly what the tensor core makes of the RAW tile it reads
            // as the B operand (kind::tf32 ignores those bits), so nothing has to be written back to shared memory --
            // the write-back of a rounded hi was a fifth of this mode's shared-memory traffic (80 KB moved per 16 KB
            // landed, against 640 of the 780 clocks a stage has at the HBM rate).  lo = x - hi is exact; the dropped
            // lo*lo term is <= 2^-20 per product with a common sign, i.e. ~3e-7 RELATIVE on a Gram entry.
            h[e] = __uint_as_float(__float_as_uint(x[e]) & 0xffffe000u);
            hh[4 * j + e] = __float_as_uint(0.5f * h[e]);
            lo[4 * j + e] = __float_as_uint(x[e] - h[e]);
          }
        }
        mbar_wait(&a_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + ((uint32_t)(q * 32) << 16) + GR_TMEM_A + (uint32_t)(as * 64);
        tmem_st_32x32b_x32(a_tmem, hh);
        tmem_st_32x32b_x32(a_tmem + 32, lo);
        tmem_st_wait();
        tc_fence_before();                                   // (no proxy fence: this mode no longer writes shared memory)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[as]);
        if (++hs == GR_NH) { hs = 0; hph ^= 1; }
        if (++as == GR_NA) { as = 0; aph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ update
// update_tc_kernel: panel <- R^T panel for one block pair and a run of 64-column tiles:
//   D[j][c] = sum_i Rt[j][i] X[i][c]     M = 128 (j), N = 64 (c), K = 128 (i)
// A = R^T (hi and lo halves) is written ONCE per CTA into tensor memory (tcgen05.st) and read from there by every
// MMA; B = the X tile exactly as it lies in HBM (rows i, columns c contiguous: an MN-major operand), landed by TMA
// with the 128B/32B-atom swizzle and split into hi/lo in place.  D goes back over the same rows of X.
constexpr int UP_NH = 5;                           // landing ring: 5 x 32 KB in flight per SM
constexpr int UP_NL = 2;                           // lo ring
constexpr int UP_TN = 64;                          // columns per tile
constexpr int UP_HI_BYTES = 128 * UP_TN * 4;       // 32 KB: two groups of 128 rows x 32 floats
constexpr int UP_LO_OFFSET = UP_NH * UP_HI_BYTES;
constexpr int UP_BAR_OFFSET = UP_LO_OFFSET + UP_NL * UP_HI_BYTES;
constexpr int UP_SMEM = UP_BAR_OFFSET + 256 + 1024;
constexpr int UP_THREADS = 384;
constexpr uint32_t UP_TMEM_A_HI = 256, UP_TMEM_A_LO = 384;   // column offsets; D buffers at 0 and 64

__device__ __forceinline__ void update_gather_rt(const float* __restrict__ Rp, uint32_t tmem_base, int q, int lane, int half,
                                                 uint64_t* a_ready) {
  const int j = q * 32 + lane;
  const float* src = Rp + j;
#pragma unroll 1
  for (int c = 2 * half; c < 2 * half + 2; ++c) {
    float x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = src[(c * 32 + e) * JK];
    uint32_t h[32], l[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float hh = rna_tf32(x[e]);
      h[e] = __float_as_uint(hh);
      l[e] = __float_as_uint(x[e] - hh);
    }
    tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + UP_TMEM_A_HI + c * 32, h);
    tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + UP_TMEM_A_LO + c * 32, l);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(a_ready);
}

__global__ void __launch_bounds__(UP_THREADS, 1)
update_tc_kernel(const __grid_constant__ CUtensorMap tmX, float* __restrict__ X, int64_t mat_stride, int ldx,
                 const int2* __restrict__ pairs, int pairs_per_mat, int nv_pad, int tiles_total, int tiles_per_cta,
                 const float* __restrict__ R, const int* __restrict__ pairflag, const int* __restrict__ done, int dbg) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + UP_BAR_OFFSET);
  uint64_t* empty = full + UP_NH;
  uint64_t* lo_ready = empty + UP_NH;
  uint64_t* lo_empty = lo_ready + UP_NL;
  uint64_t* tfull = lo_empty + UP_NL;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_ready = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int b = blockIdx.z, p = blockIdx.y;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  if (!pairflag[idx]) return;
  const int2 pr = pairs[p];
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int ntiles = min(tiles_per_cta, tiles_total - tile0);
  const int nb = nv_pad / JB, nct = tiles_total * (UP_TN / 32);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmX);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < UP_NH; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < UP_NL; ++i) { mbar_init(&lo_ready[i], 4); mbar_init(&lo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(a_ready, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int c0 = (tile0 + t) * UP_TN;
        mbar_wait(&empty[stage], phase ^ 1);
        unsigned char* hi = smem + stage * UP_HI_BYTES;
        mbar_arrive_expect_tx(&full[stage], UP_HI_BYTES);
        // box = 64 rows x 32 floats; rows 0-63 of the tile are block I, rows 64-127 block J; two 32-column groups
        for (int g = 0; g < 2; ++g) {
          tma_load_2d(hi + g * 16384, &tmX, &full[stage], 0, ((b * nb + pr.x) * nct + (c0 >> 5) + g) * JB);
          tma_load_2d(hi + g * 16384 + 8192, &tmX, &full[stage], 0, ((b * nb + pr.y) * nct + (c0 >> 5) + g) * JB);
        }
        if (++stage == UP_NH) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp, uniform control flow; one elected lane issues (see elect_one)
    constexpr uint32_t idesc = make_idesc(2, 128, UP_TN) | (1u << 16);     // B operand MN-major
    const uint64_t bh0 = make_desc_mnmajor_sw128_32b(smem_u32(smem), 16384);
    const uint64_t bl0 = make_desc_mnmajor_sw128_32b(smem_u32(smem + UP_LO_OFFSET), 16384);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0;
    int ls = 0; uint32_t lph = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t use = (uint32_t)(t >> 1);
      mbar_wait(&tempty[buf], (use & 1) ^ 1);
      mbar_wait(&lo_ready[ls], lph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * UP_TN);
        const uint64_t bh = bh0 + (uint64_t)(stage * (UP_HI_BYTES >> 4));
        const uint64_t bl = bl0 + (uint64_t)(ls * (UP_HI_BYTES >> 4));
        if (!(dbg & 2)) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {                    // K = 8 rows of the tile per step = 1024 bytes
            const uint64_t o = (uint64_t)(k * (1024 >> 4));
            mma_tf32_ts(d_tmem, tmem_base + UP_TMEM_A_HI + k * 8, bh + o, idesc, k ? 1u : 0u);
            mma_tf32_ts(d_tmem, tmem_base + UP_TMEM_A_LO + k * 8, bh + o, idesc, 1u);
            mma_tf32_ts(d_tmem, tmem_base + UP_TMEM_A_HI + k * 8, bl + o, idesc, 1u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            mma_tf32_ts(d_tmem, tmem_base + UP_TMEM_A_HI + k * 8, bh + (uint64_t)(k * (1024 >> 4)), idesc, k ? 1u : 0u);
        }
        tc_commit(&lo_empty[ls]);
        tc_commit(&tfull[buf]);          // the hi slot is handed to the epilogue, which overwrites it with D
      }
      __syncwarp();
      if (++stage == UP_NH) { stage = 0; phase ^= 1; }
      if (++ls == UP_NL) { ls = 0; lph ^= 1; }
    }
    (void)phase;
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4;
    const int j = q * 32 + lane;                           // output vector of this thread = TMEM lane
    // A = R^T: thread j gathers column j of R (coalesced across the warp) -> hi/lo -> tensor memory.  The four split
    // warps take the upper half of the columns (same TMEM lane quadrants), so the gather's L2 latency is paid twice,
    // not four times, before the first MMA.
    update_gather_rt(R + (int64_t)idx * (JK * JK), tmem_base, q, lane, 0, a_ready);
    // D tile -> the tile's own landing slot, in the layout the TMA wrote it (rows of 128 B, 32-byte chunks XOR row&3)
    // -> TMA store over the same rows of X.  The slot returns to the producer once the store has read it.
    const int row_off = j * 128;
    const int sw = j & 3;
    int stage = 0;
    int pending_stage = -1;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t use = (uint32_t)(t >> 1);
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      unsigned char* slot = smem + stage * UP_HI_BYTES;
#pragma unroll
      for (int g = 0; g < UP_TN / 32; ++g) {
        if (dbg & 4) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * UP_TN + g * 32), v);
        tmem_ld_wait();
        unsigned char* rowp = slot + g * 16384 + row_off;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {            // 32-byte chunk ch of the 128-byte row lands at chunk (ch ^ (row & 3))
          float4* dst = reinterpret_cast<float4*>(rowp + ((ch ^ sw) << 5));
          dst[0] = make_float4(__uint_as_float(v[8 * ch]), __uint_as_float(v[8 * ch + 1]), __uint_as_float(v[8 * ch + 2]),
                               __uint_as_float(v[8 * ch + 3]));
          dst[1] = make_float4(__uint_as_float(v[8 * ch + 4]), __uint_as_float(v[8 * ch + 5]), __uint_as_float(v[8 * ch + 6]),
                               __uint_as_float(v[8 * ch + 7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps
      if (warp == 4 && lane == 0) {
        const int c0 = (tile0 + t) * UP_TN;
        for (int g = 0; g < ((dbg & 8) ? 0 : 2); ++g) {
          tma_store_2d(&tmX, slot + g * 16384, 0, ((b * nb + pr.x) * nct + (c0 >> 5) + g) * JB);
          tma_store_2d(&tmX, slot + g * 16384 + 8192, 0, ((b * nb + pr.y) * nct + (c0 >> 5) + g) * JB);
        }
        tma_store_commit();
        if (pending_stage >= 0) {                              // previous tile's store has finished reading its slot
          tma_store_wait_read<1>();
          mbar_arrive(&empty[pending_stage]);
        }
        pending_stage = stage;
      }
      if (++stage == UP_NH) stage = 0;
    }
    if (warp == 4 && lane == 0) {
      tma_store_wait<0>();                                     // all stores complete (globally visible) before exit
      if (pending_stage >= 0) mbar_arrive(&empty[pending_stage]);
    }
  } else if (warp >= 8) {
    update_gather_rt(R + (int64_t)idx * (JK * JK), tmem_base, warp - 8, lane, 1, a_ready);
    const int t128 = threadIdx.x - 256;
    int stage = 0; uint32_t phase = 0;
    int ls = 0; uint32_t lph = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full[stage], phase);
      mbar_wait(&lo_empty[ls], lph ^ 1);
      float4* hi = reinterpret_cast<float4*>(smem + stage * UP_HI_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem + UP_LO_OFFSET + ls * UP_HI_BYTES);
      if (!(dbg & 1)) {
        split_tile(hi, lo, t128);
        split_tile(hi + 1024, lo + 1024, t128);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&lo_ready[ls]);
      if (++stage == UP_NH) { stage = 0; phase ^= 1; }
      if (++ls == UP_NL) { ls = 0; lph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
static int g_sms = 0;
static thread_local int g_sm_budget = 0;   // > 0: SMs the streaming passes may count on (the rest run another half's solve)
void set_sm_budget(int sms) { g_sm_budget = sms; }
static int sm_count() {
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return (g_sm_budget > 0 && g_sm_budget < g_sms) ? g_sm_budget : g_sms;
}

// tensor map over the block-tiled X of the whole batch, viewed as [batch * nv_pad * len_pad / 32 rows, 32 cols] fp32:
// a box of 64 rows x 32 floats is exactly one contiguous 8 KB tile
bool make_x_tmap(CUtensorMap* map, const float* X, int batch, int nv_pad, int len_pad) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)batch * nv_pad * (len_pad / 32), 32, 32, 64, 32);
}

cudaError_t launch_gram_tc(const CUtensorMap& tmX, const int2* pairs, int pairs_per_mat, int chunks, int chunk_cols,
                           int len_pad, int nv_pad, int batch, float* Gpart, const int* done, int precise, const int* precise_b, const int* track,
                           cudaStream_t st) {
  static bool attr[ASVD_MAX_DEVICES] = {};
  const int dev = current_device_slot();
  if (!attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const int n_items = batch * pairs_per_mat * chunks;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  gram_tc_kernel<<<grid, GR_THREADS, GR_SMEM, st>>>(tmX, pairs, pairs_per_mat, chunks, chunk_cols, len_pad, nv_pad, n_items,
                                                    Gpart, done, precise, precise_b, track);
  return cudaGetLastError();
}


// tensor map for the update: same matrix, box 64 rows x 32 floats, 128B swizzle with 32-byte atoms (MN-major operand)
bool make_x_tmap_mn(CUtensorMap* map, const float* X, int batch, int nv_pad, int len_pad) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)batch * nv_pad * (len_pad / 32), 32, 32, 64, 32,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

cudaError_t launch_update_tc(const CUtensorMap& tmX, float* X, int64_t mat_stride, int ldx, const int2* pairs,
                             int pairs_per_mat, int nv_pad, int len_pad, int batch, const float* R, const int* pairflag,
                             const int* done, cudaStream_t st) {
  static bool attr[ASVD_MAX_DEVICES] = {};
  const int dev = current_device_slot();
  if (!attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UP_SMEM);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const char* dbg_env = getenv("ASVD_B200_DBG_UPDATE");      // timing experiments only
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  const int tiles_total = len_pad / UP_TN;
  // One wave of CTAs: every CTA pays a fixed ~15 us (TMEM allocation, gathering R^T into tensor memory, pipeline fill
  // and drain), so a pair's column range is split only as far as needed to occupy the SMs once.
  int ctas_x = sm_count() / (pairs_per_mat * batch);
  if (ctas_x < 1) ctas_x = 1;
  int tiles_per_cta = (tiles_total + ctas_x - 1) / ctas_x;
  if (tiles_per_cta < 4) tiles_per_cta = tiles_total < 4 ? tiles_total : 4;
  ctas_x = (tiles_total + tiles_per_cta - 1) / tiles_per_cta;
  update_tc_kernel<<<dim3(ctas_x, pairs_per_mat, batch), UP_THREADS, UP_SMEM, st>>>(tmX, X, mat_stride, ldx, pairs, pairs_per_mat,
                                                                                  nv_pad, tiles_total, tiles_per_cta, R,
                                                                                  pairflag, done, dbg);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace asvd
