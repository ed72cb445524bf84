// Activation-scaled SVD by one-sided block Jacobi (Hestenes) — replaces modules/svd_linear.py:47-70 of the
// upstream tree (w.float() * scale -> torch.svd_lowrank -> V / scale) with an exact factorisation.
//
// Data layout (HBM): the n' = min(m,n) vectors to orthogonalise (columns of W*diag(s) when m >= n, rows
// otherwise) are the ROWS of X [nv_pad, len_pad] fp32; a block of JB=64 vectors is a contiguous slab and
// the long dimension is contiguous, so every pass below streams whole 128-byte lines.
//
// One round (nv_pad/JB - 1 rounds per sweep, nv_pad/(2*JB) disjoint block pairs per round, all pairs of all
// matrices of the batch in the same launch):
//   gram_kernel   : partial Gram matrices G_c = P_c P_c^T of each 128-vector panel over 512-column chunks
//   solve_kernel  : G = sum_c G_c in shared memory; cyclic two-sided Jacobi on the 128x128 G with the
//                   rotations accumulated in R; sort-by-norm swaps; Newton-Schulz polish of R
//   update_kernel : panel <- R^T panel  (in place, cp.async double-buffered column tiles)
// After convergence (max |cos| < tol at visit time over one whole sweep) the rows of X are sigma_j * u_j.
// The second factor is NOT accumulated: it is recovered from the ORIGINAL weight with one fp32 GEMM
// (Y = Xhat * W*diag(s)), which also anchors sigma_j = |Y_j| to the input and removes accumulated drift.
#include <stdarg.h>
#include <float.h>
#include <vector>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace asvd {

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ plan
SvdPlan make_plan(int m, int n, int batch) {
  SvdPlan p;
  memset(&p, 0, sizeof(p));
  p.m = m; p.n = n; p.batch = batch;
  p.tall = (m >= n);
  p.nv = p.tall ? n : m;
  p.len = p.tall ? m : n;
  p.nv_pad = (int)round_up(p.nv, JK);
  p.len_pad = (int)round_up(p.len, 128);
  p.ldy = (int)round_up(p.nv, 4);
  p.nb = p.nv_pad / JB;
  p.rounds = p.nb - 1;
  p.pairs = p.nb / 2;
  p.chunks = (p.len_pad + GRAM_CHUNK - 1) / GRAM_CHUNK;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (size_t)round_up((int64_t)(off + bytes), 256); return o; };
  p.off_ptrs = take(sizeof(void*) * 2 * (size_t)batch);
  p.off_pairs = take(sizeof(int2) * (size_t)p.rounds * p.pairs);
  p.off_X = take(sizeof(float) * (size_t)batch * p.nv_pad * p.len_pad);
  p.off_Y = take(sizeof(float) * (size_t)batch * p.nv_pad * p.ldy);
  p.off_G = take(sizeof(float) * (size_t)batch * p.pairs * p.chunks * JK * JK);
  p.off_R = take(sizeof(float) * (size_t)batch * p.pairs * JK * JK);
  p.off_flag = take(sizeof(int) * (size_t)batch * p.pairs);
  p.off_maxoff = take(sizeof(unsigned) * (size_t)batch);
  p.off_done = take(sizeof(int) * (size_t)batch);
  p.off_sigma = take(sizeof(float) * (size_t)batch * p.nv_pad);
  p.off_perm = take(sizeof(int) * (size_t)batch * p.nv_pad);
  p.off_status = take(sizeof(int) * (size_t)batch);
  p.off_scale = take(sizeof(float) * (size_t)batch * n);
  p.off_norm = take(sizeof(float) * (size_t)batch * p.nv_pad);
  p.bytes = off;
  return p;
}

// ------------------------------------------------------------------------------------------------ prep (a3)
// X = fp32(W) * diag(s), transposed when the vectors are the columns of W (tall case).
template <typename T>
__global__ void __launch_bounds__(256) prep_kernel(const void* const* __restrict__ Wptrs, const float* __restrict__ scale,
                                                   int64_t ldw, int m, int n, int tall, float* __restrict__ X,
                                                   int64_t mat_stride, int ldx) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const T* W = reinterpret_cast<const T*>(Wptrs[b]);
  const float* s = scale + (int64_t)b * n;
  float* Xb = X + b * mat_stride;
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (tall) {
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, j = j0 + tx;
      tile[r][tx] = (i < m && j < n) ? to_f32<T>(W[(int64_t)i * ldw + j]) * s[j] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int j = j0 + r, i = i0 + tx;
      if (j < n && i < m) Xb[(int64_t)j * ldx + i] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, j = j0 + tx;
      if (i < m && j < n) Xb[(int64_t)i * ldx + j] = to_f32<T>(W[(int64_t)i * ldw + j]) * s[j];
    }
  }
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ Gram
// Gpart[b][p][c] = P P^T over columns [c*GRAM_CHUNK, ...) of the 128-vector panel of pair p.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ X, int64_t mat_stride, int ldx,
                                                   const int2* __restrict__ pairs, int len_pad, int chunks,
                                                   int pairs_per_mat, float* __restrict__ Gpart,
                                                   const int* __restrict__ done) {
  __shared__ __align__(16) float As[2][GK][GLD];
  const int b = blockIdx.z, p = blockIdx.y, c = blockIdx.x;
  if (done[b]) return;
  const int2 pr = pairs[p];
  const float* Xb = X + b * mat_stride;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int row = t >> 1, kq = (t & 1) * 8;
  const int vec = (row < JB) ? pr.x * JB + row : pr.y * JB + (row - JB);
  const float* src = Xb + (int64_t)vec * ldx;
  const int kbeg = c * GRAM_CHUNK, kend = min(len_pad, kbeg + GRAM_CHUNK);
  const int nk = (kend - kbeg) / GK;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 r0, r1;
  auto gload = [&](int k0) {
    r0 = *reinterpret_cast<const float4*>(src + k0 + kq);
    r1 = *reinterpret_cast<const float4*>(src + k0 + kq + 4);
  };
  auto sstore = [&](int buf) {
    As[buf][kq + 0][row] = r0.x; As[buf][kq + 1][row] = r0.y; As[buf][kq + 2][row] = r0.z; As[buf][kq + 3][row] = r0.w;
    As[buf][kq + 4][row] = r1.x; As[buf][kq + 5][row] = r1.y; As[buf][kq + 6][row] = r1.z; As[buf][kq + 7][row] = r1.w;
  };
  gload(kbeg);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kbeg + (kt + 1) * GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  float* G = Gpart + (((int64_t)b * pairs_per_mat + p) * chunks + c) * (JK * JK);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    *reinterpret_cast<float4*>(&G[r * JK + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(&G[r * JK + 64 + tx * 4]) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// ------------------------------------------------------------------------------------------------ solve
constexpr int SLD = JK + 4;   // shared leading dimension (float4-aligned rows, 4-bank skew)
constexpr int SOLVE_THREADS = 1024;
constexpr size_t SOLVE_SMEM = sizeof(float) * (2 * JK * SLD) + sizeof(float2) * (JK / 2) + sizeof(float) * 64;

// round-robin (circle method) pairing of JK indices: step st in [0, JK-1), slot i in [0, JK/2)
__device__ __forceinline__ void rr_pair(int st, int i, int& p, int& q) {
  constexpr int N1 = JK - 1;
  int a, b;
  if (i == 0) { a = N1; b = st; }
  else { a = st + i; if (a >= N1) a -= N1; b = st - i; if (b < 0) b += N1; }
  p = min(a, b); q = max(a, b);
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1)
solve_kernel(const float* __restrict__ Gpart, int chunks, int pairs_per_mat, float* __restrict__ Rout,
             int* __restrict__ pairflag, unsigned* __restrict__ maxoff_bits, int* __restrict__ status,
             const int* __restrict__ done, float tol) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* G = reinterpret_cast<float*>(smem_raw);
  float* R = G + JK * SLD;
  float2* cs = reinterpret_cast<float2*>(R + JK * SLD);
  float* red = reinterpret_cast<float*>(cs + JK / 2);

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  const int tid = threadIdx.x;
  const float* Gp = Gpart + (int64_t)idx * chunks * (JK * JK);

  for (int e = tid; e < JK * JK; e += SOLVE_THREADS) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += Gp[(int64_t)c * (JK * JK) + e];
    G[(e >> 7) * SLD + (e & (JK - 1))] = s;
    R[(e >> 7) * SLD + (e & (JK - 1))] = ((e >> 7) == (e & (JK - 1))) ? 1.f : 0.f;
  }
  __syncthreads();
  // convergence measure of this pair at visit time: max |cos| between any two of its 128 vectors
  float mx = 0.f;
  int bad = 0;
  for (int e = tid; e < JK * JK; e += SOLVE_THREADS) {
    int r = e >> 7, c = e & (JK - 1);
    float g = G[r * SLD + c];
    if (!(fabsf(g) <= FLT_MAX)) bad = 1;
    if (r != c) {
      float d = G[r * SLD + r] * G[c * SLD + c];
      if (d > 0.f) mx = fmaxf(mx, fabsf(g) * rsqrtf(d));
    }
  }
  mx = warp_max(mx);
  bad = __any_sync(0xffffffffu, bad);
  if ((tid & 31) == 0) { red[tid >> 5] = mx; red[32 + (tid >> 5)] = bad ? 1.f : 0.f; }
  __syncthreads();
  if (tid < 32) {
    float v = warp_max(red[tid]);
    float bb = warp_max(red[32 + tid]);
    if (tid == 0) { red[0] = v; red[32] = bb; }
  }
  __syncthreads();
  mx = red[0];
  bad = red[32] > 0.f;
  if (bad) {
    if (tid == 0) { atomicOr(&status[b], 1); pairflag[idx] = 0; }
    return;
  }
  if (tid == 0) atomicMax(&maxoff_bits[b], __float_as_uint(mx));
  if (mx < tol) {
    if (tid == 0) pairflag[idx] = 0;
    return;
  }
  if (tid == 0) pairflag[idx] = 1;

  const int nsweeps = (mx > 1e-3f) ? 2 : 1;
  for (int sw = 0; sw < nsweeps; ++sw) {
    for (int st = 0; st < JK - 1; ++st) {
      if (tid < JK / 2) {
        int pp, qq;
        rr_pair(st, tid, pp, qq);
        float app = G[pp * SLD + pp], aqq = G[qq * SLD + qq], apq = G[pp * SLD + qq];
        float c = 1.f, s = 0.f, tt = 0.f;
        if (fabsf(apq) > 1e-8f * sqrtf(fmaxf(app, 0.f) * fmaxf(aqq, 0.f)) && apq != 0.f) {
          float tau = (aqq - app) / (2.f * apq);
          tt = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(1.f + tau * tau));
          c = rsqrtf(1.f + tt * tt);
          s = tt * c;
        }
        // keep the larger norm at the lower index (de Rijk): extra quarter turn when out of order
        if (app - tt * apq < aqq + tt * apq) { float c2 = s, s2 = -c; c = c2; s = s2; }
        cs[tid] = make_float2(c, s);
      }
      __syncthreads();
      // G <- J^T G J, one 2x2 block per (row pair, column pair)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int si = (tid >> 6) + 16 * it, ti = tid & 63;
        int ps, qs, pt, qt;
        rr_pair(st, si, ps, qs);
        rr_pair(st, ti, pt, qt);
        const float2 a = cs[si], bq = cs[ti];
        float g00 = G[ps * SLD + pt], g01 = G[ps * SLD + qt], g10 = G[qs * SLD + pt], g11 = G[qs * SLD + qt];
        float r00 = a.x * g00 - a.y * g10, r01 = a.x * g01 - a.y * g11;
        float r10 = a.y * g00 + a.x * g10, r11 = a.y * g01 + a.x * g11;
        float o00 = bq.x * r00 - bq.y * r01, o01 = bq.y * r00 + bq.x * r01;
        float o10 = bq.x * r10 - bq.y * r11, o11 = bq.y * r10 + bq.x * r11;
        if (si == ti) { o01 = 0.f; o10 = 0.f; }
        G[ps * SLD + pt] = o00; G[ps * SLD + qt] = o01; G[qs * SLD + pt] = o10; G[qs * SLD + qt] = o11;
      }
      // R <- R J
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int i = (tid >> 6) + 16 * it, ti = tid & 63;
        int pt, qt;
        rr_pair(st, ti, pt, qt);
        const float2 bq = cs[ti];
        float r0 = R[i * SLD + pt], r1 = R[i * SLD + qt];
        R[i * SLD + pt] = bq.x * r0 - bq.y * r1;
        R[i * SLD + qt] = bq.y * r0 + bq.x * r1;
      }
      __syncthreads();
    }
  }
  // Newton-Schulz polish: R <- R (1.5 I - 0.5 R^T R) restores the orthogonality lost to fp32 rounding in the
  // ~250 accumulated rotations per column (measured 1e-5 -> 4e-7), which otherwise drifts the singular values.
  const int ta = tid >> 5, tb = tid & 31;
  {
    float e[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) e[i][j] = 0.f;
    for (int l = 0; l < JK; ++l) {
      float4 a = *reinterpret_cast<const float4*>(&R[l * SLD + ta * 4]);
      float4 bb = *reinterpret_cast<const float4*>(&R[l * SLD + tb * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) e[i][j] = fmaf(av[i], bv[j], e[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&G[(ta * 4 + i) * SLD + tb * 4]) = make_float4(e[i][0], e[i][1], e[i][2], e[i][3]);
  }
  __syncthreads();
  {
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    for (int l = 0; l < JK; ++l) {
      float4 bb = *reinterpret_cast<const float4*>(&G[l * SLD + tb * 4]);
      float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = R[(ta * 4 + i) * SLD + l];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(a, bv[j], o[i][j]);
      }
    }
    float* Ro = Rout + (int64_t)idx * (JK * JK);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 r = *reinterpret_cast<const float4*>(&R[(ta * 4 + i) * SLD + tb * 4]);
      *reinterpret_cast<float4*>(&Ro[(ta * 4 + i) * JK + tb * 4]) =
          make_float4(1.5f * r.x - 0.5f * o[i][0], 1.5f * r.y - 0.5f * o[i][1], 1.5f * r.z - 0.5f * o[i][2],
                      1.5f * r.w - 0.5f * o[i][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ update
// panel <- R^T panel, i.e. out[j][c] = sum_i R[i][j] * X[i][c], in place, UPD_TILES column tiles of 128 per CTA.
constexpr int UPD_TILES = 4;
constexpr size_t UPDATE_SMEM = sizeof(float) * (JK * JK + 2 * JK * 128);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(256, 1)
update_kernel(float* __restrict__ X, int64_t mat_stride, int ldx, const int2* __restrict__ pairs, int len_pad,
              int pairs_per_mat, const float* __restrict__ Rin, const int* __restrict__ pairflag,
              const int* __restrict__ done) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Rs = reinterpret_cast<float*>(smem_raw);          // [128][128]  Rs[i][j]
  float* Xs = Rs + JK * JK;                                // [2][128][128]
  const int b = blockIdx.z, p = blockIdx.y;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  if (!pairflag[idx]) return;
  const int2 pr = pairs[p];
  float* Xb = X + b * mat_stride;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int tile0 = blockIdx.x * UPD_TILES;
  const int ntiles = min(UPD_TILES, len_pad / 128 - tile0);

  auto issue_tile = [&](int tile, int buf) {
    const int c0 = (tile0 + tile) * 128;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int e = t + 256 * i;
      int row = e >> 5, c4 = (e & 31) * 4;
      int vec = (row < JB) ? pr.x * JB + row : pr.y * JB + (row - JB);
      cp_async16(&Xs[(buf * JK + row) * 128 + c4], Xb + (int64_t)vec * ldx + c0 + c4);
    }
  };
  {
    const float* Rg = Rin + (int64_t)idx * (JK * JK);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int e = t + 256 * i;
      cp_async16(&Rs[e * 4], Rg + e * 4);
    }
  }
  issue_tile(0, 0);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) issue_tile(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* xs = Xs + buf * JK * 128;
#pragma unroll 4
    for (int i = 0; i < JK; ++i) {
      float4 a0 = *reinterpret_cast<const float4*>(&Rs[i * JK + ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&Rs[i * JK + 64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&xs[i * 128 + tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&xs[i * 128 + 64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
    }
    const int c0 = (tile0 + tile) * 128;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int j = (r < 4) ? ty * 4 + r : 64 + ty * 4 + (r - 4);
      int vec = (j < JB) ? pr.x * JB + j : pr.y * JB + (j - JB);
      float* dst = Xb + (int64_t)vec * ldx + c0;
      *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// per-row 2-norm; optionally normalises the row in place.  One CTA per row.
__global__ void __launch_bounds__(256) rownorm_kernel(float* __restrict__ X, int64_t mat_stride, int ld, int len,
                                                      int normalise, float* __restrict__ norm_out, int nv_pad,
                                                      int* __restrict__ status) {
  __shared__ float red[8];
  const int b = blockIdx.y, j = blockIdx.x;
  float* row = X + b * mat_stride + (int64_t)j * ld;
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += 256) { float v = row[i]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float nrm = sqrtf(red[0]);
  if (threadIdx.x == 0) {
    norm_out[(int64_t)b * nv_pad + j] = nrm;
    if (!(nrm <= FLT_MAX)) atomicOr(&status[b], 1);
  }
  if (normalise && nrm > 0.f) {
    const float inv = 1.f / nrm;
    for (int i = threadIdx.x; i < len; i += 256) row[i] *= inv;
  }
}

// descending bitonic sort of (sigma, index) for one matrix per CTA.  All nv_pad rows take part: the norm-sorting
// swaps inside the solver may leave a real vector in a padding slot, and the zero padding vectors sort last.
__global__ void __launch_bounds__(1024) sort_kernel(const float* __restrict__ sig_in, float* __restrict__ sig_out,
                                                    int* __restrict__ perm, int nv, int nv_pad, int P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);
  int* val = reinterpret_cast<int*>(key + P);
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    key[i] = (i < nv_pad) ? sig_in[(int64_t)b * nv_pad + i] : -1.f;   // zero padding vectors sort last (sigma 0)
    val[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int l = i ^ j;
        if (l > i) {
          bool desc = ((i & k) == 0);
          float a = key[i], c = key[l];
          // NaN keys (flagged separately through status) are treated as smallest
          bool a_lt_c = (a < c) || (a != a && c == c);
          bool swap = desc ? a_lt_c : ((c < a) || (c != c && a == a));
          if (swap) { key[i] = c; key[l] = a; int tv = val[i]; val[i] = val[l]; val[l] = tv; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < nv_pad; i += blockDim.x) {
    sig_out[(int64_t)b * nv_pad + i] = key[i];
    perm[(int64_t)b * nv_pad + i] = val[i];
  }
}

// ------------------------------------------------------------------------------------------------ extract (a5, a6)
__device__ __forceinline__ float sig_pow(float s, float e) {
  if (!(s > 0.f)) return 0.f;
  if (e == 0.f) return 1.f;
  if (e == 1.f) return s;
  if (e == 0.5f) return sqrtf(s);
  if (e == -0.5f) return rsqrtf(s);
  if (e == -1.f) return 1.f / s;
  return powf(s, e);
}

// out[i][j] = src[perm[j]][i] * sigma[j]^e   (i < rows, j < r): gathers factor columns from vector rows
template <typename TC>
__global__ void __launch_bounds__(256) extract_cols_kernel(const float* __restrict__ src, int ld, const int* __restrict__ perm,
                                                           const float* __restrict__ sigma, float e, int rows, int r,
                                                           TC* __restrict__ out, int64_t ldo) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int q = ty; q < 32; q += 8) {
    int j = j0 + q, i = i0 + tx;
    tile[q][tx] = (j < r && i < rows) ? src[(int64_t)perm[j] * ld + i] * sig_pow(sigma[j], e) : 0.f;
  }
  __syncthreads();
  for (int q = ty; q < 32; q += 8) {
    int i = i0 + q, j = j0 + tx;
    if (i < rows && j < r) out[(int64_t)i * ldo + j] = from_f32<TC>(tile[tx][q]);
  }
}

// out[j][l] = src[perm[j]][l] * sigma[j]^e / scale[l]   (j < r, l < cols)
template <typename TC>
__global__ void __launch_bounds__(256) extract_rows_kernel(const float* __restrict__ src, int ld, const int* __restrict__ perm,
                                                           const float* __restrict__ sigma, float e,
                                                           const float* __restrict__ scale, int cols, int r,
                                                           TC* __restrict__ out, int64_t ldo) {
  const int j = blockIdx.y;
  const int l = blockIdx.x * 256 + threadIdx.x;
  if (j >= r || l >= cols) return;
  float v = src[(int64_t)perm[j] * ld + l] * sig_pow(sigma[j], e) / scale[l];
  out[(int64_t)j * ldo + l] = from_f32<TC>(v);
}

// ------------------------------------------------------------------------------------------------ host side
static void build_pair_table(const SvdPlan& p, std::vector<int2>& tab) {
  // circle method over nb blocks: nb-1 rounds of nb/2 disjoint pairs (I < J)
  tab.resize((size_t)p.rounds * p.pairs);
  std::vector<int> idx(p.nb);
  for (int i = 0; i < p.nb; ++i) idx[i] = i;
  for (int r = 0; r < p.rounds; ++r) {
    for (int i = 0; i < p.pairs; ++i) {
      int a = idx[i], b = idx[p.nb - 1 - i];
      tab[(size_t)r * p.pairs + i] = make_int2(a < b ? a : b, a < b ? b : a);
    }
    int last = idx[p.nb - 1];
    for (int i = p.nb - 1; i > 1; --i) idx[i] = idx[i - 1];
    if (p.nb > 1) idx[1] = last;
  }
}

template <typename T>
static int run_svd(const SvdPlan& p, int64_t ldw, unsigned char* ws, float tol, int max_sweeps, int* sweeps_out,
                   cudaStream_t st) {
  const void* const* d_W = reinterpret_cast<const void* const*>(ws + p.off_ptrs);
  const int2* d_pairs = reinterpret_cast<const int2*>(ws + p.off_pairs);
  float* X = reinterpret_cast<float*>(ws + p.off_X);
  float* Y = reinterpret_cast<float*>(ws + p.off_Y);
  float* G = reinterpret_cast<float*>(ws + p.off_G);
  float* R = reinterpret_cast<float*>(ws + p.off_R);
  int* flag = reinterpret_cast<int*>(ws + p.off_flag);
  unsigned* maxoff = reinterpret_cast<unsigned*>(ws + p.off_maxoff);
  int* done = reinterpret_cast<int*>(ws + p.off_done);
  float* sigma = reinterpret_cast<float*>(ws + p.off_sigma);
  int* perm = reinterpret_cast<int*>(ws + p.off_perm);
  int* status = reinterpret_cast<int*>(ws + p.off_status);
  float* scale = reinterpret_cast<float*>(ws + p.off_scale);
  float* norm = reinterpret_cast<float*>(ws + p.off_norm);
  const int64_t xs = (int64_t)p.nv_pad * p.len_pad;

  static bool attrs_set = false;
  if (!attrs_set) {
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPDATE_SMEM));
    ASVD_CUDA_CHECK(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    attrs_set = true;
  }

  ASVD_CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(float) * xs * p.batch, st));
  ASVD_CUDA_CHECK(cudaMemsetAsync(done, 0, sizeof(int) * p.batch, st));
  ASVD_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int) * p.batch, st));
  {
    dim3 grid((p.n + 31) / 32, (p.m + 31) / 32, p.batch);
    prep_kernel<T><<<grid, 256, 0, st>>>(d_W, scale, ldw, p.m, p.n, p.tall, X, xs, p.len_pad);
    ASVD_CUDA_CHECK(cudaGetLastError());
  }
  std::vector<unsigned> h_maxoff(p.batch);
  std::vector<int> h_done(p.batch, 0), h_sweeps(p.batch, 0);
  int sweep = 0;
  bool all_done = false;
  for (; sweep < max_sweeps && !all_done; ++sweep) {
    ASVD_CUDA_CHECK(cudaMemsetAsync(maxoff, 0, sizeof(unsigned) * p.batch, st));
    for (int r = 0; r < p.rounds; ++r) {
      const int2* pr = d_pairs + (size_t)r * p.pairs;
      gram_kernel<<<dim3(p.chunks, p.pairs, p.batch), 256, 0, st>>>(X, xs, p.len_pad, pr, p.len_pad, p.chunks, p.pairs, G, done);
      solve_kernel<<<dim3(p.pairs, p.batch), SOLVE_THREADS, SOLVE_SMEM, st>>>(G, p.chunks, p.pairs, R, flag, maxoff, status, done, tol);
      const int ctas_x = (p.len_pad / 128 + UPD_TILES - 1) / UPD_TILES;
      update_kernel<<<dim3(ctas_x, p.pairs, p.batch), 256, UPDATE_SMEM, st>>>(X, xs, p.len_pad, pr, p.len_pad, p.pairs, R, flag, done);
    }
    ASVD_CUDA_CHECK(cudaGetLastError());
    ASVD_CUDA_CHECK(cudaMemcpyAsync(h_maxoff.data(), maxoff, sizeof(unsigned) * p.batch, cudaMemcpyDeviceToHost, st));
    ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
    all_done = true;
    bool changed = false;
    for (int b = 0; b < p.batch; ++b) {
      if (h_done[b]) continue;
      float mo;
      memcpy(&mo, &h_maxoff[b], 4);
      h_sweeps[b] = sweep + 1;
      if (mo < tol) { h_done[b] = 1; changed = true; }
      else all_done = false;
    }
    if (changed && !all_done)
      ASVD_CUDA_CHECK(cudaMemcpyAsync(done, h_done.data(), sizeof(int) * p.batch, cudaMemcpyHostToDevice, st));
  }
  // rows of X are sigma_j u_j: normalise, recover the other factor from the original weight, anchor sigma to it
  rownorm_kernel<<<dim3(p.nv_pad, p.batch), 256, 0, st>>>(X, xs, p.len_pad, p.len_pad, 1, norm, p.nv_pad, status);
  ASVD_CUDA_CHECK(cudaGetLastError());
  {
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    gb.Bptrs = d_W;
    gb.strideA = xs;
    gb.strideC = (int64_t)p.nv_pad * p.ldy;
    ASVD_CUDA_CHECK(cudaMemsetAsync(Y, 0, sizeof(float) * gb.strideC * p.batch, st));
    // per-matrix scale vectors live contiguously in the workspace: pass through pointer-free strides by
    // launching one GEMM per matrix when scales differ (batch is small); z-batched otherwise.
    for (int b = 0; b < p.batch; ++b) {
      GemmBatch g1 = gb;
      g1.Bptrs = d_W + b;
      const float* sb = scale + (int64_t)b * p.n;
      cudaError_t e;
      if (p.tall)   // Y[j][l] = sum_i Xhat[j][i] W[i][l] * s[l]
        e = launch_gemm128<float, T, float, true>(X + b * xs, p.len_pad, (const T*)nullptr, ldw, Y + b * gb.strideC, p.ldy,
                                                  p.nv_pad, p.n, p.m, nullptr, sb, nullptr, 1, g1, st);
      else          // Y[j][i] = sum_l Xhat[j][l] s[l] W[i][l]
        e = launch_gemm128<float, T, float, false>(X + b * xs, p.len_pad, (const T*)nullptr, ldw, Y + b * gb.strideC, p.ldy,
                                                   p.nv_pad, p.m, p.n, sb, nullptr, nullptr, 1, g1, st);
      ASVD_CUDA_CHECK(e);
    }
  }
  rownorm_kernel<<<dim3(p.nv_pad, p.batch), 256, 0, st>>>(Y, (int64_t)p.nv_pad * p.ldy, p.ldy, p.nv, 0, norm, p.nv_pad, status);
  ASVD_CUDA_CHECK(cudaGetLastError());
  {
    int P = 1;
    while (P < p.nv_pad) P <<= 1;
    sort_kernel<<<p.batch, 1024, 8 * (size_t)P, st>>>(norm, sigma, perm, p.nv, p.nv_pad, P);
    ASVD_CUDA_CHECK(cudaGetLastError());
  }
  std::vector<int> h_status(p.batch);
  ASVD_CUDA_CHECK(cudaMemcpyAsync(h_status.data(), status, sizeof(int) * p.batch, cudaMemcpyDeviceToHost, st));
  ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
  if (sweeps_out) for (int b = 0; b < p.batch; ++b) sweeps_out[b] = h_sweeps[b];
  for (int b = 0; b < p.batch; ++b)
    if (h_status[b]) { set_error("non-finite values in weight %d of the batch (or its scale)", b); return ASVD_ERR_NONFINITE; }
  if (!all_done) { set_error("sweep limit %d reached above tolerance %g", max_sweeps, (double)tol); return ASVD_ERR_NOT_CONVERGED; }
  return ASVD_OK;
}

}  // namespace asvd

using namespace asvd;

template <typename TC>
static int do_extract(const SvdPlan& p, const unsigned char* ws, int b, int r, int fuse, TC* A, int64_t lda, TC* B,
                      int64_t ldb, cudaStream_t st) {
  const float* X = reinterpret_cast<const float*>(ws + p.off_X) + (int64_t)b * p.nv_pad * p.len_pad;
  const float* Y = reinterpret_cast<const float*>(ws + p.off_Y) + (int64_t)b * p.nv_pad * p.ldy;
  const float* sigma = reinterpret_cast<const float*>(ws + p.off_sigma) + (int64_t)b * p.nv_pad;
  const int* perm = reinterpret_cast<const int*>(ws + p.off_perm) + (int64_t)b * p.nv_pad;
  const float* scale = reinterpret_cast<const float*>(ws + p.off_scale) + (int64_t)b * p.n;
  // exponent of sigma carried by A: UV 1/2, U 1, V 0 (modules/svd_linear.py:16-24)
  const float a = fuse == ASVD_FUSE_UV ? 0.5f : (fuse == ASVD_FUSE_U ? 1.f : 0.f);
  dim3 gA((r + 31) / 32, (p.m + 31) / 32), gB((p.n + 255) / 256, r);
  if (p.tall) {
    // X rows are unit u_j (length m); Y rows are sigma_j v_j^T diag(s) (length n)
    extract_cols_kernel<TC><<<gA, 256, 0, st>>>(X, p.len_pad, perm, sigma, a, p.m, r, A, lda);
    extract_rows_kernel<TC><<<gB, 256, 0, st>>>(Y, p.ldy, perm, sigma, -a, scale, p.n, r, B, ldb);
  } else {
    // X rows are unit v_j^T diag(s)... scaled space (length n); Y rows are sigma_j u_j (length m)
    extract_cols_kernel<TC><<<gA, 256, 0, st>>>(Y, p.ldy, perm, sigma, a - 1.f, p.m, r, A, lda);
    extract_rows_kernel<TC><<<gB, 256, 0, st>>>(X, p.len_pad, perm, sigma, 1.f - a, scale, p.n, r, B, ldb);
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  return ASVD_OK;
}

extern "C" {

int asvd_version(void) { return ASVD_B200_VERSION; }
const char* asvd_last_error(void) { return asvd::last_error(); }

int asvd_rank_for_ratio(int64_t out_features, int64_t in_features, double param_ratio, int rank_align) {
  // modules/svd_linear.py:39-44: int(n_params * ratio) // (in + out), then ceil to a multiple of rank_align
  double prod = (double)(out_features * in_features) * param_ratio;
  int64_t compressed = (int64_t)prod;   // python int() truncates toward zero
  int64_t rank = compressed / (in_features + out_features);
  if (rank_align > 1) rank = (rank + rank_align - 1) / rank_align * rank_align;
  return (int)rank;
}

size_t asvd_svd_workspace_bytes(int m, int n, int batch) {
  if (m <= 0 || n <= 0 || batch <= 0) return 0;
  return make_plan(m, n, batch).bytes;
}

int asvd_scaled_svd(const void* const* W_host_ptrs, int w_dtype, int64_t ldw, int m, int n, int batch,
                    const float* const* scale_host_ptrs, void* workspace, size_t workspace_bytes, float tol,
                    int max_sweeps, int* sweeps_out_host, void* stream) {
  ASVD_REQUIRE(W_host_ptrs && workspace, "null pointer");
  ASVD_REQUIRE(m > 0 && n > 0 && batch > 0 && ldw >= n, "bad shape m=%d n=%d batch=%d ldw=%lld", m, n, batch, (long long)ldw);
  ASVD_REQUIRE(w_dtype == ASVD_F32 || w_dtype == ASVD_F16 || w_dtype == ASVD_BF16, "bad dtype %d", w_dtype);
  ASVD_REQUIRE((m < n ? m : n) <= 16384, "min(m,n) = %d > 16384 is not supported", m < n ? m : n);
  ASVD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  SvdPlan p = make_plan(m, n, batch);
  if (workspace_bytes < p.bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, p.bytes);
    return ASVD_ERR_WORKSPACE;
  }
  if (tol <= 0.f) tol = 4e-6f;
  if (max_sweeps <= 0) max_sweeps = 30;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);

  std::vector<const void*> ptrs(2 * (size_t)batch, nullptr);
  for (int b = 0; b < batch; ++b) {
    ASVD_REQUIRE(W_host_ptrs[b] != nullptr, "weight %d is null", b);
    ASVD_REQUIRE((reinterpret_cast<uintptr_t>(W_host_ptrs[b]) & 15) == 0, "weight %d is not 16-byte aligned", b);
    ptrs[b] = W_host_ptrs[b];
  }
  ASVD_CUDA_CHECK(cudaMemcpyAsync(ws + p.off_ptrs, ptrs.data(), sizeof(void*) * 2 * batch, cudaMemcpyHostToDevice, st));
  std::vector<int2> tab;
  build_pair_table(p, tab);
  if (!tab.empty())
    ASVD_CUDA_CHECK(cudaMemcpyAsync(ws + p.off_pairs, tab.data(), sizeof(int2) * tab.size(), cudaMemcpyHostToDevice, st));
  float* scale = reinterpret_cast<float*>(ws + p.off_scale);
  for (int b = 0; b < batch; ++b) {
    const float* s = scale_host_ptrs ? scale_host_ptrs[b] : nullptr;
    if (s) ASVD_CUDA_CHECK(cudaMemcpyAsync(scale + (int64_t)b * n, s, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    else fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(scale + (int64_t)b * n, 1.f, n);
  }
  ASVD_CUDA_CHECK(cudaGetLastError());
  // the pageable host buffers above must outlive the async copies
  ASVD_CUDA_CHECK(cudaStreamSynchronize(st));
  switch (w_dtype) {
    case ASVD_F32: return run_svd<float>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st);
    case ASVD_F16: return run_svd<__half>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st);
    default: return run_svd<__nv_bfloat16>(p, ldw, ws, tol, max_sweeps, sweeps_out_host, st);
  }
}

int asvd_svd_sigma(const void* workspace, int m, int n, int batch, int b, float* sigma_out, void* stream) {
  ASVD_REQUIRE(workspace && sigma_out && b >= 0 && b < batch, "bad argument");
  SvdPlan p = make_plan(m, n, batch);
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  const float* sigma = reinterpret_cast<const float*>(ws + p.off_sigma) + (int64_t)b * p.nv_pad;
  ASVD_CUDA_CHECK(cudaMemcpyAsync(sigma_out, sigma, sizeof(float) * p.nv, cudaMemcpyDeviceToDevice,
                                  reinterpret_cast<cudaStream_t>(stream)));
  return ASVD_OK;
}


int asvd_svd_extract(const void* workspace, int m, int n, int batch, int b, int r, int sigma_fuse, int out_dtype,
                     void* A_out, int64_t lda, void* B_out, int64_t ldb, void* stream) {
  ASVD_REQUIRE(workspace && A_out && B_out && b >= 0 && b < batch, "bad argument");
  ASVD_REQUIRE(r > 0 && r <= (m < n ? m : n), "rank %d out of range (min(m,n) = %d)", r, m < n ? m : n);
  ASVD_REQUIRE(lda >= r && ldb >= n, "bad leading dimensions");
  ASVD_REQUIRE(sigma_fuse >= 0 && sigma_fuse <= 2, "bad sigma_fuse %d", sigma_fuse);
  SvdPlan p = make_plan(m, n, batch);
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (out_dtype) {
    case ASVD_F32: return do_extract<float>(p, ws, b, r, sigma_fuse, (float*)A_out, lda, (float*)B_out, ldb, st);
    case ASVD_F16: return do_extract<__half>(p, ws, b, r, sigma_fuse, (__half*)A_out, lda, (__half*)B_out, ldb, st);
    case ASVD_BF16: return do_extract<__nv_bfloat16>(p, ws, b, r, sigma_fuse, (__nv_bfloat16*)A_out, lda, (__nv_bfloat16*)B_out, ldb, st);
  }
  set_error("bad dtype %d", out_dtype);
  return ASVD_ERR_INVALID;
}

}  // extern "C"
