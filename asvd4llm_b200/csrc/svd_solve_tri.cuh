// solve_tri_g_kernel: the G half of the inner solve on the UPPER TRIANGLE of the Gram matrix, with the rotation
// parameters on a warp of their own.  (included by svd_jacobi.cu after svd_solve_quad.cuh: same quad round-robin
// ordering, same schedule table, same rotation formula, same per-pair record for solve_quad_r_kernel.)
//
// Why.  solve_quad_kernel / solve_quad_g_kernel keep all 256 patches of the symmetric 128x128 matrix and let ONE of the
// eight G warps compute every step's rotations in between its own 128 FMAs and its own parameter loads: the kernel runs
// at the pace of that warp's dependent chain (~1800 clocks per step, profiles/r02_solve_timing.log) with a third of the
// issue slots used, and at 128 registers x 256-512 threads at most two pairs share an SM.  Two observations:
//   * within a round (4 steps) the pivots of group g lie in the DIAGONAL patch (g, g), and that patch only ever sees
//     group g's own rotations.  A lane that holds nothing but the diagonal patch can run the whole round's chain
//     (parameters -> publish -> rotate its own patch, no loads, no barrier waits) ahead of everybody else;
//   * G is symmetric: patch (c, a) is the transpose of (a, c).  120 threads hold one patch of every unordered pair
//     of groups; a patch gets row rotations of its row group and column rotations of its column group as before.
// So a CTA is 5 warps: warp 0 = the diagonal patches (parameter chain; lane g computes pivots 0-1 of group g, lane g + 16
// pivots 2-3, both keep bit-identical copies of the patch), warps 1-4 = 120 off-diagonal patches, held packed for
// fma.rn.f32x2 during the four steps of a round.  The two roles run their own copies of the sweep (same barriers in the
// same order), so neither carries the other's registers.  160 threads x 128 registers and 76 KB of shared memory:
// THREE pairs per SM, whose chains interleave.
// A quad move stages the full matrix: off-diagonal threads write their moving sub-blocks in both orientations (the
// transpose is a different register selection, not a shuffle), every thread reads what its slot receives.  The diagonal
// lanes keep the UPPER triangle of their patches only (symmetric cell update, 76 FMAs instead of 128) and stage the
// lower one from its mirror image; after a move an element of an off-diagonal patch may descend from either orientation
// of its source: the two agree to rounding, and the kernel is deterministic.
constexpr int TRI_THREADS = 160;
#ifdef ASVD_SOLVE_TIMING
// timing build (scripts/tri_timing.py): clock64 marks of lane 0 of CTA (0,0); [0..7] G kernel, [8..15] replay kernel
__device__ unsigned long long g_tri_timing[16];
#define TT_MARK(cond, slot, t_last) do { if (cond) { const long long _t = clock64(); g_tri_timing[slot] += (unsigned long long)(_t - (t_last)); (t_last) = _t; } } while (0)
#else
#define TT_MARK(cond, slot, t_last) do { } while (0)
#endif
constexpr size_t SOLVET_SMEM = SOLVEQG_SMEM + sizeof(float4) * QRING * 16;

// the transposed image of the sub-blocks that move: row 8pc + j of patch (pc, pa) is column j of patch (pa, pc)
__device__ __forceinline__ void quad_stage_write_t(float* st, const float (&g)[8][8], int pa, int pc, const bool (&mv)[4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (mv[j >> 2]) *reinterpret_cast<float4*>(quad_stage(st, 8 * pc + j, pa)) = make_float4(g[0][j], g[1][j], g[2][j], g[3][j]);
    if (mv[2 + (j >> 2)])
      *reinterpret_cast<float4*>(quad_stage(st, 8 * pc + j, 16 + pa)) = make_float4(g[4][j], g[5][j], g[6][j], g[7][j]);
  }
}

// Upper-triangle view of a diagonal patch: the diagonal lanes only ever read and update entries (i, j) with i <= j.
#define TRI_U(g, i, j) ((i) <= (j) ? (g)[(i)][(j)] : (g)[(j)][(i)])

// Two-sided rotation of a SYMMETRIC diagonal patch by its own four pivots, upper triangle only.  The pivots (p_k, q_k)
// split the patch into 2x2 cells {p_a, q_a} x {p_b, q_b}; a cell is closed under the rotation (rows by pivot a, columns by
// pivot b) and cell (b, a) is the transpose of (a, b): 4 diagonal cells x 7 + 6 cells x 8 = 76 FMAs instead of 128, in
// the operation order of quad_rows / quad_cols (the upper triangle comes out bitwise as the full update would).
template <int TYPE>
__device__ __forceinline__ void tri_diag_apply(float (&g)[8][8], const float2 (&q)[4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    constexpr int dummy = 0; (void)dummy;
    const int pa_ = qp_p(TYPE, a), qa_ = qp_q(TYPE, a);
    {
      const float x = q[a].x, y = q[a].y;
      const float e = g[pa_][pa_], f = g[pa_][qa_], h = g[qa_][qa_];
      const float r00 = fmaf(x, f, e), r01 = fmaf(x, h, f), r10 = fmaf(y, e, f), r11 = fmaf(y, f, h);
      g[pa_][pa_] = fmaf(x, r01, r00);
      g[pa_][qa_] = fmaf(y, r00, r01);
      g[qa_][qa_] = fmaf(y, r10, r11);
    }
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const int pb_ = qp_p(TYPE, b), qb_ = qp_q(TYPE, b);
      const float e = TRI_U(g, pa_, pb_), f = TRI_U(g, pa_, qb_), u = TRI_U(g, qa_, pb_), v = TRI_U(g, qa_, qb_);
      // rows (pivot a): (e, u) and (f, v) are the column-wise pairs
      const float e1 = fmaf(q[a].x, u, e), u1 = fmaf(q[a].y, e, u);
      const float f1 = fmaf(q[a].x, v, f), v1 = fmaf(q[a].y, f, v);
      // columns (pivot b): (e1, f1) and (u1, v1) are the row-wise pairs
      TRI_U(g, pa_, pb_) = fmaf(q[b].x, f1, e1);
      TRI_U(g, pa_, qb_) = fmaf(q[b].y, e1, f1);
      TRI_U(g, qa_, pb_) = fmaf(q[b].x, v1, u1);
      TRI_U(g, qa_, qb_) = fmaf(q[b].y, u1, v1);
    }
  }
}

// quad_stage_write for a diagonal lane: the lower triangle is read from its mirror image
__device__ __forceinline__ void quad_stage_write_diag(float* st, const float (&g)[8][8], int pa, const bool (&mv)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (mv[2 * (i >> 2)])
      *reinterpret_cast<float4*>(quad_stage(st, 8 * pa + i, pa)) = make_float4(TRI_U(g, i, 0), TRI_U(g, i, 1), TRI_U(g, i, 2), TRI_U(g, i, 3));
    if (mv[2 * (i >> 2) + 1])
      *reinterpret_cast<float4*>(quad_stage(st, 8 * pa + i, 16 + pa)) = make_float4(TRI_U(g, i, 4), TRI_U(g, i, 5), TRI_U(g, i, 6), TRI_U(g, i, 7));
  }
}

// Branch-free form of quad_rotation (same formula; the four chains of a lane interleave instead of running one
// divergent region after the other), with one reciprocal for both scale ratios.
__device__ __forceinline__ float3 tri_rotation(float gpp, float gqq, float gpq, float dp, float dq) {
  const float inv = rcp_ftz(dp * dq);
  const float rho = dq * dq * inv, rho_inv = dp * dp * inv;
  const float delta = fmaf(rho, gqq, -(gpp * rho_inv)), h = gpq + gpq;
  const bool ok = gpq * gpq > 1e-16f * (gpp * gqq) && gpq != 0.f;         // |cos| > 1e-8
  const float s = sqrt_ftz(fmaf(delta, delta, h * h));
  float t = copysignf(fabsf(h) * rcp_ftz(fabsf(delta) + s), delta * h);
  const bool rot = ok && fabsf(t) <= 1.f;      // |t| <= 1 by construction; anything else is underflow debris: no rotation
  t = rot ? t : 0.f;
  const float c = rot ? rsqrt_ftz(fmaf(t, t, 1.f)) : 1.f;
  return make_float3(-t * rho, t * rho_inv, c);
}

// First three steps (pivots inside the quads): every lane of the parameter warp computes the four rotations of its
// group; lanes 0-15 publish.
template <int TYPE>
__device__ __forceinline__ void tri_param_step_full(float (&g)[8][8], float (&d)[8], bool pub, float2* cs_step,
                                                    float2* __restrict__ hist_step, int grp, int bar_id) {
  float3 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pl = qp_p(TYPE, k), rl = qp_q(TYPE, k);
    o[k] = tri_rotation(g[pl][pl], g[rl][rl], g[pl][rl], d[pl], d[rl]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { d[qp_p(TYPE, k)] *= o[k].z; d[qp_q(TYPE, k)] *= o[k].z; }
  if (pub) {
    const float4 a = make_float4(o[0].x, o[0].y, o[1].x, o[1].y), b = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
    *q_plane(cs_step, grp, 0) = a;
    *q_plane(cs_step, grp, 1) = b;
    *q_plane(hist_step, grp, 0) = a;
    *q_plane(hist_step, grp, 1) = b;
  }
  __syncwarp();
  asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "n"(TRI_THREADS) : "memory");
  const float2 q[4] = {make_float2(o[0].x, o[0].y), make_float2(o[1].x, o[1].y), make_float2(o[2].x, o[2].y),
                       make_float2(o[3].x, o[3].y)};
  tri_diag_apply<TYPE>(g, q);
}

// A step of the rounds (pivot k = L_k with H_(k+j)%4).  Both half-warps hold the diagonal patches; lane g computes pivots
// 0-1 of group g, lane g + 16 pivots 2-3 -- half the arithmetic and half the MUFU operations on the dependent chain --,
// each publishes its plane of the step's record (and its two c factors), and after a __syncwarp both read the whole
// record back and rotate their copies with identical operations: the copies stay bitwise equal.
template <int TYPE, bool SYNC = false>
__device__ __forceinline__ void tri_param_step(float (&g)[8][8], float (&d)[8], int half, float2* cs_step, float4* c_step,
                                               float2* __restrict__ hist_step, int grp, int bar_id) {
  constexpr int P0 = qp_p(TYPE, 0), Q0 = qp_q(TYPE, 0), P1 = qp_p(TYPE, 1), Q1 = qp_q(TYPE, 1);
  constexpr int P2 = qp_p(TYPE, 2), Q2 = qp_q(TYPE, 2), P3 = qp_p(TYPE, 3), Q3 = qp_q(TYPE, 3);
  const bool hi = half != 0;
  const float3 oa = tri_rotation(hi ? g[P2][P2] : g[P0][P0], hi ? g[Q2][Q2] : g[Q0][Q0], hi ? g[P2][Q2] : g[P0][Q0],
                                 hi ? d[P2] : d[P0], hi ? d[Q2] : d[Q0]);
  const float3 ob = tri_rotation(hi ? g[P3][P3] : g[P1][P1], hi ? g[Q3][Q3] : g[Q1][Q1], hi ? g[P3][Q3] : g[P1][Q1],
                                 hi ? d[P3] : d[P1], hi ? d[Q3] : d[Q1]);
  const float4 mine = make_float4(oa.x, oa.y, ob.x, ob.y);
  *q_plane(cs_step, grp, half) = mine;
  *q_plane(hist_step, grp, half) = mine;
  reinterpret_cast<float2*>(c_step + grp)[half] = make_float2(oa.z, ob.z);
  __syncwarp();
  // SYNC (first step of a round): wait for the off-diagonal warps, which arrive once they have read their share of the
  // staging area -- the only thing that keeps this warp from overwriting it a round early (they are there long before)
  if (SYNC) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TRI_THREADS) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "n"(TRI_THREADS) : "memory");
  float2 q[4];
  load_q4(cs_step, grp, q);
  const float4 c4 = c_step[grp];
  d[P0] *= c4.x; d[Q0] *= c4.x; d[P1] *= c4.y; d[Q1] *= c4.y;
  d[P2] *= c4.z; d[Q2] *= c4.z; d[P3] *= c4.w; d[Q3] *= c4.w;
  tri_diag_apply<TYPE>(g, q);
}

// Packed FP32 pairs (fma.rn.f32x2): measured on this part (scripts/probes/ffma2_replay_probe.cu) the replay's inner loop
// takes 325 clocks per step with scalar FFMA and 163 with FFMA2 -- the packed instruction issues at the scalar rate.
// The patch is held TRANSPOSED, rp[j][k] = (R[2k][j], R[2k+1][j]): a column rotation touches whole columns, so both
// halves of a pair see the same multiplier.  Each half is the fma.rn the scalar code performs: results are bitwise equal.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// An off-diagonal patch in packed form for the four steps of a round: gp[i][c] = (g[i][2c], g[i][2c+1]).  Row rotations
// pair along the columns (one multiplier for both halves); column rotations take the pivots two at a time -- columns
// (0,1) with their partners, columns (2,3) with theirs: with the XOR schedule the partners are the pair (4,5) or (6,7), in
// order for even steps and swapped for odd ones (the swap is an operand modifier of FFMA2, not an instruction).  Each
// half is the fma.rn of quad_rows / quad_cols: same bits, 64 packed instructions instead of 128.
__device__ __forceinline__ f32x2 swp2(f32x2 v) { float a, b; upk2(v, a, b); return pk2(b, a); }

template <int TYPE>
__device__ __forceinline__ void tri_bulk_step_packed(f32x2 (&gp)[8][4], const float2* cs_step, int pa, int pc, int bar_id) {
  static_assert(TYPE >= 3, "packed form: steps of the rounds only");
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TRI_THREADS) : "memory");
  float2 qr[4], qc[4];
  load_q4(cs_step, pa, qr);
  load_q4(cs_step, pc, qc);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = qp_p(TYPE, k), r = qp_q(TYPE, k);
    const f32x2 X = pk2(qr[k].x, qr[k].x), Y = pk2(qr[k].y, qr[k].y);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const f32x2 a = gp[p][c], b = gp[r][c];
      gp[p][c] = ffma2(X, b, a);
      gp[r][c] = ffma2(Y, a, b);
    }
  }
  constexpr int M = TYPE - 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {                    // pivots (2h, 2h+1): column pair h with partner pair cH
    constexpr bool SW = (M & 1) != 0;
    const int cH = 2 + ((h ^ (M >> 1)) & 1);
    f32x2 X = pk2(qc[2 * h].x, qc[2 * h + 1].x), Y = pk2(qc[2 * h].y, qc[2 * h + 1].y);
    if (SW) Y = swp2(Y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const f32x2 u = gp[i][h], v = gp[i][cH];
      if (!SW) {
        gp[i][h] = ffma2(X, v, u);
        gp[i][cH] = ffma2(Y, u, v);
      } else {
        gp[i][h] = ffma2(X, swp2(v), u);
        gp[i][cH] = ffma2(Y, swp2(u), v);
      }
    }
  }
}

template <int TYPE>
__device__ __forceinline__ void tri_bulk_step(float (&g)[8][8], const float2* cs_step, int pa, int pc, int bar_id) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TRI_THREADS) : "memory");
  float2 qr[4], qc[4];
  load_q4(cs_step, pa, qr);
  load_q4(cs_step, pc, qc);
  quad_rows<TYPE>(g, qr);
  quad_cols<TYPE>(g, qc);
}

__global__ void __launch_bounds__(TRI_THREADS, 3)
solve_tri_g_kernel(const float* __restrict__ Gpart, int chunks, int pairs_per_mat, float* __restrict__ aux,
                   int* __restrict__ pairflag, unsigned* __restrict__ maxoff_bits, int* __restrict__ status,
                   const int* __restrict__ done, float tol, const int2* __restrict__ pairs, int* __restrict__ track, int nb,
                   int round_stamp, const int* __restrict__ precise_b, int half_gram_tc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* G = reinterpret_cast<float*>(smem_raw);            // [JK][SLD] summed Gram, then the staging area of the moves
  float2* csh = reinterpret_cast<float2*>(G + JK * SLD);    // [QRING][64] rotations of the last steps
  float* red = reinterpret_cast<float*>(csh + QRING * 64);  // [64]
  float* gd = red + 64;                                     // [JK] final diagonal (true norms)
  float* dmov = gd + JK;                                    // [JK] scales in transit during a quad move
  float* dfold = dmov + JK;                                 // [JK] scales being folded into G
  float4* cring = reinterpret_cast<float4*>(dfold + JK);    // [QRING][16] the c factors of the last steps
  unsigned char* qsrc = reinterpret_cast<unsigned char*>(cring + QRING * 16);

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  const int tid = threadIdx.x;
  const int2 pr = pairs[p];
  int* trk = track + (int64_t)b * (nb + nb * nb);
  if (pair_is_clean(track, nb, b, pr.x, pr.y)) {
    if (tid == 0) pairflag[idx] = 0;
    return;
  }
#ifdef ASVD_SOLVE_TIMING
  const bool tt = tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
  long long tt_last = clock64();
  if (tt) for (int i = 0; i < 8; ++i) g_tri_timing[i] = 0;
#endif
  for (int i = tid; i < (QROUNDS - 1) * 8; i += TRI_THREADS) reinterpret_cast<unsigned int*>(qsrc)[i] = g_quad_src_words[i];
  const int precise = precise_b[b], half_gram = half_gram_tc && precise;
  // One chunk (every batch with at least as many pairs as SMs): nothing to sum -- the 128 rows of the Gram matrix come in
  // by 128 bulk copies of 512 bytes straight into the padded rows of G, one mbarrier.  (The generic path keeps four
  // 16-byte loads per thread in flight: 12 % of this kernel's clocks at 27 x 4096^2, profiles/r02_tri_timing.log.)
  __shared__ __align__(8) uint64_t pro_bar;
  const bool preloaded = chunks == 1;
  if (preloaded) {
    if (tid == 0) { tc::mbar_init(&pro_bar, 1); tc::fence_barrier_init(); }
    __syncthreads();
    if (tid < JK) {
      const float* src = Gpart + (int64_t)idx * (JK * JK) + tid * JK;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       tc::smem_u32(G + tid * SLD)),
                   "l"(src), "r"((uint32_t)(JK * sizeof(float))), "r"(tc::smem_u32(&pro_bar))
                   : "memory");
    }
    if (tid == 0) tc::mbar_arrive_expect_tx(&pro_bar, (uint32_t)(JK * JK * sizeof(float)));
    tc::mbar_wait(&pro_bar, 0);
  }
  if (!solve_prologue<TRI_THREADS>(Gpart, chunks, idx, b, pr, tid, G, red, pairflag, maxoff_bits, status, tol, trk, nb,
                                   round_stamp, precise, gridDim.y, half_gram, preloaded))
    return;

  float* ax = aux + (int64_t)idx * QAUX_FLOATS;
  float2* hist = reinterpret_cast<float2*>(ax);             // [QSTEPS][64]
  float* a_dhist = ax + (size_t)QSTEPS * 128;               // [QFOLDS][JK]
  float* a_dfin = a_dhist + QFOLDS * JK;                    // [JK]
  int* a_dest = reinterpret_cast<int*>(a_dfin + JK);        // [JK]

  // Roles.  Warp 0: lanes g and g + 16 hold the diagonal patch of group g (each computes two of its four pivots; lanes
  // 0-15 do the staging and record writes).  Warps 1-4: thread lt = tid - 32 holds patch (a, a + delta) of the unordered pair {a, a + delta},
  // delta = 1..7 for every a (112 threads), delta = 8 for a < 8 (8 threads); the last 8 threads shadow the delta = 8 patches (same work, no writes).
  // Eight consecutive lanes differ in a AND in c modulo 8: staging accesses in either orientation hit distinct banks.
  const bool is_param = tid < 32;
  int pa, pc;
  bool writes;
  if (is_param) {
    pa = pc = tid & 15;
    writes = tid < 16;
  } else {
    const int lt = tid - 32;
    if (lt < 112) { pa = lt & 15; pc = (pa + 1 + (lt >> 4)) & 15; }
    else          { pa = lt & 7;  pc = pa + 8; }
    writes = lt < 120;
  }
  auto bar_all = [] { asm volatile("bar.sync 2, %0;" ::"n"(TRI_THREADS) : "memory"); };
  // The two roles run their own copies of the sweep (same barriers in the same order): the parameter warp keeps a plain
  // 8x8 patch and the scales, the off-diagonal warps keep their patch PACKED for fma.rn.f32x2 -- in one loop the
  // register allocator would have to hold both forms for every thread.
  if (is_param) {
    float g[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x0 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc]);
      const float4 x1 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc + 4]);
      g[i][0] = x0.x; g[i][1] = x0.y; g[i][2] = x0.z; g[i][3] = x0.w;
      g[i][4] = x1.x; g[i][5] = x1.y; g[i][6] = x1.z; g[i][7] = x1.w;
    }
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 1.f;
    bar_all();                                              // every patch is in registers: G becomes the staging area
    TT_MARK(tt, 0, tt_last);                                // prologue
    tri_param_step_full<0>(g, d, writes, csh + 0 * 64, hist + 0 * 64, pa, 8);
    tri_param_step_full<1>(g, d, writes, csh + 1 * 64, hist + 1 * 64, pa, 9);
    tri_param_step_full<2>(g, d, writes, csh + 2 * 64, hist + 2 * 64, pa, 10);
    TT_MARK(tt, 1, tt_last);                                // first three steps
#define TRI_PSTEP(TYPE, S, BAR) \
  tri_param_step<TYPE, (TYPE) == 3>(g, d, tid >> 4, csh + ((S) & (QRING - 1)) * 64, cring + ((S) & (QRING - 1)) * 16, hist + (S) * 64, pa, (BAR))
#pragma unroll 1
    for (int r = 0; r < QROUNDS; ++r) {
      const int s0 = 3 + 4 * r;
      // what this round's move will do to this patch: looked up before the steps, off the critical path of the round's end
      const unsigned char* qs = qsrc + (r < QROUNDS - 1 ? r : 0) * 32;
      const int rsrcL = qs[2 * pa], rsrcH = qs[2 * pa + 1];
      const bool mL = rsrcL != 8 * pa, mH = rsrcH != 8 * pa + 4;
      const bool mv[4] = {mL, mL || mH, mH || mL, mH};
      TRI_PSTEP(3, s0 + 0, 4);
      TRI_PSTEP(4, s0 + 1, 5);
      TRI_PSTEP(5, s0 + 2, 6);
      TRI_PSTEP(6, s0 + 3, 7);
      TT_MARK(tt, 2, tt_last);                              // the parameter warp's four steps
      if (r == QROUNDS - 1) break;
      if ((r & 7) == 7) {
        // fold the deferred scales back into the stored values; the replay kernel folds the same values into R
        if (writes) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { dfold[8 * pa + i] = d[i]; a_dhist[(r >> 3) * JK + 8 * pa + i] = d[i]; }
        }
        bar_all();
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int jj = i; jj < 8; ++jj) g[i][jj] *= d[i] * d[jj];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 1.f;
      }
      TT_MARK(tt, 3, tt_last);                              // fold
      // ---- quad move through the staging area: only what changes place
      if (writes) {
        quad_stage_write_diag(G, g, pa, mv);
#pragma unroll
        for (int i = 0; i < 8; ++i) dmov[8 * pa + i] = d[i];
      }
      TT_MARK(tt, 4, tt_last);                              // its staging writes
      bar_all();
      TT_MARK(tt, 5, tt_last);                              // waiting for the off-diagonal warps at the move barrier
      // The parameter warp reads first: its new diagonal patches head the next round's chain, the off-diagonal warps have
      // a whole step of slack before they need theirs (they follow on barrier 3 instead of crowding the shared-memory pipe)
      quad_stage_read(G, g, rsrcL, rsrcH, rsrcL, rsrcH, mv);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = dmov[(i < 4 ? rsrcL : rsrcH) + (i & 3)];
      asm volatile("" ::"f"(g[0][0]), "f"(g[0][7]), "f"(g[7][7]), "f"(d[0]), "f"(d[7]) : "memory");   // the loads have landed
      __syncwarp();
      asm volatile("bar.arrive 3, %0;" ::"n"(TRI_THREADS) : "memory");
      TT_MARK(tt, 6, tt_last);                              // its staging reads
      // no second CTA barrier: the next write of the staging area (and of dmov) lies behind the next round's first step
      // barrier, at which this warp WAITS and every off-diagonal thread arrives after its reads
    }
#undef TRI_PSTEP
    if (writes) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { gd[8 * pa + i] = d[i] * d[i] * g[i][i]; a_dfin[8 * pa + i] = d[i]; }
    }
  } else {
    f32x2 gp[8][4];                                         // gp[i][c] = (G[8pa + i][8pc + 2c], G[8pa + i][8pc + 2c + 1])
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x0 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc]);
      const float4 x1 = *reinterpret_cast<const float4*>(&G[(8 * pa + i) * SLD + 8 * pc + 4]);
      gp[i][0] = pk2(x0.x, x0.y); gp[i][1] = pk2(x0.z, x0.w); gp[i][2] = pk2(x1.x, x1.y); gp[i][3] = pk2(x1.z, x1.w);
    }
    bar_all();                                              // every patch is in registers: G becomes the staging area
    {
      float g[8][8];                                        // the three steps inside the quads pair differently: plain form
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) upk2(gp[i][c], g[i][2 * c], g[i][2 * c + 1]);
      tri_bulk_step<0>(g, csh + 0 * 64, pa, pc, 8);
      tri_bulk_step<1>(g, csh + 1 * 64, pa, pc, 9);
      tri_bulk_step<2>(g, csh + 2 * 64, pa, pc, 10);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) gp[i][c] = pk2(g[i][2 * c], g[i][2 * c + 1]);
    }
#define TRI_BSTEP(TYPE, S, BAR) tri_bulk_step_packed<TYPE>(gp, csh + ((S) & (QRING - 1)) * 64, pa, pc, (BAR))
#pragma unroll 1
    for (int r = 0; r < QROUNDS; ++r) {
      const int s0 = 3 + 4 * r;
      const unsigned char* qs = qsrc + (r < QROUNDS - 1 ? r : 0) * 32;
      const int rsrcL = qs[2 * pa], rsrcH = qs[2 * pa + 1], csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
      const bool mrL = rsrcL != 8 * pa, mrH = rsrcH != 8 * pa + 4, mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
      const bool mv[4] = {mrL || mcL, mrL || mcH, mrH || mcL, mrH || mcH};
      TRI_BSTEP(3, s0 + 0, 4);
      TRI_BSTEP(4, s0 + 1, 5);
      TRI_BSTEP(5, s0 + 2, 6);
      TRI_BSTEP(6, s0 + 3, 7);
      if (r == QROUNDS - 1) break;
      if ((r & 7) == 7) {
        bar_all();
        float dr[8], dc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dr[i] = dfold[8 * pa + i]; dc[i] = dfold[8 * pc + i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a, bb; upk2(gp[i][c], a, bb);
            gp[i][c] = pk2(a * (dr[i] * dc[2 * c]), bb * (dr[i] * dc[2 * c + 1]));
          }
        // (dfold is next written eight rounds later, behind many blocking barriers)
      }
      float g[8][8];                                        // plain view for the staging helpers (register pairs renamed)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) upk2(gp[i][c], g[i][2 * c], g[i][2 * c + 1]);
      if (writes) {
        quad_stage_write(G, g, pa, pc, mv);
        quad_stage_write_t(G, g, pa, pc, mv);
      }
      bar_all();
      asm volatile("bar.sync 3, %0;" ::"n"(TRI_THREADS) : "memory");   // after the parameter warp's reads
      quad_stage_read(G, g, rsrcL, rsrcH, csrcL, csrcH, mv);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) gp[i][c] = pk2(g[i][2 * c], g[i][2 * c + 1]);
    }
#undef TRI_BSTEP
  }
  bar_all();
  if (tid < JK) {
    const float dd = gd[tid];
    int rank = 0;
    for (int jj = 0; jj < JK; ++jj) {
      const float e = gd[jj];
      rank += (e > dd) || (e == dd && jj < tid);
    }
    a_dest[tid] = rank;
  }
  TT_MARK(tt, 7, tt_last);                                  // tail
}

// solve_tri_r_kernel: replay of a pair's rotation record on R, without a staging area and without barriers in the loop.
// solve_quad_r_kernel copies the whole 65 KB record into shared memory next to a 66 KB staging area (one CTA per SM)
// and moves the column quads through that area with two CTA-wide barriers per round: measured, the moves were 43 % of
// its samples (MIO throttle + barrier waits) and the kernel ran at 1.1-1.5 instructions per clock.  Here
//   * rows never move in R, so a move only exchanges column quads between the 16 patches of one row group: those sit in
//     the 16 lanes of a half-warp and exchange by shuffles (32 per round when only the H quads shift, 64 in the four
//     rounds that regroup the halves);
//   * the record arrives by eight bulk copies (cp.async.bulk, 16 steps each, one mbarrier each) issued up front into the
//     buffer that later holds the sorted R -- the first steps start after one L2 round trip, the rest streams in behind
//     (reading the record from global memory one step ahead left every warp waiting ~700 clocks per step: an L1 miss
//     per step and CTA);
// so the warps never wait for each other until the final column sort.  Two CTAs per SM.
constexpr int RCHUNK = 16;                                  // steps per bulk copy
constexpr int RCHUNKS = (QSTEPS + RCHUNK - 1) / RCHUNK;
constexpr size_t SOLVETR_SMEM = sizeof(float) * (JK * SLD) + sizeof(float) * JK * (QFOLDS + 1) + sizeof(int) * JK + (QROUNDS - 1) * 32;
static_assert(sizeof(float2) * QSTEPS * 64 <= sizeof(float) * JK * SLD, "the record must fit the sorted-R buffer");

template <int TYPE>
__device__ __forceinline__ void tri_r_step(f32x2 (&rp)[8][4], const float2* hist, int step, int pc, uint64_t* mb) {
  if ((step & (RCHUNK - 1)) == 0) tc::mbar_wait(&mb[step / RCHUNK], 0);
  float2 q[4];
  load_q4(hist + step * 64, pc, q);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = qp_p(TYPE, k), r = qp_q(TYPE, k);
    const f32x2 X = pk2(q[k].x, q[k].x), Y = pk2(q[k].y, q[k].y);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const f32x2 a = rp[p][h], b = rp[r][h];
      rp[p][h] = ffma2(X, b, a);
      rp[r][h] = ffma2(Y, a, b);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
solve_tri_r_kernel(const float* __restrict__ aux, int pairs_per_mat, float* __restrict__ Rout,
                   const int* __restrict__ pairflag, const int* __restrict__ done) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Rs = reinterpret_cast<float*>(smem_raw);           // the rotation record; then [JK][SLD] R with its columns sorted
  float* dhist = Rs + JK * SLD;                             // [QFOLDS][JK]
  float* dfin = dhist + QFOLDS * JK;                        // [JK]
  int* dest = reinterpret_cast<int*>(dfin + JK);            // [JK]
  unsigned char* qsrc = reinterpret_cast<unsigned char*>(dest + JK);
  __shared__ float cnp[4][JK];
  __shared__ __align__(8) uint64_t mb[RCHUNKS];

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  if (!pairflag[idx]) return;                               // clean, converged or non-finite pair: no rotation, no R
  const int tid = threadIdx.x, lane = tid & 31;
  const float* ax = aux + (int64_t)idx * QAUX_FLOATS;
#ifdef ASVD_SOLVE_TIMING
  const bool tt = tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
  long long tt_last = clock64();
  if (tt) for (int i = 8; i < 16; ++i) g_tri_timing[i] = 0;
#endif
  if (tid == 0) {
    for (int c = 0; c < RCHUNKS; ++c) tc::mbar_init(&mb[c], 1);
    tc::fence_barrier_init();
    for (int c = 0; c < RCHUNKS; ++c) {
      const int steps = (c == RCHUNKS - 1) ? QSTEPS - c * RCHUNK : RCHUNK;
      const uint32_t bytes = (uint32_t)(steps * 64 * sizeof(float2));
      tc::mbar_arrive_expect_tx(&mb[c], bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       tc::smem_u32(Rs + (size_t)c * RCHUNK * 128)),
                   "l"(ax + (size_t)c * RCHUNK * 128), "r"(bytes), "r"(tc::smem_u32(&mb[c]))
                   : "memory");
    }
  }
  {
    const float4* src = reinterpret_cast<const float4*>(ax + (size_t)QSTEPS * 128);
    float4* dst = reinterpret_cast<float4*>(dhist);         // dhist | dfin | dest are contiguous in the record and here
    for (int i = tid; i < (QFOLDS + 2) * JK / 4; i += 256) dst[i] = src[i];
    for (int i = tid; i < (QROUNDS - 1) * 8; i += 256) reinterpret_cast<unsigned int*>(qsrc)[i] = g_quad_src_words[i];
  }
  // patch (pa, pc): the 16 column groups of a row group in the 16 lanes of a half-warp
  const int pa = 2 * (tid >> 5) + (lane >> 4), pc = lane & 15;
  const float2* hist = reinterpret_cast<const float2*>(Rs);
  f32x2 r[8][4];                                            // r[j][h] = (R[8pa + 2h][8pc + j], R[8pa + 2h + 1][8pc + j])
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int h = 0; h < 4; ++h) r[j][h] = pk2((pa == pc && 2 * h == j) ? 1.f : 0.f, (pa == pc && 2 * h + 1 == j) ? 1.f : 0.f);
  __syncthreads();                                          // barriers initialised, scales and schedule in place
  TT_MARK(tt, 8, tt_last);                                  // prologue
  tri_r_step<0>(r, hist, 0, pc, mb);
  TT_MARK(tt, 9, tt_last);                                  // first step incl. waiting for the first chunk of the record
  tri_r_step<1>(r, hist, 1, pc, mb);
  tri_r_step<2>(r, hist, 2, pc, mb);
#pragma unroll 1
  for (int rd = 0; rd < QROUNDS; ++rd) {
    const int s0 = 3 + 4 * rd;
    tri_r_step<3>(r, hist, s0 + 0, pc, mb);
    tri_r_step<4>(r, hist, s0 + 1, pc, mb);
    tri_r_step<5>(r, hist, s0 + 2, pc, mb);
    tri_r_step<6>(r, hist, s0 + 3, pc, mb);
    TT_MARK(tt, 10, tt_last);                               // four steps
    if (rd == QROUNDS - 1) break;
    if ((rd & 7) == 7) {
      const float* dh = dhist + (rd >> 3) * JK;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dc = dh[8 * pc + j];
#pragma unroll
        for (int h = 0; h < 4; ++h) { float a, b; upk2(r[j][h], a, b); r[j][h] = pk2(a * dc, b * dc); }
      }
    }
    // ---- quad move by shuffles inside the half-warp
    const unsigned char* qs = qsrc + rd * 32;
    const int csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
    const bool mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
    if (!__any_sync(0xffffffffu, mcL)) {
      // only H quads move, and they come from H slots
      const int src = (lane & 16) | (csrcH >> 3);
#pragma unroll
      for (int j = 4; j < 8; ++j)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float a, b; upk2(r[j][h], a, b);
          r[j][h] = pk2(__shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src));
        }
    } else {
      // the halves regroup: an H slot takes the L quad of its partner group, whose L slot takes this H quad
      const int src = (lane & 16) | ((mcH ? csrcH : mcL ? csrcL : 8 * pc) >> 3);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float l0, l1, h0, h1;
          upk2(r[j][h], l0, l1); upk2(r[4 + j][h], h0, h1);
          const float a0 = __shfl_sync(0xffffffffu, l0, src), a1 = __shfl_sync(0xffffffffu, l1, src);
          const float c0 = __shfl_sync(0xffffffffu, h0, src), c1 = __shfl_sync(0xffffffffu, h1, src);
          if (mcH) r[4 + j][h] = pk2(a0, a1);
          if (mcL) r[j][h] = pk2(c0, c1);
        }
    }
    TT_MARK(tt, 11, tt_last);                               // fold + move
  }
  TT_MARK(tt, 12, tt_last);
  __syncthreads();                                          // everybody is done with the record: its buffer becomes R
  TT_MARK(tt, 13, tt_last);                                 // waiting for the other warps
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = dest[8 * pc + j];
    const float dc = dfin[8 * pc + j];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float a, b; upk2(r[j][h], a, b);
      Rs[(8 * pa + 2 * h) * SLD + col] = a * dc;
      Rs[(8 * pa + 2 * h + 1) * SLD + col] = b * dc;
    }
  }
  __syncthreads();
  // unit column norms (the default tail of solve_polish_write, same partial sums in the same order), then store
  {
    const int col = tid & (JK - 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int part = (tid >> 7) + 2 * h;
      float ss = 0.f;
#pragma unroll 8
      for (int l = 0; l < JK / 4; ++l) { const float x = Rs[(part * (JK / 4) + l) * SLD + col]; ss = fmaf(x, x, ss); }
      cnp[part][col] = ss;
    }
  }
  __syncthreads();
  if (tid < JK) cnp[0][tid] = rsqrtf(cnp[0][tid] + cnp[1][tid] + cnp[2][tid] + cnp[3][tid]);
  __syncthreads();
  float* Ro = Rout + (int64_t)idx * (JK * JK);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int e = (tid + 256 * k) * 4, rr = e >> 7, c = e & (JK - 1);
    const float4 x = *reinterpret_cast<const float4*>(&Rs[rr * SLD + c]);
    const float4 n = *reinterpret_cast<const float4*>(&cnp[0][c]);
    *reinterpret_cast<float4*>(&Ro[e]) = make_float4(x.x * n.x, x.y * n.y, x.z * n.z, x.w * n.w);
  }
  TT_MARK(tt, 14, tt_last);                                 // sort, norms, store
}
