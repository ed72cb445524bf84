// solve_quad_kernel: the 128x128 two-sided Jacobi sweep of one block pair in a QUAD ROUND-ROBIN ordering.
// (included by svd_jacobi.cu after solve_kernel, whose prologue / epilogue it shares.)
//
// Why another ordering.  In the odd-even ordering of solve_kernel every step needs the diagonal and the first
// super-diagonal published through shared memory, two CTA-wide barriers, and on odd steps two boundary exchanges,
// because the pivot pairs straddle the 8x8 register patches; measured 1.5 us per step, 2/3 of it synchronisation.
// Here the 128 positions are 32 quads; patch row/column g holds the quads (g, L) = local 0-3 and (g, H) = local 4-7.
// Pairs are always taken INSIDE an 8-group, so
//   * the 2x2 pivot blocks lie in the diagonal patches: the 16 diagonal-patch threads compute the rotation
//     parameters straight from their registers (no publish step, no second barrier);
//   * a step is: parameters -> one 256-thread barrier -> 128 FMAs per thread, with no data movement at all;
//   * data moves only once per ROUND (4 steps): a recursive tournament over the 32 quads (c_quad_src) moves whole quads
//     between patches through shared memory, at most half of them per move.
// Schedule: 3 steps inside the quads ((0,1)(2,3) / (0,2)(1,3) / (0,3)(1,2)), then 31 rounds of 4 steps pairing
// L_i with H_(i^j), j = 0..3 (XOR rather than a cyclic shift: the partners of the column pairs (0,1) and (2,3) are then
// always a column PAIR, (4,5) or (6,7), in order or swapped -- what the packed FMAs of the triangular solve need):
// 127 steps, every pair of positions exactly once.
// The accumulated rotation R is kept by the other 256 threads as 8x8 patches too (column operations only).  They do
// not take part in the G barriers: every step's parameters stay in a shared-memory history and the R threads follow
// behind at their own pace (one single-use mbarrier per step says "published"), filling the issue slots the G threads leave.
// Rotations are in the scaled (fast Givens) form, one FMA per element and rotation, with the deferred scales folded
// back every 8 rounds so they stay within 0.707^36.

constexpr int QSTEPS = 127;
constexpr int QROUNDS = 31;
constexpr int QFOLDS = 3;                      // after rounds 7, 15, 23
constexpr size_t SOLVEQ_SMEM = sizeof(float) * (2 * JK * SLD) + sizeof(float2) * QSTEPS * 64 + sizeof(float) * 64 +
                               sizeof(int) * JK + sizeof(float) * JK * (3 + QFOLDS) + sizeof(uint64_t) * (QSTEPS + 1) + (QROUNDS - 1) * 32;

// local positions (p, q) of pair k in a step of the given type: 0-2 inside the quads, 3-6 = L_i with H_(i ^ (type-3))
__host__ __device__ constexpr int qp_p(int type, int k) { return type == 0 ? 2 * k : type <= 2 ? (k < 2 ? k : k + 2) : k; }
__host__ __device__ constexpr int qp_q(int type, int k) {
  return type == 0 ? 2 * k + 1
       : type == 1 ? (k < 2 ? k + 2 : k + 4)
       : type == 2 ? (k == 0 ? 3 : k == 1 ? 2 : k == 2 ? 7 : 6)
                   : 4 + (k ^ (type - 3));
}

__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Scaled Jacobi rotation of the pivot (p, q): true entries are d_i d_j ghat_ij.  Returns (x, y) = (-beta, alpha) with
// xhat_p' = xhat_p + x xhat_q, xhat_q' = xhat_q + y xhat_p (old xhat_p), and z = c, the factor both scales take.
__device__ __forceinline__ float3 quad_rotation(float gpp, float gqq, float gpq, float dp, float dq) {
  // Everything is written in the stored (scaled) entries: with rho = d_q / d_p,
  //   tau = (a_qq - a_pp) / (2 a_pq) = (rho ghat_qq - ghat_pp / rho) / (2 ghat_pq),  and the threshold test
  //   a_pq^2 > eps a_pp a_qq is scale-free.  t = sign(tau) / (|tau| + sqrt(1 + tau^2)) is evaluated with ONE sqrt and
  //   ONE reciprocal on the dependent chain: t = sign(delta h) |h| / (|delta| + sqrt(delta^2 + h^2)).
  // Raw approximate MUFU ops: any t gives an exact rotation, and the operands are far from the denormal range (the
  // scales are folded back every 8 rounds), so the range-scaling code of __fdividef / rsqrtf / sqrtf is dead weight.
  const float rho = dq * rcp_ftz(dp), rho_inv = dp * rcp_ftz(dq);
  const float delta = fmaf(rho, gqq, -(gpp * rho_inv)), h = gpq + gpq;
  float c = 1.f, t = 0.f;
  if (gpq * gpq > 1e-16f * (gpp * gqq) && gpq != 0.f) {         // |cos| > 1e-8
    const float s = sqrt_ftz(fmaf(delta, delta, h * h));
    t = copysignf(fabsf(h) * rcp_ftz(fabsf(delta) + s), delta * h);
    t = (fabsf(t) <= 1.f) ? t : 0.f;          // |t| <= 1 by construction; anything else is underflow debris (inf / NaN): no rotation
    c = rsqrt_ftz(fmaf(t, t, 1.f));
  }
  return make_float3(-t * rho, t * rho_inv, c);
}

template <int TYPE>
__device__ __forceinline__ void quad_rows(float (&g)[8][8], const float2 (&q)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    constexpr int dummy = 0; (void)dummy;
    const int p = qp_p(TYPE, k), r = qp_q(TYPE, k);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = g[p][j], b = g[r][j];
      g[p][j] = fmaf(q[k].x, b, a);
      g[r][j] = fmaf(q[k].y, a, b);
    }
  }
}
template <int TYPE>
__device__ __forceinline__ void quad_cols(float (&g)[8][8], const float2 (&q)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = qp_p(TYPE, k), r = qp_q(TYPE, k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = g[i][p], b = g[i][r];
      g[i][p] = fmaf(q[k].x, b, a);
      g[i][r] = fmaf(q[k].y, a, b);
    }
  }
}

// A step's 64 rotations (16 groups x 4 pivots x float2) are stored as TWO PLANES of 16 float4: plane 0 holds pivots 0-1
// of every group, plane 1 pivots 2-3, so that the 16 lanes of a half-warp (group = lane & 15) read 256 consecutive bytes
// per LDS.128.  With the four pivots of a group side by side (32-byte stride) every such load was a 2-way bank conflict,
// a third of the kernel's shared-memory wavefronts -- and the lead warp's dependent chain queues behind exactly that
// traffic (its own loads and shuffles share the pipe: profiles/r02_solve_timing.log).
__device__ __forceinline__ float4* q_plane(float2* step_base, int grp, int plane) {
  return reinterpret_cast<float4*>(step_base) + plane * 16 + grp;
}
__device__ __forceinline__ void load_q4(const float2* step_base, int grp, float2 (&q)[4]) {
  const float4* p4 = reinterpret_cast<const float4*>(step_base);
  const float4 a = p4[grp], b = p4[16 + grp];
  q[0] = make_float2(a.x, a.y); q[1] = make_float2(a.z, a.w); q[2] = make_float2(b.x, b.y); q[3] = make_float2(b.z, b.w);
}

// One step of the G threads.  The LEAD warp (G threads 0-31: the 16 diagonal patches and the 16 patches right of them)
// computes the four rotations of every group from its registers, publishes them and only ARRIVES at the step's named
// barrier, so it runs ahead of the other seven warps (which wait on that barrier) by up to a whole round; the barrier
// ids of a round are distinct and the blocking barriers of the quad move separate their reuse.
#ifdef ASVD_SOLVE_TIMING
__device__ unsigned long long g_solve_timing[16];
#define QT_MARK(k) do { if (gtid == 0) { const long long _t = clock64(); qt_acc[k] += _t - qt_last; qt_last = _t; } } while (0)
__shared__ long long qt_acc[12];
__shared__ long long qt_last;
#else
#define QT_MARK(k) do { } while (0)
#endif

template <int TYPE, bool LEAN = false>
__device__ __forceinline__ void quad_g_step(float (&g)[8][8], float (&d)[8], bool lead, bool is_diag, float2* cs_step, int pa,
                                            int pc, int gtid, uint64_t* mb, int step, int bar_id,
                                            float2* __restrict__ hist_step = nullptr) {
  if (lead) {
    // The parameter arithmetic is the dependent chain of the whole sweep.  Every diagonal lane computes the four
    // rotations of its group from its own registers (four independent chains of ~35 instructions interleave).  The first
    // version handed pivots 2 and 3 to lane L + 16 -- 10 shuffles out, 6 back -- to halve the arithmetic per lane; measured
    // (profiles/r02_solve_timing.log) those shuffles cost more than the whole arithmetic: they queue in the same
    // shared-memory / shuffle pipe the other fifteen warps keep busy with their rotation loads.
    QT_MARK(0);                                   // time since the end of the previous step's apply (moves, folds, ...)
    float3 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pl = qp_p(TYPE, k), rl = qp_q(TYPE, k);
      o[k] = quad_rotation(g[pl][pl], g[rl][rl], g[pl][rl], d[pl], d[rl]);
    }
    QT_MARK(2);                                   // rotation parameters
    if (is_diag) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { d[qp_p(TYPE, k)] *= o[k].z; d[qp_q(TYPE, k)] *= o[k].z; }
      *q_plane(cs_step, pa, 0) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
      *q_plane(cs_step, pa, 1) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
      if (LEAN) {                                          // the replay kernel reads the rotations from HBM / L2
        *q_plane(hist_step, pa, 0) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        *q_plane(hist_step, pa, 1) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
      }
    }
    __syncwarp();
    asm volatile("bar.arrive %0, 256;" ::"r"(bar_id) : "memory");
    if (!LEAN && gtid == 0) tc::mbar_arrive(&mb[step]);    // release: the R threads may consume this step
    QT_MARK(3);                                   // shuffles back, publish, arrive
  } else {
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
  }
  float2 qr[4], qc[4];
  load_q4(cs_step, pa, qr);
  load_q4(cs_step, pc, qc);
  if (lead) QT_MARK(4);                           // parameter loads
  quad_rows<TYPE>(g, qr);
  quad_cols<TYPE>(g, qc);
#ifdef ASVD_SOLVE_TIMING
  if (lead) { asm volatile("" ::"f"(g[0][0]), "f"(g[7][7]), "f"(g[3][4]) : "memory"); QT_MARK(5); }   // the lead warp's own 128 FMAs
#endif
}

template <int TYPE, bool WAIT = true>
__device__ __forceinline__ void quad_r_step(float (&r)[8][8], const float2* cs_step, int pc, uint64_t* mb, int step) {
  if (WAIT) tc::mbar_wait(&mb[step], 0);                   // acquire; every barrier of the array is used once
  float2 qc[4];
  load_q4(cs_step, pc, qc);
  quad_cols<TYPE>(r, qc);
}

// staging address of chunk `chunk` (float4 index 0..31) of position row `pos`.  Thread lt holds patch
// (pa, pc) = (lt & 15, (pa + (lt >> 4)) & 15): the 8 lanes of a quarter-warp differ in pa, hence in pc, so both the
// row-wise accesses (chunk = pc) and the transposed ones (chunk = pa) fall on distinct banks; and the 16 diagonal
// patches sit in lanes 0-15 of ONE warp, the only one that runs the rotation-parameter code.
__device__ __forceinline__ float* quad_stage(float* st, int pos, int chunk) { return st + pos * JK + (chunk << 2); }
// Schedule of the quad moves: c_quad_src[r][2g + h] = first position of the quad whose contents move into slot (g, h)
// after round r.  A recursive tournament over the 32 quads: 16 rounds in which slot L of group g meets the H quads of
// all 16 groups (only the H quads move, one group along), then the L quads of groups 8-15 drop into the H slots of
// groups 0-7 (and the H quads of groups 0-7 into the L slots of 8-15) and both halves repeat the scheme with 8 groups,
// then 4, 2, 1: 16 + 8 + 4 + 2 + 1 = 31 rounds, every pair of quads exactly once (checked on the host when the table
// is built), and in every move at least half of the quads stay where they are -- the staging traffic of a move is
// the dominant cost of the sweep after the dependent chain, and a sub-block whose row quad and column quad both stay
// never goes through shared memory.
__constant__ unsigned char c_quad_src[(QROUNDS - 1) * 32];
// the same table in global memory: lanes copy it to shared memory with DIFFERENT indices, which a constant bank
// serialises (8 % of the replay kernel's samples were that copy); 240 coalesced words instead
__device__ unsigned int g_quad_src_words[(QROUNDS - 1) * 8];

static bool build_quad_schedule(unsigned char* tab /* [(QROUNDS-1)*32] */) {
  int cur[32];                                   // quad id at slot 2g + h
  for (int i = 0; i < 32; ++i) cur[i] = i;
  bool met[32][32] = {};
  int round = 0;
  for (int B = 16; B >= 1; B >>= 1) {
    for (int k = 0; k < B; ++k, ++round) {
      for (int g = 0; g < 16; ++g) {
        const int a = cur[2 * g], b = cur[2 * g + 1];
        if (met[a][b]) return false;
        met[a][b] = met[b][a] = true;
      }
      if (round == QROUNDS - 1) break;
      int src[32];                               // slot -> slot its new content comes from
      for (int i = 0; i < 32; ++i) src[i] = i;
      if (k < B - 1) {                           // shift the H quads one group along inside every block of B groups
        for (int g = 0; g < 16; ++g) {
          const int b0 = g - g % B;
          src[2 * g + 1] = 2 * (b0 + (g - b0 + 1) % B) + 1;
        }
      } else {                                   // next level: first half of each block takes the L quads, second the H quads
        const int h = B / 2;
        for (int b0 = 0; b0 < 16; b0 += B)
          for (int i = 0; i < h; ++i) {
            src[2 * (b0 + i) + 1] = 2 * (b0 + h + i);          // L of the second half -> H slot of the first half
            src[2 * (b0 + h + i)] = 2 * (b0 + i) + 1;          // H of the first half  -> L slot of the second half
          }
      }
      int nxt[32];
      for (int i = 0; i < 32; ++i) { nxt[i] = cur[src[i]]; tab[round * 32 + i] = (unsigned char)(8 * (src[i] >> 1) + 4 * (src[i] & 1)); }
      for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
    }
  }
  if (round != QROUNDS - 1) return false;
  for (int a = 0; a < 32; ++a)
    for (int b = 0; b < 32; ++b)
      if (a != b && !met[a][b]) return false;
  return true;
}
static cudaError_t upload_quad_schedule() {
  static bool done[ASVD_MAX_DEVICES] = {};
  const int dev = current_device_slot();
  if (done[dev]) return cudaSuccess;
  unsigned char tab[(QROUNDS - 1) * 32];
  if (!build_quad_schedule(tab)) return cudaErrorUnknown;
  cudaError_t e = cudaMemcpyToSymbol(c_quad_src, tab, sizeof(tab));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_quad_src_words, tab, sizeof(tab));
  if (e == cudaSuccess) done[dev] = true;
  return e;
}

// staging chunk of the column quad that starts at position `pos`
__device__ __forceinline__ int quad_chunk(int pos) { return ((pos >> 2) & 1) * 16 + (pos >> 3); }

// Every thread writes the sub-blocks of its patch that leave it (row 8pa+i, column-quad chunks pc and 16+pc) ...
// mv[2 hr + hc]: the sub-block (row quad hr, column quad hc) changes its content in this move
__device__ __forceinline__ void quad_stage_write(float* st, const float (&g)[8][8], int pa, int pc, const bool (&mv)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (mv[2 * (i >> 2)]) *reinterpret_cast<float4*>(quad_stage(st, 8 * pa + i, pc)) = make_float4(g[i][0], g[i][1], g[i][2], g[i][3]);
    if (mv[2 * (i >> 2) + 1])
      *reinterpret_cast<float4*>(quad_stage(st, 8 * pa + i, 16 + pc)) = make_float4(g[i][4], g[i][5], g[i][6], g[i][7]);
  }
}
// ... and reads the sub-blocks its slot receives: rows from the source row quads (rL, rH: first positions; 8pa and
// 8pa+4 when the rows stay), columns from the source column quads
__device__ __forceinline__ void quad_stage_read(float* st, float (&g)[8][8], int rL, int rH, int cL, int cH, const bool (&mv)[4]) {
  const int kL = quad_chunk(cL), kH = quad_chunk(cH);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int sp = (i < 4 ? rL : rH) + (i & 3);
    if (mv[2 * (i >> 2)]) {
      const float4 lo = *reinterpret_cast<const float4*>(quad_stage(st, sp, kL));
      g[i][0] = lo.x; g[i][1] = lo.y; g[i][2] = lo.z; g[i][3] = lo.w;
    }
    if (mv[2 * (i >> 2) + 1]) {
      const float4 hi = *reinterpret_cast<const float4*>(quad_stage(st, sp, kH));
      g[i][4] = hi.x; g[i][5] = hi.y; g[i][6] = hi.z; g[i][7] = hi.w;
    }
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1)
solve_quad_kernel(const float* __restrict__ Gpart, int chunks, int pairs_per_mat, float* __restrict__ Rout,
                  int* __restrict__ pairflag, unsigned* __restrict__ maxoff_bits, int* __restrict__ status,
                  const int* __restrict__ done, float tol, int transpose_out, const int2* __restrict__ pairs,
                  int* __restrict__ track, int nb, int round_stamp, const int* __restrict__ precise_b, int half_gram_tc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* G = reinterpret_cast<float*>(smem_raw);            // [JK][SLD] summed Gram; staging of the G threads; later E
  float* Rs = G + JK * SLD;                                 // [JK][SLD] staging of the R threads; then R, sorted columns
  float2* csh = reinterpret_cast<float2*>(Rs + JK * SLD);   // [QSTEPS][16 groups][4 pairs] rotation history
  float* red = reinterpret_cast<float*>(csh + QSTEPS * 64); // [64]
  int* dest = reinterpret_cast<int*>(red + 64);             // [JK] output column of each position
  float* gd = reinterpret_cast<float*>(dest + JK);          // [JK] final diagonal (true norms)
  float* dmov = gd + JK;                                    // [JK] scales in transit during a quad move
  float* dfin = dmov + JK;                                  // [JK] final scales
  float* dhist = dfin + JK;                                 // [QFOLDS][JK] scales folded into G (and, later, into R)
  uint64_t* mb = reinterpret_cast<uint64_t*>(dhist + QFOLDS * JK);   // [QSTEPS + 1] one-shot "step published" barriers
  unsigned char* qsrc = reinterpret_cast<unsigned char*>(mb + QSTEPS + 1);   // [(QROUNDS-1)*32] copy of c_quad_src (lanes index it divergently)

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  const int tid = threadIdx.x;
  const int2 pr = pairs[p];
  int* trk = track + (int64_t)b * (nb + nb * nb);
  if (pair_is_clean(track, nb, b, pr.x, pr.y)) {           // untouched since it was last verified: nothing to do
    if (tid == 0) pairflag[idx] = 0;
    return;
  }
  if (tid < QSTEPS + 1) tc::mbar_init(&mb[tid], 1);         // ordered before their first use by the prologue's barriers
  for (int i = tid; i < (QROUNDS - 1) * 32; i += SOLVE_THREADS) qsrc[i] = c_quad_src[i];
  const int precise = precise_b[b], half_gram = half_gram_tc && precise;     // Gram mode of THIS matrix (see run_svd)
  if (!solve_prologue(Gpart, chunks, idx, b, pr, tid, G, red, pairflag, maxoff_bits, status, tol, trk, nb, round_stamp,
                      precise, gridDim.y, half_gram))
    return;

  // Warp roles.  A warp's scheduler is warp_id % 4; the lead warp (G warp 0) carries the dependent chain of the
  // sweep, so it shares scheduler 0 with three R warps (half the work of a G warp, and never ahead of it) while the
  // other seven G warps and five R warps fill schedulers 1-3.
  //   warp        0  1  2  3  4  5  6  7  8  9 10 11 12 13 14 15
  //   role        G0 G1 G2 G3 R0 G4 G5 G6 R1 G7 R2 R3 R4 R5 R6 R7
  const int wid = tid >> 5;
  const unsigned g_mask = 0x02EFu;                              // warps 0-3, 5-7, 9
  const bool is_g = (g_mask >> wid) & 1u;
  const int ridx = __popc((is_g ? g_mask : ~g_mask) & ((1u << wid) - 1u));   // index among the warps of the same role
  const int lt = ridx * 32 + (tid & 31);
  const int pa = lt & 15, pc = (pa + (lt >> 4)) & 15;
  if (is_g) {
    // ---------------------------------------------------------------- G threads
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
    const bool is_diag = (pa == pc);
    const bool lead = lt < 32;
    auto bar_g = [] { asm volatile("bar.sync 2, 256;" ::: "memory"); };
    bar_g();                                                 // every patch is in registers: G becomes the staging area
#ifdef ASVD_SOLVE_TIMING
    if (lt == 0) { for (int i = 0; i < 12; ++i) qt_acc[i] = 0; qt_last = clock64(); }
    const int gtid = lt;
#endif

    quad_g_step<0>(g, d, lead, is_diag, csh + 0 * 64, pa, pc, lt, mb, 0, 8);
    quad_g_step<1>(g, d, lead, is_diag, csh + 1 * 64, pa, pc, lt, mb, 1, 9);
    quad_g_step<2>(g, d, lead, is_diag, csh + 2 * 64, pa, pc, lt, mb, 2, 10);
#pragma unroll 1
    for (int r = 0; r < QROUNDS; ++r) {
      const int s0 = 3 + 4 * r;
      quad_g_step<3>(g, d, lead, is_diag, csh + (s0 + 0) * 64, pa, pc, lt, mb, s0 + 0, 4);
      quad_g_step<4>(g, d, lead, is_diag, csh + (s0 + 1) * 64, pa, pc, lt, mb, s0 + 1, 5);
      quad_g_step<5>(g, d, lead, is_diag, csh + (s0 + 2) * 64, pa, pc, lt, mb, s0 + 2, 6);
      quad_g_step<6>(g, d, lead, is_diag, csh + (s0 + 3) * 64, pa, pc, lt, mb, s0 + 3, 7);
      if (r == QROUNDS - 1) break;
      if ((r & 7) == 7) {
        // fold the deferred scales back into the stored values
        float* dh = dhist + (r >> 3) * JK;
        if (is_diag) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { dh[8 * pa + i] = d[i]; d[i] = 1.f; }
        }
        bar_g();
        float dr[8], dc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dr[i] = dh[8 * pa + i]; dc[i] = dh[8 * pc + i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) g[i][j] *= dr[i] * dc[j];
      }
      QT_MARK(6);                                  // (lead lane) fold, if any
      // ---- quad move: one pass through shared memory (rows and columns at once), only what changes place
      const unsigned char* qs = qsrc + r * 32;
      const int rsrcL = qs[2 * pa], rsrcH = qs[2 * pa + 1], csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
      const bool mrL = rsrcL != 8 * pa, mrH = rsrcH != 8 * pa + 4, mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
      const bool mv[4] = {mrL || mcL, mrL || mcH, mrH || mcL, mrH || mcH};
      quad_stage_write(G, g, pa, pc, mv);
      if (is_diag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmov[8 * pa + i] = d[i];
      }
      bar_g();
      quad_stage_read(G, g, rsrcL, rsrcH, csrcL, csrcH, mv);
      if (is_diag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = dmov[(i < 4 ? rsrcL : rsrcH) + (i & 3)];
      }
      bar_g();               // blocking for the lead warp too: nobody writes the staging area while it is being read
      QT_MARK(7);                                  // quad move incl. waiting for the other warps at its barriers
    }
#ifdef ASVD_SOLVE_TIMING
    if (lt == 0 && blockIdx.x == 0 && blockIdx.y == 0)
      for (int i = 0; i < 8; ++i) g_solve_timing[i] = (unsigned long long)qt_acc[i];
#endif
    // final diagonal (true norms) and scales, ranks for the norm sort
    if (is_diag) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { gd[8 * pa + i] = d[i] * d[i] * g[i][i]; dfin[8 * pa + i] = d[i]; }
    }
    bar_g();
    if (lt < JK) {
      const float dd = gd[lt];
      int rank = 0;
      for (int j = 0; j < JK; ++j) {
        const float e = gd[j];
        rank += (e > dd) || (e == dd && j < lt);
      }
      dest[lt] = rank;
    }
    bar_g();
    if (lt == 0) tc::mbar_arrive(&mb[QSTEPS]);
  } else {
    // ---------------------------------------------------------------- R threads
    float r[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) r[i][j] = (pa == pc && i == j) ? 1.f : 0.f;
    auto bar_r = [] { asm volatile("bar.sync 3, 256;" ::: "memory"); };
    quad_r_step<0>(r, csh + 0 * 64, pc, mb, 0);
    quad_r_step<1>(r, csh + 1 * 64, pc, mb, 1);
    quad_r_step<2>(r, csh + 2 * 64, pc, mb, 2);
#pragma unroll 1
    for (int rd = 0; rd < QROUNDS; ++rd) {
      const int s0 = 3 + 4 * rd;
      quad_r_step<3>(r, csh + (s0 + 0) * 64, pc, mb, s0 + 0);
      quad_r_step<4>(r, csh + (s0 + 1) * 64, pc, mb, s0 + 1);
      quad_r_step<5>(r, csh + (s0 + 2) * 64, pc, mb, s0 + 2);
      quad_r_step<6>(r, csh + (s0 + 3) * 64, pc, mb, s0 + 3);
      if (rd == QROUNDS - 1) break;
      if ((rd & 7) == 7) {
        // the G threads wrote this fold's scales before publishing the next step
        tc::mbar_wait(&mb[s0 + 4], 0);
        const float* dh = dhist + (rd >> 3) * JK;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dc = dh[8 * pc + j];
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i][j] *= dc;
        }
      }
      const unsigned char* qs = qsrc + rd * 32;
      const int csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
      const bool mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
      const bool mv[4] = {mcL, mcH, mcL, mcH};              // rows never move: only the column quads decide
      quad_stage_write(Rs, r, pa, pc, mv);
      bar_r();
      quad_stage_read(Rs, r, 8 * pa, 8 * pa + 4, csrcL, csrcH, mv);
      bar_r();
    }
    tc::mbar_wait(&mb[QSTEPS], 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = dest[8 * pc + j];
      const float dc = dfin[8 * pc + j];
#pragma unroll
      for (int i = 0; i < 8; ++i) Rs[(8 * pa + i) * SLD + col] = r[i][j] * dc;
    }
  }
  __syncthreads();
  solve_polish_write(G, Rs, Rout, idx, tid, transpose_out);
}


// ================================================================================================ lean variant
// EXPERIMENTAL (ASVD_B200_SOLVE=lean; written at the end of round 1 without GPU time left: compiled, not yet run).
// Why.  solve_quad_kernel is 512 threads x 128 registers: it fills an SM's register file, so a batch of four 4096^2
// occupies 128 SMs for 120 us at 27 % issue utilisation, and no other work can share those SMs (levers).
// Half of its threads only accumulate R and follow the G threads through the 65 KB rotation history.  Here the sweep is
// split in two kernels:
//   solve_quad_g_kernel  the 256 G threads alone: same steps, same barriers, same arithmetic; the lead warp ALSO
//                        streams every step's 64 rotations (512 B) to a global history and the history in shared
//                        memory shrinks to a ring of 8 steps (the lead warp is never more than 7 steps ahead: the
//                        quad move at the end of a round is a blocking barrier).  256 threads x 128 registers and
//                        ~75 KB of shared memory: TWO CTAs per SM, i.e. 296 block pairs in one wave.
//   solve_quad_r_kernel  replays the complete history on R (the R half of solve_quad_kernel without the mbarrier
//                        waits: throughput-bound, no dependent chain), sorts, normalises and writes R.
// Same operations in the same order on every element: R is expected to be bitwise the one solve_quad_kernel writes.
constexpr int QRING = 8;
// per-pair global record: rotation history, folded scales, final scales, output column of every position
constexpr size_t QAUX_FLOATS = (size_t)QSTEPS * 128 + (size_t)QFOLDS * JK + JK + JK;
static_assert(QAUX_FLOATS == QAUX_FLOATS_PLAN && QAUX_FLOATS % 4 == 0, "workspace plan and lean solve record disagree");
constexpr size_t SOLVEQG_SMEM = sizeof(float) * (JK * SLD) + sizeof(float2) * QRING * 64 + sizeof(float) * 64 +
                                sizeof(float) * JK * 3 + (QROUNDS - 1) * 32;
constexpr size_t SOLVEQR_SMEM = sizeof(float) * (JK * SLD) + sizeof(float2) * QSTEPS * 64 + sizeof(float) * JK * (QFOLDS + 1) +
                                sizeof(int) * JK + (QROUNDS - 1) * 32;

__global__ void __launch_bounds__(256, 2)
solve_quad_g_kernel(const float* __restrict__ Gpart, int chunks, int pairs_per_mat, float* __restrict__ aux,
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
  unsigned char* qsrc = reinterpret_cast<unsigned char*>(dfold + JK);

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
  for (int i = tid; i < (QROUNDS - 1) * 32; i += 256) qsrc[i] = c_quad_src[i];
  const int precise = precise_b[b], half_gram = half_gram_tc && precise;
  if (!solve_prologue<256>(Gpart, chunks, idx, b, pr, tid, G, red, pairflag, maxoff_bits, status, tol, trk, nb, round_stamp,
                           precise, gridDim.y, half_gram))
    return;

  float* ax = aux + (int64_t)idx * QAUX_FLOATS;
  float2* hist = reinterpret_cast<float2*>(ax);             // [QSTEPS][64]
  float* a_dhist = ax + (size_t)QSTEPS * 128;               // [QFOLDS][JK]
  float* a_dfin = a_dhist + QFOLDS * JK;                    // [JK]
  int* a_dest = reinterpret_cast<int*>(a_dfin + JK);        // [JK]

  const int lt = tid;
  const int pa = lt & 15, pc = (pa + (lt >> 4)) & 15;
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
  const bool is_diag = (pa == pc);
  const bool lead = lt < 32;
  auto bar_g = [] { asm volatile("bar.sync 2, 256;" ::: "memory"); };
  bar_g();                                                  // every patch is in registers: G becomes the staging area

#define QG_STEP(TYPE, S, BAR) \
  quad_g_step<TYPE, true>(g, d, lead, is_diag, csh + ((S) & (QRING - 1)) * 64, pa, pc, lt, nullptr, (S), (BAR), hist + (S) * 64)
  QG_STEP(0, 0, 8);
  QG_STEP(1, 1, 9);
  QG_STEP(2, 2, 10);
#pragma unroll 1
  for (int r = 0; r < QROUNDS; ++r) {
    const int s0 = 3 + 4 * r;
    QG_STEP(3, s0 + 0, 4);
    QG_STEP(4, s0 + 1, 5);
    QG_STEP(5, s0 + 2, 6);
    QG_STEP(6, s0 + 3, 7);
    if (r == QROUNDS - 1) break;
    if ((r & 7) == 7) {
      // fold the deferred scales back into the stored values; the replay kernel folds the same values into R
      if (is_diag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dfold[8 * pa + i] = d[i]; a_dhist[(r >> 3) * JK + 8 * pa + i] = d[i]; d[i] = 1.f; }
      }
      bar_g();
      float dr[8], dc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { dr[i] = dfold[8 * pa + i]; dc[i] = dfold[8 * pc + i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] *= dr[i] * dc[j];
      // (dfold is next written eight rounds later, behind many blocking barriers)
    }
    const unsigned char* qs = qsrc + r * 32;
    const int rsrcL = qs[2 * pa], rsrcH = qs[2 * pa + 1], csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
    const bool mrL = rsrcL != 8 * pa, mrH = rsrcH != 8 * pa + 4, mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
    const bool mv[4] = {mrL || mcL, mrL || mcH, mrH || mcL, mrH || mcH};
    quad_stage_write(G, g, pa, pc, mv);
    if (is_diag) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmov[8 * pa + i] = d[i];
    }
    bar_g();
    quad_stage_read(G, g, rsrcL, rsrcH, csrcL, csrcH, mv);
    if (is_diag) {
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = dmov[(i < 4 ? rsrcL : rsrcH) + (i & 3)];
    }
    bar_g();
  }
#undef QG_STEP
  if (is_diag) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { gd[8 * pa + i] = d[i] * d[i] * g[i][i]; a_dfin[8 * pa + i] = d[i]; }
  }
  bar_g();
  if (lt < JK) {
    const float dd = gd[lt];
    int rank = 0;
    for (int j = 0; j < JK; ++j) {
      const float e = gd[j];
      rank += (e > dd) || (e == dd && j < lt);
    }
    a_dest[lt] = rank;
  }
}

__global__ void __launch_bounds__(256, 1)
solve_quad_r_kernel(const float* __restrict__ aux, int pairs_per_mat, float* __restrict__ Rout,
                    const int* __restrict__ pairflag, const int* __restrict__ done) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Rs = reinterpret_cast<float*>(smem_raw);           // [JK][SLD] staging of the moves; then R, sorted columns
  float2* csh = reinterpret_cast<float2*>(Rs + JK * SLD);   // [QSTEPS][64] the complete rotation history
  float* dhist = reinterpret_cast<float*>(csh + QSTEPS * 64);   // [QFOLDS][JK]
  float* dfin = dhist + QFOLDS * JK;                        // [JK]
  int* dest = reinterpret_cast<int*>(dfin + JK);            // [JK]
  unsigned char* qsrc = reinterpret_cast<unsigned char*>(dest + JK);
  __shared__ float cnp[4][JK];

  const int b = blockIdx.y, p = blockIdx.x;
  if (done[b]) return;
  const int idx = b * pairs_per_mat + p;
  if (!pairflag[idx]) return;                               // clean, converged or non-finite pair: no rotation, no R
  const int tid = threadIdx.x;
  const float* ax = aux + (int64_t)idx * QAUX_FLOATS;
  {
    // history, scales and destinations are one contiguous record: float4 copies (QAUX_FLOATS is a multiple of 4)
    const float4* src = reinterpret_cast<const float4*>(ax);
    float4* dst = reinterpret_cast<float4*>(csh);           // csh | dhist | dfin | dest are contiguous in shared memory too
    for (int i = tid; i < (int)(QAUX_FLOATS / 4); i += 256) dst[i] = src[i];
    for (int i = tid; i < (QROUNDS - 1) * 32; i += 256) qsrc[i] = c_quad_src[i];
  }
  __syncthreads();

  const int pa = tid & 15, pc = (pa + (tid >> 4)) & 15;
  float r[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[i][j] = (pa == pc && i == j) ? 1.f : 0.f;
  quad_r_step<0, false>(r, csh + 0 * 64, pc, nullptr, 0);
  quad_r_step<1, false>(r, csh + 1 * 64, pc, nullptr, 1);
  quad_r_step<2, false>(r, csh + 2 * 64, pc, nullptr, 2);
#pragma unroll 1
  for (int rd = 0; rd < QROUNDS; ++rd) {
    const int s0 = 3 + 4 * rd;
    quad_r_step<3, false>(r, csh + (s0 + 0) * 64, pc, nullptr, s0 + 0);
    quad_r_step<4, false>(r, csh + (s0 + 1) * 64, pc, nullptr, s0 + 1);
    quad_r_step<5, false>(r, csh + (s0 + 2) * 64, pc, nullptr, s0 + 2);
    quad_r_step<6, false>(r, csh + (s0 + 3) * 64, pc, nullptr, s0 + 3);
    if (rd == QROUNDS - 1) break;
    if ((rd & 7) == 7) {
      const float* dh = dhist + (rd >> 3) * JK;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dc = dh[8 * pc + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i][j] *= dc;
      }
    }
    const unsigned char* qs = qsrc + rd * 32;
    const int csrcL = qs[2 * pc], csrcH = qs[2 * pc + 1];
    const bool mcL = csrcL != 8 * pc, mcH = csrcH != 8 * pc + 4;
    const bool mv[4] = {mcL, mcH, mcL, mcH};                // rows never move: only the column quads decide
    quad_stage_write(Rs, r, pa, pc, mv);
    __syncthreads();
    quad_stage_read(Rs, r, 8 * pa, 8 * pa + 4, csrcL, csrcH, mv);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = dest[8 * pc + j];
    const float dc = dfin[8 * pc + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) Rs[(8 * pa + i) * SLD + col] = r[i][j] * dc;
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
}
