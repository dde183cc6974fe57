// K2 / K4 -- transposed density contraction (split-K over the grid):
//
//   out[s][a][b] = scale * sum_terms sum_r  A_term[r][a] * M_term,s[r][b]
//
// density VJP (the XLA transpose of grad_dft/molecule.py:409,440,472-474,502 inside value_and_grad,
// grad_dft/train.py:86,147):
//   term 0 : A = ao,     M_s = rb_s*ao + 2 sum_j gb_sj*d_j ao + 2 lb_s*lap_ao
//   term j : A = d_j ao, M_s = (tb_s/2 + 2 lb_s) * d_j ao                        (j = x,y,z; tau/lapl only)
// explicit HF Fock term (grad_dft/molecule.py:606-613): A = ao, M_s = g[w,s,:] * chi[:,w,s,:], scale -1/2.
//
// One CTA per SM.  A CTA owns a (16*MT) x (2 x 16*MT) output tile (the same b-range for both spins, so the
// plane tiles are staged once) and a contiguous slice of grid rows.  Warp 8 is the TMA producer: per 16-row
// k-tile one transaction group brings the A tile, the 1..5 plane tiles and the 16x16 block of per-row
// coefficients into a ring of stages guarded by full/empty mbarriers -- there is no CTA-wide barrier in the
// main loop.  Warps 0..7 (2 along a x 4 along (spin, b-half)) never materialise M: each B fragment is formed
// in registers as sum_q coef[q][r] * plane_q[r][b] (4..5 FMAs) right before it feeds MT DMMA.8x8x4, so the
// planes are read from shared memory exactly once per warp and the tensor pipe is the only busy unit.
// Tiles are [k][cols] with a row pitch of 16*MT+4 doubles (== 4 mod 16), which makes every A/B fragment
// load (address t*pitch + g) bank-conflict-free without swizzling; the coefficient block is [coef][k] so a
// fragment's 4 k-values are consecutive.  Partial tiles go to the workspace and a second kernel adds the
// K-splits in fixed order (bitwise run-to-run reproducible).
#include "common.cuh"
#include <stdlib.h>

namespace gdft {

constexpr int BWD_BKR = 16;
constexpr int BWD_MMA_WARPS = 8;
constexpr int BWD_THREADS = 32 * (BWD_MMA_WARPS + 4);  // two consumer warpgroups + one producer warpgroup
constexpr int BWD_MAX_SLOTS = 5;
constexpr int BWD_COEF_W = 16;  // coefficient rows per grid row (planar: W[coef][Npad])
constexpr int BWD_MAX_STAGES = 6;
constexpr int BWD_MAX_MT = 5;

struct BwdTerm {
  int a_plane, nq, slot_plane0, per_spin, coef_row0;
};
// Tiles come in (at most) two sizes per dimension -- `rem` big tiles of base+1 sub-tiles first, then tc-rem small
// ones of `base` -- hence four tile classes c = 2*(a small) + (b small).  Each class has its own K-split count
// ks_class[c], proportional to the class's DMMA work per k-step, so that every CTA carries the same amount of work
// (a single K-split for all tiles makes the CTAs of the big tiles the stragglers of every wave).
struct BwdParams {
  int64_t N;
  int npad, nsub, nterms, tc, base, rem, stages, maxq, layout;
  int ks_class[4], cta_prefix[5];
  BwdTerm terms[4];
  double* part;  // [kmax][2][npad][npad]
};

// Fragments of one 4-row k-step of one warp: a[i] = A[k][i*8+g], b[j] = sum_q coef_q[k] * plane_q[k][j*8+g]
// (k = k4*4 + t).  PITCH is the shared-memory row pitch.
// PS = plane stride in slots between consecutive q: 1 (planes shared by both spins), 2 (per-spin planes, [q][spin] order)
template <int PITCH, int MI, int NJ, int NQ, int PS = 1>
__device__ __forceinline__ void bwd_load_frags(double (&a)[MI], double (&b)[NJ], const double* __restrict__ sA,
                                               const double* __restrict__ sP, const double* __restrict__ sC, int k4) {
  constexpr int SLOT_ELEMS = BWD_BKR * PITCH;
  double c[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) c[q] = sC[2 * q * BWD_BKR + k4 * 4];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    double v = c[0] * sP[k4 * 4 * PITCH + j * 8];
#pragma unroll
    for (int q = 1; q < NQ; q++) v = fma(c[q], sP[q * PS * SLOT_ELEMS + k4 * 4 * PITCH + j * 8], v);
    b[j] = v;
  }
#pragma unroll
  for (int i = 0; i < MI; i++) a[i] = sA[k4 * 4 * PITCH + i * 8];
}

template <int MI, int NJ>
__device__ __forceinline__ void bwd_mma_step(double (&acc)[MI][NJ][2], const double (&a)[MI], const double (&b)[NJ]) {
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) dmma884(acc[i][j], a[i], b[j]);
}

struct BwdWarpCtx {
  const double* sStage;
  uint64_t *full, *empty;
  int stage_elems, S, ktiles, maxq, spin, lane;
  int a_off, b_off;  // element offsets of this warp's fragment origin inside the A / plane tiles (t*PITCH + off + g)
  int a_row, b_col;  // global row / column of the warp tile's (0,0) element (+g / +2t added at the store)
  int g, t, split;
};

// One term (ktiles k-tiles starting at ring position st/ph) of one consumer warp, software-pipelined over the
// flattened 4-row k-steps: the fragments of step k+1 (LDS + the in-register combine) are issued before the DMMAs
// of step k, and at a stage boundary the wait on the next stage's full barrier and its first fragment loads are
// hoisted above the last DMMA block of the current stage, so the DMMA stream of a warp never drains between stages.
template <int PITCH, int MI, int NJ, int NQ, int PS = 1>
__device__ __forceinline__ void bwd_term_loop(double (&acc)[MI][NJ][2], const BwdWarpCtx& w, const BwdTerm& Tm, int& st, uint32_t& ph) {
  constexpr int BKR = BWD_BKR, A_ELEMS = BKR * PITCH, SLOT_ELEMS = BKR * PITCH;
  const int p_off = A_ELEMS + (Tm.per_spin ? w.spin * SLOT_ELEMS : 0) + w.b_off;
  const int c_off = A_ELEMS + w.maxq * SLOT_ELEMS + (Tm.coef_row0 + w.spin) * BKR + w.t;
  double a[2][MI], b[2][NJ];
  const double* stage = w.sStage + (size_t)st * w.stage_elems;
  mbar_wait(&w.full[st], ph);
  bwd_load_frags<PITCH, MI, NJ, NQ, PS>(a[0], b[0], stage + w.a_off, stage + p_off, stage + c_off, 0);
  for (int kt = 0; kt < w.ktiles; kt++) {
    const double* sA = stage + w.a_off;
    const double* sP = stage + p_off;
    const double* sC = stage + c_off;
    const int st_cur = st;
    bwd_load_frags<PITCH, MI, NJ, NQ, PS>(a[1], b[1], sA, sP, sC, 1);
    bwd_mma_step<MI, NJ>(acc, a[0], b[0]);
    bwd_load_frags<PITCH, MI, NJ, NQ, PS>(a[0], b[0], sA, sP, sC, 2);
    bwd_mma_step<MI, NJ>(acc, a[1], b[1]);
    bwd_load_frags<PITCH, MI, NJ, NQ, PS>(a[1], b[1], sA, sP, sC, 3);
    bwd_mma_step<MI, NJ>(acc, a[0], b[0]);
    // advance the ring; prefetch the first fragments of the next stage of this term
    if (++st == w.S) { st = 0; ph ^= 1u; }
    stage = w.sStage + (size_t)st * w.stage_elems;
    if (kt + 1 < w.ktiles) {
      mbar_wait(&w.full[st], ph);
      bwd_load_frags<PITCH, MI, NJ, NQ, PS>(a[0], b[0], stage + w.a_off, stage + p_off, stage + c_off, 0);
    }
    bwd_mma_step<MI, NJ>(acc, a[1], b[1]);
    __syncwarp();
    if (w.lane == 0) mbar_arrive(&w.empty[st_cur]);
  }
}

// The whole main loop and the partial-tile store of one consumer warp, specialised on its MI x NJ warp tile.
template <int MT, int MI, int NJ>
__device__ __forceinline__ void bwd_consumer(const BwdParams& p, const BwdWarpCtx& w) {
  constexpr int PITCH = 16 * MT + 4;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  int st = 0;
  uint32_t ph = 0;
  if (w.ktiles > 0) {
    for (int term = 0; term < p.nterms; term++) {
      const BwdTerm Tm = p.terms[term];
      if (Tm.per_spin && Tm.nq == 2) bwd_term_loop<PITCH, MI, NJ, 2, 2>(acc, w, Tm, st, ph);  // two omegas summed in the GEMM
      else if (Tm.nq == 1) bwd_term_loop<PITCH, MI, NJ, 1>(acc, w, Tm, st, ph);
      else if (Tm.nq == 4) bwd_term_loop<PITCH, MI, NJ, 4>(acc, w, Tm, st, ph);
      else bwd_term_loop<PITCH, MI, NJ, 5>(acc, w, Tm, st, ph);
    }
  }

  double* out = p.part + ((size_t)w.split * 2 + w.spin) * p.npad * p.npad;
#pragma unroll
  for (int i = 0; i < MI; i++) {
    const int a = w.a_row + i * 8 + w.g;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      const int b = w.b_col + j * 8 + 2 * w.t;
      *reinterpret_cast<double2*>(out + (size_t)a * p.npad + b) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(BWD_THREADS, 1)
density_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmP,
                   const __grid_constant__ CUtensorMap tmW, const BwdParams p) {
  constexpr int BKR = BWD_BKR, T = 16 * MT, PITCH = T + 4;
  constexpr int A_ELEMS = BKR * PITCH, SLOT_ELEMS = BKR * PITCH, COEF_ELEMS = BKR * BWD_COEF_W;
  const int stage_elems = A_ELEMS + p.maxq * SLOT_ELEMS + COEF_ELEMS;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sStage = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + (size_t)p.stages * stage_elems);
  uint64_t* empty = full + BWD_MAX_STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // CTA -> (tile class, K-split, tile): classes are laid out one after the other
  int cls = 0;
  while (cls < 3 && (int)blockIdx.x >= p.cta_prefix[cls + 1]) cls++;
  const int local = (int)blockIdx.x - p.cta_prefix[cls];
  const int ks = p.ks_class[cls];
  const int a_small = cls >> 1, b_small = cls & 1;
  const int nb_c = b_small ? p.tc - p.rem : p.rem, ntile_c = (a_small ? p.tc - p.rem : p.rem) * nb_c;
  // tiles fastest: CTAs resident together stream the same grid rows (their A / plane tiles hit in L2)
  const int split = local / ntile_c, tile = local - split * ntile_c;
  const int ta = tile / nb_c + (a_small ? p.rem : 0), tb = tile - (tile / nb_c) * nb_c + (b_small ? p.rem : 0);
  const int a_sub0 = ta * p.base + min(ta, p.rem), na = p.base + (a_small ? 0 : 1);
  const int b_sub0 = tb * p.base + min(tb, p.rem), nb = p.base + (b_small ? 0 : 1);
  const int a0 = a_sub0 * 8, b0 = b_sub0 * 8;
  const int64_t ktiles_total = (p.N + BKR - 1) / BKR;
  const int64_t kt_per = (ktiles_total + ks - 1) / ks;
  const int64_t r_begin = (int64_t)split * kt_per * BKR;
  const int64_t r_end = min(p.N, r_begin + kt_per * BKR);
  const int ktiles = r_end > r_begin ? (int)((r_end - r_begin + BKR - 1) / BKR) : 0;
  const int total = p.nterms * ktiles;
  const int S = p.stages;

  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], BWD_MMA_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= BWD_MMA_WARPS) {
    // ---- producer warpgroup: hands its registers to the consumers, then lane 0 of its first warp drives TMA ----------
    // (a lone ninth warp would sit on an SM sub-partition with two consumers and cap every thread at 168 registers;
    //  with a whole warpgroup shrunk to 40 the two consumer warpgroups grow to 232, which the 10 x 3 warp tile needs)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == BWD_MMA_WARPS && lane == 0) {
      int term = 0, kt = 0, st = 0;
      for (int it = 0; it < total; it++) {
        const BwdTerm Tm = p.terms[term];
        const int r = (int)(r_begin + (int64_t)kt * BKR);
        if (it >= S) mbar_wait(&empty[st], ((it / S) - 1) & 1);
        double* sA = sStage + (size_t)st * stage_elems;
        const int nload = Tm.per_spin ? 2 * Tm.nq : Tm.nq;
        mbar_expect_tx(&full[st], (uint32_t)(A_ELEMS + nload * SLOT_ELEMS + COEF_ELEMS) * 8u);
        tma_load_3d(sA, &tmA, &full[st], a0, r, Tm.a_plane);
        for (int q = 0; q < nload; q++) tma_load_3d(sA + A_ELEMS + q * SLOT_ELEMS, &tmP, &full[st], b0, r, Tm.slot_plane0 + q);
        tma_load_3d(sA + A_ELEMS + p.maxq * SLOT_ELEMS, &tmW, &full[st], r, 0, 0);
        if (++st == S) st = 0;
        if (++kt == ktiles) { kt = 0; term++; }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

  // ---- MMA consumers ------------------------------------------------------------------------------
  BwdWarpCtx w;
  w.sStage = sStage; w.full = full; w.empty = empty;
  w.stage_elems = stage_elems; w.S = S; w.ktiles = ktiles; w.maxq = p.maxq; w.lane = lane;
  w.g = g; w.t = t; w.split = split;
  bool ran = false;
  if (p.layout == 1) {
    // 1 x 8 warp grid: every warp spans the whole a-range of the tile (mi = na <= 2 MT A fragments per k-step) and a
    // quarter of one spin's b-range (nj <= ceil(MT/2) B fragments), so each B fragment -- the expensive one: NQ plane
    // loads and NQ-1 DFMAs on the pipe the DMMAs use -- is formed exactly once per CTA and feeds up to 2 MT DMMAs.
    // The nb sub-tiles are split into four parts in descending size order; spin 1 takes the parts in reverse, so the two
    // warps that share an SM sub-partition (warp, warp + 4) carry ceil(nb/2) sub-tiles between them.
    constexpr int NJM = (MT + 1) / 2;
    const int spin = warp >> 2, q = warp & 3;
    const int pb = nb >> 2, pr = nb & 3;
    const int qq = spin == 0 ? q : 3 - q;
    const int nj = pb + (qq < pr ? 1 : 0);
    const int start_desc = qq * pb + min(qq, pr);
    const int col_off = 8 * (spin == 0 ? start_desc : nb - start_desc - nj);
    w.spin = spin;
    w.a_off = t * PITCH + g; w.b_off = t * PITCH + col_off + g;
    w.a_row = a0; w.b_col = b0 + col_off;
    const int mi = na;
    if (nj == NJM) {
      if (mi == 2 * MT) { bwd_consumer<MT, 2 * MT, NJM>(p, w); ran = true; }
      else if (mi == 2 * MT - 1) { bwd_consumer<MT, 2 * MT - 1, NJM>(p, w); ran = true; }
      else if (MT > 1 && mi == 2 * MT - 2) { bwd_consumer<MT, (MT > 1 ? 2 * MT - 2 : 1), NJM>(p, w); ran = true; }
    } else if (NJM > 1 && nj == NJM - 1) {
      if (mi == 2 * MT) { bwd_consumer<MT, 2 * MT, (NJM > 1 ? NJM - 1 : 1)>(p, w); ran = true; }
      else if (mi == 2 * MT - 1) { bwd_consumer<MT, 2 * MT - 1, (NJM > 1 ? NJM - 1 : 1)>(p, w); ran = true; }
      else if (MT > 1 && mi == 2 * MT - 2) { bwd_consumer<MT, (MT > 1 ? 2 * MT - 2 : 1), (NJM > 1 ? NJM - 1 : 1)>(p, w); ran = true; }
    }
  } else {
    // 2 x 4 warp grid: warp tile = mi x nj sub-tiles with mi, nj in {MT, MT-1} (balanced split); each combination runs
    // its own fully unrolled loop, so ragged matrix sizes cost no predication in the hot loop
    const int wm = warp & 1, wn = warp >> 1, spin = wn >> 1, nhalf = wn & 1;
    const int mi = wm == 0 ? (na + 1) / 2 : na / 2, row_off = wm == 0 ? 0 : 8 * ((na + 1) / 2);
    const int nj = nhalf == 0 ? (nb + 1) / 2 : nb / 2, col_off = nhalf == 0 ? 0 : 8 * ((nb + 1) / 2);
    w.spin = spin;
    w.a_off = t * PITCH + row_off + g; w.b_off = t * PITCH + col_off + g;
    w.a_row = a0 + row_off; w.b_col = b0 + col_off;
    if (mi == MT && nj == MT) { bwd_consumer<MT, MT, MT>(p, w); ran = true; }
    else if (MT > 1 && mi == MT && nj == MT - 1) { bwd_consumer<MT, MT, (MT > 1 ? MT - 1 : 1)>(p, w); ran = true; }
    else if (MT > 1 && mi == MT - 1 && nj == MT) { bwd_consumer<MT, (MT > 1 ? MT - 1 : 1), MT>(p, w); ran = true; }
    else if (MT > 1 && mi == MT - 1 && nj == MT - 1) { bwd_consumer<MT, (MT > 1 ? MT - 1 : 1), (MT > 1 ? MT - 1 : 1)>(p, w); ran = true; }
  }
  if (!ran) {
    // empty warp tile (fewer sub-tiles than warps along a dimension): keep the ring moving
    for (int it = 0; it < total; it++) {
      const int st = it % S;
      mbar_wait(&full[st], (it / S) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }
}

// out[s][a][b] = scale * sum_{k < ks(tile class of (a,b))} part[k][s][a][b]   (fixed summation order)
// 32 consecutive outputs x 8 k-groups per CTA: group g sums k = g, g+8, ... (four independent loads in flight), the
// groups are then added in order through shared memory.  A serial k loop per output is a chain of dependent L2 round
// trips: 75 us for the 148 partials of an H2O-sized build, more than the GEMM it follows.
constexpr int RED_OUT = 32, RED_KG = 8;
__global__ void __launch_bounds__(RED_OUT* RED_KG) bwd_reduce_kernel(const double* __restrict__ part, int4 ks_class,
                                                                     int big_end /*first column of the small tiles*/, int npad, int n,
                                                                     double scale, double* __restrict__ out) {
  __shared__ double red[RED_KG][RED_OUT];
  const int o = threadIdx.x & (RED_OUT - 1), g = threadIdx.x / RED_OUT;
  const int idx = blockIdx.x * RED_OUT + o;
  const int total = 2 * n * n;
  double acc = 0.0;
  if (idx < total) {
    const int s = idx / (n * n), rem = idx - s * n * n, a = rem / n, b = rem - a * n;
    const int cls = (a >= big_end ? 2 : 0) + (b >= big_end ? 1 : 0);
    const int ksplit = cls == 0 ? ks_class.x : cls == 1 ? ks_class.y : cls == 2 ? ks_class.z : ks_class.w;
    const size_t stride = (size_t)2 * npad * npad;
    const double* src = part + ((size_t)s * npad + a) * npad + b;
    int k = g;
    for (; k + 3 * RED_KG < ksplit; k += 4 * RED_KG) {
      const double v0 = src[(size_t)k * stride], v1 = src[(size_t)(k + RED_KG) * stride];
      const double v2 = src[(size_t)(k + 2 * RED_KG) * stride], v3 = src[(size_t)(k + 3 * RED_KG) * stride];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; k < ksplit; k += RED_KG) acc += src[(size_t)k * stride];
  }
  red[g][o] = acc;
  __syncthreads();
  if (g == 0 && idx < total) {
    double t = red[0][o];
#pragma unroll
    for (int j = 1; j < RED_KG; j++) t += red[j][o];
    out[idx] = scale * t;
  }
}

// coefficient block, planar W[16][Npad]: row c*2+s with c=0 rho_bar, 1..3 2*grho_bar_j, 4 2*lapl_bar;
// row 10+s = tau_bar/2 + 2 lapl_bar.  Rows r in [N, Npad) are zero (the last k-tile reads them).
__global__ void bwd_coef_kernel(int64_t N, int64_t Npad, const double* __restrict__ rb, const double* __restrict__ gb,
                                const double* __restrict__ tb, const double* __restrict__ lb, double* __restrict__ W) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Npad) return;
  double w[12];
#pragma unroll
  for (int i = 0; i < 12; i++) w[i] = 0.0;
  if (r < N) {
    for (int s = 0; s < 2; s++) {
      if (rb) w[s] = rb[r * 2 + s];
      if (gb)
        for (int j = 0; j < 3; j++) w[2 * (1 + j) + s] = 2.0 * gb[(r * 2 + s) * 3 + j];
      double k = 0.0;
      if (lb) { w[8 + s] = 2.0 * lb[r * 2 + s]; k += 2.0 * lb[r * 2 + s]; }
      if (tb) k += 0.5 * tb[r * 2 + s];
      w[10 + s] = k;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) W[(size_t)i * Npad + r] = w[i];
}

// HF: W[2 q + s][r] = g[w0 + q][s][r] for q < nw
__global__ void hf_coef_kernel(int64_t N, int64_t Npad, int nw, const double* __restrict__ g_w, double* __restrict__ W) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Npad) return;
  for (int q = 0; q < 2 * nw; q++) W[(size_t)q * Npad + r] = r < N ? g_w[(size_t)q * N + r] : 0.0;
}

struct BwdPlan {
  int mt, tc, base, rem, kmax, stages, ctas, layout;
  int ks_class[4], cta_prefix[5];
  size_t smem;
};

static BwdPlan plan_bwd(int64_t N, int npad, int maxq) {
  BwdPlan pl{};
  const int nsub = npad / 8;
  // Balanced square tiling: `tc` tiles per dimension whose sizes (in 8-wide sub-tiles) differ by at most one;
  // a CTA's time goes as ceil(sa/2)*ceil(sb/2) DMMA per k-step (its busiest warp), the kernel template is the
  // largest ceil(size/2).  eff[] = measured relative DMMA issue efficiency of the MT x MT warp tile (fragment
  // loads and the in-register combine amortise better over larger tiles).
  static const double eff[6] = {0.0, 0.30, 0.53, 0.82, 0.93, 1.00};
  int best_tc = 1, best_mt = 1;
  double best_cost = 1e300;
  for (int tc = (nsub + 2 * BWD_MAX_MT - 1) / (2 * BWD_MAX_MT); tc <= nsub; tc++) {
    const int base = nsub / tc, rem = nsub - base * tc;
    const int smax = base + (rem ? 1 : 0);
    const int mt = (smax + 1) / 2;
    if (mt > BWD_MAX_MT) continue;
    const double per_dim = (double)rem * ((base + 2) / 2) + (double)(tc - rem) * ((base + 1) / 2);
    const double cost = per_dim * per_dim / eff[mt];
    if (cost < best_cost - 1e-9) { best_cost = cost; best_tc = tc; best_mt = mt; }
  }
  if (const char* e = getenv("GDFT_BWD_TILES")) {  // tuning override: tiles per dimension
    int v = atoi(e);
    if (v >= 1 && v <= nsub && (((nsub + v - 1) / v) + 1) / 2 <= BWD_MAX_MT) { best_tc = v; best_mt = (((nsub + v - 1) / v) + 1) / 2; }
  }
  pl.mt = best_mt;
  pl.tc = best_tc;
  pl.base = nsub / best_tc;
  pl.rem = nsub - pl.base * best_tc;
  const int T = 16 * best_mt;

  // Work-balanced split-K.  Class c = 2*(a small) + (b small); its per-k-step work is that of its busiest warp
  // (+1: the fragment loads and ring bookkeeping every k-step pays).  The CTA budget is a whole number of waves of
  // 148 (one resident CTA per SM), as many waves as leave >= 16 k-tiles per split (more, shorter waves: the tail is
  // ~1/waves of the run), shared among the classes in proportion to count x work.
  const int nbig = pl.rem, nsmall = best_tc - pl.rem;
  const int count[4] = {nbig * nbig, nbig * nsmall, nsmall * nbig, nsmall * nsmall};
  const int hb = (pl.base + 2) / 2, hs = (pl.base + 1) / 2;  // ceil(size/2) of a big / small tile
  pl.layout = 1;
  if (const char* e = getenv("GDFT_BWD_LAYOUT")) pl.layout = atoi(e) == 0 ? 0 : 1;
  // DMMAs per k-step on the busiest SM sub-partition: 2 x 4 grid -> 2 * ceil(sa/2) * ceil(sb/2); 1 x 8 grid -> sa * ceil(sb/2)
  const double fa_b = pl.layout ? 0.5 * (pl.base + 1) : hb, fa_s = pl.layout ? 0.5 * pl.base : hs;
  const double work[4] = {fa_b * hb + 1.0, fa_b * hs + 1.0, fa_s * hb + 1.0, fa_s * hs + 1.0};
  double total_work = 0.0;
  int ntiles = 0;
  for (int c = 0; c < 4; c++) { total_work += count[c] * work[c]; ntiles += count[c]; }
  const int64_t ktiles_total = (N + BWD_BKR - 1) / BWD_BKR;
  int waves = 8;
  if (const char* e = getenv("GDFT_BWD_WAVES")) { int v = atoi(e); if (v >= 1 && v <= 16) waves = v; }
  while (waves > 1 && (int64_t)148 * waves * 16 > ktiles_total * ntiles) waves--;
  int budget = 148 * waves;
  if (budget < ntiles) budget = ntiles;
  int ks[4], used = 0;
  double frac[4];
  for (int c = 0; c < 4; c++) {
    ks[c] = 0;
    frac[c] = -1.0;
    if (!count[c]) continue;
    const double ideal = budget * work[c] / total_work;  // K-splits per tile of this class
    ks[c] = (int)ideal;
    if (ks[c] < 1) ks[c] = 1;
    if ((int64_t)ks[c] > ktiles_total) ks[c] = (int)(ktiles_total > 0 ? ktiles_total : 1);
    frac[c] = ideal - ks[c];
    used += ks[c] * count[c];
  }
  for (;;) {  // hand the CTAs left over by the rounding to the classes with the largest fractional parts
    int bestc = -1;
    for (int c = 0; c < 4; c++)
      if (count[c] && frac[c] > 0.0 && used + count[c] <= budget && (int64_t)ks[c] < ktiles_total && (bestc < 0 || frac[c] > frac[bestc])) bestc = c;
    if (bestc < 0) break;
    ks[bestc]++;
    frac[bestc] -= 1.0;
    used += count[bestc];
  }
  pl.kmax = 1;
  pl.cta_prefix[0] = 0;
  for (int c = 0; c < 4; c++) {
    pl.ks_class[c] = ks[c];
    pl.cta_prefix[c + 1] = pl.cta_prefix[c] + ks[c] * count[c];
    if (ks[c] > pl.kmax) pl.kmax = ks[c];
  }
  pl.ctas = pl.cta_prefix[4];
  const size_t stage_bytes = (size_t)(BWD_BKR * (T + 4) * (1 + maxq) + BWD_BKR * BWD_COEF_W) * 8;
  const size_t fixed = 2 * BWD_MAX_STAGES * 8 + 128;
  const size_t smem_budget = 227 * 1024;
  pl.stages = BWD_MAX_STAGES;
  while (pl.stages > 2 && pl.stages * stage_bytes + fixed > smem_budget) pl.stages--;
  pl.smem = pl.stages * stage_bytes + fixed;
  return pl;
}

size_t density_bwd_workspace(int64_t N, int64_t n, int, int) {
  if (N <= 0 || n <= 0) return 0;
  const int npad = (int)npad_of(n);
  BwdPlan pl = plan_bwd(N, npad, BWD_MAX_SLOTS);
  size_t w_bytes = ((size_t)round_up(N, BWD_BKR) * BWD_COEF_W * 8 + 255) & ~size_t(255);
  size_t part_bytes = ((size_t)pl.kmax * 2 * npad * npad * 8 + 255) & ~size_t(255);
  return w_bytes + part_bytes + 512;
}

template <int MT>
static int launch_bwd_t(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmP, const CUtensorMap& tmW,
                        const BwdPlan& pl, const BwdParams& p) {
  GDFT_CUDA_TRY(cudaFuncSetAttribute(density_bwd_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  density_bwd_kernel<MT><<<(unsigned)pl.ctas, BWD_THREADS, pl.smem, stream>>>(tmA, tmP, tmW, p);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// Shared driver: `planes_b` is the tensor the B-side plane tiles come from (the packed basis, or chi_packed).
// The K-split / tile plan does not depend on the terms (so the workspace size does not either); only the
// stage count does (fewer plane slots per stage -> deeper ring).
static int run_bwd(cudaStream_t stream, int64_t N, int n, int nplanes_a, const double* packed, const double* planes_b,
                   int nplanes_b, const double* W, int nterms, const BwdTerm* terms, double scale, double* part, double* out) {
  const int npad = (int)npad_of(n);
  int maxq = 1;
  for (int i = 0; i < nterms; i++) {
    const int need = terms[i].per_spin ? 2 * terms[i].nq : terms[i].nq;
    maxq = need > maxq ? need : maxq;
  }
  BwdPlan pl = plan_bwd(N, npad, maxq);
  const int T = 16 * pl.mt;
  const int64_t Npad = round_up(N, BWD_BKR);
  CUtensorMap tmA, tmP, tmW;
  int rc;
  if ((rc = make_tmap_3d(&tmA, packed, npad, (uint64_t)N, nplanes_a, (uint64_t)npad * 8, (uint64_t)N * npad * 8, T + 4, BWD_BKR))) return rc;
  if ((rc = make_tmap_3d(&tmP, planes_b, npad, (uint64_t)N, nplanes_b, (uint64_t)npad * 8, (uint64_t)N * npad * 8, T + 4, BWD_BKR))) return rc;
  if ((rc = make_tmap_3d(&tmW, W, (uint64_t)Npad, BWD_COEF_W, 1, (uint64_t)Npad * 8, (uint64_t)Npad * BWD_COEF_W * 8, BWD_BKR, BWD_COEF_W)))
    return rc;
  BwdParams p{};
  p.N = N; p.npad = npad; p.nsub = npad / 8; p.nterms = nterms; p.tc = pl.tc; p.base = pl.base; p.rem = pl.rem; p.stages = pl.stages;
  p.maxq = maxq; p.layout = pl.layout;
  for (int c = 0; c < 4; c++) p.ks_class[c] = pl.ks_class[c];
  for (int c = 0; c < 5; c++) p.cta_prefix[c] = pl.cta_prefix[c];
  for (int i = 0; i < nterms; i++) p.terms[i] = terms[i];
  p.part = part;
  switch (pl.mt) {
    case 1: rc = launch_bwd_t<1>(stream, tmA, tmP, tmW, pl, p); break;
    case 2: rc = launch_bwd_t<2>(stream, tmA, tmP, tmW, pl, p); break;
    case 3: rc = launch_bwd_t<3>(stream, tmA, tmP, tmW, pl, p); break;
    case 4: rc = launch_bwd_t<4>(stream, tmA, tmP, tmW, pl, p); break;
    default: rc = launch_bwd_t<5>(stream, tmA, tmP, tmW, pl, p); break;
  }
  if (rc) return rc;
  const int total = 2 * n * n;
  bwd_reduce_kernel<<<(total + RED_OUT - 1) / RED_OUT, RED_OUT * RED_KG, 0, stream>>>(part, make_int4(pl.ks_class[0], pl.ks_class[1], pl.ks_class[2], pl.ks_class[3]),
                                                             pl.rem * (pl.base + 1) * 8, npad, n, scale, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_density_bwd(gdft_stream_t stream_, int64_t N, int64_t n, int flags, int nplanes, const double* packed,
                                const double* rho_bar, const double* grad_rho_bar, const double* tau_bar, const double* lapl_bar,
                                double* rdm1_bar, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || n <= 0 || N > (int64_t)2147483000 || n > 32768) return GDFT_BAD_SHAPE;
  if ((flags & ~(GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL)) || flags == 0) return GDFT_BAD_ARGUMENT;
  if (!packed || !rdm1_bar) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_RHO) && !rho_bar) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_GRAD) && !grad_rho_bar) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_TAU) && !tau_bar) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_LAPL) && !lapl_bar) return GDFT_BAD_ARGUMENT;
  if ((flags & (GDFT_GRAD | GDFT_TAU)) && nplanes < 4) return GDFT_BAD_SHAPE;
  if ((flags & GDFT_LAPL) && nplanes < 5) return GDFT_BAD_SHAPE;
  if (nplanes < 1 || nplanes > 5) return GDFT_BAD_SHAPE;
  if (!aligned16(packed) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < density_bwd_workspace(N, n, 0, 0)) return GDFT_WORKSPACE_TOO_SMALL;

  const int npad = (int)npad_of(n);
  BwdPlan pl = plan_bwd(N, npad, BWD_MAX_SLOTS);  // ksplit does not depend on the slot count
  Workspace wsp(ws, ws_bytes);
  double* W = wsp.take<double>((size_t)round_up(N, BWD_BKR) * BWD_COEF_W);
  double* part = wsp.take<double>((size_t)pl.kmax * 2 * npad * npad);
  if (!W || !part) return GDFT_WORKSPACE_TOO_SMALL;

  const int64_t Npad = round_up(N, BWD_BKR);
  bwd_coef_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, stream>>>(N, Npad, (flags & GDFT_RHO) ? rho_bar : nullptr,
                                                                 (flags & GDFT_GRAD) ? grad_rho_bar : nullptr,
                                                                 (flags & GDFT_TAU) ? tau_bar : nullptr,
                                                                 (flags & GDFT_LAPL) ? lapl_bar : nullptr, W);
  GDFT_LAUNCH_CHECK();

  BwdTerm terms[4];
  int nterms = 0;
  if (flags & (GDFT_RHO | GDFT_GRAD | GDFT_LAPL)) {
    int slots = (flags & GDFT_LAPL) ? 5 : (flags & GDFT_GRAD) ? 4 : 1;
    terms[nterms++] = BwdTerm{0, slots, 0, 0, 0};
  }
  if (flags & (GDFT_TAU | GDFT_LAPL)) {
    for (int j = 1; j <= 3; j++) terms[nterms++] = BwdTerm{j, 1, j, 0, 10};
  }
  return run_bwd(stream, N, (int)n, nplanes, packed, packed, nplanes, W, nterms, terms, 1.0, part, rdm1_bar);
}

extern "C" int gdft_hf_fock(gdft_stream_t stream_, int64_t N, int64_t n, int Wn, int nplanes, const double* packed,
                            const double* chi_packed, const double* g, double* fock, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || n <= 0 || N > (int64_t)2147483000 || n > 32768 || Wn <= 0 || Wn > 8) return GDFT_BAD_SHAPE;
  if (!packed || !chi_packed || !g || !fock) return GDFT_BAD_ARGUMENT;
  if (nplanes < 1 || nplanes > 5) return GDFT_BAD_SHAPE;
  if (!aligned16(packed) || !aligned16(chi_packed) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < density_bwd_workspace(N, n, 0, 0)) return GDFT_WORKSPACE_TOO_SMALL;
  const int npad = (int)npad_of(n);
  BwdPlan pl = plan_bwd(N, npad, BWD_MAX_SLOTS);  // ksplit does not depend on the slot count
  Workspace wsp(ws, ws_bytes);
  double* W = wsp.take<double>((size_t)round_up(N, BWD_BKR) * BWD_COEF_W);
  double* part = wsp.take<double>((size_t)pl.kmax * 2 * npad * npad);
  if (!W || !part) return GDFT_WORKSPACE_TOO_SMALL;
  const int64_t Npad = round_up(N, BWD_BKR);
  for (int w = 0; w < Wn; w++) {
    hf_coef_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, stream>>>(N, Npad, 1, g + (size_t)w * 2 * N, W);
    GDFT_LAUNCH_CHECK();
    BwdTerm term{0, 1, 2 * w, 1, 0};
    int rc = run_bwd(stream, N, (int)n, nplanes, packed, chi_packed, 2 * Wn, W, 1, &term, -0.5, part, fock + (size_t)w * 2 * n * n);
    if (rc) return rc;
  }
  return GDFT_OK;
}

// sum over omega of the HF Fock terms, F[s] = -1/2 sum_w ao^T diag(g[w,s]) chi[w,s], with the sum taken INSIDE the GEMM for
// pairs of omegas (M_s = g[w]*chi[w] + g[w+1]*chi[w+1] is formed in registers like the GGA combine): DM21's two omegas cost
// two GEMM units instead of four.  What dm21_hfgrads_* / b3lyp_hfgrads do with the result of gdft_hf_fock
// (grad_dft/functional.py:714-717, 755-758: vxc_hf.sum(axis=0)).  An odd omega count leaves a single at the end.
extern "C" int gdft_hf_fock_sum(gdft_stream_t stream_, int64_t N, int64_t n, int Wn, int nplanes, const double* packed,
                                const double* chi_packed, const double* g, double* fock_sum, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || n <= 0 || N > (int64_t)2147483000 || n > 32768 || Wn <= 0 || Wn > 2) return GDFT_BAD_SHAPE;
  if (!packed || !chi_packed || !g || !fock_sum) return GDFT_BAD_ARGUMENT;
  if (nplanes < 1 || nplanes > 5) return GDFT_BAD_SHAPE;
  if (!aligned16(packed) || !aligned16(chi_packed) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < density_bwd_workspace(N, n, 0, 0)) return GDFT_WORKSPACE_TOO_SMALL;
  const int npad = (int)npad_of(n);
  BwdPlan pl = plan_bwd(N, npad, BWD_MAX_SLOTS);
  Workspace wsp(ws, ws_bytes);
  double* W = wsp.take<double>((size_t)round_up(N, BWD_BKR) * BWD_COEF_W);
  double* part = wsp.take<double>((size_t)pl.kmax * 2 * npad * npad);
  if (!W || !part) return GDFT_WORKSPACE_TOO_SMALL;
  const int64_t Npad = round_up(N, BWD_BKR);
  hf_coef_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, stream>>>(N, Npad, Wn, g, W);
  GDFT_LAUNCH_CHECK();
  BwdTerm term{0, Wn, 0, 1, 0};
  return run_bwd(stream, N, (int)n, nplanes, packed, chi_packed, 2 * Wn, W, 1, &term, -0.5, part, fock_sum);
}
