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

namespace gdft {

constexpr int BWD_BKR = 16;
constexpr int BWD_MMA_WARPS = 8;
constexpr int BWD_THREADS = 32 * (BWD_MMA_WARPS + 1);
constexpr int BWD_MAX_SLOTS = 5;
constexpr int BWD_COEF_W = 16;  // coefficient rows per grid row (planar: W[coef][Npad])
constexpr int BWD_MAX_STAGES = 6;

struct BwdTerm {
  int a_plane, nq, slot_plane0, per_spin, coef_row0;
};
struct BwdParams {
  int64_t N, rows_per_split;
  int npad, nterms, tiles_b, stages, maxq;
  BwdTerm terms[4];
  double* part;  // [ksplit][2][npad][npad]
};

template <int MT, int NQ>
__device__ __forceinline__ void bwd_stage_mma(double (&acc)[MT][MT][2], const double* __restrict__ sA, const double* __restrict__ sP,
                                              const double* __restrict__ sC) {
  constexpr int PITCH = 16 * MT + 4, SLOT_ELEMS = BWD_BKR * PITCH;
#pragma unroll
  for (int k4 = 0; k4 < BWD_BKR / 4; k4++) {
    double c[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) c[q] = sC[2 * q * BWD_BKR + k4 * 4];
    double b[MT];
#pragma unroll
    for (int j = 0; j < MT; j++) {
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < NQ; q++) v = fma(c[q], sP[q * SLOT_ELEMS + k4 * 4 * PITCH + j * 8], v);
      b[j] = v;
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const double a = sA[k4 * 4 * PITCH + i * 8];
#pragma unroll
      for (int j = 0; j < MT; j++) dmma884(acc[i][j], a, b[j]);
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
  const int ta = blockIdx.x / p.tiles_b, tb = blockIdx.x - ta * p.tiles_b;
  const int a0 = ta * T, b0 = tb * T;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(p.N, r_begin + p.rows_per_split);
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

  if (warp == BWD_MMA_WARPS) {
    // ---- TMA producer ---------------------------------------------------------------------------
    if (lane == 0) {
      for (int it = 0; it < total; it++) {
        const int term = it / ktiles, kt = it - term * ktiles;
        const BwdTerm Tm = p.terms[term];
        const int r = (int)(r_begin + (int64_t)kt * BKR);
        const int st = it % S;
        if (it >= S) mbar_wait(&empty[st], ((it / S) - 1) & 1);
        double* sA = sStage + (size_t)st * stage_elems;
        const int nload = Tm.per_spin ? 2 : Tm.nq;
        mbar_expect_tx(&full[st], (uint32_t)(A_ELEMS + nload * SLOT_ELEMS + COEF_ELEMS) * 8u);
        tma_load_3d(sA, &tmA, &full[st], a0, r, Tm.a_plane);
        for (int q = 0; q < nload; q++) tma_load_3d(sA + A_ELEMS + q * SLOT_ELEMS, &tmP, &full[st], b0, r, Tm.slot_plane0 + q);
        tma_load_3d(sA + A_ELEMS + p.maxq * SLOT_ELEMS, &tmW, &full[st], r, 0, 0);
      }
    }
    return;
  }

  // ---- MMA consumers ------------------------------------------------------------------------------
  const int wm = warp & 1, wn = warp >> 1, spin = wn >> 1, nhalf = wn & 1;
  double acc[MT][MT][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < MT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < total; it++) {
    const int st = it % S;
    const BwdTerm Tm = p.terms[it / ktiles];
    const double* stage = sStage + (size_t)st * stage_elems;
    mbar_wait(&full[st], (it / S) & 1);
    const double* sA = stage + t * PITCH + wm * 8 * MT + g;
    const double* sP = stage + A_ELEMS + (Tm.per_spin ? spin * SLOT_ELEMS : 0) + t * PITCH + nhalf * 8 * MT + g;
    const double* sC = stage + A_ELEMS + p.maxq * SLOT_ELEMS + (Tm.coef_row0 + spin) * BKR + t;
    if (Tm.nq == 1) bwd_stage_mma<MT, 1>(acc, sA, sP, sC);
    else if (Tm.nq == 4) bwd_stage_mma<MT, 4>(acc, sA, sP, sC);
    else bwd_stage_mma<MT, 5>(acc, sA, sP, sC);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

  // ---- partial tile -> workspace -----------------------------------------------------------------
  double* out = p.part + ((size_t)blockIdx.y * 2 + spin) * p.npad * p.npad;
#pragma unroll
  for (int i = 0; i < MT; i++) {
    const int a = a0 + wm * 8 * MT + i * 8 + g;
#pragma unroll
    for (int j = 0; j < MT; j++) {
      const int b = b0 + nhalf * 8 * MT + j * 8 + 2 * t;
      if (a < p.npad && b < p.npad) *reinterpret_cast<double2*>(out + (size_t)a * p.npad + b) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// out[s][a][b] = scale * sum_ks part[ks][s][a][b]   (fixed summation order)
__global__ void bwd_reduce_kernel(const double* __restrict__ part, int ksplit, int npad, int n, double scale,
                                  double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 2 * n * n;
  if (idx >= total) return;
  const int s = idx / (n * n), rem = idx - s * n * n, a = rem / n, b = rem - a * n;
  const size_t off = ((size_t)s * npad + a) * npad + b;
  const size_t stride = (size_t)2 * npad * npad;
  double acc = 0.0;
  for (int k = 0; k < ksplit; k++) acc += part[k * stride + off];
  out[idx] = scale * acc;
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

// HF: W[s][r] = g[w][s][r]
__global__ void hf_coef_kernel(int64_t N, int64_t Npad, const double* __restrict__ g_w, double* __restrict__ W) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Npad) return;
  W[r] = r < N ? g_w[r] : 0.0;
  W[Npad + r] = r < N ? g_w[N + r] : 0.0;
}

struct BwdPlan {
  int mt, tiles_a, tiles_b, ksplit, stages;
  int64_t rows_per_split;
  size_t smem;
};

static BwdPlan plan_bwd(int64_t N, int npad, int maxq) {
  BwdPlan pl{};
  const int nsub = npad / 8;
  // square CTA tiles 16*c per dimension, c from {5,4,3,2,1}: least padded area, mild penalty for small
  // warp tiles (fewer DMMA per fragment load)
  int best = 1;
  double best_cost = 1e300;
  for (int c = 5; c >= 1; c--) {
    const int tiles = (nsub + 2 * c - 1) / (2 * c);
    const double cost = (double)(tiles * 2 * c) * (tiles * 2 * c) * (1.0 + 0.06 * (5 - c));
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  pl.mt = best;
  const int T = 16 * best;
  pl.tiles_a = pl.tiles_b = (npad + T - 1) / T;
  const int ntiles = pl.tiles_a * pl.tiles_b;
  const int slots = 148;  // one resident CTA per SM
  const int64_t ktiles_total = (N + BWD_BKR - 1) / BWD_BKR;
  // K-splits: fill whole waves of 148 CTAs; prefer more, shorter waves (tail effect ~ 1/waves) while each
  // split still streams >= 64 k-tiles
  int best_ks = 1;
  double best_eff = -1.0;
  for (int ks = 1; ks <= 1024; ks++) {
    if ((int64_t)ks * 64 > ktiles_total && ks > 1) break;
    const int64_t ctas = (int64_t)ks * ntiles;
    if (ctas > 4096) break;
    const double waves = (double)((ctas + slots - 1) / slots);
    const double eff = (double)ctas / (waves * slots) - 0.0005 * ks;
    if (eff > best_eff) { best_eff = eff; best_ks = ks; }
  }
  pl.ksplit = best_ks;
  pl.rows_per_split = round_up((N + best_ks - 1) / best_ks, BWD_BKR);
  const size_t stage_bytes = (size_t)(BWD_BKR * (T + 4) * (1 + maxq) + BWD_BKR * BWD_COEF_W) * 8;
  const size_t fixed = 2 * BWD_MAX_STAGES * 8 + 128;
  const size_t budget = 227 * 1024;
  pl.stages = BWD_MAX_STAGES;
  while (pl.stages > 2 && pl.stages * stage_bytes + fixed > budget) pl.stages--;
  pl.smem = pl.stages * stage_bytes + fixed;
  return pl;
}

size_t density_bwd_workspace(int64_t N, int64_t n, int, int) {
  if (N <= 0 || n <= 0) return 0;
  const int npad = (int)npad_of(n);
  BwdPlan pl = plan_bwd(N, npad, BWD_MAX_SLOTS);
  size_t w_bytes = ((size_t)round_up(N, BWD_BKR) * BWD_COEF_W * 8 + 255) & ~size_t(255);
  size_t part_bytes = ((size_t)pl.ksplit * 2 * npad * npad * 8 + 255) & ~size_t(255);
  return w_bytes + part_bytes + 512;
}

template <int MT>
static int launch_bwd_t(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmP, const CUtensorMap& tmW,
                        const BwdPlan& pl, const BwdParams& p) {
  GDFT_CUDA_TRY(cudaFuncSetAttribute(density_bwd_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid(pl.tiles_a * pl.tiles_b, pl.ksplit);
  density_bwd_kernel<MT><<<grid, BWD_THREADS, pl.smem, stream>>>(tmA, tmP, tmW, p);
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
  for (int i = 0; i < nterms; i++) maxq = terms[i].per_spin ? (maxq > 2 ? maxq : 2) : (terms[i].nq > maxq ? terms[i].nq : maxq);
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
  p.N = N; p.rows_per_split = pl.rows_per_split; p.npad = npad; p.nterms = nterms; p.tiles_b = pl.tiles_b; p.stages = pl.stages;
  p.maxq = maxq;
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
  bwd_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(part, pl.ksplit, npad, n, scale, out);
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
  double* part = wsp.take<double>((size_t)pl.ksplit * 2 * npad * npad);
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
  double* part = wsp.take<double>((size_t)pl.ksplit * 2 * npad * npad);
  if (!W || !part) return GDFT_WORKSPACE_TOO_SMALL;
  const int64_t Npad = round_up(N, BWD_BKR);
  for (int w = 0; w < Wn; w++) {
    hf_coef_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, stream>>>(N, Npad, g + (size_t)w * 2 * N, W);
    GDFT_LAUNCH_CHECK();
    BwdTerm term{0, 1, 2 * w, 1, 0};
    int rc = run_bwd(stream, N, (int)n, nplanes, packed, chi_packed, 2 * Wn, W, 1, &term, -0.5, part, fock + (size_t)w * 2 * n * n);
    if (rc) return rc;
  }
  return GDFT_OK;
}
