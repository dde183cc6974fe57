// K1 -- density family forward:  rho, grad_rho, tau, lapl_rho, e_HF from rdm1 and the packed basis.
//
// Reference semantics (grad_dft/molecule.py:409,440,472-474,502,537-541), with T_s = ao D_s and
// G_sj = (d_j ao) D_s (contraction over the FIRST index of D):
//   rho[r,s]     = sum_b T_s[r,b] ao[r,b]
//   grho[r,s,j]  = 2 sum_b T_s[r,b] d_j ao[r,b]
//   tau[r,s]     = 1/2 sum_j sum_b G_sj[r,b] d_j ao[r,b]
//   lapl[r,s]    = 4 tau[r,s] + 2 sum_b T_s[r,b] lap_ao[r,b]
//   ehf[w,s,r]   = -1/2 sum_c chi[r,w,s,c] T_s[r,c]
//
// One CTA owns 128 grid rows.  For each A-plane pass (ao, then d_x,d_y,d_z ao when tau/lapl are
// requested) and each column tile (the same 8*NTS-wide b-range for BOTH spins) it accumulates the
// 128 x (2*8*NTS) tile of [T_0 | T_1] with FP64 DMMA.8x8x4 from TMA-staged k-tiles (4-stage mbarrier
// ring), then contracts the accumulators against the planes straight from global memory in the
// epilogue; T/G never reach HBM.  Row sums live in shared memory until the CTA retires.
//
// Shared-memory tiles are [row][BK] with BK = 12 doubles: a DMMA A/B fragment load touches
// address row*12 + t (row = lane>>2 (+8..), t = lane&3), and 12*g mod 16 = {0,12,8,4} makes each
// half-warp hit 16 distinct 8-byte banks -- conflict-free without swizzling, and the box (96 B inner
// extent) is a legal dense TMA tile.  D is passed transposed (DT[s][b][a]) so the B operand has the
// same [col][k] shape as A.
#include "common.cuh"
#include <stdlib.h>

namespace gdft {

constexpr int FWD_BK = 12;
constexpr int FWD_STAGES = 4;
constexpr int FWD_THREADS = 256;
constexpr int FWD_SLOT_RHO = 0, FWD_SLOT_GRAD = 2, FWD_SLOT_LAPT = 8, FWD_SLOT_TACC = 10, FWD_SLOT_HF = 12;

struct FwdParams {
  int64_t N;
  int n, npad, nplanes, flags, W;
  int n_ctile, n_ktile, nslots;
  const double* packed;  // [C][N][npad]
  const double* chi;     // [W][2][N][npad]
  double *rho, *grho, *tau, *lapl, *ehf;
};

__global__ void transpose_pad_kernel(const double* __restrict__ D, double* __restrict__ DT, int n, int npad) {
  // DT[s][b][a] = D[s][a][b], zero padded to npad
  __shared__ double tile[32][33];
  int s = blockIdx.z;
  int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int a = a0 + i, b = b0 + threadIdx.x;
    tile[i][threadIdx.x] = (a < n && b < n) ? D[((size_t)s * n + a) * n + b] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int b = b0 + i, a = a0 + threadIdx.x;
    if (b < npad && a < npad) DT[((size_t)s * npad + b) * npad + a] = tile[threadIdx.x][i];
  }
}

// MT = 8-row m-tiles per warp: 2 -> 128 grid rows per CTA, two CTAs per SM; 1 -> 64 rows per CTA, three CTAs per SM (finer
// tiles: the last wave of CTAs wastes less, see use_64_rows).  The per-row arithmetic and its order are the same for both.
template <int NTS, int MT>
__global__ void __launch_bounds__(FWD_THREADS, MT == 2 ? 2 : 3)
density_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
  constexpr int BK = FWD_BK, BM = 64 * MT, STAGES = FWD_STAGES;
  constexpr int BN = 8 * NTS;                       // b-range per spin
  constexpr int STAGE_ELEMS = (BM + 2 * BN) * BK;   // doubles
  constexpr uint32_t STAGE_BYTES = STAGE_ELEMS * 8;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sStage = reinterpret_cast<double*>(smem_raw);
  double* rowacc = sStage + STAGES * STAGE_ELEMS;
  uint64_t* full = reinterpret_cast<uint64_t*>(rowacc + p.nslots * BM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int npad = p.npad;
  const size_t plane_stride = (size_t)p.N * npad;

  for (int i = tid; i < p.nslots * BM; i += FWD_THREADS) rowacc[i] = 0.0;
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const bool pass0 = (p.flags & (GDFT_RHO | GDFT_GRAD | GDFT_LAPL | GDFT_HF)) != 0;
  const bool passT = (p.flags & (GDFT_TAU | GDFT_LAPL)) != 0;
  const int npass = (pass0 ? 1 : 0) + (passT ? 3 : 0);
  const int iters_per_pass = p.n_ctile * p.n_ktile;
  const int total = npass * iters_per_pass;

  auto issue = [&](int it) {
    int pidx = it / iters_per_pass, rem = it - pidx * iters_per_pass;
    int ct = rem / p.n_ktile, kt = rem - ct * p.n_ktile;
    int aplane = pass0 ? pidx : pidx + 1;
    int st = it % STAGES;
    double* sA = sStage + st * STAGE_ELEMS;
    mbar_expect_tx(&full[st], STAGE_BYTES);
    tma_load_3d(sA, &tmA, &full[st], kt * BK, (int)row0, aplane);
    tma_load_3d(sA + BM * BK, &tmB, &full[st], kt * BK, ct * BN, 0);
    tma_load_3d(sA + (BM + BN) * BK, &tmB, &full[st], kt * BK, ct * BN, 1);
  };
  if (tid == 0) {
    for (int it = 0; it < STAGES - 1 && it < total; it++) issue(it);
  }

  int it = 0;
  for (int pidx = 0; pidx < npass; pidx++) {
    const int aplane = pass0 ? pidx : pidx + 1;
    for (int ct = 0; ct < p.n_ctile; ct++) {
      double acc[MT][2 * NTS][2];
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < 2 * NTS; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

      for (int kt = 0; kt < p.n_ktile; kt++, it++) {
        const int st = it % STAGES;
        mbar_wait(&full[st], (it / STAGES) & 1);
        __syncthreads();  // every warp has finished iteration it-1, whose stage is refilled next
        if (tid == 0 && it + STAGES - 1 < total) issue(it + STAGES - 1);
        const double* sA = sStage + st * STAGE_ELEMS + (warp * 8 * MT + g) * BK + t;
        const double* sB = sStage + st * STAGE_ELEMS + BM * BK + g * BK + t;
        const int ksteps = min(BK / 4, (npad - kt * BK) / 4);
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; k4++) {
          if (k4 < ksteps) {
            double a[MT];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) a[mt] = sA[mt * 8 * BK + k4 * 4];
#pragma unroll
            for (int j = 0; j < 2 * NTS; j++) {
              const double b = sB[j * 8 * BK + k4 * 4];
#pragma unroll
              for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][j], a[mt], b);
            }
          }
        }
      }

      // ---- epilogue: contract the T tile against the planes (read once, shared by both spins) ----
      const int bcol0 = ct * BN + 2 * t;
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        const int rl = warp * 8 * MT + mt * 8 + g;
        const int64_t row = row0 + rl;
        const bool rv = row < p.N;
        const double* prow = p.packed + (size_t)row * npad + bcol0;
        auto dot_plane = [&](const double* src, double& s0, double& s1) {
          s0 = 0.0; s1 = 0.0;
#pragma unroll
          for (int j = 0; j < NTS; j++) {
            double2 v = make_double2(0.0, 0.0);
            if (rv && bcol0 + j * 8 < npad) v = __ldg(reinterpret_cast<const double2*>(src + j * 8));
            s0 = fma(acc[mt][j][0], v.x, s0);       s0 = fma(acc[mt][j][1], v.y, s0);
            s1 = fma(acc[mt][NTS + j][0], v.x, s1); s1 = fma(acc[mt][NTS + j][1], v.y, s1);
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        };
        if (pass0 && pidx == 0) {
          double s0, s1;
          if (p.flags & GDFT_RHO) {
            dot_plane(prow, s0, s1);
            if (t == 0) { rowacc[(FWD_SLOT_RHO + 0) * BM + rl] += s0; rowacc[(FWD_SLOT_RHO + 1) * BM + rl] += s1; }
          }
          if (p.flags & GDFT_GRAD) {
            for (int j = 0; j < 3; j++) {
              dot_plane(prow + (size_t)(1 + j) * plane_stride, s0, s1);
              if (t == 0) { rowacc[(FWD_SLOT_GRAD + j) * BM + rl] += s0; rowacc[(FWD_SLOT_GRAD + 3 + j) * BM + rl] += s1; }
            }
          }
          if (p.flags & GDFT_LAPL) {
            dot_plane(prow + (size_t)4 * plane_stride, s0, s1);
            if (t == 0) { rowacc[(FWD_SLOT_LAPT + 0) * BM + rl] += s0; rowacc[(FWD_SLOT_LAPT + 1) * BM + rl] += s1; }
          }
          if (p.flags & GDFT_HF) {
            for (int w = 0; w < p.W; w++) {
              // chi_packed[w][s][row][col]; spin s pairs with T_s only
              const double* c0 = p.chi + ((size_t)(w * 2 + 0) * p.N + row) * npad + bcol0;
              const double* c1 = p.chi + ((size_t)(w * 2 + 1) * p.N + row) * npad + bcol0;
              double u0, u1, dummy;
              dot_plane(c0, u0, dummy);
              dot_plane(c1, dummy, u1);
              if (t == 0) { rowacc[(FWD_SLOT_HF + 2 * w) * BM + rl] += u0; rowacc[(FWD_SLOT_HF + 2 * w + 1) * BM + rl] += u1; }
            }
          }
        } else {
          double s0, s1;
          dot_plane(prow + (size_t)aplane * plane_stride, s0, s1);
          if (t == 0) { rowacc[(FWD_SLOT_TACC + 0) * BM + rl] += s0; rowacc[(FWD_SLOT_TACC + 1) * BM + rl] += s1; }
        }
      }
    }
  }
  __syncthreads();

  // ---- write the rows this CTA owns -------------------------------------------------------------
  if (tid < BM) {
    const int64_t row = row0 + tid;
    if (row < p.N) {
      const double* ra = rowacc + tid;
      if (p.flags & GDFT_RHO) {
        reinterpret_cast<double2*>(p.rho)[row] = make_double2(ra[FWD_SLOT_RHO * BM], ra[(FWD_SLOT_RHO + 1) * BM]);
      }
      if (p.flags & GDFT_GRAD) {
        double* o = p.grho + row * 6;
#pragma unroll
        for (int q = 0; q < 6; q++) o[q] = 2.0 * ra[(FWD_SLOT_GRAD + q) * BM];
      }
      const double t0 = ra[FWD_SLOT_TACC * BM], t1 = ra[(FWD_SLOT_TACC + 1) * BM];
      if (p.flags & GDFT_TAU) reinterpret_cast<double2*>(p.tau)[row] = make_double2(0.5 * t0, 0.5 * t1);
      if (p.flags & GDFT_LAPL) {
        reinterpret_cast<double2*>(p.lapl)[row] =
            make_double2(2.0 * t0 + 2.0 * ra[FWD_SLOT_LAPT * BM], 2.0 * t1 + 2.0 * ra[(FWD_SLOT_LAPT + 1) * BM]);
      }
      if (p.flags & GDFT_HF) {
        for (int w = 0; w < p.W; w++)
          for (int s = 0; s < 2; s++) p.ehf[((size_t)(w * 2 + s)) * p.N + row] = -0.5 * ra[(FWD_SLOT_HF + 2 * w + s) * BM];
      }
    }
  }
}

static int pick_nts(int nsub) {
  if (const char* e = getenv("GDFT_FWD_NTS")) { int v = atoi(e); if (v >= 1 && v <= 5) return v; }  // tuning override
  int best = 1, best_cost = 1 << 30;
  for (int c = 5; c >= 1; c--) {
    int cost = (nsub + c - 1) / c * c;
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

template <int NTS, int MT>
static int launch_fwd_mt(cudaStream_t stream, const double* DT, FwdParams p) {
  constexpr int BM = 64 * MT;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_3d(&tmA, p.packed, p.npad, (uint64_t)p.N, p.nplanes, (uint64_t)p.npad * 8, (uint64_t)p.N * p.npad * 8, FWD_BK, BM);
  if (rc) return rc;
  rc = make_tmap_3d(&tmB, DT, p.npad, p.npad, 2, (uint64_t)p.npad * 8, (uint64_t)p.npad * p.npad * 8, FWD_BK, 8 * NTS);
  if (rc) return rc;
  p.n_ctile = (p.npad / 8 + NTS - 1) / NTS;
  p.n_ktile = (p.npad + FWD_BK - 1) / FWD_BK;
  size_t smem = (size_t)FWD_STAGES * (BM + 16 * NTS) * FWD_BK * 8 + (size_t)p.nslots * BM * 8 + FWD_STAGES * 8;
  GDFT_CUDA_TRY(cudaFuncSetAttribute(density_fwd_kernel<NTS, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned grid = (unsigned)((p.N + BM - 1) / BM);
  density_fwd_kernel<NTS, MT><<<grid, FWD_THREADS, smem, stream>>>(tmA, tmB, p);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// 64-row CTAs (three per SM) for mid-size grids of the tensor-pipe-bound widths.  Measured on B200 (tools/rows_probe.py and
// the C4 bench line): at 62 500 rows x 264 AOs -- the per-GPU share of the benzene grid on 8 GPUs -- 489 128-row tiles fill the
// 296 slots 1.65 times and cost two full waves, 2.74 ms; 64-row tiles take 2.46 ms.  At 250 000 ... 500 000 rows the finer
// tiles are still ahead by 1-3 %; at 2 000 000 x 400 (C4) they lose 2 % (39.6 against 38.8 ms: DRAM traffic drops from 84 to
// 69 GB, but the halved B-fragment reuse costs more than the shorter tail gains).  Narrow matrices (n = 100, 40 000 rows:
// 0.49 -> 0.57 ms) are bound by the per-CTA chain of TMA round trips and epilogue loads, which the smaller tile lengthens.
static bool use_64_rows(int64_t N, int npad) {
  if (const char* e = getenv("GDFT_FWD_ROWS")) { int v = atoi(e); if (v == 64) return true; if (v == 128) return false; }
  const int64_t t128 = (N + 127) / 128;
  return npad >= 192 && t128 > 2 * 148 && t128 <= 32 * 148;
}

template <int NTS>
static int launch_fwd(cudaStream_t stream, const double* DT, FwdParams p) {
  return use_64_rows(p.N, p.npad) ? launch_fwd_mt<NTS, 1>(stream, DT, p) : launch_fwd_mt<NTS, 2>(stream, DT, p);
}

size_t density_fwd_workspace(int64_t n) {
  int64_t np = npad_of(n);
  return (size_t)2 * np * np * 8 + 256;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_density_fwd(gdft_stream_t stream_, int64_t N, int64_t n, int flags, int nplanes, const double* packed,
                                const double* rdm1, const double* chi_packed, int W, double* rho, double* grad_rho,
                                double* tau, double* lapl, double* ehf, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || n <= 0 || N > (int64_t)2147483000 || n > 32768) return GDFT_BAD_SHAPE;
  if ((flags & ~(GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF)) || flags == 0) return GDFT_BAD_ARGUMENT;
  if (!packed || !rdm1) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_RHO) && !rho) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_GRAD) && !grad_rho) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_TAU) && !tau) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_LAPL) && !lapl) return GDFT_BAD_ARGUMENT;
  if ((flags & GDFT_HF) && (!ehf || !chi_packed || W <= 0 || W > 8)) return GDFT_BAD_ARGUMENT;
  if ((flags & (GDFT_GRAD | GDFT_TAU)) && nplanes < 4) return GDFT_BAD_SHAPE;
  if ((flags & GDFT_LAPL) && nplanes < 5) return GDFT_BAD_SHAPE;
  if (nplanes < 1 || nplanes > 5) return GDFT_BAD_SHAPE;
  if (!aligned16(packed) || !aligned16(rho) || !aligned16(tau) || !aligned16(lapl) || !aligned16(chi_packed) || !aligned16(ws))
    return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < density_fwd_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;

  const int npad = (int)npad_of(n);
  Workspace wsp(ws, ws_bytes);
  double* DT = wsp.take<double>((size_t)2 * npad * npad);
  {
    dim3 blk(32, 8), grd((npad + 31) / 32, (npad + 31) / 32, 2);
    transpose_pad_kernel<<<grd, blk, 0, stream>>>(rdm1, DT, (int)n, npad);
    GDFT_LAUNCH_CHECK();
  }
  FwdParams p{};
  p.N = N; p.n = (int)n; p.npad = npad; p.nplanes = nplanes; p.flags = flags; p.W = (flags & GDFT_HF) ? W : 0;
  p.nslots = FWD_SLOT_HF + 2 * p.W;
  p.packed = packed; p.chi = chi_packed;
  p.rho = rho; p.grho = grad_rho; p.tau = tau; p.lapl = lapl; p.ehf = ehf;
  switch (pick_nts(npad / 8)) {
    case 1: return launch_fwd<1>(stream, DT, p);
    case 2: return launch_fwd<2>(stream, DT, p);
    case 3: return launch_fwd<3>(stream, DT, p);
    case 4: return launch_fwd<4>(stream, DT, p);
    default: return launch_fwd<5>(stream, DT, p);
  }
}
