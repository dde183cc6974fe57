// Row f4: the contraction tail of the chi generation,
//   chi[r, s, a] = sum_{b,d} rdm1[s,b,d] ao[r,b] nu[r,d,a]       grad_dft/interface/pyscf.py:1110-1111
// ("...bd,b,da->...a" vmapped over the grid points of one nu chunk, pyscf.py:1116-1119), for one range-separation
// parameter and one chunk of grid points.  nu[r] = <d| v_omega(r, r') |a> is an n x n matrix PER GRID POINT produced by
// libcint (grad_dft/external/_hf_density.py:34-103, out of path); the tail reads it exactly once, so the op is an HBM
// stream of 8 n^2 bytes per point carrying 8 n^2 FLOP (1 FLOP/B: bandwidth-bound by ~5x on B200).
//
// One CTA handles groups of 8 grid points.  Phase 1: T[pt][s][d] = sum_b ao[pt][b] rdm1[s][b][d] for the 8 points at
// once (each thread owns a column pair for 4 points, so every rdm1 element fetched from L2 feeds 4 points x 2 columns);
// T stays in shared memory.  Phase 2: warp w streams nu[r0 + w] row by row with 128-bit streaming loads (lane = column
// pair, rows unrolled by two: ~5-8 KB in flight per warp), both spins accumulated in registers across all n rows, and
// writes chi[r0 + w] itself: no cross-warp reduction, no barrier inside the stream.  With >= 2 CTAs per SM one CTA's
// phase 1 (FP64 pipe) overlaps another's phase 2 (memory pipe).
#include <stdlib.h>
#include "common.cuh"

namespace gdft {

constexpr int CHI_THREADS = 256;
constexpr int CHI_PTS = CHI_THREADS / 32;  // one warp per grid point in the streaming phase
constexpr int CHI_NJ = 8;                  // column pairs per lane per pass: 64 * NJ columns per pass

struct ChiArgs {
  int64_t Nc, ao_ld, chi_ld;
  int n;
  const double* ao;
  const double* rdm1;
  const double* nu;
  double* chi;
};

// VEC: n even and every base pointer 16-byte aligned (128-bit loads/stores); otherwise the scalar layout
// (lane = column, 32 * 2 * NJ columns per pass).
template <int NJ, bool VEC, int ROWS, int MINB>
__global__ void __launch_bounds__(CHI_THREADS, MINB) chi_contract_kernel(ChiArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, n2 = (n + 1) & ~1;
  double* sAo = sm;                          // [n][8]   ao of the group, point index fastest
  double* sT = sAo + (size_t)n * CHI_PTS;    // [8][2][n2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ngroups = (a.Nc + CHI_PTS - 1) / CHI_PTS;
  const int npairs = n2 / 2;

  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t r0 = grp * CHI_PTS;
    // ---- stage ao[r0 .. r0+7][:] transposed (rows past the chunk end are zero) ----
    for (int idx = tid; idx < CHI_PTS * n; idx += CHI_THREADS) {
      const int pt = idx / n, b = idx - pt * n;
      sAo[b * CHI_PTS + pt] = (r0 + pt < a.Nc) ? a.ao[(r0 + pt) * a.ao_ld + b] : 0.0;
    }
    __syncthreads();  // also: every warp has left the previous group's streaming phase, sT may be overwritten

    // ---- phase 1: T for the 8 points; item = (column pair, point quad) ----
    for (int item = tid; item < 2 * npairs; item += CHI_THREADS) {
      const int pr = item >> 1, quad = item & 1;
      const int d0 = 2 * pr;
      const bool has1 = d0 + 1 < n;
      double acc[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
      const double* D0 = a.rdm1 + d0;
      const double* D1 = a.rdm1 + (size_t)n * n + d0;
      const double* aop = sAo + quad * 4;
#pragma unroll 2
      for (int b = 0; b < n; b++) {
        double2 x0, x1;
        if (VEC) {
          x0 = __ldg(reinterpret_cast<const double2*>(D0 + (size_t)b * n));
          x1 = __ldg(reinterpret_cast<const double2*>(D1 + (size_t)b * n));
        } else {
          x0.x = __ldg(D0 + (size_t)b * n); x0.y = has1 ? __ldg(D0 + (size_t)b * n + 1) : 0.0;
          x1.x = __ldg(D1 + (size_t)b * n); x1.y = has1 ? __ldg(D1 + (size_t)b * n + 1) : 0.0;
        }
        const double2 a01 = *reinterpret_cast<const double2*>(aop + b * CHI_PTS);
        const double2 a23 = *reinterpret_cast<const double2*>(aop + b * CHI_PTS + 2);
        const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          acc[i][0][0] = fma(av[i], x0.x, acc[i][0][0]);
          acc[i][0][1] = fma(av[i], x0.y, acc[i][0][1]);
          acc[i][1][0] = fma(av[i], x1.x, acc[i][1][0]);
          acc[i][1][1] = fma(av[i], x1.y, acc[i][1][1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int s = 0; s < 2; s++)
          *reinterpret_cast<double2*>(sT + ((size_t)(quad * 4 + i) * 2 + s) * n2 + d0) = make_double2(acc[i][s][0], acc[i][s][1]);
    }
    __syncthreads();

    // ---- phase 2: warp `warp` streams nu[r0 + warp] ----
    const int64_t r = r0 + warp;
    if (r < a.Nc) {
      const double* nur = a.nu + (size_t)r * n * n;
      const double* t0 = sT + (size_t)(warp * 2) * n2;
      const double* t1 = t0 + n2;
      double* out = a.chi + r * a.chi_ld;
      constexpr int CPP = VEC ? 64 * NJ : 32 * 2 * NJ;  // columns per pass (the same number in both layouts)
      for (int c0 = 0; c0 < n; c0 += CPP) {
        double acc[2][2 * NJ];
#pragma unroll
        for (int j = 0; j < 2 * NJ; j++) acc[0][j] = acc[1][j] = 0.0;
        if (VEC) {
          const int cbase = c0 + 2 * lane;
          auto row = [&](int d, double2 (&v)[NJ]) {
            const double* p = nur + (size_t)d * n + cbase;
#pragma unroll
            for (int j = 0; j < NJ; j++)
              v[j] = (cbase + 64 * j < n) ? __ldcs(reinterpret_cast<const double2*>(p + 64 * j)) : make_double2(0.0, 0.0);
          };
          auto madd = [&](int d, const double2 (&v)[NJ]) {
            const double x0 = t0[d], x1 = t1[d];
#pragma unroll
            for (int j = 0; j < NJ; j++) {
              acc[0][2 * j] = fma(x0, v[j].x, acc[0][2 * j]);
              acc[0][2 * j + 1] = fma(x0, v[j].y, acc[0][2 * j + 1]);
              acc[1][2 * j] = fma(x1, v[j].x, acc[1][2 * j]);
              acc[1][2 * j + 1] = fma(x1, v[j].y, acc[1][2 * j + 1]);
            }
          };
          int d = 0;
          if (ROWS == 2) {  // two rows in flight
            for (; d + 2 <= n; d += 2) {
              double2 va[NJ], vb[NJ];
              row(d, va);
              row(d + 1, vb);
              madd(d, va);
              madd(d + 1, vb);
            }
          }
          for (; d < n; d++) {
            double2 va[NJ];
            row(d, va);
            madd(d, va);
          }
#pragma unroll
          for (int j = 0; j < NJ; j++) {
            const int c = cbase + 64 * j;
            if (c < n) {
              *reinterpret_cast<double2*>(out + c) = make_double2(acc[0][2 * j], acc[0][2 * j + 1]);
              *reinterpret_cast<double2*>(out + n + c) = make_double2(acc[1][2 * j], acc[1][2 * j + 1]);
            }
          }
        } else {
          const int cbase = c0 + lane;
          for (int d = 0; d < n; d++) {
            const double* p = nur + (size_t)d * n + cbase;
            double v[2 * NJ];
#pragma unroll
            for (int j = 0; j < 2 * NJ; j++) v[j] = (cbase + 32 * j < n) ? __ldcs(p + 32 * j) : 0.0;
            const double x0 = t0[d], x1 = t1[d];
#pragma unroll
            for (int j = 0; j < 2 * NJ; j++) {
              acc[0][j] = fma(x0, v[j], acc[0][j]);
              acc[1][j] = fma(x1, v[j], acc[1][j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 2 * NJ; j++) {
            const int c = cbase + 32 * j;
            if (c < n) { out[c] = acc[0][j]; out[n + c] = acc[1][j]; }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-fed variant (n even): one CTA per SM, 8 streaming warps + 5 warps that build T for the NEXT group of 8 points.
// The register-staged kernel above tops out at ~4.8 TB/s: its bytes in flight are bounded by registers (two rows per
// warp, 80 KB per SM) and ncu shows the warps waiting on the long scoreboard.  Here every streaming warp owns a ring of
// shared-memory stages that its lane 0 fills with 1-D bulk copies (cp.async.bulk, completion on a per-stage mbarrier):
// 100-140 KB per SM in flight, no registers tied up, and the ring keeps filling across point and group boundaries
// because nu does not depend on T.  T for group g+1 is computed by the 5 helper warps while group g streams
// (double-buffered, handed over through mbarriers), so the FP64 pipe work never stalls the stream.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CHIT_SW = 8;                          // streaming warps = points per group
constexpr int CHIT_TW = 5;                          // T-builder warps (132 column pairs at n = 264: one pass)
constexpr int CHIT_THREADS = 32 * (CHIT_SW + CHIT_TW);
constexpr int CHIT_MAX_STAGES = 8;
constexpr int CHIT_BAR_BYTES = 1024;

struct ChiTmaPlan {
  int rb, stages, stage_doubles;  // rows per stage, ring depth, doubles per stage slot
  size_t smem;
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int NJ>
__global__ void __launch_bounds__(CHIT_THREADS, 1) chi_contract_tma_kernel(ChiArgs a, ChiTmaPlan pl) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = a.n;
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw);      // [CHIT_SW][CHIT_MAX_STAGES]
  uint64_t* tfull = full + CHIT_SW * CHIT_MAX_STAGES;       // [2]
  uint64_t* tempty = tfull + 2;                             // [2]
  double* sAo = reinterpret_cast<double*>(smraw + CHIT_BAR_BYTES);  // [n][8]
  double* sT = sAo + (size_t)n * CHI_PTS;                   // [2][8][2][n]
  double* ring = sT + (size_t)2 * CHI_PTS * 2 * n;          // [CHIT_SW][stages][stage_doubles]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ngroups = (a.Nc + CHI_PTS - 1) / CHI_PTS;

  if (tid == 0) {
    for (int i = 0; i < CHIT_SW * CHIT_MAX_STAGES; i++) mbar_init(full + i, 1);
    for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 32 * CHIT_TW); mbar_init(tempty + i, CHIT_SW); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp < CHIT_SW) {
    // ================================ streaming warp: one grid point per group ================================
    const int nblk = (n + pl.rb - 1) / pl.rb;
    double* myring = ring + (size_t)warp * pl.stages * pl.stage_doubles;
    uint64_t* myfull = full + warp * CHIT_MAX_STAGES;
    // issue cursor (lane 0): next (group, row block) to request; consume cursor q counts blocks since kernel start
    int64_t iss_g = blockIdx.x;
    int iss_blk = 0, iss_stage = 0;
    auto issue_one = [&]() {
      const int64_t r = iss_g * CHI_PTS + warp;
      if (iss_g >= ngroups || r >= a.Nc) return;
      const int rows = min(pl.rb, n - iss_blk * pl.rb);
      const uint32_t bytes = (uint32_t)rows * n * 8;
      mbar_expect_tx(myfull + iss_stage, bytes);
      bulk_load_1d(myring + (size_t)iss_stage * pl.stage_doubles, a.nu + (size_t)r * n * n + (size_t)iss_blk * pl.rb * n, bytes, myfull + iss_stage);
      iss_stage = iss_stage + 1 == pl.stages ? 0 : iss_stage + 1;
      if (++iss_blk == nblk) { iss_blk = 0; iss_g += gridDim.x; }
    };
    if (lane == 0)
      for (int i = 0; i < pl.stages; i++) issue_one();
    int stage = 0;
    uint32_t phase = 0;
    int gi = 0;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
      const int buf = gi & 1;
      mbar_wait(tfull + buf, (gi >> 1) & 1);
      const int64_t r = grp * CHI_PTS + warp;
      if (r < a.Nc) {
        const double* t0 = sT + ((size_t)(buf * CHI_PTS + warp) * 2) * n;
        const double* t1 = t0 + n;
        double acc[2][2 * NJ];
#pragma unroll
        for (int j = 0; j < 2 * NJ; j++) acc[0][j] = acc[1][j] = 0.0;
        for (int blk = 0; blk < nblk; blk++) {
          mbar_wait(myfull + stage, phase);
          const double* src = myring + (size_t)stage * pl.stage_doubles + 2 * lane;
          const int rows = min(pl.rb, n - blk * pl.rb);
          for (int rr = 0; rr < rows; rr++) {
            const int d = blk * pl.rb + rr;
            const double x0 = t0[d], x1 = t1[d];
            const double* p = src + (size_t)rr * n;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
              if (2 * lane + 64 * j < n) {
                const double2 v = *reinterpret_cast<const double2*>(p + 64 * j);
                acc[0][2 * j] = fma(x0, v.x, acc[0][2 * j]);
                acc[0][2 * j + 1] = fma(x0, v.y, acc[0][2 * j + 1]);
                acc[1][2 * j] = fma(x1, v.x, acc[1][2 * j]);
                acc[1][2 * j + 1] = fma(x1, v.y, acc[1][2 * j + 1]);
              }
            }
          }
          __syncwarp();  // every lane has read the stage: lane 0 may hand it back to the copy engine
          if (lane == 0) issue_one();
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
        }
        double* out = a.chi + r * a.chi_ld;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const int c = 2 * lane + 64 * j;
          if (c < n) {
            *reinterpret_cast<double2*>(out + c) = make_double2(acc[0][2 * j], acc[0][2 * j + 1]);
            *reinterpret_cast<double2*>(out + n + c) = make_double2(acc[1][2 * j], acc[1][2 * j + 1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + buf);  // this warp no longer reads sT[buf]
    }
  } else {
    // ================================ T builders: sT[buf] for each group, one group ahead ======================
    const int tt = tid - 32 * CHIT_SW;
    constexpr int TT = 32 * CHIT_TW;
    const int npairs = n / 2;
    int gi = 0;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
      const int buf = gi & 1;
      const int64_t r0 = grp * CHI_PTS;
      named_bar_sync(1, TT);  // the previous group's items no longer read sAo
      for (int idx = tt; idx < CHI_PTS * n; idx += TT) {
        const int pt = idx / n, b = idx - pt * n;
        sAo[b * CHI_PTS + pt] = (r0 + pt < a.Nc) ? a.ao[(r0 + pt) * a.ao_ld + b] : 0.0;
      }
      mbar_wait(tempty + buf, ((gi >> 1) & 1) ^ 1);  // streamers are done with the previous use of this buffer
      named_bar_sync(1, TT);
      double* Tb = sT + (size_t)buf * CHI_PTS * 2 * n;
      // item = one column pair for all 8 points: every rdm1 element fetched from L2 feeds 8 points x 2 spins, and four
      // b-steps of loads are issued ahead of their FMAs (the loop is L2-latency-bound otherwise)
      for (int pr = tt; pr < npairs; pr += TT) {
        const int d0 = 2 * pr;
        double acc[CHI_PTS][2][2];
#pragma unroll
        for (int i = 0; i < CHI_PTS; i++) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
        const double* D0 = a.rdm1 + d0;
        const double* D1 = a.rdm1 + (size_t)n * n + d0;
        auto step = [&](int b, const double2& x0, const double2& x1) {
          const double2* ap = reinterpret_cast<const double2*>(sAo + b * CHI_PTS);
#pragma unroll
          for (int h = 0; h < CHI_PTS / 2; h++) {
            const double2 av = ap[h];
            acc[2 * h][0][0] = fma(av.x, x0.x, acc[2 * h][0][0]);
            acc[2 * h][0][1] = fma(av.x, x0.y, acc[2 * h][0][1]);
            acc[2 * h][1][0] = fma(av.x, x1.x, acc[2 * h][1][0]);
            acc[2 * h][1][1] = fma(av.x, x1.y, acc[2 * h][1][1]);
            acc[2 * h + 1][0][0] = fma(av.y, x0.x, acc[2 * h + 1][0][0]);
            acc[2 * h + 1][0][1] = fma(av.y, x0.y, acc[2 * h + 1][0][1]);
            acc[2 * h + 1][1][0] = fma(av.y, x1.x, acc[2 * h + 1][1][0]);
            acc[2 * h + 1][1][1] = fma(av.y, x1.y, acc[2 * h + 1][1][1]);
          }
        };
        int b = 0;
        for (; b + 4 <= n; b += 4) {
          double2 x0[4], x1[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            x0[u] = __ldg(reinterpret_cast<const double2*>(D0 + (size_t)(b + u) * n));
            x1[u] = __ldg(reinterpret_cast<const double2*>(D1 + (size_t)(b + u) * n));
          }
#pragma unroll
          for (int u = 0; u < 4; u++) step(b + u, x0[u], x1[u]);
        }
        for (; b < n; b++)
          step(b, __ldg(reinterpret_cast<const double2*>(D0 + (size_t)b * n)), __ldg(reinterpret_cast<const double2*>(D1 + (size_t)b * n)));
#pragma unroll
        for (int i = 0; i < CHI_PTS; i++)
#pragma unroll
          for (int sp = 0; sp < 2; sp++)
            *reinterpret_cast<double2*>(Tb + ((size_t)i * 2 + sp) * n + d0) = make_double2(acc[i][sp][0], acc[i][sp][1]);
      }
      mbar_arrive(tfull + buf);
    }
  }
}

// ring geometry for the TMA variant; stages == 0: does not fit (the register-staged kernel is used)
static ChiTmaPlan chi_tma_plan(int n) {
  ChiTmaPlan pl{};
  const size_t fixed = CHIT_BAR_BYTES + ((size_t)n * CHI_PTS + (size_t)2 * CHI_PTS * 2 * n) * 8;
  const size_t budget = (size_t)227 * 1024;
  if (fixed + (size_t)CHIT_SW * 2 * n * 8 > budget) return pl;
  const size_t per_warp = ((budget - fixed) / CHIT_SW) & ~size_t(127);
  const size_t row = (size_t)n * 8;
  pl.rb = (int)imax64(1, (int64_t)((2048 + row - 1) / row));     // stages of >= 2 KB
  while (pl.rb > 1 && per_warp / (pl.rb * row) < 2) pl.rb--;
  const size_t slot = (pl.rb * row + 127) & ~size_t(127);
  pl.stage_doubles = (int)(slot / 8);
  pl.stages = (int)imin64(CHIT_MAX_STAGES, (int64_t)(per_warp / slot));
  if (pl.stages < 2) { pl.stages = 0; return pl; }
  pl.smem = fixed + (size_t)CHIT_SW * pl.stages * slot;
  return pl;
}

template <int NJ>
static int launch_chi_tma(cudaStream_t stream, const ChiArgs& a, const ChiTmaPlan& pl, int ctas) {
  GDFT_CUDA_TRY(cudaFuncSetAttribute(chi_contract_tma_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  chi_contract_tma_kernel<NJ><<<ctas, CHIT_THREADS, pl.smem, stream>>>(a, pl);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

static size_t chi_smem(int n) {
  const int n2 = (n + 1) & ~1;
  return ((size_t)n * CHI_PTS + (size_t)CHI_PTS * 2 * n2) * 8;
}

template <int NJ, bool VEC, int ROWS, int MINB>
static int launch_chi_v(cudaStream_t stream, const ChiArgs& a, int ctas, size_t smem) {
  GDFT_CUDA_TRY(cudaFuncSetAttribute(chi_contract_kernel<NJ, VEC, ROWS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chi_contract_kernel<NJ, VEC, ROWS, MINB><<<ctas, CHI_THREADS, smem, stream>>>(a);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// occupancy variants: `per_sm` CTAs per SM (shared memory permitting).  Three resident CTAs leave 85 registers per
// thread, enough for one row in flight; two leave 128 (two rows in flight up to NJ = 5).
template <int NJ>
static int launch_chi(cudaStream_t stream, const ChiArgs& a, bool vec, int per_sm, int ctas, size_t smem) {
  if (!vec) return launch_chi_v<NJ, false, 1, 2>(stream, a, ctas, smem);
  if (per_sm >= 3) return launch_chi_v<NJ, true, 1, 3>(stream, a, ctas, smem);
  if (NJ <= 5) return launch_chi_v<NJ, true, 2, 2>(stream, a, ctas, smem);
  return launch_chi_v<NJ, true, 1, 2>(stream, a, ctas, smem);
}

}  // namespace gdft

using namespace gdft;

extern "C" int64_t gdft_chi_contract_max_n(void) { return 1152; }  // 24 n doubles of shared memory per CTA

extern "C" int gdft_chi_contract(gdft_stream_t stream_, int64_t Nc, int64_t n, const double* ao, int64_t ao_ld, const double* rdm1,
                                 const double* nu, double* chi, int64_t chi_ld) {
  if (Nc <= 0 || n <= 0 || n > gdft_chi_contract_max_n() || Nc > (int64_t)2147483000) return GDFT_BAD_SHAPE;
  if (ao_ld < n || chi_ld < 2 * n) return GDFT_BAD_SHAPE;
  if (!ao || !rdm1 || !nu || !chi) return GDFT_BAD_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(ao) | reinterpret_cast<uintptr_t>(rdm1) | reinterpret_cast<uintptr_t>(nu) |
       reinterpret_cast<uintptr_t>(chi)) & 7)
    return GDFT_BAD_ALIGNMENT;
  const bool vec = (n % 2 == 0) && (chi_ld % 2 == 0) && aligned16(rdm1) && aligned16(nu) && aligned16(chi);
  ChiArgs a{Nc, ao_ld, chi_ld, (int)n, ao, rdm1, nu, chi};
  const size_t smem = chi_smem((int)n);
  int dev = 0, sms = 148;
  GDFT_CUDA_TRY(cudaGetDevice(&dev));
  GDFT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // measured on B200 (tools/chi_probe.py): 2 CTAs/SM with two rows in flight beat 3 CTAs/SM with one (4.8 vs 4.7 TB/s at
  // n = 264); an L2 prefetch ahead of the streaming loads LOWERED the rate (3.3 TB/s) and was removed
  int per_sm = smem * 2 <= (size_t)220 * 1024 ? 2 : 1;
  if (const char* e = getenv("GDFT_CHI_PER_SM")) {  // tuning override
    const int v = atoi(e);
    if (v >= 1 && v <= 3 && smem * v <= (size_t)216 * 1024) per_sm = v;
  }
  const int64_t ngroups = (Nc + CHI_PTS - 1) / CHI_PTS;
  const int ctas = (int)imin64(ngroups, (int64_t)sms * per_sm);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // The TMA-fed kernel where it wins (measured on B200, tools/chi_probe.py, GB/s register-staged -> TMA-fed):
  // n = 264: 4830 -> 5670; n = 400: 4920 -> 4290 (ring only 4 stages deep next to the two T buffers); n = 128: 4740 -> 3140
  // and n = 44: 3050 -> 2290 (few column pairs: the T builders' serial b-loop outlasts the stream); n = 512: 4870 -> 4790.
  // So: n even, a ring of >= 6 stages, and 96..160 column pairs (one full pass of the builder warps).
  // GDFT_CHI_TMA=0 / 1 forces the register-staged / TMA-fed kernel (tuning).
  const char* tma_env = getenv("GDFT_CHI_TMA");
  if (vec && n <= 64 * CHI_NJ && !(tma_env && tma_env[0] == '0')) {
    const ChiTmaPlan pl = chi_tma_plan((int)n);
    const bool forced = tma_env && tma_env[0] == '1';
    if (pl.stages >= 2 && (forced || (pl.stages >= 6 && n / 2 >= 96 && n / 2 <= 32 * CHIT_TW))) {
      const int tctas = (int)imin64(ngroups, (int64_t)sms);
      switch ((int)((n + 63) / 64)) {
        case 1: return launch_chi_tma<1>(stream, a, pl, tctas);
        case 2: return launch_chi_tma<2>(stream, a, pl, tctas);
        case 3: return launch_chi_tma<3>(stream, a, pl, tctas);
        case 4: return launch_chi_tma<4>(stream, a, pl, tctas);
        case 5: return launch_chi_tma<5>(stream, a, pl, tctas);
        case 6: return launch_chi_tma<6>(stream, a, pl, tctas);
        case 7: return launch_chi_tma<7>(stream, a, pl, tctas);
        default: return launch_chi_tma<8>(stream, a, pl, tctas);
      }
    }
  }
  // columns per pass: one pass up to 512 columns; beyond that the fewest passes of equal width (a narrow last pass would
  // re-walk all n rows for a few lanes' worth of columns)
  const int npass = (int)((n + 64 * CHI_NJ - 1) / (64 * CHI_NJ));
  const int nj = (int)(((n + npass - 1) / npass + 63) / 64);
  switch (nj) {
    case 1: return launch_chi<1>(stream, a, vec, per_sm, ctas, smem);
    case 2: return launch_chi<2>(stream, a, vec, per_sm, ctas, smem);
    case 3: return launch_chi<3>(stream, a, vec, per_sm, ctas, smem);
    case 4: return launch_chi<4>(stream, a, vec, per_sm, ctas, smem);
    case 5: return launch_chi<5>(stream, a, vec, per_sm, ctas, smem);
    case 6: return launch_chi<6>(stream, a, vec, per_sm, ctas, smem);
    case 7: return launch_chi<7>(stream, a, vec, per_sm, ctas, smem);
    default: return launch_chi<8>(stream, a, vec, per_sm, ctas, smem);
  }
}
