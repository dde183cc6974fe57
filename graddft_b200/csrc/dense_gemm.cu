// Row f2: the Dense layers of the coefficient network (DM21's default_nn, grad_dft/functional.py:
// 793-822: 11 -> 256, 6 x (Dense + residual -> LayerNorm -> ELU), -> 3) as FP64 tensor-core GEMMs of this library, with
// the layer's elementwise tail in the GEMM epilogue, forward and reverse.  [N, 256] x [256, 256] is the shape family of
// ao . D (K1): a tall-skinny left operand streamed once, a small right operand that lives in L2.
//
//   dense_nn_kernel<NJ, EPI>   out[N, Wd] = A[N, K] B[K, Wd] (+ epilogue).  CTA = 32 rows x all Wd <= 256 columns (the
//     whole output row is in one CTA, which is what LayerNorm needs); 8 warps, warp w owns columns [32w, 32w + 32) as
//     4 x NJ DMMA.8x8x4 tiles; k-tiles of 12 doubles staged by TMA into a 4-stage mbarrier ring (row pitch 12: the
//     conflict-free fragment layout of K1; B is passed transposed, Bt[Wd][K], so that it has the A layout); 2 CTAs per SM.
//     Epilogues:
//       PLAIN        out = acc (+ bias) (+ res)                                              (x_bar of the first block)
//       LN_ELU_FWD   z = acc + bias + res;  xhat = (z - mean z) rsqrt(var z + eps);  out = elu(xhat gamma + beta);
//                    writes out, xhat, rstd                                                  (functional.py:809-819)
//       LN_ELU_BWD   ob = acc + res is the cotangent of the PREVIOUS block's output (res = this block's z_bar: the
//                    residual branch); the previous block's ELU/LayerNorm are undone right here:
//                    t = ob elu'(out_prev), gh = t gamma, z_bar_prev = rstd (gh - mean gh - xhat mean(gh xhat));
//                    writes z_bar_prev and this CTA's column sums of (t xhat, t, z_bar_prev) = partial cotangents of
//                    (LayerNorm scale, LayerNorm bias, Dense bias) -- so the reverse pass of the trunk is ONE GEMM kernel
//                    per block and no elementwise pass ever reads or writes an [N, 256] tensor on its own.
//   dense_tn_kernel            Wbar[K, Wd] = A[N, K]^T Z[N, Wd]: split-K over the rows (one K-slice and one 128 x 128 output
//     tile per CTA, 148 CTAs), operands as [16 rows][132] tiles (pitch = 4 mod 16: conflict-free transposed fragments, the
//     4 padding columns fetched by the same TMA box), fixed-order second-stage reduce (bitwise reproducible).
#include <stdlib.h>
#include "common.cuh"

namespace gdft {

constexpr int DG_BM = 32, DG_BK = 12, DG_STAGES = 4, DG_THREADS = 256;
enum { DG_PLAIN = 0, DG_LN_ELU_FWD = 1, DG_LN_ELU_BWD = 2 };

struct DenseParams {
  int64_t N;
  int K, Wd, n_ktile;
  double eps;
  const double *bias, *res, *gamma, *beta, *out_prev, *xhat_prev, *rstd_prev;
  double *out, *xhat, *rstd, *colpart;
};

// elu(o) = o > 0 ? o : expm1(o).  The library expm1 is ~45 FP64 instructions on the pipe the DMMAs of the co-resident CTA
// use (as a separate pass, K7's forward sat at 36 % of the FP64 pipe for it); for o <= 0 this one is 19: o = k ln2 + r with
// |r| <= ln2 / 2, exp(r) = 1 + r q(r) (Taylor to r^13 / 13!: remainder 0.3466^14 / 14! = 4e-18), s = 2^k, and
// expm1(o) = (s - 1) + (s r) q(r) -- exact in the leading term, full relative accuracy near 0 (k = 0: s - 1 = 0), no branch.
__device__ __forceinline__ double dg_expm1_neg(double o) {
  o = fmax(o, -64.0);  // exp(-64) = 1.6e-28: expm1 = -1 to the last bit from -37.4 down
  const double kf = rint(o * 1.4426950408889634);
  double r = fma(kf, -6.93147180369123816490e-01, o);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double q = 1.6059043836821613e-10;  // 1/13!
  q = fma(q, r, 2.08767569878681e-09);
  q = fma(q, r, 2.505210838544172e-08);
  q = fma(q, r, 2.755731922398589e-07);
  q = fma(q, r, 2.7557319223985893e-06);
  q = fma(q, r, 2.48015873015873e-05);
  q = fma(q, r, 1.984126984126984e-04);
  q = fma(q, r, 1.388888888888889e-03);
  q = fma(q, r, 8.333333333333333e-03);
  q = fma(q, r, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double sc = __longlong_as_double((long long)((int)kf + 1023) << 52);
  return fma(sc * r, q, sc - 1.0);
}
__device__ __forceinline__ double dg_elu(double o) { return o > 0.0 ? o : dg_expm1_neg(o); }
__device__ __forceinline__ void dg_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NJ, int EPI>
__global__ void __launch_bounds__(DG_THREADS, 2)
dense_nn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DenseParams p) {
  constexpr int BM = DG_BM, BK = DG_BK, STAGES = DG_STAGES, MT = BM / 8;
  constexpr int BN = 64 * NJ;  // columns staged per k-tile (8 warps x NJ tiles x 8)
  constexpr int STAGE_ELEMS = (BM + BN) * BK;
  constexpr uint32_t STAGE_BYTES = STAGE_ELEMS * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sStage = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + STAGES * STAGE_ELEMS);
  double* red = sStage;  // [2][BM][8] row partials, aliasing the ring once the main loop is over

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int total = p.n_ktile;
  auto issue = [&](int it) {
    const int st = it % STAGES;
    double* sA = sStage + st * STAGE_ELEMS;
    mbar_expect_tx(&full[st], STAGE_BYTES);
    tma_load_3d(sA, &tmA, &full[st], it * BK, (int)row0, 0);
    tma_load_3d(sA + BM * BK, &tmB, &full[st], it * BK, 0, 0);
  };
  if (tid == 0)
    for (int it = 0; it < STAGES - 1 && it < total; it++) issue(it);
  if (EPI == DG_LN_ELU_BWD) {
    // the epilogue reads this CTA's rows of xhat_prev and out_prev (64 KB each at W = 256) straight from HBM: start them
    // towards L2 now, under the main loop (issued at the point of use they cost the reverse GEMM 1.6 ms on top of 1.95)
    const int lines = (int)(((size_t)BM * p.Wd * 8) / 128);
    const int64_t nrows = p.N - row0 < BM ? p.N - row0 : BM;
    const int live = (int)(((size_t)nrows * p.Wd * 8) / 128);
    const char* a = reinterpret_cast<const char*>(p.xhat_prev + row0 * p.Wd);
    const char* b = reinterpret_cast<const char*>(p.out_prev + row0 * p.Wd);
    for (int l = tid; l < lines && l < live; l += DG_THREADS) { dg_prefetch_l2(a + (size_t)l * 128); dg_prefetch_l2(b + (size_t)l * 128); }
  }

  double acc[MT][NJ][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < total; it++) {
    const int st = it % STAGES;
    mbar_wait(&full[st], (it / STAGES) & 1);
    __syncthreads();  // every warp has finished k-tile it-1, whose stage is refilled next
    if (tid == 0 && it + STAGES - 1 < total) issue(it + STAGES - 1);
    const double* sA = sStage + st * STAGE_ELEMS + g * BK + t;
    const double* sB = sStage + st * STAGE_ELEMS + BM * BK + (warp * 8 * NJ + g) * BK + t;
    const int ksteps = min(BK / 4, (p.K - it * BK + 3) / 4);
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      if (k4 < ksteps) {
        double a[MT];
#pragma unroll
        for (int i = 0; i < MT; i++) a[i] = sA[i * 8 * BK + k4 * 4];
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const double b = sB[j * 8 * BK + k4 * 4];
#pragma unroll
          for (int i = 0; i < MT; i++) dmma884(acc[i][j], a[i], b);
        }
      }
    }
  }
  __syncthreads();  // the ring is free: `red` may alias it

  // ---- epilogue: acc[i][j][e] = C[row0 + 8i + g][32 warp + 8j + 2t + e] ------------------------------------------------
  const int Wd = p.Wd;
  const int cbase = warp * 8 * NJ + 2 * t;
  bool cv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) cv[j] = cbase + 8 * j < Wd;
  double2 bias2[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) bias2[j] = (p.bias && cv[j]) ? *reinterpret_cast<const double2*>(p.bias + cbase + 8 * j) : make_double2(0.0, 0.0);

#pragma unroll
  for (int i = 0; i < MT; i++) {
    const int64_t row = row0 + 8 * i + g;
    const bool rv = row < p.N;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      double2 r = make_double2(0.0, 0.0);
      if (p.res && rv && cv[j]) r = __ldg(reinterpret_cast<const double2*>(p.res + row * Wd + cbase + 8 * j));
      acc[i][j][0] += bias2[j].x + r.x;
      acc[i][j][1] += bias2[j].y + r.y;
    }
  }
  if (EPI == DG_PLAIN) {
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int64_t row = row0 + 8 * i + g;
      if (row < p.N) {
#pragma unroll
        for (int j = 0; j < NJ; j++)
          if (cv[j]) *reinterpret_cast<double2*>(p.out + row * Wd + cbase + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
    return;
  }

  const double inv_w = 1.0 / Wd;
  // sum over this thread's columns -> quad (t) -> per warp and row; across the 8 warps through shared memory
  auto row_reduce = [&](double (&v)[MT], int slot) {
#pragma unroll
    for (int i = 0; i < MT; i++) {
      v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
      v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
      if (t == 0) red[(slot * BM + 8 * i + g) * 8 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const double4* q = reinterpret_cast<const double4*>(red + (slot * BM + 8 * i + g) * 8);
      const double4 x = q[0], y = q[1];
      v[i] = ((x.x + x.y) + (x.z + x.w)) + ((y.x + y.y) + (y.z + y.w));
    }
  };

  if (EPI == DG_LN_ELU_FWD) {
    double s[MT];
#pragma unroll
    for (int i = 0; i < MT; i++) {
      s[i] = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; j++)
        if (cv[j]) s[i] += acc[i][j][0] + acc[i][j][1];
    }
    row_reduce(s, 0);
    double v[MT];
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const double mean = s[i] * inv_w;
      v[i] = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        acc[i][j][0] -= mean;
        acc[i][j][1] -= mean;
        if (cv[j]) v[i] = fma(acc[i][j][0], acc[i][j][0], fma(acc[i][j][1], acc[i][j][1], v[i]));
      }
    }
    row_reduce(v, 1);
    double2 gm[NJ], bt[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      gm[j] = cv[j] ? *reinterpret_cast<const double2*>(p.gamma + cbase + 8 * j) : make_double2(0.0, 0.0);
      bt[j] = cv[j] ? *reinterpret_cast<const double2*>(p.beta + cbase + 8 * j) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int64_t row = row0 + 8 * i + g;
      if (row < p.N) {
        const double rstd = 1.0 / sqrt(v[i] * inv_w + p.eps);
        if (warp == 0 && t == 0) p.rstd[row] = rstd;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          if (cv[j]) {
            const double x0 = acc[i][j][0] * rstd, x1 = acc[i][j][1] * rstd;
            *reinterpret_cast<double2*>(p.xhat + row * Wd + cbase + 8 * j) = make_double2(x0, x1);
            *reinterpret_cast<double2*>(p.out + row * Wd + cbase + 8 * j) =
                make_double2(dg_elu(fma(x0, gm[j].x, bt[j].x)), dg_elu(fma(x1, gm[j].y, bt[j].y)));
          }
        }
      }
    }
    return;
  }

  if (EPI == DG_LN_ELU_BWD) {
    double2 gm[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) gm[j] = cv[j] ? *reinterpret_cast<const double2*>(p.gamma + cbase + 8 * j) : make_double2(0.0, 0.0);
    double s1[MT], s2[MT];
    double2 cg[NJ], cb[NJ], cz[NJ];  // this thread's column sums over its MT rows
#pragma unroll
    for (int j = 0; j < NJ; j++) cg[j] = cb[j] = cz[j] = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int64_t row = row0 + 8 * i + g;
      const bool rv = row < p.N;
      s1[i] = s2[i] = 0.0;
      double2 xh[NJ], o[NJ];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        xh[j] = o[j] = make_double2(0.0, 0.0);
        if (rv && cv[j]) {
          xh[j] = __ldg(reinterpret_cast<const double2*>(p.xhat_prev + row * Wd + cbase + 8 * j));
          o[j] = __ldg(reinterpret_cast<const double2*>(p.out_prev + row * Wd + cbase + 8 * j));
        } else {
          acc[i][j][0] = acc[i][j][1] = 0.0;
        }
      }
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        // elu'(o) from the forward output: out > 0 <=> o > 0, else exp(o) = out + 1
        const double t0 = acc[i][j][0] * (o[j].x > 0.0 ? 1.0 : o[j].x + 1.0), t1 = acc[i][j][1] * (o[j].y > 0.0 ? 1.0 : o[j].y + 1.0);
        cg[j].x = fma(t0, xh[j].x, cg[j].x); cg[j].y = fma(t1, xh[j].y, cg[j].y);
        cb[j].x += t0; cb[j].y += t1;
        acc[i][j][0] = t0 * gm[j].x;
        acc[i][j][1] = t1 * gm[j].y;
        s1[i] += acc[i][j][0] + acc[i][j][1];
        s2[i] = fma(acc[i][j][0], xh[j].x, fma(acc[i][j][1], xh[j].y, s2[i]));
      }
    }
    row_reduce(s1, 0);
    row_reduce(s2, 1);
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int64_t row = row0 + 8 * i + g;
      if (row < p.N) {
        const double rstd = p.rstd_prev[row], m1 = s1[i] * inv_w, m2 = s2[i] * inv_w;
        double2 xh[NJ];  // second read of xhat (just loaded: L1 / L2) instead of 64 registers held across the reduction
#pragma unroll
        for (int j = 0; j < NJ; j++)
          xh[j] = cv[j] ? __ldg(reinterpret_cast<const double2*>(p.xhat_prev + row * Wd + cbase + 8 * j)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          if (cv[j]) {
            const double z0 = rstd * (acc[i][j][0] - m1 - xh[j].x * m2), z1 = rstd * (acc[i][j][1] - m1 - xh[j].y * m2);
            *reinterpret_cast<double2*>(p.out + row * Wd + cbase + 8 * j) = make_double2(z0, z1);
            cz[j].x += z0; cz[j].y += z1;
          }
        }
      }
    }
    // column sums over the CTA's 32 rows: this thread's MT rows are summed already; the 8 row groups g differ in lane bits 2..4
    if (p.colpart) {
      double* cp = p.colpart + (size_t)blockIdx.x * 3 * Wd;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        double vals[6] = {cg[j].x, cg[j].y, cb[j].x, cb[j].y, cz[j].x, cz[j].y};
#pragma unroll
        for (int q = 0; q < 6; q++) {
          vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], 4);
          vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], 8);
          vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], 16);
        }
        if (g == 0 && cv[j]) {
          const int c = cbase + 8 * j;
          *reinterpret_cast<double2*>(cp + c) = make_double2(vals[0], vals[1]);
          *reinterpret_cast<double2*>(cp + Wd + c) = make_double2(vals[2], vals[3]);
          *reinterpret_cast<double2*>(cp + 2 * Wd + c) = make_double2(vals[4], vals[5]);
        }
      }
    }
  }
}

// out[q][c] = sum_b part[b][q][c], b in fixed order: (LayerNorm scale_bar, LayerNorm bias_bar, Dense bias_bar) from the CTAs' partials
__global__ void dense_colsum_kernel(const double* __restrict__ part, int64_t nblocks, int cols3, double* __restrict__ o0, double* __restrict__ o1,
                                    double* __restrict__ o2, int Wd) {
  // 8 threads per output split the blocks (latency-parallel), combined in fixed order
  const int oidx = blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8, sub = threadIdx.x & 7;
  if (oidx >= cols3) return;
  double s = 0.0;
  for (int64_t b = sub; b < nblocks; b += 8) s += part[(size_t)b * cols3 + oidx];
  s += __shfl_down_sync(0xffffffffu, s, 4, 8);
  s += __shfl_down_sync(0xffffffffu, s, 2, 8);
  s += __shfl_down_sync(0xffffffffu, s, 1, 8);
  if (sub == 0) {
    const int q = oidx / Wd, c = oidx - q * Wd;
    double* o = q == 0 ? o0 : (q == 1 ? o1 : o2);
    if (o) o[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The LAST block of a trunk has no GEMM after it whose epilogue could undo its ELU / LayerNorm: one streaming pass, a warp
// per row (W <= 256: 4 double2 per lane), same formulas as the LN_ELU_BWD epilogue, per-CTA column partials.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LB_WARPS = 8, LB_MAX_CTAS = 148 * 4;

__global__ void __launch_bounds__(LB_WARPS * 32, 2)
dense_block_bwd_last_kernel(int64_t N, int W, const double* __restrict__ out_bar, const double* __restrict__ out, const double* __restrict__ xhat,
                            const double* __restrict__ rstd, const double* __restrict__ gamma, double* __restrict__ z_bar, double* __restrict__ colpart) {
  __shared__ double sred[LB_WARPS][3][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = W >> 1;
  const double inv_w = 1.0 / W;
  double2 cg[4], cb[4], cz[4], gm[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    cg[i] = cb[i] = cz[i] = make_double2(0.0, 0.0);
    gm[i] = lane + 32 * i < nv ? reinterpret_cast<const double2*>(gamma)[lane + 32 * i] : make_double2(0.0, 0.0);
  }
  for (int64_t row = (int64_t)blockIdx.x * LB_WARPS + warp; row < N; row += (int64_t)gridDim.x * LB_WARPS) {
    const double2* ob2 = reinterpret_cast<const double2*>(out_bar + row * W);
    const double2* o2 = reinterpret_cast<const double2*>(out + row * W);
    const double2* x2 = reinterpret_cast<const double2*>(xhat + row * W);
    double2 gh[4], xh[4];
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = lane + 32 * i;
      gh[i] = xh[i] = make_double2(0.0, 0.0);
      if (c < nv) {
        const double2 ob = __ldcs(ob2 + c), o = __ldcs(o2 + c);
        xh[i] = __ldcs(x2 + c);
        const double t0 = ob.x * (o.x > 0.0 ? 1.0 : o.x + 1.0), t1 = ob.y * (o.y > 0.0 ? 1.0 : o.y + 1.0);
        cg[i].x = fma(t0, xh[i].x, cg[i].x); cg[i].y = fma(t1, xh[i].y, cg[i].y);
        cb[i].x += t0; cb[i].y += t1;
        gh[i] = make_double2(t0 * gm[i].x, t1 * gm[i].y);
        s1 += gh[i].x + gh[i].y;
        s2 = fma(gh[i].x, xh[i].x, fma(gh[i].y, xh[i].y, s2));
      }
    }
    const double m1 = warp_sum(s1) * inv_w, m2 = warp_sum(s2) * inv_w, rs = rstd[row];
    double2* z2 = reinterpret_cast<double2*>(z_bar + row * W);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const double z0 = rs * (gh[i].x - m1 - xh[i].x * m2), z1 = rs * (gh[i].y - m1 - xh[i].y * m2);
        z2[c] = make_double2(z0, z1);
        cz[i].x += z0; cz[i].y += z1;
      }
    }
  }
  if (colpart == nullptr) return;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = 2 * (lane + 32 * i);
    if (c < W) {
      sred[warp][0][c] = cg[i].x; sred[warp][0][c + 1] = cg[i].y;
      sred[warp][1][c] = cb[i].x; sred[warp][1][c + 1] = cb[i].y;
      sred[warp][2][c] = cz[i].x; sred[warp][2][c + 1] = cz[i].y;
    }
  }
  __syncthreads();
  double* cp = colpart + (size_t)blockIdx.x * 3 * W;
  for (int idx = threadIdx.x; idx < 3 * W; idx += blockDim.x) {
    const int q = idx / W, c = idx - q * W;
    double s = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < LB_WARPS; w2++) s += sred[w2][q][c];
    cp[idx] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Wbar[K, Wd] = A[N, K]^T Z[N, Wd]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TN_TILE = 128, TN_KT = 16, TN_PITCH = TN_TILE + 4, TN_STAGES = 4, TN_THREADS = 256;

struct DenseTnParams {
  int64_t N;
  int K, Wd, tiles_m, tiles_n, kslices;
  int64_t rows_per_slice;  // multiple of TN_KT
  double* partial;         // [kslices][K][Wd]
};

__global__ void __launch_bounds__(TN_THREADS, 1)
dense_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmZ, const DenseTnParams p) {
  constexpr int STAGE_ELEMS = 2 * TN_KT * TN_PITCH;
  constexpr uint32_t STAGE_BYTES = STAGE_ELEMS * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sStage = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + TN_STAGES * STAGE_ELEMS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int tile = blockIdx.x % (p.tiles_m * p.tiles_n), slice = blockIdx.x / (p.tiles_m * p.tiles_n);
  const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
  const int64_t r_begin = (int64_t)slice * p.rows_per_slice;
  const int64_t r_end = (p.N < r_begin + p.rows_per_slice) ? p.N : r_begin + p.rows_per_slice;
  const int total = r_end > r_begin ? (int)((r_end - r_begin + TN_KT - 1) / TN_KT) : 0;
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmZ);
    for (int s = 0; s < TN_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int it) {
    const int st = it % TN_STAGES;
    double* sA = sStage + st * STAGE_ELEMS;
    mbar_expect_tx(&full[st], STAGE_BYTES);
    tma_load_3d(sA, &tmA, &full[st], tm * TN_TILE, (int)(r_begin + (int64_t)it * TN_KT), 0);
    tma_load_3d(sA + TN_KT * TN_PITCH, &tmZ, &full[st], tn * TN_TILE, (int)(r_begin + (int64_t)it * TN_KT), 0);
  };
  if (tid == 0)
    for (int it = 0; it < TN_STAGES - 1 && it < total; it++) issue(it);

  // 2 x 4 warp grid: warp tile 64 (m) x 32 (n) = 8 x 4 DMMA tiles
  const int wm = warp >> 2, wn = warp & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < total; it++) {
    const int st = it % TN_STAGES;
    mbar_wait(&full[st], (it / TN_STAGES) & 1);
    __syncthreads();
    if (tid == 0 && it + TN_STAGES - 1 < total) issue(it + TN_STAGES - 1);
    // rows beyond r_end inside the last k-tile belong to the next slice (or are zero-filled past N): mask them
    const int64_t rbase = r_begin + (int64_t)it * TN_KT;
    const int kvalid = (r_end - rbase < TN_KT) ? (int)(r_end - rbase) : TN_KT;
    // A^T fragment: a = A_tile[k = t][m = g]; B fragment: b = Z_tile[k = t][n = g]
    const double* sA = sStage + st * STAGE_ELEMS + t * TN_PITCH + wm * 64 + g;
    const double* sZ = sStage + st * STAGE_ELEMS + TN_KT * TN_PITCH + t * TN_PITCH + wn * 32 + g;
#pragma unroll
    for (int k4 = 0; k4 < TN_KT / 4; k4++) {
      if (k4 * 4 < kvalid) {
        const bool kv = k4 * 4 + t < kvalid;
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = kv ? sA[k4 * 4 * TN_PITCH + 8 * i] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double b = sZ[k4 * 4 * TN_PITCH + 8 * j];
#pragma unroll
          for (int i = 0; i < 8; i++) dmma884(acc[i][j], a[i], b);
        }
      }
    }
  }
  // partial tile -> partial[slice][K][Wd]
  double* out = p.partial + (size_t)slice * p.K * p.Wd;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = tm * TN_TILE + wm * 64 + 8 * i + g;
    if (m < p.K) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = tn * TN_TILE + wn * 32 + 8 * j + 2 * t;
        if (c < p.Wd) *reinterpret_cast<double2*>(out + (size_t)m * p.Wd + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

__global__ void dense_tn_reduce_kernel(const double* __restrict__ partial, int kslices, int64_t count, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  double s = 0.0;
  for (int k = 0; k < kslices; k++) s += partial[(size_t)k * count + idx];
  out[idx] = s;
}

template <int NJ, int EPI>
static int launch_dense_nn(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const DenseParams& p) {
  size_t smem = (size_t)DG_STAGES * (DG_BM + 64 * NJ) * DG_BK * 8 + DG_STAGES * 8;
  if (const char* e = getenv("GDFT_DENSE_ONE_CTA")) { if (e[0] == '1' && smem < 120 * 1024) smem = 120 * 1024; }  // tuning probe: one CTA per SM
  GDFT_CUDA_TRY((cudaFuncSetAttribute(dense_nn_kernel<NJ, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  const unsigned grid = (unsigned)((p.N + DG_BM - 1) / DG_BM);
  dense_nn_kernel<NJ, EPI><<<grid, DG_THREADS, smem, stream>>>(tmA, tmB, p);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

template <int EPI>
static int dispatch_dense_nn(cudaStream_t stream, const CUtensorMap& tmA, const double* Bt, DenseParams p) {
  const int nj = (p.Wd + 63) / 64;
  CUtensorMap tmB;
  int rc = make_tmap_3d(&tmB, Bt, (uint64_t)p.K, (uint64_t)p.Wd, 1, (uint64_t)p.K * 8, (uint64_t)p.K * p.Wd * 8, DG_BK, 64 * nj);
  if (rc) return rc;
  switch (nj) {
    case 1: return launch_dense_nn<1, EPI>(stream, tmA, tmB, p);
    case 2: return launch_dense_nn<2, EPI>(stream, tmA, tmB, p);
    case 3: return launch_dense_nn<3, EPI>(stream, tmA, tmB, p);
    default: return launch_dense_nn<4, EPI>(stream, tmA, tmB, p);
  }
}

size_t dense_workspace(int64_t N, int64_t K, int64_t Wd) {
  // the larger of: column-sum partials of the reverse GEMM (one row of 3 Wd per CTA) and the split-K partial tiles of Wbar
  const size_t colpart = (size_t)((N + DG_BM - 1) / DG_BM) * 3 * Wd * 8;
  const size_t tn = (size_t)148 * K * Wd * 8;
  return (colpart > tn ? colpart : tn) + 512;
}

}  // namespace gdft

using namespace gdft;

static bool dense_shape_ok(int64_t N, int64_t K, int64_t Wd) {
  return N > 0 && N <= (int64_t)2147483000 && K >= 2 && K <= 4096 && (K % 2) == 0 && Wd >= 8 && Wd <= 256 && (Wd % 8) == 0;
}

extern "C" int gdft_dense_supported(int64_t K, int64_t Wd) { return dense_shape_ok(1, K, Wd) ? 1 : 0; }

extern "C" int gdft_dense_fwd(gdft_stream_t stream_, int64_t N, int64_t K, int64_t Wd, const double* x, const double* kernel_t, const double* bias,
                              const double* res, double* out) {
  if (!dense_shape_ok(N, K, Wd)) return GDFT_BAD_SHAPE;
  if (!x || !kernel_t || !out) return GDFT_BAD_ARGUMENT;
  if (!aligned16(x) || !aligned16(kernel_t) || !aligned16(bias) || !aligned16(res) || !aligned16(out)) return GDFT_BAD_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tmA;
  int rc = make_tmap_3d(&tmA, x, (uint64_t)K, (uint64_t)N, 1, (uint64_t)K * 8, (uint64_t)N * K * 8, DG_BK, DG_BM);
  if (rc) return rc;
  DenseParams p{};
  p.N = N; p.K = (int)K; p.Wd = (int)Wd; p.n_ktile = (int)((K + DG_BK - 1) / DG_BK);
  p.bias = bias; p.res = res; p.out = out;
  return dispatch_dense_nn<DG_PLAIN>(stream, tmA, kernel_t, p);
}

extern "C" int gdft_dense_block_fwd(gdft_stream_t stream_, int64_t N, int64_t W, const double* x, const double* kernel_t, const double* dense_bias,
                                    const double* scale, const double* bias, double eps, double* out, double* xhat, double* rstd) {
  if (!dense_shape_ok(N, W, W)) return GDFT_BAD_SHAPE;
  if (!x || !kernel_t || !scale || !bias || !out || !xhat || !rstd) return GDFT_BAD_ARGUMENT;
  if (!aligned16(x) || !aligned16(kernel_t) || !aligned16(dense_bias) || !aligned16(scale) || !aligned16(bias) || !aligned16(out) || !aligned16(xhat))
    return GDFT_BAD_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tmA;
  int rc = make_tmap_3d(&tmA, x, (uint64_t)W, (uint64_t)N, 1, (uint64_t)W * 8, (uint64_t)N * W * 8, DG_BK, DG_BM);
  if (rc) return rc;
  DenseParams p{};
  p.N = N; p.K = (int)W; p.Wd = (int)W; p.n_ktile = (int)((W + DG_BK - 1) / DG_BK);
  p.bias = dense_bias; p.res = x; p.gamma = scale; p.beta = bias; p.eps = eps; p.out = out; p.xhat = xhat; p.rstd = rstd;
  return dispatch_dense_nn<DG_LN_ELU_FWD>(stream, tmA, kernel_t, p);
}

extern "C" int gdft_dense_block_bwd(gdft_stream_t stream_, int64_t N, int64_t W, const double* z_bar, const double* kernel, const double* prev_out,
                                    const double* prev_xhat, const double* prev_rstd, const double* prev_scale, double* prev_z_bar,
                                    double* prev_scale_bar, double* prev_bias_bar, double* prev_dense_bias_bar, void* ws, size_t ws_bytes) {
  if (!dense_shape_ok(N, W, W)) return GDFT_BAD_SHAPE;
  if (!z_bar || !kernel || !prev_out || !prev_xhat || !prev_rstd || !prev_scale || !prev_z_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(z_bar) || !aligned16(kernel) || !aligned16(prev_out) || !aligned16(prev_xhat) || !aligned16(prev_scale) || !aligned16(prev_z_bar) ||
      !aligned16(ws))
    return GDFT_BAD_ALIGNMENT;
  const bool pg = prev_scale_bar || prev_bias_bar || prev_dense_bias_bar;
  const int64_t nblocks = (N + DG_BM - 1) / DG_BM;
  if (pg && ws_bytes < (size_t)nblocks * 3 * W * 8) return GDFT_WORKSPACE_TOO_SMALL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tmA;
  int rc = make_tmap_3d(&tmA, z_bar, (uint64_t)W, (uint64_t)N, 1, (uint64_t)W * 8, (uint64_t)N * W * 8, DG_BK, DG_BM);
  if (rc) return rc;
  DenseParams p{};
  p.N = N; p.K = (int)W; p.Wd = (int)W; p.n_ktile = (int)((W + DG_BK - 1) / DG_BK);
  p.res = z_bar; p.gamma = prev_scale; p.out_prev = prev_out; p.xhat_prev = prev_xhat; p.rstd_prev = prev_rstd; p.out = prev_z_bar;
  p.colpart = pg ? static_cast<double*>(ws) : nullptr;
  // x_bar = z_bar kernel^T: the transposed right operand of that product is the kernel as stored
  rc = dispatch_dense_nn<DG_LN_ELU_BWD>(stream, tmA, kernel, p);
  if (rc) return rc;
  if (pg) {
    const int cols3 = 3 * (int)W;
    dense_colsum_kernel<<<(cols3 + 31) / 32, 256, 0, stream>>>(p.colpart, nblocks, cols3, prev_scale_bar, prev_bias_bar, prev_dense_bias_bar, (int)W);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}

extern "C" int gdft_dense_bwd_weight(gdft_stream_t stream_, int64_t N, int64_t K, int64_t Wd, const double* x, const double* z_bar, double* kernel_bar,
                                     void* ws, size_t ws_bytes) {
  if (!dense_shape_ok(N, K, Wd) || K % 8 != 0) return GDFT_BAD_SHAPE;
  if (!x || !z_bar || !kernel_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(x) || !aligned16(z_bar) || !aligned16(kernel_bar) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DenseTnParams p{};
  p.N = N; p.K = (int)K; p.Wd = (int)Wd;
  p.tiles_m = (int)((K + TN_TILE - 1) / TN_TILE);
  p.tiles_n = (int)((Wd + TN_TILE - 1) / TN_TILE);
  const int tiles = p.tiles_m * p.tiles_n;
  int ks = 148 / tiles;
  if (ks < 1) ks = 1;
  const int64_t ktiles = (N + TN_KT - 1) / TN_KT;
  if (ks > ktiles) ks = (int)ktiles;
  p.rows_per_slice = ((ktiles + ks - 1) / ks) * TN_KT;
  p.kslices = (int)((N + p.rows_per_slice - 1) / p.rows_per_slice);
  if (ws_bytes < (size_t)p.kslices * K * Wd * 8) return GDFT_WORKSPACE_TOO_SMALL;
  p.partial = static_cast<double*>(ws);
  CUtensorMap tmA, tmZ;
  int rc = make_tmap_3d(&tmA, x, (uint64_t)K, (uint64_t)N, 1, (uint64_t)K * 8, (uint64_t)N * K * 8, TN_PITCH, TN_KT);
  if (rc) return rc;
  rc = make_tmap_3d(&tmZ, z_bar, (uint64_t)Wd, (uint64_t)N, 1, (uint64_t)Wd * 8, (uint64_t)N * Wd * 8, TN_PITCH, TN_KT);
  if (rc) return rc;
  const size_t smem = (size_t)TN_STAGES * 2 * TN_KT * TN_PITCH * 8 + TN_STAGES * 8;
  GDFT_CUDA_TRY(cudaFuncSetAttribute(dense_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dense_tn_kernel<<<(unsigned)(tiles * p.kslices), TN_THREADS, smem, stream>>>(tmA, tmZ, p);
  GDFT_LAUNCH_CHECK();
  const int64_t count = K * Wd;
  dense_tn_reduce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(p.partial, p.kslices, count, kernel_bar);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_dense_block_bwd_last(gdft_stream_t stream_, int64_t N, int64_t W, const double* out_bar, const double* out, const double* xhat,
                                         const double* rstd, const double* scale, double* z_bar, double* scale_bar, double* bias_bar,
                                         double* dense_bias_bar, void* ws, size_t ws_bytes) {
  if (!dense_shape_ok(N, W, W)) return GDFT_BAD_SHAPE;
  if (!out_bar || !out || !xhat || !rstd || !scale || !z_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(out_bar) || !aligned16(out) || !aligned16(xhat) || !aligned16(scale) || !aligned16(z_bar) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  const bool pg = scale_bar || bias_bar || dense_bias_bar;
  int64_t ctas = (N + LB_WARPS - 1) / LB_WARPS;
  if (ctas > LB_MAX_CTAS) ctas = LB_MAX_CTAS;
  if (pg && ws_bytes < (size_t)ctas * 3 * W * 8) return GDFT_WORKSPACE_TOO_SMALL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  double* colpart = pg ? static_cast<double*>(ws) : nullptr;
  dense_block_bwd_last_kernel<<<(unsigned)ctas, LB_WARPS * 32, 0, stream>>>(N, (int)W, out_bar, out, xhat, rstd, scale, z_bar, colpart);
  GDFT_LAUNCH_CHECK();
  if (pg) {
    const int cols3 = 3 * (int)W;
    dense_colsum_kernel<<<(cols3 + 31) / 32, 256, 0, stream>>>(colpart, ctas, cols3, scale_bar, bias_bar, dense_bias_bar, (int)W);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}
