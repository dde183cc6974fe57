// Row f1: the two reductions of the CDIIS step that the host framework runs as degenerate GEMMs
// (grad_dft/evaluate.py:1041-1205).  Inside the graph-replayed SCF iteration of a small molecule the iteration time IS
// the sum of its kernels' durations, and these two were 21 us each at n = 43 (a 10 x 10 output with K = n^2 on ONE CTA
// of a 32 x 64-tile GEMM; a 1 x n^2 output with K = 10):
//   gram[s][i][j] = sum_kl e[i][s][k][l] e[j][s][k][l]          "iskl,jskl->sij"   (evaluate.py:1165)
//   out[s][k][l]  = sum_i x[s][i] f[i][s][k][l]                 "si,isjk->sjk"     (evaluate.py:1198)
// One warp per gram entry (fixed lane-strided order + the warp tree: deterministic), one thread per output element.
#include "common.cuh"
#include <math_constants.h>

namespace gdft {

// BORDERED: write the CDIIS matrix B[2, m+1, m+1] of evaluate.py:1167-1181 instead of the bare Gram matrix:
// B[0,0] = 0, B[0,1+i] = B[1+i,0] = live_i, B[1+i,1+j] = G_ij, and B[1+i,1+i] = 1 for the slots that are not live yet
// (live_i = i <= cycle).
template <bool BORDERED>
__global__ void __launch_bounds__(1024) diis_gram_kernel(int m, int64_t nn, int cycle, const double* __restrict__ e, double* __restrict__ gram) {
  // one CTA of 32 warps per entry (s, i <= j) of the symmetric Gram matrix: each warp takes a 32nd of the n^2 elements, then
  // a fixed-order sum.  (One WARP per entry, the first version, left a 10 x 10 x 2 problem on 200 warps: 515 us at n = 264,
  // 7 % of the benzene-shaped iteration on 8 GPUs, where this n x n work is replicated on every rank.)
  const int npairs = m * (m + 1) / 2;
  const int s = blockIdx.x / npairs;
  int rem = blockIdx.x - s * npairs, i = 0;
  while (rem >= m - i) { rem -= m - i; ++i; }
  const int j = i + rem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double part[32];
  const double* a = e + ((size_t)i * 2 + s) * nn;
  const double* b = e + ((size_t)j * 2 + s) * nn;
  const int64_t chunk = ((nn + 31) / 32 + 31) & ~int64_t(31);
  const int64_t k1 = (warp + 1) * chunk < nn ? (warp + 1) * chunk : nn;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int64_t k = warp * chunk + lane;
  for (; k + 96 < k1; k += 128) {
    double av[4], bv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { av[u] = a[k + 32 * u]; bv[u] = b[k + 32 * u]; }
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = fma(av[u], bv[u], acc[u]);
  }
  for (; k < k1; k += 32) acc[0] = fma(a[k], b[k], acc[0]);
  const double w = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
  if (lane == 0) part[warp] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int t = 0; t < 32; t++) v += part[t];
    if (BORDERED) {
      const int mb = m + 1;
      double* B = gram + (size_t)s * mb * mb;
      const bool live = i <= cycle;
      B[(1 + i) * mb + 1 + j] = (i == j && !live) ? 1.0 : v;
      B[(1 + j) * mb + 1 + i] = (i == j && !live) ? 1.0 : v;
      if (i == j) {
        B[1 + i] = live ? 1.0 : 0.0;
        B[(1 + i) * mb] = live ? 1.0 : 0.0;
        if (i == 0) B[0] = 0.0;
      }
    } else {
      gram[((size_t)s * m + i) * m + j] = v;
      gram[((size_t)s * m + j) * m + i] = v;
    }
  }
}

__global__ void __launch_bounds__(256) diis_combine_kernel(int m, int64_t nn, const double* __restrict__ x, const double* __restrict__ f,
                                                          double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * nn) return;
  const int s = (int)(idx / nn);
  const int64_t kl = idx - (int64_t)s * nn;
  double acc = 0.0;
  for (int i = 0; i < m; i++) acc = fma(x[s * m + i], f[((size_t)i * 2 + s) * nn + kl], acc);
  out[idx] = acc;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_diis_gram(gdft_stream_t stream, int m, int64_t n, const double* err /*[m,2,n,n]*/, double* gram /*[2,m,m]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!err || !gram) return GDFT_BAD_ARGUMENT;
  diis_gram_kernel<false><<<m * (m + 1), 1024, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, 0, err, gram);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_diis_matrix(gdft_stream_t stream, int m, int64_t n, int cycle, const double* err /*[m,2,n,n]*/,
                                double* B /*[2,m+1,m+1]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!err || !B) return GDFT_BAD_ARGUMENT;
  diis_gram_kernel<true><<<m * (m + 1), 1024, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, cycle, err, B);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_diis_combine(gdft_stream_t stream, int m, int64_t n, const double* x /*[2,m]*/, const double* fock_vec /*[m,2,n,n]*/,
                                 double* out /*[2,n,n]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!x || !fock_vec || !out) return GDFT_BAD_ARGUMENT;
  const int64_t total = 2 * n * n;
  diis_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, x, fock_vec, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// abs_clip (grad_dft/molecule.py:687-689): out = |src| > thr ? x : 0.  With x = src it is the clip itself; with x = the
// incoming cotangent it is its VJP (and, applied again, the VJP of that).  The host-framework composite is four launches
// (abs, compare, zeros_like, where) per call and three more in its reverse pass; the predictor calls it on the
// densities and on the Fock matrix of every build.
namespace gdft {
__global__ void __launch_bounds__(256) abs_clip_kernel(int64_t count, const double* __restrict__ x, const double* __restrict__ src, double thr,
                                                      double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fabs(src[i]) > thr ? x[i] : 0.0;
}
}  // namespace gdft

extern "C" int gdft_abs_clip(gdft_stream_t stream, int64_t count, const double* x, const double* src, double thr, double* out) {
  if (count < 0) return GDFT_BAD_SHAPE;
  if (count == 0) return GDFT_OK;
  if (!x || !src || !out) return GDFT_BAD_ARGUMENT;
  const unsigned grid = (unsigned)imin64((count + 255) / 256, 148 * 16);
  gdft::abs_clip_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(count, x, src, thr, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The n x n tail of one DIIS SCF iteration of a small molecule (n <= 64) as TWO kernels around the eigensolver, one CTA per
// spin (grad_dft/evaluate.py:983-1016, 1111-1205; grad_dft/utils/eigenproblem.py:110-129; grad_dft/molecule.py:815-889).
// In the graph-replayed H2O-shaped iteration this tail was ~55 host-framework launches (two 43 x 43 GEMMs on a 32 x 64-tile
// kernel, ring-buffer copies, a pivoted LU + inverse of an 11 x 11 matrix in five library launches, a radix sort for the
// aufbau ranks ...) at 2-3 us each: ~140 of the 480 us.
//
//   scf_diis_kernel   fds = F D S, err = fds - fds^T (evaluate.py:1130-1134); ring-buffer slot written in place (the slot
//                     rotates instead of the whole buffer shifting: logical entry i lives in physical slot (head + i) % m);
//                     the ONE new row/column of the Gram matrix (evaluate.py:1165; the others cannot have changed);
//                     bordered CDIIS matrix (1167-1181) in logical order; x = B^-1 e_0 by LU with partial pivoting in one
//                     warp (upstream: jnp.linalg.inv(B) @ C); F' = sum_i x_i F_i (1198); C = L^-1 F' L^-T (eigenproblem.py:
//                     127) -> the reduced matrix the eigensolver takes.
//   scf_occupy_kernel mo_coeff = L^-T V (eigenproblem.py:129); aufbau occupations: the nelec lowest orbitals by stable rank
//                     (molecule.py:851-889; nelec = round(sum of the previous occupations)); rdm1 = C occ C^T (815-846).
// ---------------------------------------------------------------------------------------------------------------------
namespace gdft {

constexpr int SCF_MAX_N = 64;
constexpr int SCF_MAX_M = 16;
constexpr int SCF_THREADS = 1024;  // one CTA per spin: the steps are serial, each is n^2-parallel -- as many threads as a CTA takes

// Shared-memory matrices are [np8][p]: np8 = n rounded up to 8 (rows and columns beyond n are ZERO, so whole 8 x 8 x 4 tensor
// tiles can be used without predication), pitch p = np8 + 4 == 4 or 12 (mod 16): the DMMA fragment loads (row = lane >> 2,
// col = lane & 3 and the transposed pattern) then touch 16 distinct 8-byte banks per half-warp -- conflict-free, as in K1.
__device__ __host__ __forceinline__ int scf_np8(int n) { return (n + 7) & ~7; }
__device__ __host__ __forceinline__ int scf_pitch(int n) { return scf_np8(n) + 4; }

// C = op(A) op(B) with FP64 tensor tiles (one 8 x 8 output tile per warp and round, k in steps of 4); `kscale` (optional)
// multiplies column k of op(A).  As scalar code this product is shared-memory-bandwidth-bound (two 8-byte loads per FMA:
// n^3 / 8 wavefronts, 12 000 clocks at n = 43); a DMMA needs two fragment loads per 256 FMAs.
template <bool TRANS_A, bool TRANS_B>
__device__ __forceinline__ void scf_matmul(int n, int p, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                           const double* __restrict__ kscale = nullptr) {
  const int np8 = scf_np8(n), nt = np8 >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int tile = warp; tile < nt * nt; tile += SCF_THREADS / 32) {
    const int i0 = (tile / nt) << 3, j0 = (tile % nt) << 3;
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
    const double* ap = TRANS_A ? A + t * p + i0 + g : A + (i0 + g) * p + t;
    const double* bp = TRANS_B ? B + (j0 + g) * p + t : B + t * p + j0 + g;
    const int as = TRANS_A ? 4 * p : 4, bs = TRANS_B ? 4 : 4 * p;
    for (int k0 = 0; k0 < np8; k0 += 8) {  // two independent accumulator chains
      double a0 = ap[0], a1 = ap[as];
      const double b0 = bp[0], b1 = bp[bs];
      if (kscale) { a0 *= kscale[k0 + t]; a1 *= kscale[k0 + 4 + t]; }
      dmma884(c0, a0, b0);
      dmma884(c1, a1, b1);
      ap += 2 * as;
      bp += 2 * bs;
    }
    C[(i0 + g) * p + j0 + 2 * t] = c0[0] + c1[0];
    C[(i0 + g) * p + j0 + 2 * t + 1] = c0[1] + c1[1];
  }
}
// global [n][n] -> shared [np8][p], padding zeroed
__device__ __forceinline__ void scf_load(int n, int p, const double* __restrict__ g, double* __restrict__ s) {
  const int np8 = scf_np8(n);
  for (int o = threadIdx.x; o < np8 * np8; o += SCF_THREADS) {
    const int i = o / np8, j = o - i * np8;
    s[i * p + j] = (i < n && j < n) ? g[i * n + j] : 0.0;
  }
}

struct DiisArgs {
  int n, m, cycle;
  const double *fock, *rdm1, *overlap, *L_inv;  // [2,n,n], [2,n,n], [n,n], [n,n]
  double *fock_vec, *err_vec, *gram;            // [m,2,n,n], [m,2,n,n], [2,m,m] (persistent between cycles)
  double *x_out, *fock_out, *C_out;             // [2,m] (diagnostic), [2,n,n], [2,n,n]
};

__global__ void __launch_bounds__(SCF_THREADS) scf_diis_kernel(const DiisArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, m = a.m, p = scf_pitch(n), nn = n * n, s = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int msz = scf_np8(n) * p;
  double* sA = sm;              // F, later F'
  double* sB = sA + msz;        // D, later (F D) S, later C
  double* sC = sB + msz;        // S, later L_inv
  double* sD = sC + msz;        // products
  __shared__ double sBmat[(SCF_MAX_M + 1) * (SCF_MAX_M + 2)];
  __shared__ double sx[SCF_MAX_M + 1];
  __shared__ double sPart[SCF_THREADS / 32];
  __shared__ double sRpiv, sRdiag[SCF_MAX_M + 1];
  __shared__ int sPiv;

  // ring-buffer bookkeeping of evaluate.py:1111-1128: cycle < m writes slot `cycle`; cycle == m is dropped (out-of-bounds
  // .at[].set(), kept); cycle > m shifts by one and appends -- here the slot rotates: head = number of shifts so far
#ifdef GDFT_SCF_PROF
  long long stamps[12]; int ns = 0;
#define STAMP() do { __syncthreads(); if (tid == 0) stamps[ns] = clock64(); ns++; } while (0)
#else
#define STAMP() do {} while (0)
#endif
  const int cyc = a.cycle & 0xffff;
  STAMP();
  const bool store = cyc != m;
  const int head = cyc > m ? (cyc - m) % m : 0;
  const int slot = cyc < m ? cyc : (head + m - 1) % m;  // physical slot of the new entry (logical cyc, or logical m - 1)

  scf_load(n, p, a.fock + (size_t)s * nn, sA);
  scf_load(n, p, a.rdm1 + (size_t)s * nn, sB);
  scf_load(n, p, a.overlap, sC);
  __syncthreads();
  STAMP();
  scf_matmul<false, false>(n, p, sA, sB, sD);  // F D
  __syncthreads();
  scf_matmul<false, false>(n, p, sD, sC, sB);  // (F D) S
  __syncthreads();
  scf_load(n, p, a.L_inv, sC);  // S is dead: L^-1 takes its place (needed at the very end; the load overlaps what follows)
  STAMP();
  if (store) {
    double* eg = a.err_vec + ((size_t)slot * 2 + s) * nn;
    double* fg = a.fock_vec + ((size_t)slot * 2 + s) * nn;
    for (int o = tid; o < nn; o += SCF_THREADS) {
      const int i = o / n, j = o - i * n;
      eg[o] = sB[i * p + j] - sB[j * p + i];
      fg[o] = sA[i * p + j];
    }
    __syncthreads();  // the new slot is read back from global memory below (same CTA: visible after the barrier)
    STAMP();
    // the new row / column of the Gram matrix: <err_slot, err_q> for every physical slot q
    double* G = a.gram + (size_t)s * m * m;
    {
      constexpr int NW = SCF_THREADS / 32;
      const int wpq = NW / m > 0 ? NW / m : 1;  // warps per dot product (m = 10: three thirds of the n^2 elements each)
      const int chunk = ((nn + wpq - 1) / wpq + 31) & ~31;
      for (int q0 = 0; q0 < m; q0 += NW / wpq) {
        const int q = q0 + warp / wpq, part = warp % wpq;
        double acc = 0.0;
        if (q < m && warp / wpq < NW / wpq) {
          const double* b = a.err_vec + ((size_t)q * 2 + s) * nn;
          const int k1 = min(nn, (part + 1) * chunk);
          double ac[4] = {0.0, 0.0, 0.0, 0.0};
          int k = part * chunk + lane;
          for (; k + 96 < k1; k += 128) {
            double ev[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { ev[u] = eg[k + 32 * u]; bv[u] = b[k + 32 * u]; }
#pragma unroll
            for (int u = 0; u < 4; u++) ac[u] = fma(ev[u], bv[u], ac[u]);
          }
          for (; k < k1; k += 32) ac[0] = fma(eg[k], b[k], ac[0]);
          acc = warp_sum((ac[0] + ac[1]) + (ac[2] + ac[3]));
          if (lane == 0) sPart[warp] = acc;
        }
        __syncthreads();
        if (tid < NW / wpq && q0 + tid < m) {  // fixed-order sum of the parts
          double v = 0.0;
          for (int t = 0; t < wpq; t++) v += sPart[tid * wpq + t];
          G[slot * m + q0 + tid] = v;
          G[(q0 + tid) * m + slot] = v;
        }
        __syncthreads();
      }
    }
  }
  STAMP();
  // bordered matrix in logical order: entry i <-> physical (head + i) % m; live_i = i <= cycle
  const int mb = m + 1;
  {
    const double* G = a.gram + (size_t)s * m * m;
    for (int o = tid; o < mb * mb; o += SCF_THREADS) {
      const int r = o / mb, c = o - r * mb;
      double v;
      if (r == 0 && c == 0) v = 0.0;
      else if (r == 0) v = (c - 1 <= cyc) ? 1.0 : 0.0;
      else if (c == 0) v = (r - 1 <= cyc) ? 1.0 : 0.0;
      else {
        const int i = r - 1, j = c - 1;
        v = (i == j && i > cyc) ? 1.0 : G[((head + i) % m) * m + (head + j) % m];
      }
      sBmat[r * (mb + 1) + c] = v;
    }
    if (tid < mb) sBmat[tid * (mb + 1) + mb] = tid == 0 ? 1.0 : 0.0;  // right-hand side e_0
  }
  __syncthreads();
  STAMP();
  // LU with partial pivoting on [B | e_0] (LAPACK's pivot rule: the first largest |entry|; rows scaled by the reciprocal pivot
  // as dgetf2 does), then back substitution; one thread per element of the trailing block.  Only the first LU_THREADS threads
  // take part and they meet at a named barrier: the ~60 barriers of the elimination cost a third of a CTA-wide one each
  constexpr int LU_THREADS = 320;  // >= (SCF_MAX_M) * (SCF_MAX_M + 2) trailing elements
#define LU_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(LU_THREADS) : "memory")
  if (tid < LU_THREADS) {
    const int w = mb + 1;
    for (int k = 0; k < mb; k++) {
      if (warp == 0) {
        // a NaN counts as the largest magnitude, so that the choice stays inside rows k .. mb-1 and uniform over the warp
        // (upstream's inv() of a singular CDIIS matrix yields NaN too; the SCF result is then NaN on both sides)
        double best = -1.0;
        if (lane >= k && lane < mb) { const double v = sBmat[lane * w + k]; best = (v != v) ? CUDART_INF : fabs(v); }
        int arg = lane;
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
          if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) { sPiv = arg; sRpiv = 1.0 / sBmat[arg * w + k]; }
      }
      LU_SYNC();
      const int arg = sPiv;
      if (arg != k) {  // uniform branch
        if (tid < w) { const double t = sBmat[k * w + tid]; sBmat[k * w + tid] = sBmat[arg * w + tid]; sBmat[arg * w + tid] = t; }
        LU_SYNC();
      }
      const int wk = w - k - 1;                    // trailing columns k+1 .. w-1 (the right-hand side included)
      const int r = k + 1 + tid / wk, c = k + 1 + tid % wk;
      double v = 0.0;
      if (r < mb) v = fma(-(sBmat[r * w + k] * sRpiv), sBmat[k * w + c], sBmat[r * w + c]);
      LU_SYNC();
      if (r < mb) sBmat[r * w + c] = v;
      if (tid == 0) sRdiag[k] = sRpiv;
      LU_SYNC();
    }
    for (int r = mb - 1; r >= 0; r--) {
      if (tid == 0) sx[r] = sBmat[r * w + mb] * sRdiag[r];
      LU_SYNC();
      if (tid < r) sBmat[tid * w + mb] = fma(-sBmat[tid * w + r], sx[r], sBmat[tid * w + mb]);
      LU_SYNC();
    }
  }
#undef LU_SYNC
  __syncthreads();
  STAMP();
  if (tid < m) a.x_out[s * m + tid] = sx[1 + tid];
  // F' = sum_i x_i F_i in logical order (evaluate.py:1198)
  for (int o = tid; o < nn; o += SCF_THREADS) {
    double acc = 0.0;
    for (int i0 = 0; i0 < m; i0 += 8) {  // eight ring-buffer loads in flight, summed in logical order
      double fv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) fv[u] = i0 + u < m ? a.fock_vec[((size_t)((head + i0 + u) % m) * 2 + s) * nn + o] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (i0 + u < m) acc = fma(sx[1 + i0 + u], fv[u], acc);
    }
    a.fock_out[(size_t)s * nn + o] = acc;
    sA[(o / n) * p + (o % n)] = acc;
  }
  __syncthreads();
  STAMP();
  scf_matmul<false, false>(n, p, sC, sA, sD);  // L^-1 F'
  __syncthreads();
  scf_matmul<false, true>(n, p, sD, sC, sB);   // (L^-1 F') L^-T
  __syncthreads();
  for (int o = tid; o < nn; o += SCF_THREADS) a.C_out[(size_t)s * nn + o] = sB[(o / n) * p + (o % n)];
  STAMP();
#ifdef GDFT_SCF_PROF
  if (tid == 0 && s == 0 && (a.cycle >> 16)) for (int i = 0; i < ns; i++) a.x_out[i] = (double)(stamps[i] - stamps[0]);
#endif
}

struct OccupyArgs {
  int n;
  const double *evals, *V, *L_inv, *occ_prev;  // [2,n], [2,n,n], [n,n], [2,n]
  double *mo_coeff, *mo_occ, *rdm1;            // [2,n,n], [2,n], [2,n,n]
};

__global__ void __launch_bounds__(SCF_THREADS) scf_occupy_kernel(const OccupyArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, p = scf_pitch(n), nn = n * n, s = blockIdx.x, tid = threadIdx.x;
  const int msz = scf_np8(n) * p;
  double* sL = sm;
  double* sV = sL + msz;
  double* sC = sV + msz;
  __shared__ double socc[SCF_MAX_N], sev[SCF_MAX_N];
  __shared__ double snel;
  scf_load(n, p, a.L_inv, sL);
  scf_load(n, p, a.V + (size_t)s * nn, sV);
  if (tid < n) sev[tid] = a.evals[s * n + tid];
  if (tid < 32) {  // occupations are 0 / 1: the sum is exact in any order
    double t = 0.0;
    for (int j = tid; j < n; j += 32) t += a.occ_prev[s * n + j];
    t = warp_sum(t);
    if (tid == 0) snel = rint(t);
  }
  __syncthreads();
  scf_matmul<true, false>(n, p, sL, sV, sC);  // L^-T V
  if (tid < n) {
    // stable ascending rank (argsort(stable=True) then inverse permutation): ties keep index order; NaN sorts last
    const double e = sev[tid];
    int rank = 0;
    for (int i = 0; i < n; i++) {
      const double o = sev[i];
      const bool less = (o < e) || (e != e && o == o) || ((o == e || (o != o && e != e)) && i < tid);
      rank += less ? 1 : 0;
    }
    const double oc = (double)rank < snel ? 1.0 : 0.0;
    socc[tid] = oc;
    a.mo_occ[s * n + tid] = oc;
  }
  __syncthreads();
  for (int o = tid; o < nn; o += SCF_THREADS) a.mo_coeff[(size_t)s * nn + o] = sC[(o / n) * p + (o % n)];
  // rdm1[i][k] = sum_j C[i][j] occ[j] C[k][j]: the same tensor-tile product with the occupations scaling the k index
  if (tid >= n && tid < scf_np8(n)) socc[tid] = 0.0;
  __syncthreads();
  scf_matmul<false, true>(n, p, sC, sC, sV, socc);
  __syncthreads();
  for (int o = tid; o < nn; o += SCF_THREADS) a.rdm1[(size_t)s * nn + o] = sV[(o / n) * p + (o % n)];
}

}  // namespace gdft

namespace gdft {
// occ[s][j] = 1 for the nelec_s lowest eigenvalues by stable ascending rank (ties keep index order, NaN sorts last), else 0;
// nelec_s = round(sum_j occ_prev[s][j]) -- grad_dft/molecule.py:851-889 without the sort: n comparisons per orbital
__global__ void __launch_bounds__(256) aufbau_occ_kernel(int n, const double* __restrict__ evals, const double* __restrict__ occ_prev,
                                                        double* __restrict__ occ) {
  const int s = blockIdx.y;
  __shared__ double snel;
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int j = threadIdx.x; j < n; j += 32) t += occ_prev[s * n + j];
    t = warp_sum(t);
    if (threadIdx.x == 0) snel = rint(t);
  }
  extern __shared__ double sev[];
  const bool staged = n <= 4096;
  if (staged)
    for (int i = threadIdx.x; i < n; i += blockDim.x) sev[i] = evals[(size_t)s * n + i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double* ev = staged ? sev : evals + (size_t)s * n;
  const double e = ev[j];
  int rank = 0;
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const double o = ev[i];
    rank += ((o < e) || (e != e && o == o) || ((o == e || (o != o && e != e)) && i < j)) ? 1 : 0;
  }
  occ[(size_t)s * n + j] = (double)rank < snel ? 1.0 : 0.0;
}
}  // namespace gdft

extern "C" int gdft_aufbau_occupations(gdft_stream_t stream, int64_t n, const double* evals, const double* occ_prev, double* occ) {
  if (n <= 0 || n > 65536) return GDFT_BAD_SHAPE;
  if (!evals || !occ_prev || !occ) return GDFT_BAD_ARGUMENT;
  dim3 grid((unsigned)((n + 255) / 256), 2);
  gdft::aufbau_occ_kernel<<<grid, 256, (n <= 4096 ? (size_t)n * 8 : 0), static_cast<cudaStream_t>(stream)>>>((int)n, evals, occ_prev, occ);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_scf_stage_max_n(void) { return gdft::SCF_MAX_N; }

extern "C" int gdft_scf_diis_step(gdft_stream_t stream, int64_t n, int m, int cycle, const double* fock, const double* rdm1, const double* overlap,
                                  const double* L_inv, double* fock_vec, double* err_vec, double* gram, double* x_out, double* fock_out,
                                  double* C_out) {
  if (n <= 0 || n > gdft::SCF_MAX_N || m <= 0 || m > gdft::SCF_MAX_M || cycle < 0) return GDFT_BAD_SHAPE;
  if (!fock || !rdm1 || !overlap || !L_inv || !fock_vec || !err_vec || !gram || !x_out || !fock_out || !C_out) return GDFT_BAD_ARGUMENT;
  gdft::DiisArgs a{(int)n, m, cycle, fock, rdm1, overlap, L_inv, fock_vec, err_vec, gram, x_out, fock_out, C_out};
  const size_t smem = (size_t)4 * gdft::scf_np8((int)n) * gdft::scf_pitch((int)n) * sizeof(double);
  GDFT_CUDA_TRY(cudaFuncSetAttribute(gdft::scf_diis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gdft::scf_diis_kernel<<<2, gdft::SCF_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_scf_occupy(gdft_stream_t stream, int64_t n, const double* evals, const double* V, const double* L_inv, const double* occ_prev,
                               double* mo_coeff, double* mo_occ, double* rdm1) {
  if (n <= 0 || n > gdft::SCF_MAX_N) return GDFT_BAD_SHAPE;
  if (!evals || !V || !L_inv || !occ_prev || !mo_coeff || !mo_occ || !rdm1) return GDFT_BAD_ARGUMENT;
  gdft::OccupyArgs a{(int)n, evals, V, L_inv, occ_prev, mo_coeff, mo_occ, rdm1};
  const size_t smem = (size_t)3 * gdft::scf_np8((int)n) * gdft::scf_pitch((int)n) * sizeof(double);
  GDFT_CUDA_TRY(cudaFuncSetAttribute(gdft::scf_occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gdft::scf_occupy_kernel<<<2, gdft::SCF_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
