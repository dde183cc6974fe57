// Row f1: the symmetric eigenproblem inside the SCF iteration
// (grad_dft/utils/eigenproblem.py:26-149, jnp.linalg.eigh per spin after the Cholesky reduction) for the small
// matrices of the H2O/H2-class molecules, where the library path (cuSOLVER syevd: a chain of ~300 tiny kernels,
// ~1.4 ms for two 43 x 43 matrices, plus a host synchronisation for its status word) dominates the iteration and
// cannot be captured in a CUDA graph.  One CTA per matrix; A and the accumulated rotations V live in shared memory
// (n <= 90).  Parallel-order cyclic Jacobi in the Brent-Luk arrangement: the m = 2*ceil(n/2) indices sit in m/2
// adjacent position pairs (2k, 2k+1); a round rotates every pair at once, A <- J^T A J and V <- V J, and then moves
// rows/columns by ONE FIXED position permutation (the round-robin tournament step), so that after m-1 rounds every
// index pair has met once and every index is back where it started.  Consequences for the kernel:
//   - thread <-> 2 x 2 block (k, l) is static: the block is read with two 128-bit shared loads from fixed addresses,
//     no pair tables, no index arithmetic in the loop;
//   - the owner of the diagonal block (k, k) computes the rotation (c_k, s_k) from the registers it has just loaded;
//   - every thread reads its blocks into registers BEFORE the first barrier and writes the rotated blocks to their
//     permuted positions AFTER it: in place, two barriers per round, no second copy of A.
// Sweeps repeat until the off-diagonal mass is below (n eps)^2 of the Frobenius norm (quadratic convergence: 5-8 sweeps).
// Eigenvalues are returned ascending with the matching eigenvector columns (the convention of jnp.linalg.eigh);
// eigenvector signs are arbitrary there as here.  No host synchronisation, no status word: the iteration count is
// bounded and a non-finite input gives non-finite output.  Odd n: index n is a padding row/column of zeros, which no
// rotation ever mixes with the rest (a zero off-diagonal element means "no rotation").
#include <stdlib.h>
#include "common.cuh"

namespace gdft {

#ifdef GDFT_EIG_PROFILE  // tools/eigh_prof.cu: per-phase clock counters of one A thread and one V thread
__device__ long long g_eig_prof[8];
#define EIG_TICK(var) const long long var = clock64()
#define EIG_ACC(slot, t0, t1, who) do { if (who) g_eig_prof[slot] += (t1) - (t0); } while (0)
#else
#define EIG_TICK(var)
#define EIG_ACC(slot, t0, t1, who)
#endif

constexpr int EIG_THREADS = 512;
constexpr int EIG_MAX_N = 90;  // beyond: cuSOLVER through the host framework is faster (n = 104: 3.5 ms vs 6.6 ms)
constexpr int EIG_MAX_SWEEPS = 40;

// Stop when off(A)^2 <= (n eps)^2 ||A||_F^2: below that the off-diagonal mass is rounding noise of the rotations themselves
// (n^2 elements of relative size eps) and further sweeps only churn it -- with a fixed 1e-30 the n = 43 solve ran on for
// twice the sweeps it needed.  Eigenvalue errors are second order in the remaining off-diagonal mass.
__device__ __forceinline__ double eig_tol(int n) {
  const double ne = n * 2.220446049250313e-16;
  return fmax(ne * ne, 1e-30);
}

// where the row/column at position `pos` goes after a round (npair >= 2); position 0 never moves
__device__ __forceinline__ int eig_next_pos(int pos, int npair) {
  if (pos == 0) return 0;
  if (pos & 1) return pos >= 3 ? pos - 2 : 2;
  return (pos >> 1) < npair - 1 ? pos + 2 : pos + 1;
}

// NB / NV: 2 x 2 blocks of A and (row, pair) items of V per thread
template <int NB, int NV>
__global__ void __launch_bounds__(EIG_THREADS) sym_eig_jacobi_kernel(int n, const double* __restrict__ A_in, double* __restrict__ evals,
                                                                     double* __restrict__ evecs, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int npair = (n + 1) / 2, m = 2 * npair;  // m even: every row of sA / sV starts 16-byte aligned
  double* sA = sm;                       // [m][m]
  double* sV = sA + (size_t)m * m;       // [n][m]
  double* sc = sV + (size_t)n * m;       // [npair] cos
  double* ss = sc + npair;               // [npair] sin
  __shared__ double red[EIG_THREADS / 32];
  __shared__ double s_off, s_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* A = A_in + (size_t)blockIdx.x * n * n;

  // symmetrised load (the caller's matrix is symmetric up to round-off; eigh reads one triangle)
  for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
    const int i = idx / m, j = idx - i * m;
    sA[idx] = (i < n && j < n) ? 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]) : 0.0;
  }
  for (int idx = tid; idx < n * m; idx += EIG_THREADS) {
    const int i = idx / m, j = idx - i * m;
    sV[idx] = (i == j) ? 1.0 : 0.0;
  }

  // static work assignment
  int b_src[NB], b_row[NB], b_col[NB], b_k[NB], b_l[NB];  // source offset, destination row offsets / columns (packed), pair ids
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int b = tid + j * EIG_THREADS;
    if (b < npair * npair) {
      const int k = b / npair, l = b - k * npair;
      b_k[j] = k; b_l[j] = l;
      b_src[j] = 2 * k * m + 2 * l;
      const int ra = npair > 1 ? eig_next_pos(2 * k, npair) : 2 * k, rb = npair > 1 ? eig_next_pos(2 * k + 1, npair) : 2 * k + 1;
      const int ca = npair > 1 ? eig_next_pos(2 * l, npair) : 2 * l, cb = npair > 1 ? eig_next_pos(2 * l + 1, npair) : 2 * l + 1;
      b_row[j] = (ra << 16) | rb;
      b_col[j] = (ca << 16) | cb;
    } else {
      b_k[j] = -1; b_l[j] = 0; b_src[j] = 0; b_row[j] = 0; b_col[j] = 0;
    }
  }
  int v_src[NV], v_dst[NV], v_l[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int v = tid + j * EIG_THREADS;
    if (v < n * npair) {
      const int i = v / npair, l = v - i * npair;  // consecutive threads -> consecutive pairs of one row: conflict-free
      v_l[j] = l;
      v_src[j] = i * m + 2 * l;
      const int ca = npair > 1 ? eig_next_pos(2 * l, npair) : 2 * l, cb = npair > 1 ? eig_next_pos(2 * l + 1, npair) : 2 * l + 1;
      v_dst[j] = ((i * m + ca) << 16) | (i * m + cb);
    } else {
      v_l[j] = -1; v_src[j] = 0; v_dst[j] = 0;
    }
  }
  __syncthreads();

  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; sweep++) {
    // off-diagonal and total mass (every index is back at its own position at a sweep boundary)
    double off = 0.0, tot = 0.0;
    for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
      const int i = idx / m, j = idx - i * m;
      const double v = sA[idx];
      tot += v * v;
      if (i != j) off += v * v;
    }
    off = warp_sum(off);
    tot = warp_sum(tot);
    if (lane == 0) red[warp] = off;
    __syncthreads();
    if (tid == 0) { double s = 0; for (int w = 0; w < EIG_THREADS / 32; w++) s += red[w]; s_off = s; }
    __syncthreads();
    if (lane == 0) red[warp] = tot;
    __syncthreads();
    if (tid == 0) { double s = 0; for (int w = 0; w < EIG_THREADS / 32; w++) s += red[w]; s_tot = s; }
    __syncthreads();
    if (!(s_off > eig_tol(n) * s_tot)) {  // also leaves on NaN
      if (tid == 0 && info != nullptr) info[blockIdx.x] = (s_off - s_off == 0.0 && s_tot - s_tot == 0.0) ? sweep : -2  /* inf - inf and NaN - NaN are NaN */;
      break;
    }
    if (tid == 0 && info != nullptr && sweep == EIG_MAX_SWEEPS - 1) info[blockIdx.x] = -1;  // the bound was hit

    for (int r = 0; r < m - 1; r++) {
      // ---- read phase: own blocks and V items into registers; diagonal-block owners publish the rotations ----
      double2 a0[NB], a1[NB], vv[NV];
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (b_k[j] >= 0) {
          a0[j] = *reinterpret_cast<const double2*>(sA + b_src[j]);
          a1[j] = *reinterpret_cast<const double2*>(sA + b_src[j] + m);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; j++)
        if (v_l[j] >= 0) vv[j] = *reinterpret_cast<const double2*>(sV + v_src[j]);
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (b_k[j] >= 0 && b_k[j] == b_l[j]) {
          const double apq = a0[j].y;
          double c = 1.0, s = 0.0;
          if (apq != 0.0) {
            // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = d / b, with one sqrt, one divide and one rsqrt
            const double d = a1[j].y - a0[j].x, b = 2.0 * apq;
            const double den = fabs(d) + sqrt(fma(d, d, b * b));
            double t = den > 0.0 ? fabs(b) / den : 1.0;
            if ((d < 0.0) != (b < 0.0)) t = -t;
            c = rsqrt(fma(t, t, 1.0));
            s = t * c;
          }
          sc[b_k[j]] = c;
          ss[b_k[j]] = s;
        }
      }
      __syncthreads();
      // ---- write phase: rotate (columns by pair l, then rows by pair k) and store at the permuted positions ----
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (b_k[j] >= 0) {
          const double ck = sc[b_k[j]], sk = ss[b_k[j]], cl = sc[b_l[j]], sl = ss[b_l[j]];
          const double tpP = cl * a0[j].x - sl * a0[j].y, tpQ = sl * a0[j].x + cl * a0[j].y;
          const double tqP = cl * a1[j].x - sl * a1[j].y, tqQ = sl * a1[j].x + cl * a1[j].y;
          const int ra = (b_row[j] >> 16) * m, rb = (b_row[j] & 0xffff) * m, ca = b_col[j] >> 16, cb = b_col[j] & 0xffff;
          sA[ra + ca] = ck * tpP - sk * tqP;
          sA[ra + cb] = ck * tpQ - sk * tqQ;
          sA[rb + ca] = sk * tpP + ck * tqP;
          sA[rb + cb] = sk * tpQ + ck * tqQ;
        }
      }
#pragma unroll
      for (int j = 0; j < NV; j++) {
        if (v_l[j] >= 0) {
          const double cl = sc[v_l[j]], sl = ss[v_l[j]];
          sV[v_dst[j] >> 16] = cl * vv[j].x - sl * vv[j].y;
          sV[v_dst[j] & 0xffff] = sl * vv[j].x + cl * vv[j].y;
        }
      }
      __syncthreads();
    }
  }

  // ascending order (ties by index): rank_i = #{j : lambda_j < lambda_i or (== and j < i)}
  int* rank = reinterpret_cast<int*>(sc);  // 2 * npair doubles >= n ints
  __syncthreads();
  for (int i = tid; i < n; i += EIG_THREADS) {
    const double li = sA[i * m + i];
    int rk = 0;
    for (int j = 0; j < n; j++) {
      const double lj = sA[j * m + j];
      rk += (lj < li || (lj == li && j < i)) ? 1 : 0;
    }
    rank[i] = rk;
  }
  __syncthreads();
  double* ev = evals + (size_t)blockIdx.x * n;
  double* vec = evecs + (size_t)blockIdx.x * n * n;
  for (int i = tid; i < n; i += EIG_THREADS) ev[rank[i]] = sA[i * m + i];
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int row = idx / n, col = idx - row * n;
    vec[(size_t)row * n + rank[col]] = sV[row * m + col];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// n <= 64 (the H2O / H2 class): the same Brent-Luk rounds, restructured around what the shared-memory version actually
// spends (ncu: ~200 instructions per warp per round on 4 warps per scheduler, 2000 clocks per round; timing ablation:
// a round's skeleton of two barriers cost 720 clocks, the V update 390, the rotation chain itself almost nothing):
//   - ONE barrier per round.  A is kept in two copies; a round reads one and writes the rotated, permuted blocks into the
//     other, so there is no read/write hazard, and every thread derives the rotations it needs (pair k and pair l of its
//     block) from the diagonal blocks of the copy it reads instead of waiting for a broadcast;
//   - warp specialisation: 8 warps carry the blocks of A, 8 warps carry V;
//   - V never touches shared memory: row i lives in the registers of one warp, lane l holding the position pair
//     (2l, 2l+1); the rotation is local to the lane, the position permutation of a round is one shuffle up and one down,
//     and the rows of a warp are independent chains the scheduler interleaves (no per-row branches);
//   - only the upper triangle of A is stored and updated (blocks k <= l: half the shared-memory traffic and FLOPs);
//     a rotated element whose permuted position falls below the diagonal is stored at the transposed address;
//   - the rotation comes from two rsqrt and no sqrt / divide (measured latencies on B200, tools/fp64_latency.cu: DFMA 9
//     clocks, rsqrt 77, sqrt 102, divide 134, CTA barrier 45): with h = d^2 + b^2, r = rsqrt(h), x = |d| r = cos 2theta,
//     |b| r = sin 2theta:  c = sqrt((1 + x)/2) = u rsqrt(u) with u = (1 + x)/2,  |s| = sin 2theta / (2c) = |b| r rsqrt(u) / 2
//     (no cancellation anywhere; c^2 + s^2 = u + (1 - x^2)/(4u) = 1).
// ---------------------------------------------------------------------------------------------------------------
// rotation of position pair k from the current copy of A: (c, s) with c^2 + s^2 = 1 that annihilates A[2k][2k+1]
__device__ __forceinline__ void eig_pair_rotation(const double* __restrict__ src, int m, int k, double& c, double& s) {
  const int p = 2 * k;
  const double2 top = *reinterpret_cast<const double2*>(src + p * m + p);  // (a_pp, a_pq)
  const double aqq = src[(p + 1) * m + p + 1];
  const double d = aqq - top.x, b = 2.0 * top.y;
  const double h = fma(d, d, b * b);
  c = 1.0; s = 0.0;
  if (top.y != 0.0 && h > 1e-290) {
    const double rh = rsqrt(h);
    const double u = fma(0.5 * fabs(d), rh, 0.5);  // (1 + cos 2theta) / 2 in [1/2, 1]
    const double ru = rsqrt(u);
    c = u * ru;
    s = 0.5 * fabs(b) * rh * ru;
    if ((d < 0.0) != (b < 0.0)) s = -s;
  }
}

// n x n products of the warm start as FP64 tensor tiles (one 8 x 8 tile per warp and round, predicated fragment loads, so
// neither padding nor alignment is asked of the operands): C(i,j) = sum_k opA(i,k) opB(k,j).  As scalar loops the three
// products of a warm start were shared-memory-bandwidth-bound (two 8-byte loads per FMA, n^3 / 8 wavefronts: ~5 us each at
// n = 43, a fifth of the whole solve).  `out(i, j, value)` receives every element with i, j < n.
template <bool TA, bool TB, typename Out>
__device__ __forceinline__ void eig_mma(int n, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, Out out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nt = (n + 7) >> 3;
  for (int tile = warp; tile < nt * nt; tile += EIG_THREADS / 32) {
    const int i = ((tile / nt) << 3) + g, j = ((tile % nt) << 3) + g;
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
    for (int k0 = 0; k0 < n; k0 += 8) {
      const int ka = k0 + t, kb = k0 + 4 + t;
      const double a0 = (i < n && ka < n) ? (TA ? A[ka * lda + i] : A[i * lda + ka]) : 0.0;
      const double a1 = (i < n && kb < n) ? (TA ? A[kb * lda + i] : A[i * lda + kb]) : 0.0;
      const double b0 = (j < n && ka < n) ? (TB ? B[j * ldb + ka] : B[ka * ldb + j]) : 0.0;
      const double b1 = (j < n && kb < n) ? (TB ? B[j * ldb + kb] : B[kb * ldb + j]) : 0.0;
      dmma884(c0, a0, b0);
      dmma884(c1, a1, b1);
    }
    const int jc = ((tile % nt) << 3) + 2 * t;
    if (i < n && jc < n) out(i, jc, c0[0] + c1[0]);
    if (i < n && jc + 1 < n) out(i, jc + 1, c0[1] + c1[1]);
  }
}

template <int NB, int NR>
__global__ void __launch_bounds__(EIG_THREADS) sym_eig_jacobi_small_kernel(int n, const double* __restrict__ A_in, const double* __restrict__ V0_in,
                                                                           double* __restrict__ evals, double* __restrict__ evecs,
                                                                           int* __restrict__ info) {
  // warps 0..7 own the 2 x 2 blocks of A, warps 8..15 own the rows of V (NR rows per warp, in registers).
  // Warm start (V0_in != NULL): the sweeps run on A' = V0^T A V0 for an orthogonal V0 -- the eigenvectors of the previous
  // SCF cycle, which leave A' nearly diagonal (off^2/||A||^2 = 8e-3, 1e-4, 1e-6, ... over the cycles of the H2O-shaped
  // loop against ~0.9 cold: 2-4 sweeps instead of 7) -- and the eigenvectors returned are V0 V'.
  constexpr int NW = EIG_THREADS / 32, NAW = NW / 2, NVW = NW - NAW;
  constexpr int ATHREADS = 32 * NAW;
  extern __shared__ __align__(16) double sm[];
  const int npair = (n + 1) / 2, m = 2 * npair;
  double* sA = sm;  // [2][m][m]: the round reads one copy and writes the other (upper triangles live)
  __shared__ double red[NW];
  __shared__ double s_off, s_tot;
  __shared__ int rank[64];
  __shared__ double slog[128];  // [2][2][32]: (cos, sin) per pair, double-buffered over rounds
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool a_warp = warp < NAW;
  const double* A = A_in + (size_t)blockIdx.x * n * n;

  for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
    const int i = idx / m, j = idx - i * m;
    sA[idx] = (i < n && j < n) ? 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]) : 0.0;
  }
  double* sV0 = sA + (size_t)2 * m * m;  // [n][n], only with a warm start
  const bool warm = V0_in != nullptr;
  if (warm) {
    const double* V0 = V0_in + (size_t)blockIdx.x * n * n;
    double* sT = sA + (size_t)m * m;  // the second copy of A as scratch: T = A V0, [n][m]
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) sV0[idx] = V0[idx];
    __syncthreads();
    // V0 is handed from cycle to cycle as V0 V', which drifts from orthogonality by a few ulps per cycle: one Newton-Schulz
    // step V0 <- V0 + 1/2 V0 (I - V0^T V0) squares the defect away (1e-13 -> 1e-26), so no periodic cold start is needed
    eig_mma<true, false>(n, sV0, n, sV0, n, [&](int i, int j, double v) { sT[i * m + j] = (i == j ? 1.0 : 0.0) - v; });
    __syncthreads();
    {
      constexpr int MAXT = (64 / 8) * (64 / 8) / (EIG_THREADS / 32);  // tiles per warp at n = 64
      double corr[2 * MAXT];
      int cnt = 0;
      eig_mma<false, false>(n, sV0, n, sT, m, [&](int, int, double v) { if (cnt < 2 * MAXT) corr[cnt] = v; cnt++; });
      __syncthreads();
      cnt = 0;
      const int nt = (n + 7) >> 3, g = lane >> 2, t = lane & 3;
      for (int tile = warp; tile < nt * nt; tile += EIG_THREADS / 32) {  // the same walk as eig_mma's, same order of emission
        const int i = ((tile / nt) << 3) + g, jc = ((tile % nt) << 3) + 2 * t;
        if (i < n && jc < n) { sV0[i * n + jc] = fma(0.5, corr[cnt], sV0[i * n + jc]); cnt++; }
        if (i < n && jc + 1 < n) { sV0[i * n + jc + 1] = fma(0.5, corr[cnt], sV0[i * n + jc + 1]); cnt++; }
      }
    }
    __syncthreads();
    eig_mma<false, false>(n, sA, m, sV0, n, [&](int i, int j, double v) { sT[i * m + j] = v; });  // T = A V0
    __syncthreads();
    eig_mma<true, false>(n, sV0, n, sT, m, [&](int i, int j, double v) { if (i <= j) sA[i * m + j] = v; });  // A' = V0^T T (upper triangle)
  }
  // V rows: row i = (warp - NAW) + NVW * q, lane l holds the position pair (2l, 2l+1); rows >= n stay zero
  double vt[NR], vb[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) {
    const int i = (warp - NAW) + NVW * q;
    vt[q] = (!a_warp && i < n && 2 * lane == i) ? 1.0 : 0.0;
    vb[q] = (!a_warp && i < n && 2 * lane + 1 == i) ? 1.0 : 0.0;
  }
  // static block assignment over the upper triangle of the pair grid (A warps)
  int b_k[NB], b_l[NB], b_d[NB][4];  // pair ids; destination offsets of the block's four elements
  const int nblk = npair * (npair + 1) / 2;
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int b = tid + j * ATHREADS;
    b_k[j] = -1; b_l[j] = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) b_d[j][e] = 0;
    if (a_warp && b < nblk) {
      int k = 0, rem = b;
      while (rem >= npair - k) { rem -= npair - k; k++; }
      const int l = k + rem;
      b_k[j] = k; b_l[j] = l;
      const int ra = npair > 1 ? eig_next_pos(2 * k, npair) : 2 * k, rb = npair > 1 ? eig_next_pos(2 * k + 1, npair) : 2 * k + 1;
      const int ca = npair > 1 ? eig_next_pos(2 * l, npair) : 2 * l, cb = npair > 1 ? eig_next_pos(2 * l + 1, npair) : 2 * l + 1;
      const int rr[4] = {ra, ra, rb, rb}, cc[4] = {ca, cb, ca, cb};
#pragma unroll
      for (int e = 0; e < 4; e++) b_d[j][e] = rr[e] <= cc[e] ? rr[e] * m + cc[e] : cc[e] * m + rr[e];
    }
  }
  const bool v_first = lane == 0, v_last = lane == npair - 1, single_pair = npair == 1;
  int cur = 0;
  __syncthreads();

  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; sweep++) {
    const double* Ac = sA + (size_t)cur * m * m;
    double off = 0.0, tot = 0.0;
    for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
      const int i = idx / m, j = idx - i * m;
      if (i <= j) {
        const double v = Ac[idx];
        if (i == j) tot += v * v;
        else off += 2.0 * v * v;
      }
    }
    tot += off;
    off = warp_sum(off);
    tot = warp_sum(tot);
    if (lane == 0) red[warp] = off;
    __syncthreads();
    if (tid == 0) { double s = 0; for (int w = 0; w < NW; w++) s += red[w]; s_off = s; }
    __syncthreads();
    if (lane == 0) red[warp] = tot;
    __syncthreads();
    if (tid == 0) { double s = 0; for (int w = 0; w < NW; w++) s += red[w]; s_tot = s; }
    __syncthreads();
    if (!(s_off > eig_tol(n) * s_tot)) {  // also leaves on NaN
      if (tid == 0 && info != nullptr) info[blockIdx.x] = (s_off - s_off == 0.0 && s_tot - s_tot == 0.0) ? sweep : -2  /* inf - inf and NaN - NaN are NaN */;
      break;
    }
    if (tid == 0 && info != nullptr && sweep == EIG_MAX_SWEEPS - 1) info[blockIdx.x] = -1;  // the bound was hit

    // Round r: (1) the rotation warp (warp 0, lane = pair) derives the npair rotations from the diagonal blocks of the
    // copy of A being read and publishes them; barrier 0 (all warps); (2) the A warps write the rotated, permuted blocks
    // into the other copy while the V warps rotate and permute their rows; barrier 1 (A warps only: the V warps run on
    // into the next round and meet the others again at its barrier 0; the rotations are double-buffered for that).
    for (int r = 0; r < m - 1; r++) {
      const double* src = sA + (size_t)cur * m * m;
      double* dst = sA + (size_t)(cur ^ 1) * m * m;
      double* lc = slog + (r & 1) * 64;
      double* ls = lc + 32;
      EIG_TICK(t_begin);
      if (warp == 0 && lane < npair) {
        double c, s2;
        eig_pair_rotation(src, m, lane, c, s2);
        lc[lane] = c;
        ls[lane] = s2;
      }
      EIG_TICK(t_rot);
      asm volatile("bar.sync 0, %0;" ::"n"(EIG_THREADS) : "memory");
      EIG_TICK(t_bar0);
      if (a_warp) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
          if (b_k[j] >= 0) {
            const int k = b_k[j], l = b_l[j];
            const int o = 2 * k * m + 2 * l;
            const double2 a0 = *reinterpret_cast<const double2*>(src + o);
            double2 a1 = *reinterpret_cast<const double2*>(src + o + m);
            if (k == l) a1.x = a0.y;  // the mirror image of (p, q): the lower triangle is not maintained
            const double ck = lc[k], sk = ls[k], cl = lc[l], sl = ls[l];
            const double tpP = cl * a0.x - sl * a0.y, tpQ = sl * a0.x + cl * a0.y;
            const double tqP = cl * a1.x - sl * a1.y, tqQ = sl * a1.x + cl * a1.y;
            dst[b_d[j][0]] = ck * tpP - sk * tqP;
            dst[b_d[j][1]] = ck * tpQ - sk * tqQ;
            if (k != l) dst[b_d[j][2]] = sk * tpP + ck * tqP;  // diagonal block: the same slot as element 1
            dst[b_d[j][3]] = sk * tpQ + ck * tqQ;
          }
        }
        EIG_TICK(t_work);
        asm volatile("bar.sync 1, %0;" ::"n"(ATHREADS) : "memory");
        EIG_TICK(t_bar1);
        EIG_ACC(0, t_begin, t_rot, tid == 0 && blockIdx.x == 0);   // rotation warp
        EIG_ACC(1, t_rot, t_bar0, tid == 0 && blockIdx.x == 0);    // barrier 0
        EIG_ACC(2, t_bar0, t_work, tid == 0 && blockIdx.x == 0);   // A update
        EIG_ACC(3, t_work, t_bar1, tid == 0 && blockIdx.x == 0);   // barrier 1
        EIG_ACC(4, 0, 1, tid == 0 && blockIdx.x == 0);             // rounds
      } else {
        // V <- V J, then the position permutation: top'_0 = top_0, top'_1 = bot_0, top'_l = top_{l-1};
        // bot'_l = bot_{l+1}, bot'_{npair-1} = top_{npair-1}: one shuffle up (of bot at lane 0, top elsewhere), one down
        const double cl = lane < npair ? lc[lane] : 1.0, sl = lane < npair ? ls[lane] : 0.0;
        // straight-line over the rows (npair == 1 never permutes: v_first == v_last == lane 0 keeps t, and the value
        // shuffled down into bot is overwritten only where it matters): the NR chains interleave in the scheduler
        double t[NR], b[NR], up[NR], dn[NR];
#pragma unroll
        for (int q = 0; q < NR; q++) { t[q] = cl * vt[q] - sl * vb[q]; b[q] = sl * vt[q] + cl * vb[q]; }
#pragma unroll
        for (int q = 0; q < NR; q++) { up[q] = __shfl_up_sync(0xffffffffu, v_first ? b[q] : t[q], 1); dn[q] = __shfl_down_sync(0xffffffffu, b[q], 1); }
#pragma unroll
        for (int q = 0; q < NR; q++) {
          vt[q] = v_first ? t[q] : up[q];
          vb[q] = single_pair ? b[q] : (v_last ? t[q] : dn[q]);
        }
        EIG_TICK(t_work);
        EIG_ACC(5, t_bar0, t_work, tid == 32 * NAW && blockIdx.x == 0);   // V update
        EIG_ACC(6, t_begin, t_bar0, tid == 32 * NAW && blockIdx.x == 0);  // V wait at barrier 0
      }
      cur ^= 1;
    }
    __syncthreads();  // sweep boundary: the V warps rejoin; A complete for the convergence test
  }

  const double* Ac = sA + (size_t)cur * m * m;
  for (int i = tid; i < n; i += EIG_THREADS) {
    const double li = Ac[i * m + i];
    int rk = 0;
    for (int j = 0; j < n; j++) {
      const double lj = Ac[j * m + j];
      rk += (lj < li || (lj == li && j < i)) ? 1 : 0;
    }
    rank[i] = rk;
  }
  __syncthreads();
  double* ev = evals + (size_t)blockIdx.x * n;
  double* vec = evecs + (size_t)blockIdx.x * n * n;
  for (int i = tid; i < n; i += EIG_THREADS) ev[rank[i]] = Ac[i * m + i];
  // eigenvectors, columns in ascending order: straight to global memory, or (warm start) into the free copy of A first
  double* sVp = sA + (size_t)(cur ^ 1) * m * m;  // [n][n]
  if (!a_warp) {
#pragma unroll
    for (int q = 0; q < NR; q++) {
      const int i = (warp - NAW) + NVW * q;
      if (i < n) {
        double* dstrow = warm ? sVp + (size_t)i * n : vec + (size_t)i * n;
        if (2 * lane < n) dstrow[rank[2 * lane]] = vt[q];
        if (2 * lane + 1 < n) dstrow[rank[2 * lane + 1]] = vb[q];
      }
    }
  }
  if (warm) {
    __syncthreads();
    eig_mma<false, false>(n, sV0, n, sVp, n, [&](int i, int j, double v) { vec[(size_t)i * n + j] = v; });  // V = V0 V'
  }
}

template <int NB, int NR>
static int launch_eig_small(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  const int m = 2 * ((n + 1) / 2);
  const size_t smem = ((size_t)2 * m * m + (V0 ? (size_t)n * n : 0)) * 8;
  GDFT_CUDA_TRY((cudaFuncSetAttribute(sym_eig_jacobi_small_kernel<NB, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)2 * m * m + (size_t)n * n) * 8))));
  sym_eig_jacobi_small_kernel<NB, NR><<<(unsigned)batch, EIG_THREADS, smem, stream>>>(n, A, V0, evals, evecs, info);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

static size_t eig_smem(int n) {
  const int npair = (n + 1) / 2, m = 2 * npair;
  return ((size_t)m * m + (size_t)n * m + 2 * npair) * 8 + 64;
}

template <int NB, int NV>
static int launch_eig(cudaStream_t stream, int64_t batch, int n, const double* A, double* evals, double* evecs, int* info) {
  const size_t smem = eig_smem(n);
  GDFT_CUDA_TRY(cudaFuncSetAttribute(sym_eig_jacobi_kernel<NB, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sym_eig_jacobi_kernel<NB, NV><<<(unsigned)batch, EIG_THREADS, smem, stream>>>(n, A, evals, evecs, info);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

}  // namespace gdft

using namespace gdft;

namespace gdft {
int sym_eigh_cluster(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info);
int sym_eigh_cluster_max_n();
}

extern "C" int gdft_sym_eigh_max_n(void) { return sym_eigh_cluster_max_n(); }

extern "C" int gdft_sym_eigh_ex(gdft_stream_t stream_, int64_t batch, int64_t n, const double* A, const double* V0, double* evals,
                                double* evecs, int* info) {
  if (batch <= 0 || n <= 0 || n > sym_eigh_cluster_max_n() || batch > 8191) return GDFT_BAD_SHAPE;
  if (!A || !evals || !evecs) return GDFT_BAD_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // upper blocks per A thread (256 threads): npair (npair + 1) / 2 <= 253 up to n = 44, <= 528 up to n = 64; V rows per V warp: n / 8
  if (n <= 16) return launch_eig_small<1, 2>(stream, batch, (int)n, A, V0, evals, evecs, info);
  if (n <= 32) return launch_eig_small<1, 4>(stream, batch, (int)n, A, V0, evals, evecs, info);
  if (n <= 44) return launch_eig_small<1, 6>(stream, batch, (int)n, A, V0, evals, evecs, info);
  if (n <= 48) return launch_eig_small<3, 6>(stream, batch, (int)n, A, V0, evals, evecs, info);
  if (n <= 64) return launch_eig_small<3, 8>(stream, batch, (int)n, A, V0, evals, evecs, info);
  if (n <= EIG_MAX_N) {
    const char* e = getenv("GDFT_EIGH_ONE_CTA");  // 65..90: the one-CTA kernel (A and V in shared memory, cold start only) on request
    if (e && e[0] == '1') return launch_eig<4, 8>(stream, batch, (int)n, A, evals, evecs, info);
  }
  return sym_eigh_cluster(stream, batch, (int)n, A, V0, evals, evecs, info);  // one 8-CTA cluster per matrix (eigh_cluster.cu)
}

extern "C" int gdft_sym_eigh_warm(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, const double* V0, double* evals,
                                  double* evecs) {
  return gdft_sym_eigh_ex(stream, batch, n, A, V0, evals, evecs, nullptr);
}

extern "C" int gdft_sym_eigh(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, double* evals, double* evecs) {
  return gdft_sym_eigh_ex(stream, batch, n, A, nullptr, evals, evecs, nullptr);
}
