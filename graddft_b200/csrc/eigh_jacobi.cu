// Row f1: the symmetric eigenproblem inside the SCF iteration
// (grad_dft/utils/eigenproblem.py:26-149, jnp.linalg.eigh per spin after the Cholesky reduction) for the small
// matrices of the H2O/H2-class molecules, where the library path (cuSOLVER syevd: a chain of ~300 tiny kernels,
// ~1.4 ms for two 43 x 43 matrices, plus a host synchronisation for its status word) dominates the iteration and
// cannot be captured in a CUDA graph.  One CTA per matrix; A and the accumulated rotations V live in shared memory
// (n <= 104).  Parallel-order cyclic Jacobi in the Brent-Luk arrangement: the m = 2*ceil(n/2) indices sit in m/2
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

constexpr int EIG_THREADS = 512;
constexpr int EIG_MAX_N = 104;
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
                                                                     double* __restrict__ evecs) {
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
    if (!(s_off > eig_tol(n) * s_tot)) break;  // also leaves on NaN

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
// n <= 64 (the H2O / H2 class): the same Brent-Luk rounds with more savings (ncu on the shared-memory version: ~200
// instructions per warp per round on 4 warps per scheduler, 2000 clocks per round against a ~250-clock rotation chain).
//   - warp specialisation: 8 warps carry A (the serial chain of a round), 8 warps carry V and trail by one barrier: the
//     V update of round r overlaps the read phase and rotation chain of round r+1 (measured DFMA latency 9 clocks,
//     rsqrt 77, sqrt 102, divide 134, CTA barrier 45: tools/fp64_latency.cu);
//   - V never touches shared memory: row i lives in the registers of one warp, lane l holding the position pair
//     (2l, 2l+1); the rotation is local to the lane and the position permutation of a round is two warp shuffles;
//   - only the upper triangle of A is stored and updated (blocks k <= l: half the shared-memory traffic and FLOPs);
//     a rotated element whose permuted position falls below the diagonal is stored at the transposed address;
//   - the rotation comes from two rsqrt and no sqrt / divide: with h = d^2 + b^2, r = rsqrt(h), x = |d| r = cos 2theta
//     and |b| r = sin 2theta:  c = sqrt((1 + x)/2) = u rsqrt(u) with u = (1 + x)/2,  |s| = sin 2theta / (2c) = |b| r rsqrt(u) / 2
//     (no cancellation anywhere; c^2 + s^2 = u + (1 - x^2)/(4u) = 1).  The rotation is the serial part of a round.
// ---------------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(EIG_THREADS) sym_eig_jacobi_small_kernel(int n, const double* __restrict__ A_in, double* __restrict__ evals,
                                                                           double* __restrict__ evecs, int dbg) {
  // warps 0..7 own the 2 x 2 blocks of A (the latency chain of a round); warps 8..15 own V and follow one barrier behind
  constexpr int NW = EIG_THREADS / 32, NAW = NW / 2, NVW = NW - NAW, NR = 8;  // NR: V rows per V warp (n <= 64)
  constexpr int ATHREADS = 32 * NAW;
  extern __shared__ __align__(16) double sm[];
  const int npair = (n + 1) / 2, m = 2 * npair;
  double* sA = sm;                  // [m][m], upper triangle live
  double* slog = sA + (size_t)m * m;  // [2][2][32]: (cos, sin) of the round, double-buffered over rounds
  __shared__ double red[NW];
  __shared__ double s_off, s_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool a_warp = warp < NAW;
  const double* A = A_in + (size_t)blockIdx.x * n * n;

  for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
    const int i = idx / m, j = idx - i * m;
    sA[idx] = (i < n && j < n) ? 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]) : 0.0;
  }
  // V rows in registers (V warps): row i = (warp - NAW) + NVW * q, lane l holds the position pair (2l, 2l+1)
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
  __syncthreads();

  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; sweep++) {
    double off = 0.0, tot = 0.0;
    for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
      const int i = idx / m, j = idx - i * m;
      if (i <= j) {
        const double v = sA[idx];
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
    if (dbg ? sweep >= 8 : !(s_off > eig_tol(n) * s_tot)) break;  // also leaves on NaN

    for (int r = 0; r < m - 1; r++) {
      double* lc = slog + (r & 1) * 64;
      double* ls = lc + 32;
      if (a_warp) {
        // ---- read phase; the owners of the diagonal blocks publish the rotations of this round ----
        double2 a0[NB], a1[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) {
          if (b_k[j] >= 0) {
            const int src = 2 * b_k[j] * m + 2 * b_l[j];
            a0[j] = *reinterpret_cast<const double2*>(sA + src);
            a1[j] = *reinterpret_cast<const double2*>(sA + src + m);
            if (b_k[j] == b_l[j]) {
              a1[j].x = a0[j].y;  // the mirror image of (p, q): the lower triangle is not maintained
              const double apq = a0[j].y;
              double c = 1.0, s = 0.0;
              const double d = a1[j].y - a0[j].x, b = 2.0 * apq;
              const double h = fma(d, d, b * b);
              if (dbg & 2) { c = 0.8; s = 0.6; }
              else if (apq != 0.0 && h > 1e-290) {
                const double rh = rsqrt(h);
                const double u = fma(0.5 * fabs(d), rh, 0.5);  // (1 + cos 2theta) / 2 in [1/2, 1]
                const double ru = rsqrt(u);
                c = u * ru;
                s = 0.5 * fabs(b) * rh * ru;
                if ((d < 0.0) != (b < 0.0)) s = -s;
              }
              lc[b_k[j]] = c;
              ls[b_k[j]] = s;
            }
          }
        }
        asm volatile("bar.sync 0, %0;" ::"n"(EIG_THREADS) : "memory");  // barrier 0, all warps: rotations visible, every block read
        // ---- write phase: rotate (columns by pair l, then rows by pair k), store at the permuted positions ----
#pragma unroll
        for (int j = 0; j < NB; j++) {
          if (b_k[j] >= 0 && !(dbg & 4)) {
            const double ck = lc[b_k[j]], sk = ls[b_k[j]], cl = lc[b_l[j]], sl = ls[b_l[j]];
            const double tpP = cl * a0[j].x - sl * a0[j].y, tpQ = sl * a0[j].x + cl * a0[j].y;
            const double tqP = cl * a1[j].x - sl * a1[j].y, tqQ = sl * a1[j].x + cl * a1[j].y;
            sA[b_d[j][0]] = ck * tpP - sk * tqP;
            sA[b_d[j][1]] = ck * tpQ - sk * tqQ;
            if (b_k[j] != b_l[j]) sA[b_d[j][2]] = sk * tpP + ck * tqP;  // diagonal block: the same slot as element 1
            sA[b_d[j][3]] = sk * tpQ + ck * tqQ;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ATHREADS) : "memory");  // A warps only: the next round may read A
      } else {
        asm volatile("bar.sync 0, %0;" ::"n"(EIG_THREADS) : "memory");  // barrier 0 (the V warps' side of it)
        // V <- V J, then the position permutation: top'_0 = top_0, top'_1 = bot_0, top'_l = top_{l-1};
        // bot'_l = bot_{l+1}, bot'_{npair-1} = top_{npair-1}.  Runs while the A warps are already in the next round's
        // read phase: the rotations of round r+1 go to the other half of slog, and the A warps cannot pass barrier 0 of
        // round r+1 (after which half r is rewritten) before this warp arrives there.
        const double cl = lane < npair ? lc[lane] : 1.0, sl = lane < npair ? ls[lane] : 0.0;
#pragma unroll
        for (int q = 0; q < NR; q++) {
          if ((warp - NAW) + NVW * q < n && !(dbg & 1)) {  // warp-uniform
            const double t = cl * vt[q] - sl * vb[q], b = sl * vt[q] + cl * vb[q];
            if (npair > 1) {
              const double t_up = __shfl_up_sync(0xffffffffu, t, 1), b_up = __shfl_up_sync(0xffffffffu, b, 1);
              const double b_dn = __shfl_down_sync(0xffffffffu, b, 1);
              vt[q] = lane == 0 ? t : lane == 1 ? b_up : t_up;
              vb[q] = lane == npair - 1 ? t : b_dn;
            } else {
              vt[q] = t; vb[q] = b;
            }
          }
        }
      }
    }
    __syncthreads();  // sweep boundary: A complete for the convergence test
  }

  int* rank = reinterpret_cast<int*>(slog);  // 128 doubles >= 64 ints
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
  if (!a_warp) {
#pragma unroll
    for (int q = 0; q < NR; q++) {
      const int i = (warp - NAW) + NVW * q;
      if (i < n) {
        if (2 * lane < n) vec[(size_t)i * n + rank[2 * lane]] = vt[q];
        if (2 * lane + 1 < n) vec[(size_t)i * n + rank[2 * lane + 1]] = vb[q];
      }
    }
  }
}

template <int NB>
static int launch_eig_small(cudaStream_t stream, int64_t batch, int n, const double* A, double* evals, double* evecs) {
  const int m = 2 * ((n + 1) / 2);
  const size_t smem = ((size_t)m * m + 128) * 8;
  GDFT_CUDA_TRY(cudaFuncSetAttribute(sym_eig_jacobi_small_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dbg = 0;
  if (const char* e = getenv("GDFT_EIG_DBG")) dbg = atoi(e);  // timing experiments only (results are wrong when set)
  sym_eig_jacobi_small_kernel<NB><<<(unsigned)batch, EIG_THREADS, smem, stream>>>(n, A, evals, evecs, dbg);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

static size_t eig_smem(int n) {
  const int npair = (n + 1) / 2, m = 2 * npair;
  return ((size_t)m * m + (size_t)n * m + 2 * npair) * 8 + 64;
}

template <int NB, int NV>
static int launch_eig(cudaStream_t stream, int64_t batch, int n, const double* A, double* evals, double* evecs) {
  const size_t smem = eig_smem(n);
  GDFT_CUDA_TRY(cudaFuncSetAttribute(sym_eig_jacobi_kernel<NB, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sym_eig_jacobi_kernel<NB, NV><<<(unsigned)batch, EIG_THREADS, smem, stream>>>(n, A, evals, evecs);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_sym_eigh_max_n(void) { return EIG_MAX_N; }

extern "C" int gdft_sym_eigh(gdft_stream_t stream_, int64_t batch, int64_t n, const double* A, double* evals, double* evecs) {
  if (batch <= 0 || n <= 0 || n > EIG_MAX_N || batch > 65535) return GDFT_BAD_SHAPE;
  if (!A || !evals || !evecs) return GDFT_BAD_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // items per thread: ceil(npair^2 / 512) blocks, ceil(n * npair / 512) V items
  if (n <= 44) return launch_eig_small<1>(stream, batch, (int)n, A, evals, evecs);  // npair <= 22: 253 upper blocks on 256 threads
  if (n <= 64) return launch_eig_small<3>(stream, batch, (int)n, A, evals, evecs);  // npair <= 32: 528 blocks
  if (n <= 90) return launch_eig<4, 8>(stream, batch, (int)n, A, evals, evecs);
  return launch_eig<6, 11>(stream, batch, (int)n, A, evals, evecs);
}
