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
// Sweeps repeat until the off-diagonal mass is below 1e-30 of the Frobenius norm (quadratic convergence: 6-10 sweeps).
// Eigenvalues are returned ascending with the matching eigenvector columns (the convention of jnp.linalg.eigh);
// eigenvector signs are arbitrary there as here.  No host synchronisation, no status word: the iteration count is
// bounded and a non-finite input gives non-finite output.  Odd n: index n is a padding row/column of zeros, which no
// rotation ever mixes with the rest (a zero off-diagonal element means "no rotation").
#include "common.cuh"

namespace gdft {

constexpr int EIG_THREADS = 512;
constexpr int EIG_MAX_N = 104;
constexpr int EIG_MAX_SWEEPS = 40;

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
    if (!(s_off > 1e-30 * s_tot)) break;  // also leaves on NaN

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
  if (n <= 44) return launch_eig<1, 2>(stream, batch, (int)n, A, evals, evecs);
  if (n <= 64) return launch_eig<2, 4>(stream, batch, (int)n, A, evals, evecs);
  if (n <= 90) return launch_eig<4, 8>(stream, batch, (int)n, A, evals, evecs);
  return launch_eig<6, 11>(stream, batch, (int)n, A, evals, evecs);
}
