// Row f1: the symmetric eigenproblem inside the SCF iteration
// (grad_dft/utils/eigenproblem.py:26-149, jnp.linalg.eigh per spin after the Cholesky reduction) for the small
// matrices of the H2O/H2-class molecules, where the library path (cuSOLVER syevd: a chain of ~300 tiny kernels,
// ~1.4 ms for two 43 x 43 matrices, plus a host synchronisation for its status word) dominates the iteration and
// cannot be captured in a CUDA graph.  One CTA per matrix; A and the accumulated rotations V live in shared memory
// (n <= 104); parallel-order cyclic Jacobi: every round applies n/2 disjoint Givens rotations (round-robin pairing)
// as a column pass A <- A J, V <- V J and a row pass A <- J^T A; sweeps repeat until the off-diagonal mass is below
// 1e-30 of the Frobenius norm (quadratic convergence: 6-9 sweeps).  Eigenvalues are returned ascending with the matching
// eigenvector columns (the convention of jnp.linalg.eigh); eigenvector signs are arbitrary there as here.  No host
// synchronisation, no status word: the iteration count is bounded and a non-finite input gives non-finite output.
#include "common.cuh"

namespace gdft {

constexpr int EIG_THREADS = 512;
constexpr int EIG_MAX_N = 104;
constexpr int EIG_MAX_SWEEPS = 40;

__global__ void __launch_bounds__(EIG_THREADS) sym_eig_jacobi_kernel(int n, const double* __restrict__ A_in, double* __restrict__ evals,
                                                                     double* __restrict__ evecs) {
  extern __shared__ __align__(16) double sm[];
  const int pitch = n | 1;  // odd pitch: column walks are bank-conflict-free
  double* sA = sm;
  double* sV = sA + (size_t)n * pitch;
  double* sc = sV + (size_t)n * pitch;   // [npair] cos
  double* ss = sc + (n + 1) / 2 + 1;     // [npair] sin
  int* sp = reinterpret_cast<int*>(ss + (n + 1) / 2 + 1);  // [npair] p index, then [npair] q index
  __shared__ double red[EIG_THREADS / 32];
  __shared__ double s_off, s_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* A = A_in + (size_t)blockIdx.x * n * n;
  const int npair = (n + 1) / 2, m = 2 * npair;
  int* sq = sp + npair;

  // symmetrised load (the caller's matrix is symmetric up to round-off; eigh reads one triangle)
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int i = idx / n, j = idx - i * n;
    sA[i * pitch + j] = 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]);
    sV[i * pitch + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();

  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; sweep++) {
    // off-diagonal and total mass
    double off = 0.0, tot = 0.0;
    for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
      const int i = idx / n, j = idx - i * n;
      const double v = sA[i * pitch + j];
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
      // round-robin pairing of m players (player m-1 fixed); a pair touching the padding index (>= n) is skipped
      if (tid < npair) {
        int p, q;
        if (tid == 0) { p = m - 1; q = r; }
        else { p = (r + tid) % (m - 1); q = (r - tid + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = sA[p * pitch + q];
          if (apq != 0.0) {
            const double tau = (sA[q * pitch + q] - sA[p * pitch + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = p;  // marks "no rotation"
        }
        sc[tid] = c; ss[tid] = s; sp[tid] = p; sq[tid] = q;
      }
      __syncthreads();
      // column pass: (x_ip, x_iq) <- (c x_ip - s x_iq, s x_ip + c x_iq) for X = A and X = V
      for (int idx = tid; idx < npair * n; idx += EIG_THREADS) {
        const int k = idx / n, i = idx - k * n;
        const int p = sp[k], q = sq[k];
        if (p == q) continue;
        const double c = sc[k], s = ss[k];
        const double ap = sA[i * pitch + p], aq = sA[i * pitch + q];
        sA[i * pitch + p] = c * ap - s * aq;
        sA[i * pitch + q] = s * ap + c * aq;
        const double vp = sV[i * pitch + p], vq = sV[i * pitch + q];
        sV[i * pitch + p] = c * vp - s * vq;
        sV[i * pitch + q] = s * vp + c * vq;
      }
      __syncthreads();
      // row pass: (a_pj, a_qj) <- (c a_pj - s a_qj, s a_pj + c a_qj)
      for (int idx = tid; idx < npair * n; idx += EIG_THREADS) {
        const int k = idx / n, j = idx - k * n;
        const int p = sp[k], q = sq[k];
        if (p == q) continue;
        const double c = sc[k], s = ss[k];
        const double ap = sA[p * pitch + j], aq = sA[q * pitch + j];
        sA[p * pitch + j] = c * ap - s * aq;
        sA[q * pitch + j] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }

  // ascending order (ties by index): rank_i = #{j : lambda_j < lambda_i or (== and j < i)}
  for (int i = tid; i < n; i += EIG_THREADS) {
    const double li = sA[i * pitch + i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const double lj = sA[j * pitch + j];
      rank += (lj < li || (lj == li && j < i)) ? 1 : 0;
    }
    reinterpret_cast<int*>(sc)[i] = rank;
  }
  __syncthreads();
  const int* rank = reinterpret_cast<const int*>(sc);
  double* ev = evals + (size_t)blockIdx.x * n;
  double* vec = evecs + (size_t)blockIdx.x * n * n;
  for (int i = tid; i < n; i += EIG_THREADS) ev[rank[i]] = sA[i * pitch + i];
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int row = idx / n, col = idx - row * n;
    vec[(size_t)row * n + rank[col]] = sV[row * pitch + col];
  }
}

static size_t eig_smem(int n) {
  const int pitch = n | 1, npair = (n + 1) / 2;
  return (size_t)2 * n * pitch * 8 + (size_t)2 * (npair + 1) * 8 + (size_t)2 * npair * 4 + 64;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_sym_eigh_max_n(void) { return EIG_MAX_N; }

extern "C" int gdft_sym_eigh(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, double* evals, double* evecs) {
  if (batch <= 0 || n <= 0 || n > EIG_MAX_N || batch > 65535) return GDFT_BAD_SHAPE;
  if (!A || !evals || !evecs) return GDFT_BAD_ARGUMENT;
  const size_t smem = eig_smem((int)n);
  GDFT_CUDA_TRY(cudaFuncSetAttribute(sym_eig_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sym_eig_jacobi_kernel<<<(unsigned)batch, EIG_THREADS, smem, static_cast<cudaStream_t>(stream)>>>((int)n, A, evals, evecs);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
