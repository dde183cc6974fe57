// K3 (ERI sweep: Coulomb J, optional K, E_J), K6 (XC quadrature fwd/bwd) and the n x n predictor glue.
// All of these are HBM-streaming kernels: 128-bit coalesced loads, warp-shuffle reductions, fixed-order
// second-stage reductions (bitwise reproducible).
#include "common.cuh"

namespace gdft {

// ---------------------------------------------------------------------------------------------------
// J[row] = sum_col eri[row][col] P[col],  row = (p,q), col = (r,t)        grad_dft/molecule.py:811
// 8 warps x 4 rows per CTA pass; P streamed through shared memory in 2048-double chunks so the
// ERI stream is the only global traffic that matters.
// ---------------------------------------------------------------------------------------------------
constexpr int ERI_THREADS = 256;
constexpr int ERI_CHUNK = 2048;

// RPW rows per warp: 4 for large tensors (P chunk reused by four rows per shared-memory read); 1 for the small ones
// (n = 43: 1849 rows are 58 CTAs at four rows per warp, a latency-bound 59 us for 27 MB; one row per warp gives 232 CTAs).
// The per-row summation order (lane-strided columns in chunk order, then the warp tree) is the same for both.
template <bool VEC, int RPW>
__global__ void __launch_bounds__(ERI_THREADS) eri_j_kernel(int64_t R, int64_t C, const double* __restrict__ eri,
                                                           const double* __restrict__ P, double* __restrict__ J) {
  constexpr int ROWS_PER_CTA = 8 * RPW;
  __shared__ __align__(16) double sP[ERI_CHUNK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nblocks = (R + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t row_base = blk * ROWS_PER_CTA + warp * RPW;
    double acc[RPW];
    const double* rowp[RPW];
#pragma unroll
    for (int i = 0; i < RPW; i++) {
      acc[i] = 0.0;
      const int64_t row = row_base + i < R ? row_base + i : R - 1;  // clamp: duplicates are discarded below
      rowp[i] = eri + row * C;
    }
    for (int64_t c0 = 0; c0 < C; c0 += ERI_CHUNK) {
      const int len = (int)min((int64_t)ERI_CHUNK, C - c0);
      __syncthreads();
      for (int i = tid; i < len; i += ERI_THREADS) sP[i] = P[c0 + i];
      __syncthreads();
      if (VEC) {
        const int len2 = len >> 1;  // C even => len even
#pragma unroll(RPW == 1 ? 8 : 2)
        for (int i = lane; i < len2; i += 32) {
          const double2 pv = reinterpret_cast<const double2*>(sP)[i];
          double2 e[RPW];
#pragma unroll
          for (int q = 0; q < RPW; q++) e[q] = __ldcs(reinterpret_cast<const double2*>(rowp[q] + c0) + i);
#pragma unroll
          for (int q = 0; q < RPW; q++) acc[q] = fma(e[q].x, pv.x, fma(e[q].y, pv.y, acc[q]));
        }
      } else {
#pragma unroll(RPW == 1 ? 8 : 2)
        for (int i = lane; i < len; i += 32) {
          const double pv = sP[i];
#pragma unroll
          for (int q = 0; q < RPW; q++) acc[q] = fma(__ldcs(rowp[q] + c0 + i), pv, acc[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const double s = warp_sum(acc[q]);
      if (lane == 0 && row_base + q < R) J[row_base + q] = s;
    }
  }
}

// Pbar[col] = sum_row Jbar[row] eri[row][col]; grid (col chunks, row splits) -> part[split][C]
constexpr int ERIT_THREADS = 256;
template <bool VEC>
__global__ void __launch_bounds__(ERIT_THREADS) eri_jt_kernel(int64_t R, int64_t C, int64_t rows_per_split,
                                                             const double* __restrict__ eri, const double* __restrict__ Jbar,
                                                             double* __restrict__ part) {
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split, r1 = min(R, r0 + rows_per_split);
  if (VEC) {
    const int64_t c = ((int64_t)blockIdx.x * ERIT_THREADS + threadIdx.x) * 2;
    if (c >= C) return;
    double2 acc = make_double2(0.0, 0.0);
    int64_t r = r0;
    for (; r + 4 <= r1; r += 4) {
      double2 e[4];
#pragma unroll
      for (int q = 0; q < 4; q++) e[q] = __ldcs(reinterpret_cast<const double2*>(eri + (r + q) * C + c));
#pragma unroll
      for (int q = 0; q < 4; q++) { const double j = __ldg(Jbar + r + q); acc.x = fma(j, e[q].x, acc.x); acc.y = fma(j, e[q].y, acc.y); }
    }
    for (; r < r1; r++) {
      const double2 e = __ldcs(reinterpret_cast<const double2*>(eri + r * C + c));
      const double j = __ldg(Jbar + r);
      acc.x = fma(j, e.x, acc.x); acc.y = fma(j, e.y, acc.y);
    }
    *reinterpret_cast<double2*>(part + (int64_t)blockIdx.y * C + c) = acc;
  } else {
    const int64_t c = (int64_t)blockIdx.x * ERIT_THREADS + threadIdx.x;
    if (c >= C) return;
    double acc = 0.0;
    for (int64_t r = r0; r < r1; r++) acc = fma(__ldg(Jbar + r), __ldcs(eri + r * C + c), acc);
    part[(int64_t)blockIdx.y * C + c] = acc;
  }
}

__global__ void sum_splits_kernel(int64_t C, int splits, const double* __restrict__ part, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int s = 0; s < splits; s++) acc += part[(int64_t)s * C + c];
  out[c] = acc;
}

// out[0] = scale * sum_i a[i] b[i]   (single CTA, fixed order)
__global__ void __launch_bounds__(1024) dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                  double scale, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc = fma(a[i], b[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = scale * v;
  }
}

// out[0] = E_nuc + sum_i P[i] (h[i] + J[i] / 2): the non-XC energy of grad_dft/molecule.py:697-733 (E_nuc + E_1 + E_J) from the
// spin-summed density matrix, the core Hamiltonian and the Coulomb matrix in one pass (single CTA, fixed order)
__global__ void __launch_bounds__(1024) nonxc_energy_kernel(int64_t n, const double* __restrict__ P, const double* __restrict__ h,
                                                           const double* __restrict__ J, const double* __restrict__ enuc, double* __restrict__ out) {
  __shared__ double red[32];
  double a4[4] = {0.0, 0.0, 0.0, 0.0};
  int64_t i = threadIdx.x;
  for (; i + 3072 < n; i += 4096) {  // four independent loads of each operand in flight
    double p[4], j[4], hh[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { p[u] = P[i + 1024 * u]; j[u] = J[i + 1024 * u]; hh[u] = h[i + 1024 * u]; }
#pragma unroll
    for (int u = 0; u < 4; u++) a4[u] = fma(p[u], fma(0.5, j[u], hh[u]), a4[u]);
  }
  for (; i < n; i += 1024) a4[0] = fma(P[i], fma(0.5, J[i], h[i]), a4[0]);
  double acc = warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3]));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = enuc[0] + v;
  }
}

// ---------------------------------------------------------------------------------------------------
// J AND K from ONE pass over the tensor (BASELINE.json north_star "J/K"; K is not in the reference):
//   J[p][q] = sum_{r,t} eri[p][q][r][t] P[r][t]      K[p][r] = sum_{q,t} eri[p][q][r][t] P[q][t]
// The (p,q) block [r][t] is n x n contiguous doubles.  CTA = (p, chunk of q), two q per step.  A warp takes four rows r of
// both blocks at a time (eight 128-bit streaming loads per lane and t-slice in flight); P[r][:] is read once for the two
// blocks (L2), P[q][:] and P[q+1][:] sit in shared memory.  J: lane-local over the whole block, one warp + CTA reduction per
// block.  K: one lane-partial per row, the four rows reduced together by a butterfly (6 shuffles per 4 rows instead of 20),
// accumulated over q in shared memory by the warp that owns the row; the q-chunks' partial K are summed in fixed order by
// sum_splits_kernel.  8 n^4 bytes for both results.
// ---------------------------------------------------------------------------------------------------
constexpr int JK_THREADS = 256;
constexpr int JK_ROWS = 4;
constexpr int JK_MAX_SPLIT = 16;

template <bool VEC>
__global__ void __launch_bounds__(JK_THREADS, 2) eri_jk_kernel(int n, int qsplit, int qper, const double* __restrict__ eri,
                                                              const double* __restrict__ P, double* __restrict__ J,
                                                              double* __restrict__ Kpart) {
  extern __shared__ __align__(16) double jk_smem[];
  const int npad = (n + 1) & ~1;
  double* sPq0 = jk_smem;
  double* sPq1 = jk_smem + npad;
  double* sK = jk_smem + 2 * npad;
  __shared__ double sJ[JK_THREADS / 32][2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p = blockIdx.x / qsplit, s = blockIdx.x % qsplit;
  const int q0 = s * qper, q1 = min(n, q0 + qper);
  const int ngroups = (n + JK_ROWS - 1) / JK_ROWS;
  const size_t nn = (size_t)n * n;
  for (int r = tid; r < n; r += JK_THREADS) sK[r] = 0.0;
  for (int q = q0; q < q1; q += 2) {
    const bool has1 = q + 1 < q1;
    __syncthreads();
    for (int i = tid; i < npad; i += JK_THREADS) {
      sPq0[i] = i < n ? P[(size_t)q * n + i] : 0.0;
      sPq1[i] = (has1 && i < n) ? P[(size_t)(q + 1) * n + i] : 0.0;
    }
    __syncthreads();
    const double* b0 = eri + ((size_t)p * n + q) * nn;
    const double* b1 = has1 ? b0 + nn : b0;  // odd tail: the second block repeats the first with zero weights
    double j0 = 0.0, j1 = 0.0;
    for (int g = warp; g < ngroups; g += JK_THREADS / 32) {
      const int r0 = g * JK_ROWS;
      double k[JK_ROWS];
#pragma unroll
      for (int u = 0; u < JK_ROWS; u++) k[u] = 0.0;
      if (VEC) {
        const int n2 = n >> 1;
        for (int i = lane; i < n2; i += 32) {
          const double2 w0 = reinterpret_cast<const double2*>(sPq0)[i], w1 = reinterpret_cast<const double2*>(sPq1)[i];
          double2 e0[JK_ROWS], e1[JK_ROWS], pr[JK_ROWS];
#pragma unroll
          for (int u = 0; u < JK_ROWS; u++) {
            if (r0 + u < n) {  // warp-uniform
              const size_t off = (size_t)(r0 + u) * n;
              e0[u] = __ldcs(reinterpret_cast<const double2*>(b0 + off) + i);
              e1[u] = __ldcs(reinterpret_cast<const double2*>(b1 + off) + i);
              pr[u] = __ldg(reinterpret_cast<const double2*>(P + off) + i);
            } else {
              e0[u] = e1[u] = pr[u] = make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int u = 0; u < JK_ROWS; u++) {
            j0 = fma(e0[u].x, pr[u].x, fma(e0[u].y, pr[u].y, j0));
            j1 = fma(e1[u].x, pr[u].x, fma(e1[u].y, pr[u].y, j1));
            k[u] = fma(e0[u].x, w0.x, fma(e0[u].y, w0.y, fma(e1[u].x, w1.x, fma(e1[u].y, w1.y, k[u]))));
          }
        }
      } else {
        for (int i = lane; i < n; i += 32) {
          const double w0 = sPq0[i], w1 = sPq1[i];
          double e0[JK_ROWS], e1[JK_ROWS], pr[JK_ROWS];
#pragma unroll
          for (int u = 0; u < JK_ROWS; u++) {
            if (r0 + u < n) {
              const size_t off = (size_t)(r0 + u) * n + i;
              e0[u] = __ldcs(b0 + off);
              e1[u] = __ldcs(b1 + off);
              pr[u] = __ldg(P + off);
            } else {
              e0[u] = e1[u] = pr[u] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < JK_ROWS; u++) {
            j0 = fma(e0[u], pr[u], j0);
            j1 = fma(e1[u], pr[u], j1);
            k[u] = fma(e0[u], w0, fma(e1[u], w1, k[u]));
          }
        }
      }
      // four lane-partials per lane -> one row total per group of eight lanes: halve the value count at xor 16 and xor 8
      // (each lane passes on the half it does not keep), then plain butterflies
      {
        const bool hi = lane & 16;
        const double a0 = (hi ? k[2] : k[0]) + __shfl_xor_sync(0xffffffffu, hi ? k[0] : k[2], 16);
        const double a1 = (hi ? k[3] : k[1]) + __shfl_xor_sync(0xffffffffu, hi ? k[1] : k[3], 16);
        const bool hi2 = lane & 8;
        double v = (hi2 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi2 ? a0 : a1, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int r = r0 + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        if ((lane & 7) == 0 && r < n) sK[r] += v;  // row r belongs to this warp for every q: no race
      }
    }
    j0 = warp_sum(j0);
    j1 = warp_sum(j1);
    if (lane == 0) { sJ[warp][0] = j0; sJ[warp][1] = j1; }
    __syncthreads();
    if (tid < 2 && (tid == 0 || has1)) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < JK_THREADS / 32; w++) acc += sJ[w][tid];
      J[(size_t)p * n + q + tid] = acc;
    }
  }
  __syncthreads();
  for (int r = tid; r < n; r += JK_THREADS) Kpart[((size_t)s * n + p) * n + r] = sK[r];
}

// Cotangent of K wrt P for ANY tensor: Pbar[q][t] = sum_{p,r} eri[p][q][r][t] Kbar[p][r].  CTA = (q, chunk of p); a thread owns
// one t-slice (the CTA spans whole rows, so the stream is contiguous), eight rows in flight; p-chunks summed in fixed order.
template <bool VEC>
__global__ void __launch_bounds__(256) eri_kt_kernel(int n, int psplit, int pper, const double* __restrict__ eri,
                                                    const double* __restrict__ Kbar, double* __restrict__ part) {
  const int q = blockIdx.x / psplit, s = blockIdx.x % psplit;
  const int p0 = s * pper, p1 = min(n, p0 + pper);
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  const size_t nn = (size_t)n * n;
  if (VEC) {
    if (i >= (n >> 1)) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int p = p0; p < p1; p++) {
      const double2* base = reinterpret_cast<const double2*>(eri + ((size_t)p * n + q) * nn) + i;
      const double* kb = Kbar + (size_t)p * n;
      const int n2 = n >> 1;
      int r = 0;
      for (; r + 8 <= n; r += 8) {
        double2 e[8];
#pragma unroll
        for (int u = 0; u < 8; u++) e[u] = __ldcs(base + (size_t)(r + u) * n2);
#pragma unroll
        for (int u = 0; u < 8; u++) { const double w = __ldg(kb + r + u); acc.x = fma(w, e[u].x, acc.x); acc.y = fma(w, e[u].y, acc.y); }
      }
      for (; r < n; r++) {
        const double2 e = __ldcs(base + (size_t)r * n2);
        const double w = __ldg(kb + r);
        acc.x = fma(w, e.x, acc.x); acc.y = fma(w, e.y, acc.y);
      }
    }
    reinterpret_cast<double2*>(part + ((size_t)s * n + q) * n)[i] = acc;
  } else {
    if (i >= n) return;
    double acc = 0.0;
    for (int p = p0; p < p1; p++) {
      const double* base = eri + ((size_t)p * n + q) * nn + i;
      const double* kb = Kbar + (size_t)p * n;
      int r = 0;
      for (; r + 8 <= n; r += 8) {
        double e[8];
#pragma unroll
        for (int u = 0; u < 8; u++) e[u] = __ldcs(base + (size_t)(r + u) * n);
#pragma unroll
        for (int u = 0; u < 8; u++) acc = fma(__ldg(kb + r + u), e[u], acc);
      }
      for (; r < n; r++) acc = fma(__ldg(kb + r), __ldcs(base + (size_t)r * n), acc);
    }
    part[((size_t)s * n + q) * n + i] = acc;
  }
}

size_t eri_workspace(int64_t n) {
  if (n <= 0) return 0;
  const int64_t C = n * n;
  return (size_t)64 * C * 8 + 512;  // row-split partials of the transposed sweep
}

// ---------------------------------------------------------------------------------------------------
// XC quadrature                                       grad_dft/functional.py:251-253, 342; molecule.py:687-689
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double aclip(double x, double c) { return fabs(x) > c ? x : 0.0; }

constexpr int INT_THREADS = 256;
constexpr int INT_MAX_BLOCKS = 1184;  // 8 x 148

__global__ void __launch_bounds__(INT_THREADS) integrate_fwd_kernel(int64_t N, int F, int64_t c_rows, const double* __restrict__ c,
                                                                   const double* __restrict__ d, const double* __restrict__ w,
                                                                   double clip, double* __restrict__ partial) {
  __shared__ double red[INT_THREADS / 32];
  double acc = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * INT_THREADS + threadIdx.x; r < N; r += (int64_t)gridDim.x * INT_THREADS) {
    const double* cr = c + (c_rows == 1 ? 0 : r * F);
    const double* dr = d + r * F;
    double e = 0.0;
    for (int f = 0; f < F; f++) e = fma(cr[f], dr[f], e);
    e = aclip(aclip(e, clip), clip);
    acc = fma(aclip(w[r], clip), e, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < INT_THREADS / 32; i++) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(int count, int stride, int ncol, const double* __restrict__ partial,
                                                           double* __restrict__ out) {
  // out[col] = sum_i partial[i*stride + col]; one CTA, fixed order
  __shared__ double red[32];
  for (int col = 0; col < ncol; col++) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += 1024) acc += partial[(size_t)i * stride + col];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = red[threadIdx.x];
      v = warp_sum(v);
      if (threadIdx.x == 0) out[col] = v;
    }
    __syncthreads();
  }
}

// e_bar_r = E_bar * aclip(w_r) * [|e_r| > clip];  d_bar[r,f] = e_bar_r c[r,f];  c_bar[r,f] = e_bar_r d[r,f]
// (c_rows == 1: c_bar[f] = sum_r e_bar_r d[r,f] via per-block partials)
__global__ void __launch_bounds__(INT_THREADS) integrate_bwd_kernel(int64_t N, int F, int64_t c_rows, const double* __restrict__ c,
                                                                   const double* __restrict__ d, const double* __restrict__ w,
                                                                   double clip, const double* __restrict__ E_bar,
                                                                   double* __restrict__ c_bar, double* __restrict__ d_bar,
                                                                   double* __restrict__ partial /*[grid][F] when c_rows==1 && c_bar*/) {
  __shared__ double red[INT_THREADS / 32];
  const double Eb = E_bar[0];
  double cacc[32];
  const bool reduce_c = (c_rows == 1) && (c_bar != nullptr);
  if (reduce_c)
    for (int f = 0; f < F; f++) cacc[f] = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * INT_THREADS + threadIdx.x; r < N; r += (int64_t)gridDim.x * INT_THREADS) {
    const double* cr = c + (c_rows == 1 ? 0 : r * F);
    const double* dr = d + r * F;
    double e = 0.0;
    for (int f = 0; f < F; f++) e = fma(cr[f], dr[f], e);
    const double eb = (fabs(e) > clip) ? Eb * aclip(w[r], clip) : 0.0;
    if (d_bar)
      for (int f = 0; f < F; f++) d_bar[r * F + f] = eb * cr[f];
    if (c_bar) {
      if (reduce_c) {
        for (int f = 0; f < F; f++) cacc[f] = fma(eb, dr[f], cacc[f]);
      } else {
        for (int f = 0; f < F; f++) c_bar[r * F + f] = eb * dr[f];
      }
    }
  }
  if (reduce_c) {
    for (int f = 0; f < F; f++) {
      double v = warp_sum(cacc[f]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < INT_THREADS / 32; i++) s += red[i];
        partial[(size_t)blockIdx.x * F + f] = s;
      }
      __syncthreads();
    }
  }
}

size_t integrate_workspace(int64_t) { return (size_t)INT_MAX_BLOCKS * 32 * 8 + 512; }

// ---------------------------------------------------------------------------------------------------
// predictor glue                                                        grad_dft/train.py:148-163, 205-215
// ---------------------------------------------------------------------------------------------------
__global__ void fock_assemble_kernel(int n, const double* __restrict__ h1e, const double* __restrict__ J,
                                     const double* __restrict__ Dbar, double clip, double* __restrict__ fock) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n * n) return;
  const int s = idx / (n * n), rem = idx - s * n * n, i = rem / n, j = rem - i * n;
  const int ij = i * n + j, ji = j * n + i;
  const double x = aclip(h1e[ij] + J[ij] + Dbar[s * n * n + ij], clip);
  const double xt = aclip(h1e[ji] + J[ji] + Dbar[s * n * n + ji], clip);
  fock[idx] = aclip(0.5 * (x + xt), clip);
}
__global__ void fock_add_sym_kernel(int n, const double* __restrict__ V, double clip, double* __restrict__ fock) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n * n) return;
  const int s = idx / (n * n), rem = idx - s * n * n, i = rem / n, j = rem - i * n;
  fock[idx] = aclip(fock[idx] + (V[idx] + V[s * n * n + j * n + i]), clip);
}

// ---------------------------------------------------------------------------------------------------
// Packed rep_tensor.  (pq|rt) is symmetric under p<->q and r<->t (grad_dft/interface/pyscf.py builds it with
// mol.intor("int2e"), aosym s1), so J_pq = sum_{r>=t} (pq|rt) (P_rt + P_tr)(1 - delta_rt/2) needs only the
// NP = n(n+1)/2 pair rows x NP pair columns: a quarter of the 8 n^4 bytes, re-laid-out ONCE per molecule like the
// packed basis (packed[pair(p,q)][pair(r,t)], pair(i,j) = i(i+1)/2 + j for i >= j, row pitch NPP = NP rounded up to even,
// padding column zero).  The sweep is then the plain row-times-vector kernel above on contiguous rows, and because the
// packed matrix is itself symmetric the transposed sweep (VJP of J w.r.t. P) is the same call.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t eri_npair(int64_t n) { return n * (n + 1) / 2; }
__host__ __device__ inline int64_t eri_npair_pad(int64_t n) { return (eri_npair(n) + 1) & ~int64_t(1); }

__device__ __forceinline__ void pair_decode(int64_t pr, int& i, int& j) {
  // i = floor((sqrt(8 pr + 1) - 1) / 2), corrected for rounding
  int64_t ii = (int64_t)((sqrt(8.0 * (double)pr + 1.0) - 1.0) * 0.5);
  while (ii * (ii + 1) / 2 > pr) --ii;
  while ((ii + 1) * (ii + 2) / 2 <= pr) ++ii;
  i = (int)ii;
  j = (int)(pr - ii * (ii + 1) / 2);
}

// one CTA per pair row: packed[pr - pair0][pair(k,l)] = src_row(pr)[k][l], k >= l
// src_mode 0: src holds the rows (p,q) = src_row0 .. of the full tensor, 1: src holds exactly the pair rows pair0 ..
__global__ void __launch_bounds__(256) eri_pack_kernel(int n, int64_t pair0, int64_t pairs, int src_mode, int64_t src_row0,
                                                       const double* __restrict__ src, double* __restrict__ packed) {
  const int64_t NPP = eri_npair_pad(n), NP = eri_npair(n);
  for (int64_t lp = blockIdx.x; lp < pairs; lp += gridDim.x) {
    int i, j;
    pair_decode(pair0 + lp, i, j);
    const double* row = src + (src_mode ? lp : ((int64_t)i * n + j - src_row0)) * (int64_t)n * n;
    double* out = packed + lp * NPP;
    for (int k = threadIdx.x >> 5; k < n; k += 8) {
      const double* rk = row + (int64_t)k * n;
      double* ok = out + (int64_t)k * (k + 1) / 2;
      for (int l = threadIdx.x & 31; l <= k; l += 32) ok[l] = __ldcs(rk + l);
    }
    if (threadIdx.x == 0 && NPP > NP) out[NP] = 0.0;
  }
}

// part[blockIdx.x] = {max |asymmetry|, max |value|} over the rows this CTA visits (rows (p,q) = row0 .. row0+rows-1)
__global__ void __launch_bounds__(256) eri_symmetry_kernel(int n, int64_t row0, int64_t rows, const double* __restrict__ src,
                                                           double* __restrict__ part) {
  double asym = 0.0, big = 0.0;
  const int64_t R1 = row0 + rows;
  for (int64_t lr = blockIdx.x; lr < rows; lr += gridDim.x) {
    const int64_t r = row0 + lr;
    const int p = (int)(r / n), q = (int)(r % n);
    const double* row = src + lr * (int64_t)n * n;
    const int64_t rt = (int64_t)q * n + p;  // the (q,p) row, when this block holds it
    const double* rowT = (rt >= row0 && rt < R1 && q < p) ? src + (rt - row0) * (int64_t)n * n : nullptr;
    for (int64_t c = threadIdx.x; c < (int64_t)n * n; c += 256) {
      const int k = (int)(c / n), l = (int)(c % n);
      const double v = row[c];
      big = fmax(big, fabs(v));
      if (l < k) asym = fmax(asym, fabs(v - row[(int64_t)l * n + k]));
      if (rowT) asym = fmax(asym, fabs(v - rowT[c]));
    }
  }
  __shared__ double sa[256], sb[256];
  sa[threadIdx.x] = asym; sb[threadIdx.x] = big;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] = fmax(sa[threadIdx.x], sa[threadIdx.x + o]); sb[threadIdx.x] = fmax(sb[threadIdx.x], sb[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = sa[0]; part[2 * blockIdx.x + 1] = sb[0]; }
}
__global__ void eri_symmetry_final_kernel(int nparts, const double* __restrict__ part, double* __restrict__ out2) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 32) { a = fmax(a, part[2 * i]); b = fmax(b, part[2 * i + 1]); }
  for (int o = 16; o > 0; o >>= 1) { a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); }
  if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
}

// Ppk[pair(k,l)] = P[k][l] + P[l][k] (k > l), P[k][k] (k == l); padding 0
__global__ void eri_pack_p_kernel(int n, const double* __restrict__ P, double* __restrict__ Ppk) {
  const int64_t NPP = eri_npair_pad(n), NP = eri_npair(n);
  const int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= NPP) return;
  if (pr >= NP) { Ppk[pr] = 0.0; return; }
  int k, l;
  pair_decode(pr, k, l);
  Ppk[pr] = (k == l) ? P[(int64_t)k * n + k] : P[(int64_t)k * n + l] + P[(int64_t)l * n + k];
}
// J[i][j] = J[j][i] = Jpk[pair(i,j) - pair0] for the local pairs; every other entry 0 (so that partial J's sum to the whole)
__global__ void eri_unpack_j_kernel(int n, int64_t pair0, int64_t pairs, const double* __restrict__ Jpk, double* __restrict__ J) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int a = (int)(idx / n), b = (int)(idx % n);
  const int i = a > b ? a : b, j = a > b ? b : a;
  const int64_t pr = (int64_t)i * (i + 1) / 2 + j - pair0;
  J[idx] = (pr >= 0 && pr < pairs) ? Jpk[pr] : 0.0;
}

}  // namespace gdft

using namespace gdft;

// Row block [row0, row0+rows) of the (pq) x (rt) sweep: `eri_rows` points at the block's first row.  The per-row
// summation order does not depend on the blocking, so a row-sharded J is bitwise equal to the unsharded one.
static int eri_gemv_launch(cudaStream_t stream, int64_t C, int64_t rows, const double* eri_rows, const double* P, double* J_rows);
static int eri_j_rows_launch(cudaStream_t stream, int64_t n, int64_t rows, const double* eri_rows, const double* P, double* J_rows) {
  return eri_gemv_launch(stream, n * n, rows, eri_rows, P, J_rows);
}
static int eri_gemv_launch(cudaStream_t stream, int64_t C, int64_t rows, const double* eri_rows, const double* P, double* J_rows) {
  const bool vec = (C % 2 == 0) && aligned16(eri_rows);
  const bool small = (rows + 31) / 32 < 2 * 148;  // fewer than two CTAs per SM at four rows per warp
  const int64_t nblocks = small ? (rows + 7) / 8 : (rows + 31) / 32;
  const unsigned grid = (unsigned)imin64(nblocks, 148 * 8);
  if (small) {
    if (vec) eri_j_kernel<true, 1><<<grid, ERI_THREADS, 0, stream>>>(rows, C, eri_rows, P, J_rows);
    else eri_j_kernel<false, 1><<<grid, ERI_THREADS, 0, stream>>>(rows, C, eri_rows, P, J_rows);
  } else {
    if (vec) eri_j_kernel<true, 4><<<grid, ERI_THREADS, 0, stream>>>(rows, C, eri_rows, P, J_rows);
    else eri_j_kernel<false, 4><<<grid, ERI_THREADS, 0, stream>>>(rows, C, eri_rows, P, J_rows);
  }
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_nonxc_energy(gdft_stream_t stream, int64_t n, const double* P, const double* h1e, const double* J,
                                 const double* nuclear_repulsion, double* out) {
  if (n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!P || !h1e || !J || !nuclear_repulsion || !out) return GDFT_BAD_ARGUMENT;
  nonxc_energy_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(n * n, P, h1e, J, nuclear_repulsion, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// ---- packed rep_tensor entry points -----------------------------------------------------------------------------------
extern "C" int64_t gdft_eri_npair(int64_t n) { return n > 0 ? eri_npair(n) : 0; }
extern "C" size_t gdft_eri_packed_bytes(int64_t n, int64_t pairs) { return (n > 0 && pairs > 0) ? (size_t)pairs * (size_t)eri_npair_pad(n) * 8 : 0; }
extern "C" size_t gdft_eri_packed_workspace(int64_t n) { return n > 0 ? (size_t)(2 * eri_npair_pad(n) + 2 * 148 * 8 + 64) * 8 : 0; }

extern "C" int gdft_eri_symmetry_defect(gdft_stream_t stream_, int64_t n, int64_t row0, int64_t rows, const double* eri_rows, double* out2,
                                        void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048 || row0 < 0 || rows <= 0 || row0 + rows > n * n) return GDFT_BAD_SHAPE;
  if (!eri_rows || !out2) return GDFT_BAD_ARGUMENT;
  if (ws_bytes < gdft_eri_packed_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
  const int nparts = (int)imin64(rows, 148 * 8);
  double* part = static_cast<double*>(ws);
  eri_symmetry_kernel<<<nparts, 256, 0, stream>>>((int)n, row0, rows, eri_rows, part);
  GDFT_LAUNCH_CHECK();
  eri_symmetry_final_kernel<<<1, 32, 0, stream>>>(nparts, part, out2);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_eri_pack(gdft_stream_t stream_, int64_t n, int src_is_pair_rows, int64_t src_row0, int64_t src_rows, const double* src,
                             int64_t pair0, int64_t pairs, double* packed) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048 || pair0 < 0 || pairs <= 0 || pair0 + pairs > eri_npair(n)) return GDFT_BAD_SHAPE;
  if (!src || !packed) return GDFT_BAD_ARGUMENT;
  if (!aligned16(packed)) return GDFT_BAD_ALIGNMENT;
  if (src_is_pair_rows) {
    if (src_rows != pairs) return GDFT_BAD_SHAPE;
  } else {
    // first and last pair rows must lie inside the source block (rows in between do, the map pair -> row is increasing)
    auto row_of = [&](int64_t pr) { int64_t i = (int64_t)((sqrt(8.0 * (double)pr + 1.0) - 1.0) * 0.5); while (i * (i + 1) / 2 > pr) --i; while ((i + 1) * (i + 2) / 2 <= pr) ++i; return i * n + (pr - i * (i + 1) / 2); };
    if (src_row0 < 0 || src_rows <= 0 || row_of(pair0) < src_row0 || row_of(pair0 + pairs - 1) >= src_row0 + src_rows) return GDFT_BAD_SHAPE;
  }
  eri_pack_kernel<<<(unsigned)imin64(pairs, 148 * 16), 256, 0, stream>>>((int)n, pair0, pairs, src_is_pair_rows ? 1 : 0, src_row0, src, packed);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// J[n,n] from the packed pair rows [pair0, pair0 + pairs): entries of other pairs are written as zero.  EJ (optional) is
// 1/2 <P, J> of what was written (the whole E_J when pairs == npair).
extern "C" int gdft_eri_j_packed(gdft_stream_t stream_, int64_t n, int64_t pair0, int64_t pairs, const double* packed, const double* P, double* J,
                                 double* EJ, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048 || pair0 < 0 || pairs < 0 || pair0 + pairs > eri_npair(n)) return GDFT_BAD_SHAPE;
  if (!packed || !P || !J) return GDFT_BAD_ARGUMENT;
  if (!aligned16(packed) || !aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < gdft_eri_packed_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
  const int64_t NPP = eri_npair_pad(n);
  double* Ppk = static_cast<double*>(ws);
  double* Jpk = Ppk + NPP;
  eri_pack_p_kernel<<<(unsigned)((NPP + 255) / 256), 256, 0, stream>>>((int)n, P, Ppk);
  GDFT_LAUNCH_CHECK();
  if (pairs > 0)
    if (int rc = eri_gemv_launch(stream, NPP, pairs, packed, Ppk, Jpk)) return rc;
  eri_unpack_j_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, stream>>>((int)n, pair0, pairs, Jpk, J);
  GDFT_LAUNCH_CHECK();
  if (EJ) {
    dot_kernel<<<1, 1024, 0, stream>>>(n * n, P, J, 0.5, EJ);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}

// chunks of the second index for the J+K sweep / the K transpose: enough CTAs for a few waves, at most JK_MAX_SPLIT partials
static void jk_split(int64_t n, int unit, int* split, int* per) {
  int want = (int)imin64(JK_MAX_SPLIT, imax64(1, (8 * 148 + n - 1) / n));
  int per_ = (int)((n + want - 1) / want);
  per_ = ((per_ + unit - 1) / unit) * unit;
  *per = per_;
  *split = (int)((n + per_ - 1) / per_);
}

extern "C" int gdft_eri_jk(gdft_stream_t stream_, int64_t n, const double* eri, const double* P, double* J, double* K,
                           double* EJ, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048) return GDFT_BAD_SHAPE;
  if (!eri || !P || !J) return GDFT_BAD_ARGUMENT;
  const int64_t R = n * n;
  if (K) {
    // both contractions from one pass over the tensor
    if (!ws || ws_bytes < eri_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
    int qsplit, qper;
    jk_split(n, 2, &qsplit, &qper);
    double* Kpart = static_cast<double*>(ws);
    const size_t smem = (size_t)3 * ((n + 1) & ~(int64_t)1) * sizeof(double);
    const bool vec = (n % 2 == 0) && aligned16(eri) && aligned16(P);
    const unsigned grid = (unsigned)(n * qsplit);
    if (vec) eri_jk_kernel<true><<<grid, JK_THREADS, smem, stream>>>((int)n, qsplit, qper, eri, P, J, Kpart);
    else eri_jk_kernel<false><<<grid, JK_THREADS, smem, stream>>>((int)n, qsplit, qper, eri, P, J, Kpart);
    GDFT_LAUNCH_CHECK();
    sum_splits_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, qsplit, Kpart, K);
    GDFT_LAUNCH_CHECK();
  } else {
    if (int rc = eri_j_rows_launch(stream, n, R, eri, P, J)) return rc;
  }
  if (EJ) {
    dot_kernel<<<1, 1024, 0, stream>>>(R, P, J, 0.5, EJ);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}

extern "C" int gdft_eri_k_transpose(gdft_stream_t stream_, int64_t n, const double* eri, const double* Kbar, double* Pbar, void* ws,
                                    size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048) return GDFT_BAD_SHAPE;
  if (!eri || !Kbar || !Pbar) return GDFT_BAD_ARGUMENT;
  if (!ws || ws_bytes < eri_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
  int psplit, pper;
  jk_split(n, 1, &psplit, &pper);
  double* part = static_cast<double*>(ws);
  const bool vec = (n % 2 == 0) && aligned16(eri);
  const int slices = vec ? (int)(n / 2) : (int)n;
  const int threads = (int)imin64(256, ((slices + 31) / 32) * 32);
  dim3 grid((unsigned)(n * psplit), (unsigned)((slices + threads - 1) / threads));
  if (vec) eri_kt_kernel<true><<<grid, threads, 0, stream>>>((int)n, psplit, pper, eri, Kbar, part);
  else eri_kt_kernel<false><<<grid, threads, 0, stream>>>((int)n, psplit, pper, eri, Kbar, part);
  GDFT_LAUNCH_CHECK();
  const int64_t C = n * n;
  sum_splits_kernel<<<(unsigned)((C + 255) / 256), 256, 0, stream>>>(C, psplit, part, Pbar);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_eri_j_rows(gdft_stream_t stream_, int64_t n, int64_t rows, const double* eri_rows, const double* P,
                               double* J_rows) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048 || rows < 0 || rows > n * n) return GDFT_BAD_SHAPE;
  if (rows == 0) return GDFT_OK;
  if (!eri_rows || !P || !J_rows) return GDFT_BAD_ARGUMENT;
  return eri_j_rows_launch(stream, n, rows, eri_rows, P, J_rows);
}

static int eri_jt_rows_launch(cudaStream_t stream, int64_t n, int64_t rows, const double* eri_rows, const double* Jbar_rows,
                              double* Pbar, void* ws) {
  const int64_t R = rows, C = n * n;
  const bool vec = (C % 2 == 0) && aligned16(eri_rows);
  int splits = (int)imin64(64, imax64(1, R / 64));
  const int64_t rows_per_split = (R + splits - 1) / splits;
  double* part = static_cast<double*>(ws);
  const int64_t cols_per_cta = vec ? 2 * ERIT_THREADS : ERIT_THREADS;
  dim3 grid((unsigned)((C + cols_per_cta - 1) / cols_per_cta), (unsigned)splits);
  if (vec) eri_jt_kernel<true><<<grid, ERIT_THREADS, 0, stream>>>(R, C, rows_per_split, eri_rows, Jbar_rows, part);
  else eri_jt_kernel<false><<<grid, ERIT_THREADS, 0, stream>>>(R, C, rows_per_split, eri_rows, Jbar_rows, part);
  GDFT_LAUNCH_CHECK();
  sum_splits_kernel<<<(unsigned)((C + 255) / 256), 256, 0, stream>>>(C, splits, part, Pbar);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_eri_j_transpose(gdft_stream_t stream_, int64_t n, const double* eri, const double* Jbar, double* Pbar,
                                    void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048) return GDFT_BAD_SHAPE;
  if (!eri || !Jbar || !Pbar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < eri_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
  return eri_jt_rows_launch(stream, n, n * n, eri, Jbar, Pbar, ws);
}

extern "C" int gdft_eri_j_transpose_rows(gdft_stream_t stream_, int64_t n, int64_t rows, const double* eri_rows,
                                         const double* Jbar_rows, double* Pbar, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 2048 || rows <= 0 || rows > n * n) return GDFT_BAD_SHAPE;
  if (!eri_rows || !Jbar_rows || !Pbar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(ws)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < eri_workspace(n)) return GDFT_WORKSPACE_TOO_SMALL;
  return eri_jt_rows_launch(stream, n, rows, eri_rows, Jbar_rows, Pbar, ws);
}

extern "C" int gdft_xc_integrate_fwd(gdft_stream_t stream_, int64_t N, int F, int64_t c_rows, const double* c, const double* d,
                                     const double* w, double clip, double* E, void* ws, size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || F <= 0 || F > 32 || (c_rows != 1 && c_rows != N)) return GDFT_BAD_SHAPE;
  if (!c || !d || !w || !E) return GDFT_BAD_ARGUMENT;
  if (ws_bytes < integrate_workspace(N)) return GDFT_WORKSPACE_TOO_SMALL;
  const int grid = (int)imin64(INT_MAX_BLOCKS, (N + INT_THREADS - 1) / INT_THREADS);
  double* partial = static_cast<double*>(ws);
  integrate_fwd_kernel<<<grid, INT_THREADS, 0, stream>>>(N, F, c_rows, c, d, w, clip, partial);
  GDFT_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 1024, 0, stream>>>(grid, 1, 1, partial, E);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_xc_integrate_bwd(gdft_stream_t stream_, int64_t N, int F, int64_t c_rows, const double* c, const double* d,
                                     const double* w, double clip, const double* E_bar, double* c_bar, double* d_bar, void* ws,
                                     size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N <= 0 || F <= 0 || F > 32 || (c_rows != 1 && c_rows != N)) return GDFT_BAD_SHAPE;
  if (!c || !d || !w || !E_bar) return GDFT_BAD_ARGUMENT;
  if (ws_bytes < integrate_workspace(N)) return GDFT_WORKSPACE_TOO_SMALL;
  const int grid = (int)imin64(INT_MAX_BLOCKS, (N + INT_THREADS - 1) / INT_THREADS);
  double* partial = static_cast<double*>(ws);
  integrate_bwd_kernel<<<grid, INT_THREADS, 0, stream>>>(N, F, c_rows, c, d, w, clip, E_bar, c_bar, d_bar, partial);
  GDFT_LAUNCH_CHECK();
  if (c_bar && c_rows == 1) {
    sum_partials_kernel<<<1, 1024, 0, stream>>>(grid, F, F, partial, c_bar);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}

extern "C" int gdft_fock_assemble(gdft_stream_t stream, int64_t n, const double* h1e, const double* J, const double* rdm1_bar,
                                  double clip, double* fock) {
  if (n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!h1e || !J || !rdm1_bar || !fock) return GDFT_BAD_ARGUMENT;
  const int total = (int)(2 * n * n);
  fock_assemble_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>((int)n, h1e, J, rdm1_bar, clip, fock);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_fock_add_sym(gdft_stream_t stream, int64_t n, const double* V, double clip, double* fock) {
  if (n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!V || !fock) return GDFT_BAD_ARGUMENT;
  const int total = (int)(2 * n * n);
  fock_add_sym_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>((int)n, V, clip, fock);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
