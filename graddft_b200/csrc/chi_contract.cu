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
template <int NJ, bool VEC>
__global__ void __launch_bounds__(CHI_THREADS, 2) chi_contract_kernel(ChiArgs a) {
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
          if (NJ <= 5) {  // two rows in flight; wider passes already carry >= 3 KB per warp in one row
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

static size_t chi_smem(int n) {
  const int n2 = (n + 1) & ~1;
  return ((size_t)n * CHI_PTS + (size_t)CHI_PTS * 2 * n2) * 8;
}

template <int NJ>
static int launch_chi(cudaStream_t stream, const ChiArgs& a, bool vec, int ctas, size_t smem) {
  if (vec) {
    GDFT_CUDA_TRY(cudaFuncSetAttribute(chi_contract_kernel<NJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chi_contract_kernel<NJ, true><<<ctas, CHI_THREADS, smem, stream>>>(a);
  } else {
    GDFT_CUDA_TRY(cudaFuncSetAttribute(chi_contract_kernel<NJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chi_contract_kernel<NJ, false><<<ctas, CHI_THREADS, smem, stream>>>(a);
  }
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
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
  const int per_sm = smem * 2 <= (size_t)220 * 1024 ? 2 : 1;
  const int64_t ngroups = (Nc + CHI_PTS - 1) / CHI_PTS;
  const int ctas = (int)imin64(ngroups, (int64_t)sms * per_sm);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // columns per pass: one pass up to 512 columns; beyond that the fewest passes of equal width (a narrow last pass would
  // re-walk all n rows for a few lanes' worth of columns)
  const int npass = (int)((n + 64 * CHI_NJ - 1) / (64 * CHI_NJ));
  const int nj = (int)(((n + npass - 1) / npass + 63) / 64);
  switch (nj) {
    case 1: return launch_chi<1>(stream, a, vec, ctas, smem);
    case 2: return launch_chi<2>(stream, a, vec, ctas, smem);
    case 3: return launch_chi<3>(stream, a, vec, ctas, smem);
    case 4: return launch_chi<4>(stream, a, vec, ctas, smem);
    case 5: return launch_chi<5>(stream, a, vec, ctas, smem);
    case 6: return launch_chi<6>(stream, a, vec, ctas, smem);
    case 7: return launch_chi<7>(stream, a, vec, ctas, smem);
    default: return launch_chi<8>(stream, a, vec, ctas, smem);
  }
}
