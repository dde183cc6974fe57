// Row f1, matrices beyond the one-CTA Jacobi kernels (64 < n <= 320: the benzene/def2-TZVP class,
// n = 264): the symmetric eigenproblem of the SCF iteration (grad_dft/utils/eigenproblem.py:110-149, jnp.linalg.eigh per
// spin after the Cholesky reduction) as ONE thread-block cluster per matrix, stream-ordered, with no host-visible status
// word -- so that `make_jitted_scf_loop` at this size is a single CUDA graph like the small molecules, and the replicated
// eigensolve stops being the serial part of the grid-sharded iteration (the library path, cuSOLVER syevd, runs ~1000
// launch-bound kernels: 2.8 ms per 264 x 264 matrix, two matrices back to back).
//
// Algorithm: one-sided (Hestenes) Jacobi on W = (C + sigma I) V0.  With sigma = 1.5 ||C||_F the shifted matrix C' is positive
// definite with a condition number <= 5, so at convergence -- the columns of W mutually orthogonal -- W = C' V = V diag(lambda +
// sigma): the eigenvectors are the normalised columns of W and the eigenvalues their norms minus sigma; V itself is never
// carried.  V0 is the identity (cold) or the eigenvectors of the previous SCF cycle (warm: W0 is then nearly orthogonal and
// two or three sweeps do, against ten cold).  Each rotation touches two columns only, so there is no two-sided update and
// no rotation broadcast: a warp owns one column pair per round, computes the three dot products (a, b, g) = (|w_p|^2,
// |w_q|^2, w_p.w_q) with a shuffle all-reduce, rotates in registers (the rsqrt-only formulas of eigh_jacobi.cu) and hands
// its two columns on.
//
// Layout: a cluster of 8 CTAs, `warps` warps each; warp w of CTA c is slot k = c * warps + w of the Brent-Luk round-robin
// arrangement (M = 8 * warps >= ceil(n/2) slots; 2M - 1 rounds per sweep; surplus columns are zero and never rotate).
// Columns live in double-buffered shared-memory mailboxes [buffer][slot][top|bottom][32 * EPL]; a round reads its pair from
// the current buffer, and writes the rotated columns straight into the mailboxes of the slots that own them next round
// (top -> slot k+1, bottom -> slot k-1, the two ends turn around) -- across a CTA boundary that is a distributed-shared-
// memory store (mapa + st.shared::cluster) -- followed by ONE cluster barrier per round.  Lane l holds rows l + 32 e.
// Convergence: a sweep in which no rotation exceeded |g| > tol sqrt(a b) (every warp publishes one flag per sweep to all
// eight CTAs).  `info[b]` (optional, device) receives the sweep count, -1 if the bound was hit, -2 for a non-finite result.
#include "common.cuh"

namespace gdft {

constexpr int HC_CLUSTER = 8;
constexpr int HC_MAX_WARPS = 20;
constexpr int HC_MAX_N = 320;
constexpr int HC_MAX_SWEEPS = 40;

__device__ __forceinline__ uint32_t hc_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t hc_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void hc_st_f64(uint32_t caddr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(caddr), "d"(v) : "memory"); }
__device__ __forceinline__ void hc_st_u32(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void hc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int EPL>
__global__ void __launch_bounds__(HC_MAX_WARPS * 32, 1)
sym_eig_hestenes_cluster_kernel(int n, int warps, const double* __restrict__ A_in, const double* __restrict__ V0_in, double* __restrict__ evals,
                                double* __restrict__ evecs, int* __restrict__ info, int max_sweeps) {
  constexpr int COL = 32 * EPL;
  extern __shared__ __align__(16) unsigned char hc_smem[];
  const int M = HC_CLUSTER * warps;
  double* mail = reinterpret_cast<double*>(hc_smem);               // [2][warps][2][COL]
  double* lam_all = mail + (size_t)2 * warps * 2 * COL;            // [2M]
  uint32_t* flags = reinterpret_cast<uint32_t*>(lam_all + 2 * M);  // [M]
  __shared__ double red[HC_MAX_WARPS];
  __shared__ double s_sigma;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t cta = hc_ctarank();
  const int mat = blockIdx.x / HC_CLUSTER;
  const int k = (int)cta * warps + w;  // slot
  const double* A = A_in + (size_t)mat * n * n;
  const double* V0 = V0_in ? V0_in + (size_t)mat * n * n : nullptr;

  // ---- sigma = 1.5 ||A||_F (every CTA computes it, identically) ------------------------------------------------------
  {
    double s = 0.0;
    for (int idx = tid; idx < n * n; idx += blockDim.x) { const double v = A[idx]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) red[w] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < warps; i++) t += red[i];
      s_sigma = t > 0.0 ? 1.5 * sqrt(t) : 1.0;
    }
    __syncthreads();
  }
  const double sigma = s_sigma;

  // ---- W0 = (A + sigma I) V0, this slot's two columns p = 2k, q = 2k + 1 --------------------------------------------
  double wp[EPL], wq[EPL];
  const int p = 2 * k, q = 2 * k + 1;
  if (V0 == nullptr) {
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int i = lane + 32 * e;
      wp[e] = (p < n && i < n) ? A[(size_t)p * n + i] + (i == p ? sigma : 0.0) : 0.0;  // A symmetric: column p read as row p
      wq[e] = (q < n && i < n) ? A[(size_t)q * n + i] + (i == q ? sigma : 0.0) : 0.0;
    }
  } else {
    double* vp = mail + ((size_t)(1 * warps + w) * 2 + 0) * COL;  // buffer 1 of this slot as staging for the V0 columns
    double* vq = vp + COL;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int i = lane + 32 * e;
      vp[i] = (p < n && i < n) ? V0[(size_t)i * n + p] : 0.0;
      vq[i] = (q < n && i < n) ? V0[(size_t)i * n + q] : 0.0;
      wp[e] = 0.0;
      wq[e] = 0.0;
    }
    __syncwarp();
    if (p < n) {
      for (int j = 0; j < n; j++) {
        const double xp = vp[j], xq = vq[j];
        const double* row = A + (size_t)j * n + lane;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const double c = (lane + 32 * e < n) ? __ldg(row + 32 * e) : 0.0;
          wp[e] = fma(c, xp, wp[e]);
          wq[e] = fma(c, xq, wq[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        wp[e] = fma(sigma, vp[lane + 32 * e], wp[e]);
        wq[e] = fma(sigma, vq[lane + 32 * e], wq[e]);
      }
    }
    __syncwarp();
  }
  {
    double* dst = mail + ((size_t)(0 * warps + w) * 2) * COL;
#pragma unroll
    for (int e = 0; e < EPL; e++) { dst[lane + 32 * e] = wp[e]; dst[COL + lane + 32 * e] = wq[e]; }
  }

  // ---- where this slot's columns go after a round (Brent-Luk): top_0 stays; top_1 <- bottom_0; top_k <- top_{k-1};
  //      bottom_k <- bottom_{k+1}; bottom_{M-1} <- top_{M-1} ---------------------------------------------------------------
  int top_slot, top_pos, bot_slot, bot_pos;
  if (k == 0) { top_slot = 0; top_pos = 0; bot_slot = 1; bot_pos = 0; }
  else if (k == M - 1) { top_slot = M - 1; top_pos = 1; bot_slot = M - 2; bot_pos = 1; }
  else { top_slot = k + 1; top_pos = 0; bot_slot = k - 1; bot_pos = 1; }
  const uint32_t mail_s = smem_u32(mail);
  auto dest = [&](int slot, int pos, int buf) -> uint32_t {
    const int dc = slot / warps, dw = slot - dc * warps;
    const uint32_t off = (uint32_t)((((size_t)buf * warps + dw) * 2 + pos) * COL + lane) * 8u;
    return hc_mapa(mail_s + off, (uint32_t)dc);
  };
  const uint32_t dtop0 = dest(top_slot, top_pos, 0), dtop1 = dest(top_slot, top_pos, 1);
  const uint32_t dbot0 = dest(bot_slot, bot_pos, 0), dbot1 = dest(bot_slot, bot_pos, 1);
  const uint32_t flags_s = smem_u32(flags), lam_s = smem_u32(lam_all);

  const double tol = sqrt((double)n) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  hc_cluster_sync();  // every CTA of the cluster is resident and has written its initial columns

  int buf = 0, sweeps = 0;
  bool converged = false;
  const int rounds = 2 * M - 1;
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    bool any = false;
    for (int r = 0; r < rounds; r++) {
      const double* src = mail + ((size_t)(buf * warps + w) * 2) * COL + lane;
      double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        wp[e] = src[32 * e];
        wq[e] = src[COL + 32 * e];
        a = fma(wp[e], wp[e], a);
        b = fma(wq[e], wq[e], b);
        g = fma(wp[e], wq[e], g);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      double c = 1.0, s = 0.0;
      const double d = 0.5 * (b - a);
      const double h = fma(d, d, g * g);
      if (g * g > tol2 * (a * b) && h > 1e-290) {
        any = true;
        const double rh = rsqrt(h);
        const double u = fma(0.5 * fabs(d), rh, 0.5);  // (1 + cos 2theta) / 2 in [1/2, 1]
        const double ru = rsqrt(u);
        c = u * ru;
        s = 0.5 * g * rh * ru;
        if (d < 0.0) s = -s;
      }
      const uint32_t dt = buf ? dtop0 : dtop1, db = buf ? dbot0 : dbot1;  // the other buffer
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        hc_st_f64(dt + 256u * e, fma(c, wp[e], -s * wq[e]));
        hc_st_f64(db + 256u * e, fma(s, wp[e], c * wq[e]));
      }
      if (r == rounds - 1 && lane < HC_CLUSTER) hc_st_u32(hc_mapa(flags_s + 4u * k, (uint32_t)lane), any ? 1u : 0u);
      hc_cluster_sync();
      buf ^= 1;
    }
    sweeps = sweep + 1;
    uint32_t f = 0;
    for (int i = lane; i < M; i += 32) f |= flags[i];
    converged = !__any_sync(0xffffffffu, f != 0);
  }

  // ---- eigenvalues = column norms - sigma; ascending order; eigenvectors = normalised columns ----------------------------
  const double* src = mail + ((size_t)(buf * warps + w) * 2) * COL + lane;
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; e++) {
    wp[e] = src[32 * e];
    wq[e] = src[COL + 32 * e];
    a = fma(wp[e], wp[e], a);
    b = fma(wq[e], wq[e], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double na = sqrt(a), nb = sqrt(b);
  const double lp = a > 0.0 ? na - sigma : (a == 0.0 ? inf : a), lq = b > 0.0 ? nb - sigma : (b == 0.0 ? inf : b);  // zero column = padding; NaN stays NaN
  if (lane < HC_CLUSTER) hc_st_f64(hc_mapa(lam_s + 8u * (2 * k), (uint32_t)lane), lp);
  else if (lane < 2 * HC_CLUSTER) hc_st_f64(hc_mapa(lam_s + 8u * (2 * k + 1), (uint32_t)(lane - HC_CLUSTER)), lq);
  hc_cluster_sync();
  int rp = 0, rq = 0, bad = 0;
  for (int j = lane; j < 2 * M; j += 32) {
    const double lj = lam_all[j];
    rp += (lj < lp || (lj == lp && j < 2 * k)) ? 1 : 0;
    rq += (lj < lq || (lj == lq && j < 2 * k + 1)) ? 1 : 0;
    bad |= (lj != lj) ? 1 : 0;
  }
  rp = __reduce_add_sync(0xffffffffu, rp);
  rq = __reduce_add_sync(0xffffffffu, rq);
  bad = __any_sync(0xffffffffu, bad);
  double* ev = evals + (size_t)mat * n;
  double* vec = evecs + (size_t)mat * n * n;
  if (lp != inf && rp < n) {
    const double ia = 1.0 / na;
    if (lane == 0) ev[rp] = lp;
#pragma unroll
    for (int e = 0; e < EPL; e++)
      if (lane + 32 * e < n) vec[(size_t)(lane + 32 * e) * n + rp] = wp[e] * ia;
  }
  if (lq != inf && rq < n) {
    const double ib = 1.0 / nb;
    if (lane == 0) ev[rq] = lq;
#pragma unroll
    for (int e = 0; e < EPL; e++)
      if (lane + 32 * e < n) vec[(size_t)(lane + 32 * e) * n + rq] = wq[e] * ib;
  }
  if (info != nullptr && k == 0 && lane == 0) info[mat] = bad ? -2 : (converged ? sweeps : -1);
}

template <int EPL>
static int launch_hestenes(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  const int m = (n + 1) / 2;
  const int warps = (m + HC_CLUSTER - 1) / HC_CLUSTER;
  const int M = HC_CLUSTER * warps;
  const size_t smem = (size_t)2 * warps * 2 * 32 * EPL * 8 + (size_t)2 * M * 8 + (size_t)M * 4 + 16;
  auto kern = sym_eig_hestenes_cluster_kernel<EPL>;
  GDFT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(batch * HC_CLUSTER));
  cfg.blockDim = dim3((unsigned)(warps * 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = HC_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GDFT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, n, warps, A, V0, evals, evecs, info, (int)HC_MAX_SWEEPS));
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

int sym_eigh_cluster(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  const int epl = (n + 31) / 32;
  switch (epl) {
    case 1: case 2: case 3: return launch_hestenes<3>(stream, batch, n, A, V0, evals, evecs, info);
    case 4: return launch_hestenes<4>(stream, batch, n, A, V0, evals, evecs, info);
    case 5: return launch_hestenes<5>(stream, batch, n, A, V0, evals, evecs, info);
    case 6: return launch_hestenes<6>(stream, batch, n, A, V0, evals, evecs, info);
    case 7: return launch_hestenes<7>(stream, batch, n, A, V0, evals, evecs, info);
    case 8: return launch_hestenes<8>(stream, batch, n, A, V0, evals, evecs, info);
    case 9: return launch_hestenes<9>(stream, batch, n, A, V0, evals, evecs, info);
    case 10: return launch_hestenes<10>(stream, batch, n, A, V0, evals, evecs, info);
    default: return GDFT_BAD_SHAPE;
  }
}

int sym_eigh_cluster_max_n() { return HC_MAX_N; }

}  // namespace gdft
