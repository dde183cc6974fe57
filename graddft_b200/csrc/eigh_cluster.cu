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
// Columns live in double-buffered shared-memory mailboxes [buffer][slot][top|bottom][64 * EP2]; lane l holds rows 2l, 2l+1
// (+ 64 e) as one double2.  A round reads its pair from the current buffer and sends the rotated columns straight into the
// mailboxes of the slots that own them next round (top -> slot k+1, bottom -> slot k-1, the two ends turn around) with
// st.async: 16-byte distributed-shared-memory stores that complete transaction bytes on the RECEIVER's mbarrier (one per slot
// and buffer, armed by the receiver with the byte count of two columns).  There is no cluster-wide barrier in a round: a
// slot waits only for its own two incoming columns, so the rounds run as a wavefront between neighbours (a cluster barrier
// costs ~450 clocks; the first version of this kernel, with one per round, spent 1060 clocks per round outside the
// arithmetic).  Why two buffers suffice without "empty" barriers: slot k's two producers are exactly the two slots it sends
// to; having received both columns of round r it knows both have finished reading their round r-1 buffers, which are the
// ones it now writes.  Convergence: a sweep in which no rotation exceeded |g| > 1e-9 sqrt(a b) (rotations down to
// sqrt(n) eps are still applied during that sweep; Jacobi converges quadratically, so what is left afterwards is far below
// rounding) -- every warp publishes one flag per sweep to all eight CTAs, followed by the sweep's only cluster barrier.
// `info[b]` (optional, device) receives the sweep count, -1 if the bound was hit, -2 for a non-finite result.
#include <stdlib.h>
#include "common.cuh"

namespace gdft {

constexpr int HC_MAX_CLUSTER = 16;  // 8 is the portable maximum; 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed
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

__device__ __forceinline__ void hc_st_async_v2(uint32_t caddr, double x, double y, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(caddr), "d"(x), "d"(y), "r"(cbar)
               : "memory");
}

// EP2 = double2 per lane and column: n <= 64 * EP2
// MAXW: launch bound in warps (11: what a 16-CTA cluster needs up to n = 320, 170 registers per thread; 20: an 8-CTA cluster)
template <int EP2, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
sym_eig_hestenes_cluster_kernel(int n, int warps, int CL, int tile_rows, const double* __restrict__ A_in, const double* __restrict__ V0_in, double* __restrict__ evals,
                                double* __restrict__ evecs, int* __restrict__ info, int max_sweeps) {
  constexpr int COL = 64 * EP2;
  constexpr uint32_t PAIR_BYTES = 2u * COL * 8u;
  extern __shared__ __align__(16) unsigned char hc_smem[];
  const int M = CL * warps;
  double* mail = reinterpret_cast<double*>(hc_smem);               // [2][warps][2][COL]
  double* lam_all = mail + (size_t)2 * warps * 2 * COL;            // [2M]
  uint64_t* bars = reinterpret_cast<uint64_t*>(lam_all + 2 * M);   // [2][warps]
  uint32_t* flags = reinterpret_cast<uint32_t*>(bars + 2 * warps);  // [M]
  double* atile = reinterpret_cast<double*>(hc_smem + ((((size_t)2 * warps * 2 * COL + 2 * M + 2 * warps) * 8 + (size_t)M * 4 + 15) & ~(size_t)15));  // [tile_rows][COL]
  __shared__ double red[HC_MAX_WARPS];
  __shared__ double s_sigma;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t cta = hc_ctarank();
  const int mat = blockIdx.x / CL;
  const int k = (int)cta * warps + w;  // slot
  const double* A = A_in + (size_t)mat * n * n;
  const double* V0 = V0_in ? V0_in + (size_t)mat * n * n : nullptr;

  // who sends this slot its columns (the inverse of the movement below), and whether they sit in this CTA: a local producer
  // writes with plain 128-bit shared stores and ARRIVES on the slot's mbarrier; a producer in a neighbouring CTA sends
  // st.async stores that complete transaction bytes on it (st.async is a ~30-clock instruction per warp: used for every
  // column it made a round cost 5100 clocks at n = 264 against 3200 with st.shared::cluster.f64; only the two columns per
  // CTA and round that cross a CTA boundary take that path now)
  const int src_top = k <= 1 ? 0 : k - 1, src_bot = k == M - 1 ? M - 1 : k + 1;
  const int n_local_src = ((src_top / warps == (int)cta) ? 1 : 0) + ((src_bot / warps == (int)cta) ? 1 : 0);
  if (lane == 0) { mbar_init(&bars[w], 1 + n_local_src); mbar_init(&bars[warps + w], 1 + n_local_src); }
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
      mbar_fence_init();
    }
    __syncthreads();
  }
  const double sigma = s_sigma;

  // ---- W0 = (A + sigma I) V0, this slot's two columns p = 2k, q = 2k + 1; lane l holds rows i0 = 2l + 64e and i0 + 1 -----
  double2 wp[EP2], wq[EP2];
  const int p = 2 * k, q = 2 * k + 1;
  if (V0 == nullptr) {
#pragma unroll
    for (int e = 0; e < EP2; e++) {
      const int i = 2 * lane + 64 * e;
      // A symmetric: column p read as row p
      wp[e].x = (p < n && i < n) ? A[(size_t)p * n + i] + (i == p ? sigma : 0.0) : 0.0;
      wp[e].y = (p < n && i + 1 < n) ? A[(size_t)p * n + i + 1] + (i + 1 == p ? sigma : 0.0) : 0.0;
      wq[e].x = (q < n && i < n) ? A[(size_t)q * n + i] + (i == q ? sigma : 0.0) : 0.0;
      wq[e].y = (q < n && i + 1 < n) ? A[(size_t)q * n + i + 1] + (i + 1 == q ? sigma : 0.0) : 0.0;
    }
  } else {
    double* vp = mail + ((size_t)(1 * warps + w) * 2 + 0) * COL;  // buffer 1 of this slot as staging for the V0 columns
    double* vq = vp + COL;
    for (int i = lane; i < COL; i += 32) {
      vp[i] = (p < n && i < n) ? V0[(size_t)i * n + p] : 0.0;
      vq[i] = (q < n && i < n) ? V0[(size_t)i * n + q] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < EP2; e++) wp[e] = wq[e] = make_double2(0.0, 0.0);
    // A passes through shared memory in tiles of `tile_rows` rows, loaded once per CTA and used by all its warps (read
    // straight from L2 by every warp, the product cost 0.3 ms of a 0.8 ms warm solve at n = 264: 676 KB per warp)
    for (int j0 = 0; j0 < n; j0 += tile_rows) {
      const int rows = min(tile_rows, n - j0);
      __syncthreads();  // the previous tile is no longer read (first pass: the staged V0 columns are complete)
      for (int idx = tid; idx < rows * COL; idx += blockDim.x) {
        const int jj = idx / COL, i = idx - jj * COL;
        atile[idx] = i < n ? __ldg(A + (size_t)(j0 + jj) * n + i) : 0.0;
      }
      __syncthreads();
      if (p < n) {
        for (int jj = 0; jj < rows; jj++) {
          const double xp = vp[j0 + jj], xq = vq[j0 + jj];
          const double2* row = reinterpret_cast<const double2*>(atile + (size_t)jj * COL) + lane;
#pragma unroll
          for (int e = 0; e < EP2; e++) {
            const double2 c = row[32 * e];
            wp[e].x = fma(c.x, xp, wp[e].x); wp[e].y = fma(c.y, xp, wp[e].y);
            wq[e].x = fma(c.x, xq, wq[e].x); wq[e].y = fma(c.y, xq, wq[e].y);
          }
        }
      }
    }
    if (p < n) {
#pragma unroll
      for (int e = 0; e < EP2; e++) {
        const double2 a = *reinterpret_cast<const double2*>(vp + 2 * lane + 64 * e), b = *reinterpret_cast<const double2*>(vq + 2 * lane + 64 * e);
        wp[e].x = fma(sigma, a.x, wp[e].x); wp[e].y = fma(sigma, a.y, wp[e].y);
        wq[e].x = fma(sigma, b.x, wq[e].x); wq[e].y = fma(sigma, b.y, wq[e].y);
      }
    }
    __syncwarp();
  }
  {
    double2* dst = reinterpret_cast<double2*>(mail + ((size_t)(0 * warps + w) * 2) * COL) + lane;
#pragma unroll
    for (int e = 0; e < EP2; e++) { dst[32 * e] = wp[e]; dst[COL / 2 + 32 * e] = wq[e]; }
  }

  // ---- where this slot's columns go after a round (Brent-Luk): top_0 stays; top_1 <- bottom_0; top_k <- top_{k-1};
  //      bottom_k <- bottom_{k+1}; bottom_{M-1} <- top_{M-1} ---------------------------------------------------------------
  int top_slot, top_pos, bot_slot, bot_pos;
  if (k == 0) { top_slot = 0; top_pos = 0; bot_slot = 1; bot_pos = 0; }
  else if (k == M - 1) { top_slot = M - 1; top_pos = 1; bot_slot = M - 2; bot_pos = 1; }
  else { top_slot = k + 1; top_pos = 0; bot_slot = k - 1; bot_pos = 1; }
  const uint32_t mail_s = smem_u32(mail), bars_s = smem_u32(bars);
  auto dest = [&](int slot, int pos, int buf) -> uint32_t {
    const int dc = slot / warps, dw = slot - dc * warps;
    const uint32_t off = (uint32_t)((((size_t)buf * warps + dw) * 2 + pos) * COL + 2 * lane) * 8u;
    return hc_mapa(mail_s + off, (uint32_t)dc);
  };
  auto dest_bar = [&](int slot, int buf) -> uint32_t {
    const int dc = slot / warps, dw = slot - dc * warps;
    return hc_mapa(bars_s + 8u * (uint32_t)(buf * warps + dw), (uint32_t)dc);
  };
  const bool top_local = top_slot / warps == (int)cta, bot_local = bot_slot / warps == (int)cta;
  auto local_col = [&](int slot, int pos, int buf) -> double2* {
    return reinterpret_cast<double2*>(mail + (((size_t)buf * warps + (slot - (int)cta * warps)) * 2 + pos) * COL) + lane;
  };
  const uint32_t dtop0 = dest(top_slot, top_pos, 0), dtop1 = dest(top_slot, top_pos, 1);
  const uint32_t dbot0 = dest(bot_slot, bot_pos, 0), dbot1 = dest(bot_slot, bot_pos, 1);
  const uint32_t btop0 = dest_bar(top_slot, 0), btop1 = dest_bar(top_slot, 1), bbot0 = dest_bar(bot_slot, 0), bbot1 = dest_bar(bot_slot, 1);
  const uint32_t flags_s = smem_u32(flags), lam_s = smem_u32(lam_all);

  const double tol = sqrt((double)n) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  const double theta2 = 1e-18;  // (1e-9)^2: a sweep without a rotation above this is the last one
  hc_cluster_sync();  // every CTA of the cluster is resident, its mbarriers initialised, its initial columns written

  int sweeps = 0;
  unsigned R = 0;  // rounds done so far: round R reads buffer R & 1, completed by the (R-1)>>1-th phase of its mbarrier
  bool converged = false;
  const int rounds = 2 * M - 1;
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    bool big = false;
    for (int r = 0; r < rounds; r++, R++) {
      const int buf = R & 1;
      if (lane == 0) {  // arm the other buffer for round R + 1: this slot's own arrival, with the bytes the remote producers will send
        if (n_local_src == 2) mbar_arrive(&bars[(buf ^ 1) * warps + w]);
        else mbar_expect_tx(&bars[(buf ^ 1) * warps + w], (uint32_t)(2 - n_local_src) * (PAIR_BYTES / 2));
      }
      if (R > 0) mbar_wait(&bars[buf * warps + w], ((R - 1) >> 1) & 1);
      const double2* src = reinterpret_cast<const double2*>(mail + ((size_t)(buf * warps + w) * 2) * COL) + lane;
      double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
      for (int e = 0; e < EP2; e++) {
        wp[e] = src[32 * e];
        wq[e] = src[COL / 2 + 32 * e];
        a = fma(wp[e].x, wp[e].x, a); a = fma(wp[e].y, wp[e].y, a);
        b = fma(wq[e].x, wq[e].x, b); b = fma(wq[e].y, wq[e].y, b);
        g = fma(wp[e].x, wq[e].x, g); g = fma(wp[e].y, wq[e].y, g);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      double c = 1.0, s = 0.0;
      const double d = 0.5 * (b - a);
      const double h = fma(d, d, g * g), g2 = g * g, ab = a * b;
      if (g2 > tol2 * ab && h > 1e-290) {
        big = big || g2 > theta2 * ab;
        const double rh = rsqrt(h);
        const double u = fma(0.5 * fabs(d), rh, 0.5);  // (1 + cos 2theta) / 2 in [1/2, 1]
        const double ru = rsqrt(u);
        c = u * ru;
        s = 0.5 * g * rh * ru;
        if (d < 0.0) s = -s;
      }
      const uint32_t dt = buf ? dtop0 : dtop1, db = buf ? dbot0 : dbot1;  // the other buffer
      const uint32_t bt = buf ? btop0 : btop1, bb = buf ? bbot0 : bbot1;
      if (top_local) {
        double2* o = local_col(top_slot, top_pos, buf ^ 1);
#pragma unroll
        for (int e = 0; e < EP2; e++) o[32 * e] = make_double2(fma(c, wp[e].x, -s * wq[e].x), fma(c, wp[e].y, -s * wq[e].y));
      } else {
#pragma unroll
        for (int e = 0; e < EP2; e++) hc_st_async_v2(dt + 512u * e, fma(c, wp[e].x, -s * wq[e].x), fma(c, wp[e].y, -s * wq[e].y), bt);
      }
      if (bot_local) {
        double2* o = local_col(bot_slot, bot_pos, buf ^ 1);
#pragma unroll
        for (int e = 0; e < EP2; e++) o[32 * e] = make_double2(fma(s, wp[e].x, c * wq[e].x), fma(s, wp[e].y, c * wq[e].y));
      } else {
#pragma unroll
        for (int e = 0; e < EP2; e++) hc_st_async_v2(db + 512u * e, fma(s, wp[e].x, c * wq[e].x), fma(s, wp[e].y, c * wq[e].y), bb);
      }
      if (top_local || bot_local) {
        __syncwarp();  // all lanes' stores before lane 0's releasing arrive
        if (lane == 0) {
          if (top_local) mbar_arrive(&bars[(buf ^ 1) * warps + (top_slot - (int)cta * warps)]);
          if (bot_local) mbar_arrive(&bars[(buf ^ 1) * warps + (bot_slot - (int)cta * warps)]);
        }
      }
    }
    sweeps = sweep + 1;
    if (lane < CL) hc_st_u32(hc_mapa(flags_s + 4u * k, (uint32_t)lane), big ? 1u : 0u);
    hc_cluster_sync();
    uint32_t f = 0;
    for (int i = lane; i < M; i += 32) f |= flags[i];
    converged = !__any_sync(0xffffffffu, f != 0);
    hc_cluster_sync();  // nobody overwrites a flag of the next sweep before everyone has read this one's
  }

  // ---- eigenvalues = column norms - sigma; ascending order; eigenvectors = normalised columns ----------------------------
  const int buf = R & 1;
  if (R > 0) mbar_wait(&bars[buf * warps + w], ((R - 1) >> 1) & 1);
  const double2* src = reinterpret_cast<const double2*>(mail + ((size_t)(buf * warps + w) * 2) * COL) + lane;
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int e = 0; e < EP2; e++) {
    wp[e] = src[32 * e];
    wq[e] = src[COL / 2 + 32 * e];
    a = fma(wp[e].x, wp[e].x, a); a = fma(wp[e].y, wp[e].y, a);
    b = fma(wq[e].x, wq[e].x, b); b = fma(wq[e].y, wq[e].y, b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double na = sqrt(a), nb = sqrt(b);
  const double lp = a > 0.0 ? na - sigma : (a == 0.0 ? inf : a), lq = b > 0.0 ? nb - sigma : (b == 0.0 ? inf : b);  // zero column = padding; NaN stays NaN
  if (lane < CL) hc_st_f64(hc_mapa(lam_s + 8u * (2 * k), (uint32_t)lane), lp);
  else if (lane < 2 * CL) hc_st_f64(hc_mapa(lam_s + 8u * (2 * k + 1), (uint32_t)(lane - CL)), lq);
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
    for (int e = 0; e < EP2; e++) {
      const int i = 2 * lane + 64 * e;
      if (i < n) vec[(size_t)i * n + rp] = wp[e].x * ia;
      if (i + 1 < n) vec[(size_t)(i + 1) * n + rp] = wp[e].y * ia;
    }
  }
  if (lq != inf && rq < n) {
    const double ib = 1.0 / nb;
    if (lane == 0) ev[rq] = lq;
#pragma unroll
    for (int e = 0; e < EP2; e++) {
      const int i = 2 * lane + 64 * e;
      if (i < n) vec[(size_t)i * n + rq] = wq[e].x * ib;
      if (i + 1 < n) vec[(size_t)(i + 1) * n + rq] = wq[e].y * ib;
    }
  }
  if (info != nullptr && k == 0 && lane == 0) info[mat] = bad ? -2 : (converged ? sweeps : -1);
}

template <int EP2, int MAXW>
static int launch_hestenes_w(cudaStream_t stream, int64_t batch, int n, int CL, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  const int m = (n + 1) / 2;
  const int warps = (m + CL - 1) / CL;
  if (warps > MAXW) return GDFT_BAD_SHAPE;
  auto kern = sym_eig_hestenes_cluster_kernel<EP2, MAXW>;
  {
    const int M = CL * warps;
    const size_t base = ((((size_t)2 * warps * 2 * 64 * EP2 + 2 * M + 2 * warps) * 8 + (size_t)M * 4 + 15) & ~(size_t)15);
    int tile_rows = 0;
    if (V0 != nullptr) {  // rows of A staged per pass of the W0 = (A + sigma I) V0 product: what fits next to the mailboxes, at most 32
      tile_rows = (int)((size_t)(227 * 1024 - 1024 - base) / ((size_t)64 * EP2 * 8));
      tile_rows = tile_rows > 32 ? 32 : tile_rows;
      if (tile_rows < 1) return GDFT_BAD_SHAPE;
    }
    const size_t smem = base + (size_t)tile_rows * 64 * EP2 * 8;
    GDFT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) GDFT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(batch * CL));
    cfg.blockDim = dim3((unsigned)(warps * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CL > 8) {  // can this device co-schedule a 16-CTA cluster of this shape at all?  If not, the caller falls back to 8
      static std::atomic<int> can16[HC_MAX_WARPS + 1];  // capability cache per warp count (0 unknown, 1 yes, 2 no): a pure function of the device
      int st = can16[warps].load(std::memory_order_relaxed);
      if (st == 0) {
        int nclusters = 0;
        const bool ok = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1;
        if (!ok) (void)cudaGetLastError();
        st = ok ? 1 : 2;
        can16[warps].store(st, std::memory_order_relaxed);
      }
      if (st == 2) return -1;
    }
    GDFT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, n, warps, CL, tile_rows, A, V0, evals, evecs, info, (int)HC_MAX_SWEEPS));
    GDFT_LAUNCH_CHECK();
    return GDFT_OK;
  }
}

template <int EP2>
static int launch_hestenes(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  // cluster of 16 CTAs (non-portable size, supported on B200) from n = 160 on: half the columns, hence half the FP64-pipe and
  // shared-memory time, per SM and round (n = 264: 0.36 -> 0.25 ms per sweep); GDFT_EIGH_CLUSTER = 8 | 16 overrides
  int CL = n >= 160 ? 16 : 8;
  if (const char* e = getenv("GDFT_EIGH_CLUSTER")) { const int v = atoi(e); if (v == 8 || v == 16) CL = v; }
  const int m = (n + 1) / 2;
  if (CL == 16 && (m + 15) / 16 <= 11) {
    const int rc = launch_hestenes_w<EP2, 11>(stream, batch, n, 16, A, V0, evals, evecs, info);
    if (rc != -1) return rc;
  }
  if ((m + 7) / 8 <= 11) return launch_hestenes_w<EP2, 11>(stream, batch, n, 8, A, V0, evals, evecs, info);
  return launch_hestenes_w<EP2, 20>(stream, batch, n, 8, A, V0, evals, evecs, info);
}

int sym_eigh_cluster(cudaStream_t stream, int64_t batch, int n, const double* A, const double* V0, double* evals, double* evecs, int* info) {
  switch ((n + 63) / 64) {
    case 1: case 2: return launch_hestenes<2>(stream, batch, n, A, V0, evals, evecs, info);
    case 3: return launch_hestenes<3>(stream, batch, n, A, V0, evals, evecs, info);
    case 4: return launch_hestenes<4>(stream, batch, n, A, V0, evals, evecs, info);
    case 5: return launch_hestenes<5>(stream, batch, n, A, V0, evals, evecs, info);
    default: return GDFT_BAD_SHAPE;
  }
}

int sym_eigh_cluster_max_n() { return HC_MAX_N; }

}  // namespace gdft
