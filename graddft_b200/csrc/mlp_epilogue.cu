// Row f2: the residual block of the DM21 coefficient network, grad_dft/functional.py:809-819,
//   z = dense(x) + x ;  n = (z - mean z) / sqrt(var z + eps) ;  o = n * scale + bias ;  out = elu(o),
// where dense(x) = x K + k: the library GEMM delivers y = x K, the Dense bias k is added here (gdft_dense_ln_elu_*), and its
// cotangent -- the column sum of z_bar -- comes out of the reverse pass with the LayerNorm parameter cotangents,
// as ONE streaming pass over the [N, W] activations after the (library) FP64 GEMM, forward and reverse.  The host
// framework would run it as ~10 elementwise/reduction kernels per block, each a full HBM round trip of a 1 GB tensor
// at the benzene shape; here a row lives in the registers of one warp (W <= 512 doubles: 16 per lane), the two
// LayerNorm moments are warp-shuffle reductions, and HBM sees 2 reads + 1 write (forward) / 3 reads + 1 write (reverse).
// HBM-bound: 8*W*3 B per row forward, 8*W*4 B per row reverse.
// The reverse pass recomputes z from its two inputs (cheaper than saving it) and reads the per-row (mean, rstd) the
// forward left behind; the parameter cotangents (scale, bias) are accumulated per lane over a fixed set of rows,
// reduced across the CTA's warps in shared memory and then across CTAs in fixed order (bitwise reproducible).
#include "common.cuh"

namespace gdft {

constexpr int LN_WARPS = 8;
constexpr int LN_THREADS = 32 * LN_WARPS;
constexpr int LN_MAXV = 8;  // double2 per lane -> W <= 512
constexpr int LN_MAX_CTAS = 148 * 4;

struct LnArgs {
  int64_t N;
  int W;
  double eps;
  const double *y, *ybias, *res, *gamma, *beta, *stats_in, *dout, *fwd_out;
  double *out, *stats_out, *dz, *partial;
};

__device__ __forceinline__ double elu_val(double o) { return o > 0.0 ? o : expm1(o); }
__device__ __forceinline__ double elu_der(double o) { return o > 0.0 ? 1.0 : exp(o); }

// MAXV = double2 per lane the row needs (W <= 64 * MAXV): sizing the register arrays for W = 512 when W = 256 cost the
// reverse pass 205 registers per thread, one CTA per SM and a latency-bound 41 % of DRAM bandwidth (ncu)
template <int MAXV>
__global__ void __launch_bounds__(LN_THREADS) ln_elu_fwd_kernel(const LnArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = a.W >> 1;
  const double inv_w = 1.0 / a.W;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < a.N; row += (int64_t)gridDim.x * LN_WARPS) {
    const double2* y2 = reinterpret_cast<const double2*>(a.y + row * a.W);
    const double2* r2 = a.res ? reinterpret_cast<const double2*>(a.res + row * a.W) : nullptr;
    double2 z[MAXV];
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      z[i] = make_double2(0.0, 0.0);
      if (c < nv) {
        z[i] = __ldcs(y2 + c);
        if (a.ybias) { const double2 yb = reinterpret_cast<const double2*>(a.ybias)[c]; z[i].x += yb.x; z[i].y += yb.y; }
        if (r2) { const double2 r = __ldcs(r2 + c); z[i].x += r.x; z[i].y += r.y; }
        s += z[i].x + z[i].y;
      }
    }
    const double mean = warp_sum(s) * inv_w;
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      if (lane + 32 * i < nv) { const double dx = z[i].x - mean, dy = z[i].y - mean; v += dx * dx + dy * dy; }
    }
    const double rstd = 1.0 / sqrt(warp_sum(v) * inv_w + a.eps);
    double2* o2 = reinterpret_cast<double2*>(a.out + row * a.W);
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const double2 g = reinterpret_cast<const double2*>(a.gamma)[c], b = reinterpret_cast<const double2*>(a.beta)[c];
        o2[c] = make_double2(elu_val((z[i].x - mean) * rstd * g.x + b.x), elu_val((z[i].y - mean) * rstd * g.y + b.y));
      }
    }
    if (lane == 0 && a.stats_out) reinterpret_cast<double2*>(a.stats_out)[row] = make_double2(mean, rstd);
  }
}

template <bool PARAM_GRADS, int MAXV>
__global__ void __launch_bounds__(LN_THREADS, MAXV <= 4 ? 2 : 1) ln_elu_bwd_kernel(const LnArgs a) {
  extern __shared__ __align__(16) double sred[];  // [LN_WARPS][3][W] when PARAM_GRADS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = a.W >> 1;
  const double inv_w = 1.0 / a.W;
  double2 pg[MAXV], pb[MAXV], pz[MAXV];  // column sums: scale_bar, bias_bar, and z_bar (= the Dense bias cotangent)
#pragma unroll
  for (int i = 0; i < MAXV; i++) pg[i] = pb[i] = pz[i] = make_double2(0.0, 0.0);
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < a.N; row += (int64_t)gridDim.x * LN_WARPS) {
    const double2* y2 = reinterpret_cast<const double2*>(a.y + row * a.W);
    const double2* r2 = a.res ? reinterpret_cast<const double2*>(a.res + row * a.W) : nullptr;
    const double2* d2 = reinterpret_cast<const double2*>(a.dout + row * a.W);
    const double2 st = reinterpret_cast<const double2*>(a.stats_in)[row];
    const double mean = st.x, rstd = st.y;
    double2 nrm[MAXV], dn[MAXV];
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      nrm[i] = dn[i] = make_double2(0.0, 0.0);
      if (c < nv) {
        double2 z = __ldcs(y2 + c);
        if (a.ybias) { const double2 yb = reinterpret_cast<const double2*>(a.ybias)[c]; z.x += yb.x; z.y += yb.y; }
        if (r2) { const double2 r = __ldcs(r2 + c); z.x += r.x; z.y += r.y; }
        const double2 g = reinterpret_cast<const double2*>(a.gamma)[c], b = reinterpret_cast<const double2*>(a.beta)[c];
        const double2 dy = __ldcs(d2 + c);
        nrm[i] = make_double2((z.x - mean) * rstd, (z.y - mean) * rstd);
        double dox, doy;
        if (a.fwd_out) {
          // elu'(o) from the forward OUTPUT: out > 0 <=> o > 0, and for o <= 0 out = expm1(o), so exp(o) = out + 1 -- one more
          // 8-byte read per element instead of an FP64 exp (the reverse pass is FP64-pipe-bound, not HBM-bound, with it)
          const double2 fo = __ldcs(reinterpret_cast<const double2*>(a.fwd_out + row * a.W) + c);
          dox = dy.x * (fo.x > 0.0 ? 1.0 : fo.x + 1.0);
          doy = dy.y * (fo.y > 0.0 ? 1.0 : fo.y + 1.0);
        } else {
          dox = dy.x * elu_der(nrm[i].x * g.x + b.x);
          doy = dy.y * elu_der(nrm[i].y * g.y + b.y);
        }
        if (PARAM_GRADS) {
          pg[i].x = fma(dox, nrm[i].x, pg[i].x); pg[i].y = fma(doy, nrm[i].y, pg[i].y);
          pb[i].x += dox; pb[i].y += doy;
        }
        dn[i] = make_double2(dox * g.x, doy * g.y);
        s1 += dn[i].x + dn[i].y;
        s2 += dn[i].x * nrm[i].x + dn[i].y * nrm[i].y;
      }
    }
    const double m1 = warp_sum(s1) * inv_w, m2 = warp_sum(s2) * inv_w;
    double2* z2 = reinterpret_cast<double2*>(a.dz + row * a.W);
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const double2 zb = make_double2(rstd * (dn[i].x - m1 - nrm[i].x * m2), rstd * (dn[i].y - m1 - nrm[i].y * m2));
        z2[c] = zb;
        if (PARAM_GRADS) { pz[i].x += zb.x; pz[i].y += zb.y; }
      }
    }
  }
  if (PARAM_GRADS) {
    double2* s2p = reinterpret_cast<double2*>(sred);
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c = lane + 32 * i;
      if (c < nv) { s2p[(warp * 3 + 0) * nv + c] = pg[i]; s2p[(warp * 3 + 1) * nv + c] = pb[i]; s2p[(warp * 3 + 2) * nv + c] = pz[i]; }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * a.W; idx += LN_THREADS) {  // idx = which * W + col
      double acc = 0.0;
      for (int w = 0; w < LN_WARPS; w++) acc += sred[(size_t)w * 3 * a.W + idx];
      a.partial[(size_t)blockIdx.x * 3 * a.W + idx] = acc;
    }
  }
}

// out[which][col] = sum_b partial[b][which][col]  (fixed order).  32 outputs x 8 block-groups per CTA: group g sums
// b = g, g+8, ... with four independent loads in flight, the groups are added in order through shared memory (a serial
// loop over the ~600 partial rows is a chain of dependent L2 round trips: 34 us per call).
constexpr int LNR_OUT = 32, LNR_G = 8;
__global__ void __launch_bounds__(LNR_OUT* LNR_G) ln_param_reduce_kernel(int nblocks, int W, const double* __restrict__ partial,
                                                                         double* __restrict__ dgamma, double* __restrict__ dbeta,
                                                                         double* __restrict__ dybias) {
  __shared__ double red[LNR_G][LNR_OUT];
  const int o = threadIdx.x & (LNR_OUT - 1), g = threadIdx.x / LNR_OUT;
  const int idx = blockIdx.x * LNR_OUT + o;
  const size_t stride = (size_t)3 * W;
  double acc = 0.0;
  if (idx < 3 * W) {
    const double* src = partial + idx;
    int b = g;
    for (; b + 3 * LNR_G < nblocks; b += 4 * LNR_G) {
      const double v0 = src[(size_t)b * stride], v1 = src[(size_t)(b + LNR_G) * stride];
      const double v2 = src[(size_t)(b + 2 * LNR_G) * stride], v3 = src[(size_t)(b + 3 * LNR_G) * stride];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; b < nblocks; b += LNR_G) acc += src[(size_t)b * stride];
  }
  red[g][o] = acc;
  __syncthreads();
  if (g == 0 && idx < 3 * W) {
    double t = red[0][o];
#pragma unroll
    for (int j = 1; j < LNR_G; j++) t += red[j][o];
    double* dst = idx < W ? dgamma : idx < 2 * W ? dbeta : dybias;
    if (dst) dst[idx % W] = t;
  }
}

static int ln_grid(int64_t N) { return (int)imin64(LN_MAX_CTAS, (N + LN_WARPS - 1) / LN_WARPS); }

size_t ln_elu_workspace(int64_t N, int64_t W) {
  if (N <= 0 || W <= 0) return 0;
  return (size_t)ln_grid(N) * 3 * (size_t)W * 8 + 256;
}

static int ln_check(int64_t N, int64_t W) {
  if (N <= 0 || W <= 0 || W > 2 * 32 * LN_MAXV || (W & 1) || N > (int64_t)1 << 40) return GDFT_BAD_SHAPE;
  return GDFT_OK;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_dense_ln_elu_fwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* ybias, const double* res,
                                     const double* scale, const double* bias, double eps, double* out, double* stats) {
  if (int rc = ln_check(N, W)) return rc;
  if (!y || !scale || !bias || !out) return GDFT_BAD_ARGUMENT;
  if (!aligned16(y) || !aligned16(ybias) || !aligned16(res) || !aligned16(scale) || !aligned16(bias) || !aligned16(out) || !aligned16(stats))
    return GDFT_BAD_ALIGNMENT;
  LnArgs a{};
  a.N = N; a.W = (int)W; a.eps = eps; a.y = y; a.ybias = ybias; a.res = res; a.gamma = scale; a.beta = bias; a.out = out; a.stats_out = stats;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (W <= 128) ln_elu_fwd_kernel<2><<<ln_grid(N), LN_THREADS, 0, st>>>(a);
  else if (W <= 256) ln_elu_fwd_kernel<4><<<ln_grid(N), LN_THREADS, 0, st>>>(a);
  else ln_elu_fwd_kernel<8><<<ln_grid(N), LN_THREADS, 0, st>>>(a);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_ln_elu_fwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* res, const double* scale,
                               const double* bias, double eps, double* out, double* stats) {
  return gdft_dense_ln_elu_fwd(stream, N, W, y, nullptr, res, scale, bias, eps, out, stats);
}

extern "C" int gdft_dense_ln_elu_bwd(gdft_stream_t stream_, int64_t N, int64_t W, const double* y, const double* ybias, const double* res,
                                     const double* scale, const double* bias, const double* stats, const double* fwd_out,
                                     const double* out_bar, double* z_bar, double* scale_bar, double* bias_bar, double* ybias_bar, void* ws,
                                     size_t ws_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = ln_check(N, W)) return rc;
  if (!y || !scale || !bias || !stats || !out_bar || !z_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(y) || !aligned16(ybias) || !aligned16(res) || !aligned16(scale) || !aligned16(bias) || !aligned16(stats) ||
      !aligned16(fwd_out) || !aligned16(out_bar) || !aligned16(z_bar) || !aligned16(ws))
    return GDFT_BAD_ALIGNMENT;
  const bool pgrads = scale_bar || bias_bar || ybias_bar;
  if (pgrads && ws_bytes < ln_elu_workspace(N, W)) return GDFT_WORKSPACE_TOO_SMALL;
  LnArgs a{};
  a.N = N; a.W = (int)W; a.y = y; a.ybias = ybias; a.res = res; a.gamma = scale; a.beta = bias; a.stats_in = stats; a.dout = out_bar;
  a.dz = z_bar; a.fwd_out = fwd_out;
  a.partial = static_cast<double*>(ws);
  const int grid = ln_grid(N);
  if (pgrads) {
    const size_t smem = (size_t)LN_WARPS * 3 * W * 8;
    if (W <= 128) {
      GDFT_CUDA_TRY((cudaFuncSetAttribute(ln_elu_bwd_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
      ln_elu_bwd_kernel<true, 2><<<grid, LN_THREADS, smem, stream>>>(a);
    } else if (W <= 256) {
      GDFT_CUDA_TRY((cudaFuncSetAttribute(ln_elu_bwd_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
      ln_elu_bwd_kernel<true, 4><<<grid, LN_THREADS, smem, stream>>>(a);
    } else {
      GDFT_CUDA_TRY((cudaFuncSetAttribute(ln_elu_bwd_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
      ln_elu_bwd_kernel<true, 8><<<grid, LN_THREADS, smem, stream>>>(a);
    }
    GDFT_LAUNCH_CHECK();
    ln_param_reduce_kernel<<<(unsigned)((3 * W + LNR_OUT - 1) / LNR_OUT), LNR_OUT * LNR_G, 0, stream>>>(grid, (int)W, a.partial, scale_bar,
                                                                                                      bias_bar, ybias_bar);
    GDFT_LAUNCH_CHECK();
  } else {
    if (W <= 128) ln_elu_bwd_kernel<false, 2><<<grid, LN_THREADS, 0, stream>>>(a);
    else if (W <= 256) ln_elu_bwd_kernel<false, 4><<<grid, LN_THREADS, 0, stream>>>(a);
    else ln_elu_bwd_kernel<false, 8><<<grid, LN_THREADS, 0, stream>>>(a);
    GDFT_LAUNCH_CHECK();
  }
  return GDFT_OK;
}

extern "C" int gdft_ln_elu_bwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* res, const double* scale,
                               const double* bias, const double* stats, const double* out_bar, double* z_bar, double* scale_bar,
                               double* bias_bar, void* ws, size_t ws_bytes) {
  return gdft_dense_ln_elu_bwd(stream, N, W, y, nullptr, res, scale, bias, stats, nullptr, out_bar, z_bar, scale_bar, bias_bar, nullptr, ws,
                               ws_bytes);
}
