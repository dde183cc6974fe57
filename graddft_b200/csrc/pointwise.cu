// K5 -- closed-form per-point features, forward and reverse (one thread per grid point; HBM-streaming).
// The reverse pass evaluates the same template code on forward-mode dual numbers over the six local
// variables (rho_a, rho_b, sigma_aa, sigma_bb, x_a, x_b) with x = laplacian (LYP) or tau (DM21 MGGA),
// contracts with the output cotangent and applies d sigma_ss / d grad_rho_s = 2 grad_rho_s.
// The reverse pass of the reverse pass (gdft_pointwise_bwd2: what differentiating V_xc once more needs, i.e.
// training through the SCF loop, grad_dft/evaluate.py:917-1038) runs the same code on dual numbers whose
// components are themselves dual numbers carrying the incoming direction.
#include "common.cuh"
#include "pointwise_math.h"

namespace gdft {

using pw::Dual;

__host__ __device__ inline int pointwise_ncols(int id) {
  switch (id) {
    case GDFT_PW_LSDA_X: case GDFT_PW_B88_X: case GDFT_PW_VWN_C: case GDFT_PW_LYP_C: case GDFT_PW_PW92_C: case GDFT_PW_DM21_LDA: return 1;
    case GDFT_PW_B88_SET: case GDFT_PW_DM21_GGA: return 2;
    case GDFT_PW_B3LYP_SET: case GDFT_PW_DM21_MGGA: return 4;
    case GDFT_PW_DM21_INPUTS: return 7;
    case GDFT_PW_FEAT_LDA: return 4;
    case GDFT_PW_FEAT_GGA: return 8;
    case GDFT_PW_FEAT_MGGA: return 16;
    default: return 0;
  }
}
// bit 0: grad_rho, bit 1: lapl, bit 2: tau
__host__ __device__ inline int pointwise_needs(int id) {
  switch (id) {
    case GDFT_PW_B88_X: case GDFT_PW_B88_SET: case GDFT_PW_DM21_GGA: case GDFT_PW_FEAT_GGA: return 1;
    case GDFT_PW_LYP_C: case GDFT_PW_B3LYP_SET: return 1 | 2;
    case GDFT_PW_DM21_MGGA: case GDFT_PW_FEAT_MGGA: return 1 | 4;
    case GDFT_PW_DM21_INPUTS: return 1 | 4;
    default: return 0;
  }
}

// output columns / columns that are actually evaluated (the rest are the upstream-zero correlation features)
template <int ID> struct PwCols {
  static constexpr int F = (ID == GDFT_PW_B88_SET || ID == GDFT_PW_DM21_GGA) ? 2 : (ID == GDFT_PW_B3LYP_SET || ID == GDFT_PW_DM21_MGGA) ? 4
                         : ID == GDFT_PW_FEAT_LDA ? 4 : ID == GDFT_PW_FEAT_GGA ? 8 : ID == GDFT_PW_FEAT_MGGA ? 16 : 1;
  static constexpr int FE = (ID == GDFT_PW_FEAT_LDA || ID == GDFT_PW_FEAT_GGA || ID == GDFT_PW_FEAT_MGGA) ? F / 2 : F;
};

// feats[] for every id except DM21_INPUTS; v = {ra, rb, saa, sbb, xa, xb}
template <int ID, typename T>
__device__ __forceinline__ void eval_features(const T (&v)[6], double clip, T* feats) {
  if (ID == GDFT_PW_LSDA_X) feats[0] = pw::lsda_x(v[0], v[1], clip);
  if (ID == GDFT_PW_B88_X) feats[0] = pw::b88_x(v[0], v[1], v[2], v[3], clip);
  if (ID == GDFT_PW_VWN_C) feats[0] = pw::vwn_c(v[0], v[1], clip);
  if (ID == GDFT_PW_LYP_C) feats[0] = pw::lyp_c(v[0], v[1], v[2], v[3], v[4], v[5], clip);
  if (ID == GDFT_PW_PW92_C) feats[0] = pw::pw92_c(v[0], v[1], clip);
  if (ID == GDFT_PW_B88_SET) {
    feats[0] = pw::lsda_x(v[0], v[1], clip);
    feats[1] = pw::b88_x(v[0], v[1], v[2], v[3], clip);
  }
  if (ID == GDFT_PW_B3LYP_SET) {
    feats[0] = pw::lsda_x(v[0], v[1], clip);
    feats[1] = pw::b88_x(v[0], v[1], v[2], v[3], clip);
    feats[2] = pw::vwn_c(v[0], v[1], clip);
    feats[3] = pw::lyp_c(v[0], v[1], v[2], v[3], v[4], v[5], clip);
  }
  if (ID == GDFT_PW_DM21_LDA || ID == GDFT_PW_DM21_GGA || ID == GDFT_PW_DM21_MGGA) {
    const int nu = (ID == GDFT_PW_DM21_LDA) ? 1 : 2, nw = (ID == GDFT_PW_DM21_MGGA) ? 2 : 1;
    int col = 0;
    for (int i = 0; i < nu; i++)
      for (int j = 0; j < nw; j++) {
        T term = pw::dm21_term_spin(v[0], v[2], v[4], i, j, clip) + pw::dm21_term_spin(v[1], v[3], v[5], i, j, clip);
        if (i == 0 && j == 0) term = term * (-2.0 * pw::PI * pow(3.0 / (4.0 * pw::PI), 4.0 / 3.0));
        feats[col++] = term;
      }
  }
  if (ID == GDFT_PW_FEAT_LDA || ID == GDFT_PW_FEAT_GGA || ID == GDFT_PW_FEAT_MGGA) {
    const int nu = (ID == GDFT_PW_FEAT_LDA) ? 1 : 2, nw = (ID == GDFT_PW_FEAT_MGGA) ? 2 : 1;
    int col = 0;
    for (int i = 0; i < nu; i++)
      for (int j = 0; j < nw; j++) {
        feats[col++] = pw::mgga_term_spin(v[0], v[2], v[4], i, j, clip);
        feats[col++] = pw::mgga_term_spin(v[1], v[3], v[5], i, j, clip);
      }
  }
}

struct PwArgs {
  int64_t N;
  double clip;
  const double *rho, *grho, *tau, *lapl, *out_bar;
  double *out, *rho_bar, *grho_bar, *tau_bar, *lapl_bar;
};

template <int ID>
__global__ void __launch_bounds__(128) pointwise_fwd_kernel(const PwArgs a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  constexpr int F = PwCols<ID>::F, FE = PwCols<ID>::FE;
  const int needs = pointwise_needs(ID);
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  double v[6] = {rho.x, rho.y, 0, 0, 0, 0};
  if (needs & 1) {
    const double* g = a.grho + r * 6;
    v[2] = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    v[3] = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
  }
  if (needs & 2) { const double2 l = reinterpret_cast<const double2*>(a.lapl)[r]; v[4] = l.x; v[5] = l.y; }
  if (needs & 4) { const double2 l = reinterpret_cast<const double2*>(a.tau)[r]; v[4] = l.x; v[5] = l.y; }
  double feats[FE];
  eval_features<ID, double>(v, a.clip, feats);
#pragma unroll
  for (int f = 0; f < F; f++) a.out[r * F + f] = f < FE ? feats[f < FE ? f : 0] : 0.0;
}

template <int ID>
__global__ void __launch_bounds__(128) pointwise_bwd_kernel(const PwArgs a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  constexpr int F = PwCols<ID>::F, FE = PwCols<ID>::FE;
  const int needs = pointwise_needs(ID);
  typedef Dual<6> D6;
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  double g[6] = {0, 0, 0, 0, 0, 0};
  double x[6] = {rho.x, rho.y, 0, 0, 0, 0};
  if (needs & 1) {
#pragma unroll
    for (int q = 0; q < 6; q++) g[q] = a.grho[r * 6 + q];
    x[2] = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    x[3] = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
  }
  if (needs & 2) { const double2 l = reinterpret_cast<const double2*>(a.lapl)[r]; x[4] = l.x; x[5] = l.y; }
  if (needs & 4) { const double2 l = reinterpret_cast<const double2*>(a.tau)[r]; x[4] = l.x; x[5] = l.y; }
  D6 v[6];
#pragma unroll
  for (int q = 0; q < 6; q++) v[q] = pw::Make<D6>::variable(x[q], q);
  D6 feats[FE];
  eval_features<ID, D6>(v, a.clip, feats);
  double d[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int f = 0; f < FE; f++) {
    const double ob = a.out_bar[r * F + f];
#pragma unroll
    for (int q = 0; q < 6; q++) d[q] = fma(ob, feats[f].d[q], d[q]);
  }
  if (a.rho_bar) reinterpret_cast<double2*>(a.rho_bar)[r] = make_double2(d[0], d[1]);
  if (a.grho_bar) {
    double* o = a.grho_bar + r * 6;
#pragma unroll
    for (int j = 0; j < 3; j++) { o[j] = 2.0 * d[2] * g[j]; o[3 + j] = 2.0 * d[3] * g[3 + j]; }
  }
  if (a.lapl_bar) reinterpret_cast<double2*>(a.lapl_bar)[r] = (needs & 2) ? make_double2(d[4], d[5]) : make_double2(0.0, 0.0);
  if (a.tau_bar) reinterpret_cast<double2*>(a.tau_bar)[r] = (needs & 4) ? make_double2(d[4], d[5]) : make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------------------------------
// First-order XC build of a closed-form functional in ONE pass per grid point (value + VJP): features on dual numbers (their
// values ARE the forward pass), the optional exact-exchange column h = sum_{w,s} e_HF[w,s,r] (popular_functionals.py:330-338),
// abs_clip of the densities (functional.py:160-185), e = sum_f c_f d_f with a constant coefficient row, abs_clip of e and of the
// quadrature weight, E = sum_r w_r e_r (functional.py:219-253, 316-342) -- and, because the seed of the reverse pass is known
// per point (dE/dd_f = w_r [|e_r| > clip] c_f [|d_f| > clip]), the cotangents of rho / grad_rho / tau / lapl and of e_HF in the
// same thread.  Replaces pointwise_fwd + cat + abs_clip + integrate_fwd + integrate_bwd + abs_clip' + slice copies + pointwise_bwd
// (ten launches, two evaluations of the formulas) inside the predictor's first-order path; same per-point arithmetic as those
// kernels (same templates, same fma order over the columns).
// ---------------------------------------------------------------------------------------------------------------------
struct XcPointArgs {
  int64_t N;
  double clip;
  int W;  // number of omegas of the exact-exchange column (0: none)
  double coef[8];
  const double *rho, *grho, *tau, *lapl, *ehf, *w;
  double *partial, *rho_bar, *grho_bar, *tau_bar, *lapl_bar, *ehf_bar;
};

template <int ID>
__global__ void __launch_bounds__(128) xc_point_kernel(const XcPointArgs a) {
  constexpr int F = PwCols<ID>::F, FE = PwCols<ID>::FE;
  const int needs = pointwise_needs(ID);
  typedef Dual<6> D6;
  __shared__ double red[4];
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0.0;
  if (r < a.N) {
    const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
    double g[6] = {0, 0, 0, 0, 0, 0};
    double x[6] = {rho.x, rho.y, 0, 0, 0, 0};
    if (needs & 1) {
#pragma unroll
      for (int q = 0; q < 6; q++) g[q] = a.grho[r * 6 + q];
      x[2] = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
      x[3] = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
    }
    if (needs & 2) { const double2 l = reinterpret_cast<const double2*>(a.lapl)[r]; x[4] = l.x; x[5] = l.y; }
    if (needs & 4) { const double2 l = reinterpret_cast<const double2*>(a.tau)[r]; x[4] = l.x; x[5] = l.y; }
    D6 v[6];
#pragma unroll
    for (int q = 0; q < 6; q++) v[q] = pw::Make<D6>::variable(x[q], q);
    D6 feats[FE];
    eval_features<ID, D6>(v, a.clip, feats);
    double e = 0.0, seed[FE];
#pragma unroll
    for (int f = 0; f < FE; f++) {
      const double raw = feats[f].v;
      const bool m = fabs(raw) > a.clip;
      e = fma(a.coef[f], m ? raw : 0.0, e);
      seed[f] = m ? a.coef[f] : 0.0;
    }
    double h = 0.0;
    bool mh = false;
    if (a.W > 0) {
      for (int q = 0; q < 2 * a.W; q++) h += a.ehf[(size_t)q * a.N + r];
      mh = fabs(h) > a.clip;
      e = fma(a.coef[F], mh ? h : 0.0, e);
    }
    const double wr = a.w[r];
    const double wc = fabs(wr) > a.clip ? wr : 0.0;
    const bool live = fabs(e) > a.clip;
    contrib = live ? wc * e : 0.0;
    const double eb = live ? wc : 0.0;
    double d[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int f = 0; f < FE; f++) {
      const double ob = eb * seed[f];
#pragma unroll
      for (int q = 0; q < 6; q++) d[q] = fma(ob, feats[f].d[q], d[q]);
    }
    if (a.rho_bar) reinterpret_cast<double2*>(a.rho_bar)[r] = make_double2(d[0], d[1]);
    if (a.grho_bar) {
      double* o = a.grho_bar + r * 6;
#pragma unroll
      for (int j = 0; j < 3; j++) { o[j] = 2.0 * d[2] * g[j]; o[3 + j] = 2.0 * d[3] * g[3 + j]; }
    }
    if (a.lapl_bar) reinterpret_cast<double2*>(a.lapl_bar)[r] = (needs & 2) ? make_double2(d[4], d[5]) : make_double2(0.0, 0.0);
    if (a.tau_bar) reinterpret_cast<double2*>(a.tau_bar)[r] = (needs & 4) ? make_double2(d[4], d[5]) : make_double2(0.0, 0.0);
    if (a.W > 0 && a.ehf_bar) {
      const double hb = mh ? eb * a.coef[F] : 0.0;
      for (int q = 0; q < 2 * a.W; q++) a.ehf_bar[(size_t)q * a.N + r] = hb;
    }
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) a.partial[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

// E = sum of the per-CTA partial sums: one CTA, fixed order
__global__ void __launch_bounds__(1024) xc_point_sum_kernel(int64_t count, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < count; i += 1024) acc += partial[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v2 = red[threadIdx.x];
    v2 = warp_sum(v2);
    if (threadIdx.x == 0) out[0] = v2;
  }
}

// VJP of pointwise_bwd_kernel.  With X = (rho, grad_rho, x) the inputs, ob the output cotangent and
// Xbar(X, ob) = J(X)^T ob the first-order result, this kernel receives the cotangent U of Xbar and returns
//   ob_bar[f] = (J U)_f                      (a directional derivative: the inner dual part of the value)
//   X_t       = d/dX <U, Xbar(X, ob)>        (mixed second derivatives: inner dual part of the outer derivatives,
//                                             plus the explicit dependence of d sigma/d grad_rho on grad_rho)
struct Pw2Args {
  int64_t N;
  double clip;
  const double *rho, *grho, *tau, *lapl, *out_bar;
  const double *u_rho, *u_grho, *u_tau, *u_lapl;
  double *out_bar_bar, *rho_t, *grho_t, *tau_t, *lapl_t;
};

template <int ID>
__global__ void __launch_bounds__(128) pointwise_bwd2_kernel(const Pw2Args a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  constexpr int F = PwCols<ID>::F, FE = PwCols<ID>::FE;
  const int needs = pointwise_needs(ID);
  typedef Dual<1> B1;
  typedef Dual<6, B1> H;
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  double g[6] = {0, 0, 0, 0, 0, 0}, ug[6] = {0, 0, 0, 0, 0, 0};
  double x[6] = {rho.x, rho.y, 0, 0, 0, 0};
  double w[6] = {0, 0, 0, 0, 0, 0};
  if (a.u_rho) { const double2 u = reinterpret_cast<const double2*>(a.u_rho)[r]; w[0] = u.x; w[1] = u.y; }
  if (needs & 1) {
#pragma unroll
    for (int q = 0; q < 6; q++) g[q] = a.grho[r * 6 + q];
    x[2] = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    x[3] = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
    if (a.u_grho) {
#pragma unroll
      for (int q = 0; q < 6; q++) ug[q] = a.u_grho[r * 6 + q];
      w[2] = 2.0 * (g[0] * ug[0] + g[1] * ug[1] + g[2] * ug[2]);
      w[3] = 2.0 * (g[3] * ug[3] + g[4] * ug[4] + g[5] * ug[5]);
    }
  }
  if (needs & 2) {
    const double2 l = reinterpret_cast<const double2*>(a.lapl)[r]; x[4] = l.x; x[5] = l.y;
    if (a.u_lapl) { const double2 u = reinterpret_cast<const double2*>(a.u_lapl)[r]; w[4] = u.x; w[5] = u.y; }
  }
  if (needs & 4) {
    const double2 l = reinterpret_cast<const double2*>(a.tau)[r]; x[4] = l.x; x[5] = l.y;
    if (a.u_tau) { const double2 u = reinterpret_cast<const double2*>(a.u_tau)[r]; w[4] = u.x; w[5] = u.y; }
  }
  H v[6];
#pragma unroll
  for (int q = 0; q < 6; q++) {
    B1 b; b.v = x[q]; b.d[0] = w[q];
    v[q] = pw::Make<H>::variable(b, q);
  }
  H feats[FE];
  eval_features<ID, H>(v, a.clip, feats);
  double d[6] = {0, 0, 0, 0, 0, 0}, h[6] = {0, 0, 0, 0, 0, 0};
  if (a.out_bar_bar)
    for (int f = FE; f < F; f++) a.out_bar_bar[r * F + f] = 0.0;
#pragma unroll
  for (int f = 0; f < FE; f++) {
    const double ob = a.out_bar[r * F + f];
    if (a.out_bar_bar) a.out_bar_bar[r * F + f] = feats[f].v.d[0];
#pragma unroll
    for (int q = 0; q < 6; q++) { d[q] = fma(ob, feats[f].d[q].v, d[q]); h[q] = fma(ob, feats[f].d[q].d[0], h[q]); }
  }
  if (a.rho_t) reinterpret_cast<double2*>(a.rho_t)[r] = make_double2(h[0], h[1]);
  if (a.grho_t) {
    double* o = a.grho_t + r * 6;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      o[j] = 2.0 * (h[2] * g[j] + d[2] * ug[j]);
      o[3 + j] = 2.0 * (h[3] * g[3 + j] + d[3] * ug[3 + j]);
    }
  }
  if (a.lapl_t) reinterpret_cast<double2*>(a.lapl_t)[r] = (needs & 2) ? make_double2(h[4], h[5]) : make_double2(0.0, 0.0);
  if (a.tau_t) reinterpret_cast<double2*>(a.tau_t)[r] = (needs & 4) ? make_double2(h[4], h[5]) : make_double2(0.0, 0.0);
}

// functional.py:520-531: [rho'_a, rho'_b, |g_a+g_b|^2, |g_a|^2, |g_b|^2, tau_a, tau_b], rho' = max(|rho|,clip) sign(rho)
__global__ void __launch_bounds__(128) dm21_inputs_fwd_kernel(const PwArgs a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  const double2 tau = reinterpret_cast<const double2*>(a.tau)[r];
  const double* g = a.grho + r * 6;
  auto squash = [&](double x) { const double s = (x > 0.0) - (x < 0.0); return fmax(fabs(x), a.clip) * s; };
  double* o = a.out + r * 7;
  o[0] = squash(rho.x); o[1] = squash(rho.y);
  double st = 0, sa = 0, sb = 0;
#pragma unroll
  for (int j = 0; j < 3; j++) { const double t = g[j] + g[3 + j]; st += t * t; sa += g[j] * g[j]; sb += g[3 + j] * g[3 + j]; }
  o[2] = st; o[3] = sa; o[4] = sb; o[5] = tau.x; o[6] = tau.y;
}
__global__ void __launch_bounds__(128) dm21_inputs_bwd_kernel(const PwArgs a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  const double* g = a.grho + r * 6;
  const double* ob = a.out_bar + r * 7;
  // d/dx [max(|x|,c) sign(x)] = 1 where |x| > c (ties follow torch.maximum: 1/2), else 0
  auto dsq = [&](double x) { const double ax = fabs(x); return ax > a.clip ? 1.0 : (ax == a.clip ? 0.5 : 0.0); };
  if (a.rho_bar) reinterpret_cast<double2*>(a.rho_bar)[r] = make_double2(ob[0] * dsq(rho.x), ob[1] * dsq(rho.y));
  if (a.grho_bar) {
    double* o = a.grho_bar + r * 6;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double t = 2.0 * (g[j] + g[3 + j]) * ob[2];
      o[j] = t + 2.0 * g[j] * ob[3];
      o[3 + j] = t + 2.0 * g[3 + j] * ob[4];
    }
  }
  if (a.tau_bar) reinterpret_cast<double2*>(a.tau_bar)[r] = make_double2(ob[5], ob[6]);
  if (a.lapl_bar) reinterpret_cast<double2*>(a.lapl_bar)[r] = make_double2(0.0, 0.0);
}

// VJP of dm21_inputs_bwd_kernel: the squash derivative is piecewise constant and the gradient columns are quadratic
// forms, so only the (grad_rho, out_bar) cross terms survive.
__global__ void __launch_bounds__(128) dm21_inputs_bwd2_kernel(const Pw2Args a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  const double2 rho = reinterpret_cast<const double2*>(a.rho)[r];
  const double* g = a.grho + r * 6;
  const double* ob = a.out_bar + r * 7;
  auto dsq = [&](double x) { const double ax = fabs(x); return ax > a.clip ? 1.0 : (ax == a.clip ? 0.5 : 0.0); };
  double ur[2] = {0, 0}, ug[6] = {0, 0, 0, 0, 0, 0}, ut[2] = {0, 0};
  if (a.u_rho) { const double2 u = reinterpret_cast<const double2*>(a.u_rho)[r]; ur[0] = u.x; ur[1] = u.y; }
  if (a.u_grho)
    for (int q = 0; q < 6; q++) ug[q] = a.u_grho[r * 6 + q];
  if (a.u_tau) { const double2 u = reinterpret_cast<const double2*>(a.u_tau)[r]; ut[0] = u.x; ut[1] = u.y; }
  if (a.out_bar_bar) {
    double* o = a.out_bar_bar + r * 7;
    double st = 0, sa = 0, sb = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      st += 2.0 * (g[j] + g[3 + j]) * (ug[j] + ug[3 + j]);
      sa += 2.0 * g[j] * ug[j];
      sb += 2.0 * g[3 + j] * ug[3 + j];
    }
    o[0] = dsq(rho.x) * ur[0]; o[1] = dsq(rho.y) * ur[1];
    o[2] = st; o[3] = sa; o[4] = sb; o[5] = ut[0]; o[6] = ut[1];
  }
  if (a.rho_t) reinterpret_cast<double2*>(a.rho_t)[r] = make_double2(0.0, 0.0);
  if (a.grho_t) {
    double* o = a.grho_t + r * 6;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double t = 2.0 * ob[2] * (ug[j] + ug[3 + j]);
      o[j] = t + 2.0 * ob[3] * ug[j];
      o[3 + j] = t + 2.0 * ob[4] * ug[3 + j];
    }
  }
  if (a.tau_t) reinterpret_cast<double2*>(a.tau_t)[r] = make_double2(0.0, 0.0);
  if (a.lapl_t) reinterpret_cast<double2*>(a.lapl_t)[r] = make_double2(0.0, 0.0);
}

template <int ID>
static void launch_pw2(cudaStream_t st, const Pw2Args& a) {
  pointwise_bwd2_kernel<ID><<<(unsigned)((a.N + 127) / 128), 128, 0, st>>>(a);
}

static int dispatch_pw2(cudaStream_t st, int id, const Pw2Args& a) {
  switch (id) {
    case GDFT_PW_LSDA_X: launch_pw2<GDFT_PW_LSDA_X>(st, a); break;
    case GDFT_PW_B88_X: launch_pw2<GDFT_PW_B88_X>(st, a); break;
    case GDFT_PW_VWN_C: launch_pw2<GDFT_PW_VWN_C>(st, a); break;
    case GDFT_PW_LYP_C: launch_pw2<GDFT_PW_LYP_C>(st, a); break;
    case GDFT_PW_PW92_C: launch_pw2<GDFT_PW_PW92_C>(st, a); break;
    case GDFT_PW_B3LYP_SET: launch_pw2<GDFT_PW_B3LYP_SET>(st, a); break;
    case GDFT_PW_B88_SET: launch_pw2<GDFT_PW_B88_SET>(st, a); break;
    case GDFT_PW_DM21_LDA: launch_pw2<GDFT_PW_DM21_LDA>(st, a); break;
    case GDFT_PW_DM21_GGA: launch_pw2<GDFT_PW_DM21_GGA>(st, a); break;
    case GDFT_PW_DM21_MGGA: launch_pw2<GDFT_PW_DM21_MGGA>(st, a); break;
    case GDFT_PW_FEAT_LDA: launch_pw2<GDFT_PW_FEAT_LDA>(st, a); break;
    case GDFT_PW_FEAT_GGA: launch_pw2<GDFT_PW_FEAT_GGA>(st, a); break;
    case GDFT_PW_FEAT_MGGA: launch_pw2<GDFT_PW_FEAT_MGGA>(st, a); break;
    case GDFT_PW_DM21_INPUTS: dm21_inputs_bwd2_kernel<<<(unsigned)((a.N + 127) / 128), 128, 0, st>>>(a); break;
    default: return GDFT_BAD_ARGUMENT;
  }
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

template <int ID>
static void launch_pw(bool bwd, cudaStream_t st, const PwArgs& a) {
  const unsigned grid = (unsigned)((a.N + 127) / 128);
  if (bwd) pointwise_bwd_kernel<ID><<<grid, 128, 0, st>>>(a);
  else pointwise_fwd_kernel<ID><<<grid, 128, 0, st>>>(a);
}

static int dispatch_pw(bool bwd, cudaStream_t st, int id, const PwArgs& a) {
  const unsigned grid = (unsigned)((a.N + 127) / 128);
  switch (id) {
    case GDFT_PW_LSDA_X: launch_pw<GDFT_PW_LSDA_X>(bwd, st, a); break;
    case GDFT_PW_B88_X: launch_pw<GDFT_PW_B88_X>(bwd, st, a); break;
    case GDFT_PW_VWN_C: launch_pw<GDFT_PW_VWN_C>(bwd, st, a); break;
    case GDFT_PW_LYP_C: launch_pw<GDFT_PW_LYP_C>(bwd, st, a); break;
    case GDFT_PW_PW92_C: launch_pw<GDFT_PW_PW92_C>(bwd, st, a); break;
    case GDFT_PW_B3LYP_SET: launch_pw<GDFT_PW_B3LYP_SET>(bwd, st, a); break;
    case GDFT_PW_B88_SET: launch_pw<GDFT_PW_B88_SET>(bwd, st, a); break;
    case GDFT_PW_DM21_LDA: launch_pw<GDFT_PW_DM21_LDA>(bwd, st, a); break;
    case GDFT_PW_DM21_GGA: launch_pw<GDFT_PW_DM21_GGA>(bwd, st, a); break;
    case GDFT_PW_DM21_MGGA: launch_pw<GDFT_PW_DM21_MGGA>(bwd, st, a); break;
    case GDFT_PW_FEAT_LDA: launch_pw<GDFT_PW_FEAT_LDA>(bwd, st, a); break;
    case GDFT_PW_FEAT_GGA: launch_pw<GDFT_PW_FEAT_GGA>(bwd, st, a); break;
    case GDFT_PW_FEAT_MGGA: launch_pw<GDFT_PW_FEAT_MGGA>(bwd, st, a); break;
    case GDFT_PW_DM21_INPUTS:
      if (bwd) dm21_inputs_bwd_kernel<<<grid, 128, 0, st>>>(a);
      else dm21_inputs_fwd_kernel<<<grid, 128, 0, st>>>(a);
      break;
    default: return GDFT_BAD_ARGUMENT;
  }
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

static int check_pw_inputs(int64_t N, int id, const double* rho, const double* grho, const double* tau, const double* lapl) {
  if (N <= 0) return GDFT_BAD_SHAPE;
  if (id < 0 || id >= GDFT_PW_COUNT) return GDFT_BAD_ARGUMENT;
  const int needs = pointwise_needs(id);
  if (!rho || ((needs & 1) && !grho) || ((needs & 2) && !lapl) || ((needs & 4) && !tau)) return GDFT_BAD_ARGUMENT;
  if (!aligned16(rho) || !aligned16(tau) || !aligned16(lapl)) return GDFT_BAD_ALIGNMENT;
  return GDFT_OK;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_pointwise_ncols(int id) { return pointwise_ncols(id); }

extern "C" int gdft_pointwise_fwd(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho, const double* grad_rho,
                                  const double* tau, const double* lapl, double* out) {
  int rc = check_pw_inputs(N, id, rho, grad_rho, tau, lapl);
  if (rc) return rc;
  if (!out) return GDFT_BAD_ARGUMENT;
  PwArgs a{};
  a.N = N; a.clip = clip; a.rho = rho; a.grho = grad_rho; a.tau = tau; a.lapl = lapl; a.out = out;
  return dispatch_pw(false, static_cast<cudaStream_t>(stream), id, a);
}

extern "C" int gdft_pointwise_bwd(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho, const double* grad_rho,
                                  const double* tau, const double* lapl, const double* out_bar, double* rho_bar,
                                  double* grad_rho_bar, double* tau_bar, double* lapl_bar) {
  int rc = check_pw_inputs(N, id, rho, grad_rho, tau, lapl);
  if (rc) return rc;
  if (!out_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(rho_bar) || !aligned16(tau_bar) || !aligned16(lapl_bar)) return GDFT_BAD_ALIGNMENT;
  PwArgs a{};
  a.N = N; a.clip = clip; a.rho = rho; a.grho = grad_rho; a.tau = tau; a.lapl = lapl; a.out_bar = out_bar;
  a.rho_bar = rho_bar; a.grho_bar = grad_rho_bar; a.tau_bar = tau_bar; a.lapl_bar = lapl_bar;
  return dispatch_pw(true, static_cast<cudaStream_t>(stream), id, a);
}

extern "C" int gdft_pointwise_bwd2(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho, const double* grad_rho,
                                   const double* tau, const double* lapl, const double* out_bar, const double* u_rho,
                                   const double* u_grad_rho, const double* u_tau, const double* u_lapl, double* out_bar_bar,
                                   double* rho_t, double* grad_rho_t, double* tau_t, double* lapl_t) {
  int rc = check_pw_inputs(N, id, rho, grad_rho, tau, lapl);
  if (rc) return rc;
  if (!out_bar) return GDFT_BAD_ARGUMENT;
  if (!aligned16(u_rho) || !aligned16(u_tau) || !aligned16(u_lapl) || !aligned16(rho_t) || !aligned16(tau_t) || !aligned16(lapl_t))
    return GDFT_BAD_ALIGNMENT;
  Pw2Args a{};
  a.N = N; a.clip = clip; a.rho = rho; a.grho = grad_rho; a.tau = tau; a.lapl = lapl; a.out_bar = out_bar;
  a.u_rho = u_rho; a.u_grho = u_grad_rho; a.u_tau = u_tau; a.u_lapl = u_lapl;
  a.out_bar_bar = out_bar_bar; a.rho_t = rho_t; a.grho_t = grad_rho_t; a.tau_t = tau_t; a.lapl_t = lapl_t;
  return dispatch_pw2(static_cast<cudaStream_t>(stream), id, a);
}

// First-order XC build of a closed-form functional in one pass (see xc_point_kernel): E_xc and the cotangents of the grid
// quantities (and of e_HF when an exact-exchange column with W omegas is present).  coef: F feature coefficients followed,
// when W > 0, by the coefficient of the exact-exchange column.  Workspace: ceil(N / 128) doubles.
extern "C" size_t gdft_xc_point_workspace(int64_t N) { return N > 0 ? (size_t)((N + 127) / 128) * 8 + 256 : 0; }

extern "C" int gdft_xc_point_fused(gdft_stream_t stream_, int64_t N, int id, double clip, const double* coef, int ncoef, const double* rho,
                                   const double* grad_rho, const double* tau, const double* lapl, const double* ehf, int W, const double* w,
                                   double* E, double* rho_bar, double* grad_rho_bar, double* tau_bar, double* lapl_bar, double* ehf_bar,
                                   void* ws, size_t ws_bytes) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  int rc = check_pw_inputs(N, id, rho, grad_rho, tau, lapl);
  if (rc) return rc;
  const int F = pointwise_ncols(id);
  if (id == GDFT_PW_DM21_INPUTS || F <= 0 || F > 7 || W < 0 || W > 8 || ncoef != F + (W > 0 ? 1 : 0)) return GDFT_BAD_SHAPE;
  if (!coef || !w || !E || !rho_bar || (W > 0 && (!ehf || !ehf_bar))) return GDFT_BAD_ARGUMENT;
  if (!aligned16(rho_bar) || !aligned16(tau_bar) || !aligned16(lapl_bar)) return GDFT_BAD_ALIGNMENT;
  if (ws_bytes < gdft_xc_point_workspace(N)) return GDFT_WORKSPACE_TOO_SMALL;
  XcPointArgs a{};
  a.N = N; a.clip = clip; a.W = W;
  for (int i = 0; i < ncoef; i++) a.coef[i] = coef[i];
  a.rho = rho; a.grho = grad_rho; a.tau = tau; a.lapl = lapl; a.ehf = ehf; a.w = w;
  a.partial = static_cast<double*>(ws);
  a.rho_bar = rho_bar; a.grho_bar = grad_rho_bar; a.tau_bar = tau_bar; a.lapl_bar = lapl_bar; a.ehf_bar = ehf_bar;
  const unsigned grid = (unsigned)((N + 127) / 128);
  switch (id) {
    case GDFT_PW_LSDA_X: xc_point_kernel<GDFT_PW_LSDA_X><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_B88_X: xc_point_kernel<GDFT_PW_B88_X><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_VWN_C: xc_point_kernel<GDFT_PW_VWN_C><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_LYP_C: xc_point_kernel<GDFT_PW_LYP_C><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_PW92_C: xc_point_kernel<GDFT_PW_PW92_C><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_B3LYP_SET: xc_point_kernel<GDFT_PW_B3LYP_SET><<<grid, 128, 0, st>>>(a); break;
    case GDFT_PW_B88_SET: xc_point_kernel<GDFT_PW_B88_SET><<<grid, 128, 0, st>>>(a); break;
    default: return GDFT_BAD_ARGUMENT;
  }
  GDFT_LAUNCH_CHECK();
  xc_point_sum_kernel<<<1, 1024, 0, st>>>((int64_t)grid, a.partial, E);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
