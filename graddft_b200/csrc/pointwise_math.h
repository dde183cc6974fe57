// Closed-form per-grid-point functionals of Grad DFT, written once as scalar-type-generic templates:
// T = double gives values, T = Dual<NV> (forward-mode dual numbers) gives values + exact first
// derivatives with the reference's jnp.clip / jnp.where sub-gradient conventions
// (clip: zero derivative where clipped; where: derivative of the selected branch only).
// Compiles as CUDA device code (pointwise.cu) and as plain host C++ (tests/native, CPU unit tests of
// the formulas against the oracle -- the product only ever runs the device build).
//
// Formulas follow, line by line:
//   grad_dft/functional.py:950-979   exchange_polarization_correction
//   grad_dft/functional.py:982-1045  correlation_polarization_correction
//   grad_dft/popular_functionals.py:29-50 lsda_x_e, 52-103 b88_x_e, 105-139 pw92_c_e,
//                                   141-195 vwn_c_e, 197-269 lyp_c_e
//   grad_dft/functional.py:504-531   dm21_coefficient_inputs, 534-626 dm21_densities
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GDFT_HD __host__ __device__ __forceinline__
#else
#define GDFT_HD inline
#endif

namespace gdft {
namespace pw {

constexpr double LN2 = 0.693147180559945309417232121458;
constexpr double PI = 3.14159265358979323846264338328;

// Forward-mode dual number over a base scalar B: B = double gives first derivatives; B = Dual<1> (itself a dual
// number carrying one direction) gives, in addition, the derivative of every first derivative along that
// direction -- what the VJP of the VJP needs (training through the SCF loop differentiates V_xc once more).
template <int NV, typename B = double>
struct Dual {
  B v;
  B d[NV];
};

// ---- construction ---------------------------------------------------------------------------------
template <typename T> struct Make;
template <> struct Make<double> {
  static GDFT_HD double constant(double c) { return c; }
};
template <int NV, typename B> struct Make<Dual<NV, B>> {
  static GDFT_HD Dual<NV, B> constant(double c) {
    Dual<NV, B> r; r.v = Make<B>::constant(c);
#pragma unroll
    for (int i = 0; i < NV; i++) r.d[i] = Make<B>::constant(0.0);
    return r;
  }
  static GDFT_HD Dual<NV, B> variable(const B& x, int idx) {
    Dual<NV, B> r; r.v = x;
#pragma unroll
    for (int i = 0; i < NV; i++) r.d[i] = Make<B>::constant((i == idx) ? 1.0 : 0.0);
    return r;
  }
};

GDFT_HD double val(double x) { return x; }
template <int NV, typename B> GDFT_HD double val(const Dual<NV, B>& x) { return val(x.v); }

// chain rule helper: f(x) with value fv and derivative fd (both of the base type)
template <int NV, typename B> GDFT_HD Dual<NV, B> chain(const Dual<NV, B>& x, const B& fv, const B& fd) {
  Dual<NV, B> r; r.v = fv;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = fd * x.d[i];
  return r;
}

// ---- arithmetic -----------------------------------------------------------------------------------
template <int NV, typename B> GDFT_HD Dual<NV, B> operator+(const Dual<NV, B>& a, const Dual<NV, B>& b) {
  Dual<NV, B> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator-(const Dual<NV, B>& a, const Dual<NV, B>& b) {
  Dual<NV, B> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator*(const Dual<NV, B>& a, const Dual<NV, B>& b) {
  Dual<NV, B> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator/(const Dual<NV, B>& a, const Dual<NV, B>& b) {
  Dual<NV, B> r; r.v = a.v / b.v;
  const B inv = 1.0 / b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator-(const Dual<NV, B>& a) {
  Dual<NV, B> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = -a.d[i];
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator+(const Dual<NV, B>& a, double c) { Dual<NV, B> r = a; r.v = r.v + c; return r; }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator+(double c, const Dual<NV, B>& a) { return a + c; }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator-(const Dual<NV, B>& a, double c) { Dual<NV, B> r = a; r.v = r.v - c; return r; }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator-(double c, const Dual<NV, B>& a) { return (-a) + c; }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator*(const Dual<NV, B>& a, double c) {
  Dual<NV, B> r; r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < NV; i++) r.d[i] = a.d[i] * c;
  return r;
}
template <int NV, typename B> GDFT_HD Dual<NV, B> operator*(double c, const Dual<NV, B>& a) { return a * c; }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator/(const Dual<NV, B>& a, double c) { return a * (1.0 / c); }
template <int NV, typename B> GDFT_HD Dual<NV, B> operator/(double c, const Dual<NV, B>& a) {
  const B q = c / a.v;
  return chain(a, q, -(q / a.v));
}

// ---- elementary functions (double overloads first) --------------------------------------------------
GDFT_HD double f_log2(double x) { return log2(x); }
GDFT_HD double f_exp2(double x) { return exp2(x); }
GDFT_HD double f_log(double x) { return log(x); }
GDFT_HD double f_exp(double x) { return exp(x); }
GDFT_HD double f_sqrt(double x) { return sqrt(x); }
GDFT_HD double f_atan(double x) { return atan(x); }
GDFT_HD double f_asinh(double x) { return asinh(x); }
GDFT_HD double f_pow(double x, double p) { return pow(x, p); }
GDFT_HD double f_clip_min(double x, double c) { return x >= c ? x : c; }
GDFT_HD double f_select(bool cond, double a, double b) { return cond ? a : b; }

// Each derivative is written in base-type arithmetic, so that with B = Dual<1> it is itself differentiated.
template <int NV, typename B> GDFT_HD Dual<NV, B> f_log2(const Dual<NV, B>& x) { return chain(x, f_log2(x.v), (1.0 / LN2) / x.v); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_exp2(const Dual<NV, B>& x) { const B e = f_exp2(x.v); return chain(x, e, e * LN2); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_log(const Dual<NV, B>& x) { return chain(x, f_log(x.v), 1.0 / x.v); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_exp(const Dual<NV, B>& x) { const B e = f_exp(x.v); return chain(x, e, e); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_sqrt(const Dual<NV, B>& x) { const B s = f_sqrt(x.v); return chain(x, s, 0.5 / s); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_atan(const Dual<NV, B>& x) { return chain(x, f_atan(x.v), 1.0 / (1.0 + x.v * x.v)); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_asinh(const Dual<NV, B>& x) { return chain(x, f_asinh(x.v), 1.0 / f_sqrt(1.0 + x.v * x.v)); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_pow(const Dual<NV, B>& x, double p) { return chain(x, f_pow(x.v, p), p * f_pow(x.v, p - 1.0)); }
// jnp.clip(x, a_min=c): value max(x,c); derivative passes where x >= c (torch.clamp convention at the tie)
template <int NV, typename B> GDFT_HD Dual<NV, B> f_clip_min(const Dual<NV, B>& x, double c) {
  return val(x) >= c ? x : Make<Dual<NV, B>>::constant(c);
}
template <int NV, typename B> GDFT_HD Dual<NV, B> f_select(bool cond, const Dual<NV, B>& a, const Dual<NV, B>& b) { return cond ? a : b; }

// 2^(4 log2(x) / 3), the reference's log-domain x^{4/3} (functional.py:1014-1016).  Value exactly as written
// there; derivative (4/3) x^{1/3}, which stays finite (0) at x == 0 where differentiating through log2
// would give 0 * inf (fully spin-polarised points).  One level further down (the derivative of that derivative,
// (4/9) x^{-2/3}, is infinite at x == 0) the exact-zero point is treated as a constant.
GDFT_HD double f_pow43_log2(double x) { return exp2(4.0 * log2(x) / 3.0); }
GDFT_HD double f_cbrt43(double x) { return (4.0 / 3.0) * exp2(log2(x) / 3.0); }
template <int NV, typename B> GDFT_HD Dual<NV, B> f_cbrt43(const Dual<NV, B>& x) {
  if (val(x) == 0.0) return Make<Dual<NV, B>>::constant(0.0);
  return chain(x, f_cbrt43(x.v), (4.0 / 9.0) * f_exp2(-2.0 * f_log2(x.v) / 3.0));
}
template <int NV, typename B> GDFT_HD Dual<NV, B> f_pow43_log2(const Dual<NV, B>& x) {
  return chain(x, f_pow43_log2(x.v), f_cbrt43(x.v));
}

template <typename T> GDFT_HD T cst(double c) { return Make<T>::constant(c); }

// ---- spin interpolation -------------------------------------------------------------------------------
constexpr double FZ_DEN = 0.51984209978974632953442121455650;  // 2 (2^{1/3} - 1)
constexpr double FZ_PP0 = 1.70992093416136561756560043006;     // f''(0) = 8 / (9 * FZ_DEN)
constexpr double LOG2_RS0 = -0.688844542917054;              // log2((3/(4 pi))^{1/3})

// functional.py:973-979
template <typename T> GDFT_HD T exchange_polarization(const T& eP, const T& eF, const T& ra, const T& rb) {
  const T zeta = (ra - rb) / (ra + rb);
  const T fz = (f_pow(1.0 - zeta, 4.0 / 3.0) + f_pow(1.0 + zeta, 4.0 / 3.0) - 2.0) / FZ_DEN;
  return eP + (eF - eP) * fz;
}

// The PW92 "G" function in the reference's exp2/log2 form: 2A(1+a1 rs) ln(1 + 1/(2A(b1 rs^1/2 + b2 rs + b3 rs^3/2 + b4 rs^2)))
template <typename T> GDFT_HD T pw_G(const T& log_rs, double A, double a1, double b1, double b2, double b3, double b4) {
  const T ars = f_exp2(log2(a1) + log_rs);
  const T brs_1_2 = f_exp2(log2(b1) + log_rs / 2.0);
  const T brs = f_exp2(log2(b2) + log_rs);
  const T brs_3_2 = f_exp2(log2(b3) + 3.0 * log_rs / 2.0);
  const T brs2 = f_exp2(log2(b4) + 2.0 * log_rs);
  return 2.0 * A * (1.0 + ars) * f_log(1.0 + (1.0 / (2.0 * A)) / (brs_1_2 + brs + brs_3_2 + brs2));
}

// functional.py:1008-1045
template <typename T> GDFT_HD T correlation_polarization(const T& eP, const T& eF, const T& ra, const T& rb, double clip) {
  const T rt = ra + rb;
  const T log_rho = f_log2(f_clip_min(rt, clip));
  const T log_rs = LOG2_RS0 - log_rho / 3.0;
  const T zeta = f_select(val(rt) > clip, (ra - rb) / rt, cst<T>(0.0));
  const T alphac = pw_G(log_rs, 0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  const T zm = f_pow43_log2(1.0 - zeta);
  const T zp = f_pow43_log2(1.0 + zeta);
  const T fz = (zm + zp - 2.0) / FZ_DEN;
  const T z2 = zeta * zeta;
  const T z4 = z2 * z2;
  return eP + alphac * (fz / FZ_PP0) * (1.0 - z4) + (eF - eP) * fz * z4;
}

// ---- energy densities ---------------------------------------------------------------------------------
// popular_functionals.py:41-50
template <typename T> GDFT_HD T lsda_x(const T& ra_in, const T& rb_in, double clip) {
  const T ra = f_clip_min(ra_in, clip), rb = f_clip_min(rb_in, clip);
  const T rt43 = f_pow(ra + rb, 4.0 / 3.0);
  const double cP = -0.75 * cbrt(3.0 / PI), cF = -0.75 * cbrt(6.0 / PI);
  return exchange_polarization(cP * rt43, cF * rt43, ra, rb);
}

// popular_functionals.py:70-97, one spin channel's contribution (positive; caller negates the sum)
template <typename T> GDFT_HD T b88_x_spin(const T& r_in, const T& sigma, double clip) {
  const double beta = 0.0042;
  const T r = f_clip_min(r_in, clip);
  const T log_rho = f_log2(f_clip_min(r, clip));
  const T log_g = f_log2(f_clip_min(sigma, clip)) / 2.0;
  const T log_x = log_g - (4.0 / 3.0) * log_rho;
  const T x = f_exp2(log_x);
  return beta * f_exp2(4.0 * log_rho / 3.0 + 2.0 * log_x - f_log2(1.0 + 6.0 * beta * x * f_asinh(x)));
}
template <typename T> GDFT_HD T b88_x(const T& ra, const T& rb, const T& saa, const T& sbb, double clip) {
  return -(b88_x_spin(ra, saa, clip) + b88_x_spin(rb, sbb, clip));
}

// popular_functionals.py:120-139
template <typename T> GDFT_HD T pw92_c(const T& ra, const T& rb, double clip) {
  const T rt = ra + rb;
  const T log_rho = f_log2(f_clip_min(rt, clip));
  const T log_rs = LOG2_RS0 - log_rho / 3.0;
  const T eP = -pw_G(log_rs, 0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const T eF = -pw_G(log_rs, 0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  return correlation_polarization(eP, eF, ra, rb, clip) * rt;
}

template <typename T> GDFT_HD T vwn_ePF(const T& x, const T& log_x, double A, double b, double c, double x0) {
  const T X = f_exp2(2.0 * log_x) + f_exp2(log_x + log2(b)) + c;
  const double X0 = x0 * x0 + b * x0 + c;
  const double Q = sqrt(4.0 * c - b * b);
  const T at = f_atan(Q / (2.0 * x + b));
  const T xm = x - x0;
  return (A / 2.0) * (2.0 * f_log(x) - f_log(X) + (2.0 * b / Q) * at -
                      (b * x0 / X0) * (f_log(xm * xm / X) + (2.0 * (2.0 * x0 + b) / Q) * at));
}
// popular_functionals.py:158-195
template <typename T> GDFT_HD T vwn_c(const T& ra_in, const T& rb_in, double clip) {
  const T ra = f_select(val(ra_in) > clip, ra_in, cst<T>(0.0));
  const T rb = f_select(val(rb_in) > clip, rb_in, cst<T>(0.0));
  const T rt = ra + rb;
  const T log_rho = f_log2(f_clip_min(rt, clip));
  const T log_rs = LOG2_RS0 - log_rho / 3.0;
  const T log_x = log_rs / 2.0;
  const T x = f_exp2(log_x);
  const T eP = vwn_ePF(x, log_x, 0.0621814, 3.72744, 12.9352, -0.10498);
  const T eF = vwn_ePF(x, log_x, 0.0621814 / 2, 7.06042, 18.0578, -0.325);
  return correlation_polarization(eP, eF, ra, rb, clip) * rt;
}

// popular_functionals.py:229-269
template <typename T>
GDFT_HD T lyp_c(const T& ra_in, const T& rb_in, const T& saa, const T& sbb, const T& la, const T& lb, double clip) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const double CF = 0.3 * pow(3.0 * PI * PI, 2.0 / 3.0);
  const T ra = f_clip_min(ra_in, clip), rb = f_clip_min(rb_in, clip);
  const T zero = cst<T>(0.0);
  const T ta = (f_select(val(ra) > clip, saa / ra, zero) - la) / 8.0;
  const T tb = (f_select(val(rb) > clip, sbb / rb, zero) - lb) / 8.0;
  const T rt = ra + rb;
  const bool live = val(rt) > clip;
  const T frac = f_select(live, (ra * ra + rb * rb) / (rt * rt), cst<T>(1.0));
  const T gamma = 2.0 * (1.0 - frac);
  const T rhos_ts = rt * (ta + tb);
  const T rho_t = ra * ta + rb * tb;
  const T rho_lap = ra * la + rb * lb;
  const T rhom1_3 = f_pow(rt, -1.0 / 3.0);
  const T rho8_3 = f_pow(ra, 8.0 / 3.0) + f_pow(rb, 8.0 / 3.0);
  const T rhom5_3 = f_pow(rt, -5.0 / 3.0);
  const T expf = f_select(val(rt) > 0.0, f_exp(-c * rhom1_3), zero);
  const T par = pow(2.0, 2.0 / 3.0) * CF * rho8_3 - rhos_ts + rho_t / 9.0 + rho_lap / 18.0;
  const T brk = f_select(live, 2.0 * b * rhom5_3 * par * expf, zero);
  return -a * f_select(live, gamma / (1.0 + d * rhom1_3) * (rt + brk), zero);
}

// functional.py:594-623: one spin channel of the u^i w^j expansion, unscaled
template <typename T> GDFT_HD T dm21_term_spin(const T& r, const T& sigma, const T& tau, int i, int j, double clip) {
  const double beta = 1.0 / 1024.0;
  const T log_rho = f_log2(f_clip_min(r, clip));
  const bool live = val(log_rho) > log2(clip);
  T expo = (4.0 / 3.0) * log_rho;
  if (i > 0) {
    const T log_g = f_log2(f_clip_min(sigma, clip)) / 2.0;
    const T log_x = log_g - (4.0 / 3.0) * log_rho;
    const T log_u = f_select(live, log_x - f_log2(1.0 + beta * f_exp2(log_x)) + log2(beta), cst<T>(0.0));
    expo = expo + (double)i * log_u;
  }
  if (j > 0) {
    const T log_tau = f_log2(f_clip_min(tau, clip));
    const T log_1t = -((5.0 / 3.0) * log_rho - log_tau + (2.0 / 3.0) * log2(6.0 * PI * PI) + log2(3.0 / 5.0));
    const T log_w = f_select(live, log_1t - f_log2(1.0 + beta * f_exp2(log_1t)) + log2(beta), cst<T>(0.0));
    expo = expo + (double)j * log_w;
  }
  return f_exp2(expo);
}

// functional.py:1087-1135 (`densities`, the MGGA feature library): one spin channel of rho^{4/3} u^i w^j.  Same u as
// dm21_term_spin; w is built from log2(tau) - 5/3 log2(rho) without the Thomas-Fermi constant (functional.py:1122).
template <typename T> GDFT_HD T mgga_term_spin(const T& r, const T& sigma, const T& tau, int i, int j, double clip) {
  const double beta = 1.0 / 1024.0;
  const T log_rho = f_log2(f_clip_min(r, clip));
  const bool live = val(log_rho) > log2(clip);
  T expo = (4.0 / 3.0) * log_rho;
  if (i > 0) {
    const T log_g = f_log2(f_clip_min(sigma, clip)) / 2.0;
    const T log_x = log_g - (4.0 / 3.0) * log_rho;
    const T log_u = f_select(live, log_x - f_log2(1.0 + beta * f_exp2(log_x)) + log2(beta), cst<T>(0.0));
    expo = expo + (double)i * log_u;
  }
  if (j > 0) {
    const T log_1t = f_log2(f_clip_min(tau, clip)) - (5.0 / 3.0) * log_rho;
    const T log_w = f_select(live, log_1t - f_log2(1.0 + beta * f_exp2(log_1t)) + log2(beta), cst<T>(0.0));
    expo = expo + (double)j * log_w;
  }
  return f_exp2(expo);
}

}  // namespace pw
}  // namespace gdft
