// Host build of graddft_b200/csrc/pointwise_math.h for CPU unit tests of the closed-form formulas
// (TEST INFRASTRUCTURE: the product only ever runs the device build in pointwise.cu).
#include <stdint.h>
#include "../../graddft_b200/csrc/pointwise_math.h"
using namespace gdft::pw;

// id 0..4 = lsda_x, b88_x, vwn_c, lyp_c, pw92_c; id 100 + 10*i + j = DM21 u^i w^j column (x = tau)
template <typename T> static T eval(int id, const T (&v)[6], double clip) {
  switch (id) {
    case 0: return lsda_x(v[0], v[1], clip);
    case 1: return b88_x(v[0], v[1], v[2], v[3], clip);
    case 2: return vwn_c(v[0], v[1], clip);
    case 3: return lyp_c(v[0], v[1], v[2], v[3], v[4], v[5], clip);
    case 4: return pw92_c(v[0], v[1], clip);
    default: {
      const int i = (id - 100) / 10, j = (id - 100) % 10;
      T term = dm21_term_spin(v[0], v[2], v[4], i, j, clip) + dm21_term_spin(v[1], v[3], v[5], i, j, clip);
      if (i == 0 && j == 0) term = term * (-2.0 * PI * pow(3.0 / (4.0 * PI), 4.0 / 3.0));
      return term;
    }
  }
}
// v[N][6] = (rho_a, rho_b, sigma_aa, sigma_bb, x_a, x_b) -> out[N], dout[N][6]
extern "C" void pw_host_eval(int id, int64_t N, double clip, const double* v, double* out, double* dout) {
  for (int64_t r = 0; r < N; r++) {
    double x[6];
    Dual<6> d[6];
    for (int q = 0; q < 6; q++) { x[q] = v[r * 6 + q]; d[q] = Make<Dual<6>>::variable(x[q], q); }
    out[r] = eval<double>(id, x, clip);
    const Dual<6> fd = eval<Dual<6>>(id, d, clip);
    for (int q = 0; q < 6; q++) dout[r * 6 + q] = fd.d[q];
  }
}
