// XLA custom-call adapters (legacy "API_VERSION_STATUS_RETURNING"-free ABI: void(cudaStream_t, void** buffers,
// const char* opaque, size_t opaque_len)), so that the kernels can be registered as JAX custom calls with
// jax.ffi.register_ffi_target(..., api_version=0) / xla_client.register_custom_call_target and wrapped in
// jax.custom_vjp (graddft_b200/jax_ffi.py).  No XLA headers are needed for this ABI.  `buffers` lists the operands
// in order, then the results; `opaque` is a packed GdftXlaDims.  Errors cannot be returned through this ABI, so a
// failing call poisons its first result with NaN (visible downstream) and records the status for
// gdft_xla_last_status().  Nothing here contains arithmetic.
//
// NOTE: JAX is not installed in the build/test environment of this repository; every adapter is exercised through
// ctypes with hand-packed buffer lists exactly as XLA's thunk would pass them (tests/test_xla_adapters_gpu.py), driven
// by the same call plans graddft_b200/jax_ffi.py hands to jax.ffi.ffi_call.
#include "common.cuh"

namespace gdft {
struct XlaDims {
  int64_t N, n, F, c_rows;
  int32_t flags, nplanes, W, id;
  double clip;
  uint64_t ws_bytes;
};
thread_local int g_xla_status = 0;

__global__ void poison_kernel(double* p) { *p = __longlong_as_double(0x7ff8000000000000LL); }

static void finish(int rc, cudaStream_t s, void* first_result) {
  g_xla_status = rc;
  if (rc != GDFT_OK && first_result) poison_kernel<<<1, 1, 0, s>>>(static_cast<double*>(first_result));
}
static bool dims_ok(const char* opaque, size_t len, XlaDims* d) {
  if (!opaque || len != sizeof(XlaDims)) { g_xla_status = GDFT_BAD_ARGUMENT; return false; }
  memcpy(d, opaque, sizeof(XlaDims));
  return true;
}
}  // namespace gdft
using namespace gdft;

extern "C" int gdft_xla_last_status(void) { return g_xla_status; }
extern "C" size_t gdft_xla_dims_size(void) { return sizeof(XlaDims); }

// operands: ao, grad_ao, grad2_ao (flags bit 0: grad_ao present, bit 1: grad2_ao present) | results: packed[nplanes, N, npad]
extern "C" void gdft_pack_basis_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_pack_basis(s, d.N, d.n, (const double*)b[0], (d.flags & 1) ? (const double*)b[1] : nullptr,
                           (d.flags & 2) ? (const double*)b[2] : nullptr, (double*)b[3], d.nplanes);
  finish(rc, (cudaStream_t)s, b[3]);
}
// operands: chi[N, W, 2, n] | results: chi_packed[W, 2, N, npad]
extern "C" void gdft_pack_chi_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_pack_chi(s, d.N, d.n, d.W, (const double*)b[0], (double*)b[1]);
  finish(rc, (cudaStream_t)s, b[1]);
}
// operands: packed, rdm1, chi_packed | results: rho, grad_rho, tau, lapl, ehf, ws   (unused ones are 1-element dummies)
extern "C" void gdft_density_fwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  const int f = d.flags;
  int rc = gdft_density_fwd(s, d.N, d.n, f, d.nplanes, (const double*)b[0], (const double*)b[1], (f & GDFT_HF) ? (const double*)b[2] : nullptr,
                            d.W, (f & GDFT_RHO) ? (double*)b[3] : nullptr, (f & GDFT_GRAD) ? (double*)b[4] : nullptr,
                            (f & GDFT_TAU) ? (double*)b[5] : nullptr, (f & GDFT_LAPL) ? (double*)b[6] : nullptr,
                            (f & GDFT_HF) ? (double*)b[7] : nullptr, b[8], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[3]);
}
// operands: packed, rho_bar, grad_rho_bar, tau_bar, lapl_bar | results: rdm1_bar, ws
extern "C" void gdft_density_bwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  const int f = d.flags;
  int rc = gdft_density_bwd(s, d.N, d.n, f, d.nplanes, (const double*)b[0], (f & GDFT_RHO) ? (const double*)b[1] : nullptr,
                            (f & GDFT_GRAD) ? (const double*)b[2] : nullptr, (f & GDFT_TAU) ? (const double*)b[3] : nullptr,
                            (f & GDFT_LAPL) ? (const double*)b[4] : nullptr, (double*)b[5], b[6], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[5]);
}
// operands: packed, chi_packed, g | results: fock, ws
extern "C" void gdft_hf_fock_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_hf_fock(s, d.N, d.n, d.W, d.nplanes, (const double*)b[0], (const double*)b[1], (const double*)b[2], (double*)b[3], b[4], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[3]);
}
// operands: eri, P | results: J, EJ
extern "C" void gdft_eri_j_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_jk(s, d.n, (const double*)b[0], (const double*)b[1], (double*)b[2], nullptr, (double*)b[3], nullptr, 0);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: eri, P | results: J, K, ws   (both contractions from one pass over the tensor)
extern "C" void gdft_eri_jk_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_jk(s, d.n, (const double*)b[0], (const double*)b[1], (double*)b[2], (double*)b[3], nullptr, b[4], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: eri, Kbar | results: Pbar, ws
extern "C" void gdft_eri_k_transpose_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_k_transpose(s, d.n, (const double*)b[0], (const double*)b[1], (double*)b[2], b[3], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: eri, Jbar | results: Pbar, ws
extern "C" void gdft_eri_j_transpose_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_j_transpose(s, d.n, (const double*)b[0], (const double*)b[1], (double*)b[2], b[3], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: c, d, w | results: E, ws
extern "C" void gdft_xc_integrate_fwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_xc_integrate_fwd(s, d.N, (int)d.F, d.c_rows, (const double*)b[0], (const double*)b[1], (const double*)b[2], d.clip, (double*)b[3],
                                 b[4], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[3]);
}
// operands: c, d, w, E_bar | results: c_bar, d_bar, ws
extern "C" void gdft_xc_integrate_bwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_xc_integrate_bwd(s, d.N, (int)d.F, d.c_rows, (const double*)b[0], (const double*)b[1], (const double*)b[2], d.clip,
                                 (const double*)b[3], (double*)b[4], (double*)b[5], b[6], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[5]);
}
// operands: rho, grad_rho, tau, lapl | results: out
extern "C" void gdft_pointwise_fwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_pointwise_fwd(s, d.N, d.id, d.clip, (const double*)b[0], (d.flags & 1) ? (const double*)b[1] : nullptr,
                              (d.flags & 4) ? (const double*)b[2] : nullptr, (d.flags & 2) ? (const double*)b[3] : nullptr, (double*)b[4]);
  finish(rc, (cudaStream_t)s, b[4]);
}
// operands: rho, grad_rho, tau, lapl, out_bar | results: rho_bar, grad_rho_bar, tau_bar, lapl_bar
extern "C" void gdft_pointwise_bwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_pointwise_bwd(s, d.N, d.id, d.clip, (const double*)b[0], (d.flags & 1) ? (const double*)b[1] : nullptr,
                              (d.flags & 4) ? (const double*)b[2] : nullptr, (d.flags & 2) ? (const double*)b[3] : nullptr, (const double*)b[4],
                              (double*)b[5], (d.flags & 1) ? (double*)b[6] : nullptr, (d.flags & 4) ? (double*)b[7] : nullptr,
                              (d.flags & 2) ? (double*)b[8] : nullptr);
  finish(rc, (cudaStream_t)s, b[5]);
}
// operands: rho, grad_rho, tau, lapl, out_bar, u_rho, u_grad_rho, u_tau, u_lapl | results: out_bar_bar, rho_t, grad_rho_t, tau_t, lapl_t
// (flags: bit 0 grad_rho, bit 1 lapl, bit 2 tau are present -- operands and results of absent quantities are 1-element dummies)
extern "C" void gdft_pointwise_bwd2_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  const bool g = d.flags & 1, l = d.flags & 2, t = d.flags & 4;
  int rc = gdft_pointwise_bwd2(s, d.N, d.id, d.clip, (const double*)b[0], g ? (const double*)b[1] : nullptr, t ? (const double*)b[2] : nullptr,
                               l ? (const double*)b[3] : nullptr, (const double*)b[4], (const double*)b[5], g ? (const double*)b[6] : nullptr,
                               t ? (const double*)b[7] : nullptr, l ? (const double*)b[8] : nullptr, (double*)b[9], (double*)b[10],
                               g ? (double*)b[11] : nullptr, t ? (double*)b[12] : nullptr, l ? (double*)b[13] : nullptr);
  finish(rc, (cudaStream_t)s, b[9]);
}
// operands: eri_rows [rows, n, n] (rows passed as N), P | results: J_rows
extern "C" void gdft_eri_j_rows_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_j_rows(s, d.n, d.N, (const double*)b[0], (const double*)b[1], (double*)b[2]);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: eri_rows, Jbar_rows | results: Pbar, ws
extern "C" void gdft_eri_j_transpose_rows_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_eri_j_transpose_rows(s, d.n, d.N, (const double*)b[0], (const double*)b[1], (double*)b[2], b[3], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[2]);
}
// operands: y, res, scale, bias (N rows, width passed as n, eps as clip; flags bit 0: res present) | results: out, stats
extern "C" void gdft_ln_elu_fwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_ln_elu_fwd(s, d.N, d.n, (const double*)b[0], (d.flags & 1) ? (const double*)b[1] : nullptr, (const double*)b[2],
                           (const double*)b[3], d.clip, (double*)b[4], (double*)b[5]);
  finish(rc, (cudaStream_t)s, b[4]);
}
// operands: y, res, scale, bias, stats, out_bar | results: z_bar, scale_bar, bias_bar, ws
extern "C" void gdft_ln_elu_bwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_ln_elu_bwd(s, d.N, d.n, (const double*)b[0], (d.flags & 1) ? (const double*)b[1] : nullptr, (const double*)b[2],
                           (const double*)b[3], (const double*)b[4], (const double*)b[5], (double*)b[6], (double*)b[7], (double*)b[8], b[9],
                           d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[6]);
}
// operands: y, ybias, res, scale, bias (flags bit 0: res present, bit 1: ybias present) | results: out, stats
extern "C" void gdft_dense_ln_elu_fwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_dense_ln_elu_fwd(s, d.N, d.n, (const double*)b[0], (d.flags & 2) ? (const double*)b[1] : nullptr,
                                 (d.flags & 1) ? (const double*)b[2] : nullptr, (const double*)b[3], (const double*)b[4], d.clip,
                                 (double*)b[5], (double*)b[6]);
  finish(rc, (cudaStream_t)s, b[5]);
}
// operands: y, ybias, res, scale, bias, stats, fwd_out, out_bar | results: z_bar, scale_bar, bias_bar, ybias_bar, ws
extern "C" void gdft_dense_ln_elu_bwd_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_dense_ln_elu_bwd(s, d.N, d.n, (const double*)b[0], (d.flags & 2) ? (const double*)b[1] : nullptr,
                                 (d.flags & 1) ? (const double*)b[2] : nullptr, (const double*)b[3], (const double*)b[4],
                                 (const double*)b[5], (d.flags & 4) ? (const double*)b[6] : nullptr, (const double*)b[7], (double*)b[8],
                                 (double*)b[9], (double*)b[10], (d.flags & 2) ? (double*)b[11] : nullptr, b[12], d.ws_bytes);
  finish(rc, (cudaStream_t)s, b[8]);
}
// operands: A[batch = N, n, n] | results: evals[N, n], evecs[N, n, n]
extern "C" void gdft_sym_eigh_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_sym_eigh(s, d.N, d.n, (const double*)b[0], (double*)b[1], (double*)b[2]);
  finish(rc, (cudaStream_t)s, b[1]);
}
// operands: ao[N, n] (rows of the chunk), rdm1[2, n, n], nu[N, n, n] | results: chi[N, 2, n]   (one omega, one chunk)
extern "C" void gdft_chi_contract_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_chi_contract(s, d.N, d.n, (const double*)b[0], d.n, (const double*)b[1], (const double*)b[2], (double*)b[3], 2 * d.n);
  finish(rc, (cudaStream_t)s, b[3]);
}
// operands: err_vec[m = W, 2, n, n] | results: gram[2, m, m]
extern "C" void gdft_diis_gram_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_diis_gram(s, d.W, d.n, (const double*)b[0], (double*)b[1]);
  finish(rc, (cudaStream_t)s, b[1]);
}
// operands: x[2, m = W], fock_vec[m, 2, n, n] | results: out[2, n, n]
extern "C" void gdft_diis_combine_xla(gdft_stream_t s, void** b, const char* opaque, size_t len) {
  XlaDims d;
  if (!dims_ok(opaque, len, &d)) return;
  int rc = gdft_diis_combine(s, d.W, d.n, (const double*)b[0], (const double*)b[1], (double*)b[2]);
  finish(rc, (cudaStream_t)s, b[2]);
}
