/*
 * gdft_b200.h -- C-ABI of libgdft_b200.so: the B200 (sm_100a) kernels behind Grad DFT's
 * per-SCF-iteration hot path.
 *
 * The reference (XanaduAI/GradDFT) has no plugin/FFI registry; its extension point is the set of
 * jit-wrapped free functions in grad_dft/molecule.py that `Molecule` methods forward to, plus
 * `Functional.xc_energy` (grad_dft/functional.py:219-253).  Each entry point below replaces the XLA
 * lowering of one of those functions (cited per function).  Bindings (torch.autograd.Function via
 * ctypes; jax.ffi + custom_vjp where JAX exists) live in graddft_b200/{ops,jax_ffi}.py and contain
 * no arithmetic; INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; all arrays are IEEE float64,
 *     row-major, last index fastest, in the reference's own layouts (grad_dft/molecule.py:76-102);
 *   - every call is stream-ordered on `stream` (a cudaStream_t), never synchronises the device,
 *     never allocates: scratch comes from the caller's `ws` (size from gdft_workspace_bytes);
 *   - return value: 0 OK, 1 bad shape, 2 bad alignment, 3 workspace too small, 4 CUDA error
 *     (code via gdft_last_cuda_error(), thread-local), 5 bad argument.  No exceptions; the only
 *     process-wide state is a relaxed atomic launch counter (statistics, gdft_launch_count) and the
 *     lazily resolved libnccl entry points: safe to call from any host thread on any stream/device;
 *   - there is NO CPU implementation in this library.
 */
#ifndef GDFT_B200_H
#define GDFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gdft_stream_t; /* cudaStream_t */

enum gdft_status {
  GDFT_OK = 0,
  GDFT_BAD_SHAPE = 1,
  GDFT_BAD_ALIGNMENT = 2,
  GDFT_WORKSPACE_TOO_SMALL = 3,
  GDFT_CUDA_ERROR = 4,
  GDFT_BAD_ARGUMENT = 5
};

/* which grid quantities a density call produces / receives cotangents for */
enum gdft_density_flags {
  GDFT_RHO = 1,   /* rho[N,2]          grad_dft/molecule.py:388-409  */
  GDFT_GRAD = 2,  /* grad_rho[N,2,3]   grad_dft/molecule.py:414-440  */
  GDFT_TAU = 4,   /* tau[N,2]          grad_dft/molecule.py:479-502  */
  GDFT_LAPL = 8,  /* lapl_rho[N,2]     grad_dft/molecule.py:445-474  */
  GDFT_HF = 16    /* e_HF[W,2,N]       grad_dft/molecule.py:507-541  */
};

enum gdft_op {
  GDFT_OP_DENSITY_FWD = 1,
  GDFT_OP_DENSITY_BWD = 2,
  GDFT_OP_HF_FOCK = 3,
  GDFT_OP_ERI_J = 4,
  GDFT_OP_XC_INTEGRATE = 5,
  GDFT_OP_LN_ELU = 6, /* N = rows, n = width */
  GDFT_OP_DENSE = 7   /* N = rows, n = K, flags = Wd */
};

/* closed-form per-point feature sets (grad_dft/popular_functionals.py, grad_dft/functional.py) */
enum gdft_pointwise_id {
  GDFT_PW_LSDA_X = 0,      /* 1 column   popular_functionals.py:29-50            */
  GDFT_PW_B88_X = 1,       /* 1 column   popular_functionals.py:52-103           */
  GDFT_PW_VWN_C = 2,       /* 1 column   popular_functionals.py:141-195          */
  GDFT_PW_LYP_C = 3,       /* 1 column   popular_functionals.py:197-269          */
  GDFT_PW_PW92_C = 4,      /* 1 column   popular_functionals.py:105-139          */
  GDFT_PW_B3LYP_SET = 5,   /* 4 columns [lsda,b88,vwn,lyp]  popular_functionals.py:306-326 */
  GDFT_PW_B88_SET = 6,     /* 2 columns [lsda,b88]          popular_functionals.py:277-284 */
  GDFT_PW_DM21_INPUTS = 7, /* 7 columns  functional.py:504-531                   */
  GDFT_PW_DM21_LDA = 8,    /* 1 column   functional.py:534-626 (functional_type="LDA") */
  GDFT_PW_DM21_GGA = 9,    /* 2 columns  functional.py:534-626 ("GGA")           */
  GDFT_PW_DM21_MGGA = 10,  /* 4 columns  functional.py:534-626 ("MGGA")          */
  /* `densities` (the MGGA feature library the article's neural functionals train on), functional.py:1048-1202:
   * per-spin rho^{4/3} u^i w^j columns [(i,j) major, spin minor], followed by as many correlation columns that are
   * identically zero upstream (jnp.round(e_PW92, -30) == 0 and the `> clip` test on it). */
  GDFT_PW_FEAT_LDA = 11,   /* 2 + 2 columns */
  GDFT_PW_FEAT_GGA = 12,   /* 4 + 4 columns */
  GDFT_PW_FEAT_MGGA = 13,  /* 8 + 8 columns */
  GDFT_PW_COUNT = 14
};

int gdft_version(void);
int gdft_last_cuda_error(void);
/* number of CUDA kernels this library has launched in this process so far (statistics for bench.py) */
unsigned long long gdft_launch_count(void);
const char* gdft_status_string(int status);
/* 1 if the current device is compute capability 10.x (the only supported target), else 0 */
int gdft_device_supported(void);

/* ---- packed basis -------------------------------------------------------------------------
 * ao / grad_ao / grad_n_ao[2] are constant across SCF iterations and training steps, so they are
 * re-laid-out ONCE into planar, column-padded planes packed[C][N][npad], npad = gdft_npad(n):
 *   plane 0 = ao, 1..3 = d/dx,d/dy,d/dz ao (reference layout grad_ao[N,n,3] has xyz innermost,
 *   grad_dft/interface/pyscf.py:826), plane 4 = sum_i grad_n_ao[2][:,:,i] (only that sum is ever
 *   used: grad_dft/molecule.py:474).  C = 4, or 5 when grad2_ao != NULL.  Padding columns are zero. */
int64_t gdft_npad(int64_t n);
size_t gdft_packed_basis_bytes(int64_t N, int64_t n, int nplanes);
int gdft_pack_basis(gdft_stream_t stream, int64_t N, int64_t n, const double* ao /*[N,n]*/,
                    const double* grad_ao /*[N,n,3] or NULL*/, const double* grad2_ao /*[N,n,3] or NULL*/,
                    double* packed /*[C,N,npad]*/, int nplanes);
/* chi[N,W,2,n] (grad_dft/molecule.py:92) -> chi_packed[W,2,N,npad] */
int gdft_pack_chi(gdft_stream_t stream, int64_t N, int64_t n, int W, const double* chi, double* chi_packed);

size_t gdft_workspace_bytes(int op, int64_t N, int64_t n, int flags, int W);

/* ---- K1: density family forward (linear map L: rdm1 -> grid quantities) --------------------
 * replaces density / grad_density / kinetic_density / lapl_density / HF_energy_density
 * (grad_dft/molecule.py:409,440,502,472-474,537-541).  T_s = ao D_s is formed tile-wise by FP64
 * DMMA from TMA-staged tiles and contracted against the planes in the epilogue; T never reaches HBM.
 * Outputs not selected by `flags` may be NULL. */
int gdft_density_fwd(gdft_stream_t stream, int64_t N, int64_t n, int flags, int nplanes,
                     const double* packed, const double* rdm1 /*[2,n,n]*/,
                     const double* chi_packed /*[W,2,N,npad] or NULL*/, int W,
                     double* rho /*[N,2]*/, double* grad_rho /*[N,2,3]*/, double* tau /*[N,2]*/,
                     double* lapl /*[N,2]*/, double* ehf /*[W,2,N]*/, void* ws, size_t ws_bytes);

/* ---- K2: density family transpose (L^T: grid cotangents -> rdm1 cotangent) ------------------
 * the XLA transpose of the above inside value_and_grad (grad_dft/train.py:86,147):
 *   Dbar_s = ao^T (rb_s*ao + 2 sum_j gb_sj*dj_ao + 2 lb_s*lap_ao) + sum_j dj_ao^T ((tb_s/2 + 2 lb_s)*dj_ao)
 * un-symmetrised.  Cotangents not selected by `flags` may be NULL.  Split-K over the grid with a
 * deterministic two-stage reduction.  Its own VJP is gdft_density_fwd. */
int gdft_density_bwd(gdft_stream_t stream, int64_t N, int64_t n, int flags, int nplanes,
                     const double* packed, const double* rho_bar, const double* grad_rho_bar,
                     const double* tau_bar, const double* lapl_bar, double* rdm1_bar /*[2,n,n]*/,
                     void* ws, size_t ws_bytes);

/* ---- K4: explicit exact-exchange Fock term ---------------------------------------------------
 * F[w,s,a,c] = -1/2 sum_r ao[r,a] g[w,s,r] chi[r,w,s,c]   (grad_dft/molecule.py:606-613, 678-685) */
int gdft_hf_fock(gdft_stream_t stream, int64_t N, int64_t n, int W, int nplanes, const double* packed,
                 const double* chi_packed, const double* g /*[W,2,N]*/, double* fock /*[W,2,n,n]*/,
                 void* ws, size_t ws_bytes);
/* sum over omega of the above, F[s] = sum_w F[w,s] (what grad_dft/functional.py:714-717, 755-758 do with it: vxc_hf.sum(axis=0)),
 * with the sum taken inside the GEMM: W <= 2 omegas cost two GEMM units, not 2 W. */
int gdft_hf_fock_sum(gdft_stream_t stream, int64_t N, int64_t n, int W, int nplanes, const double* packed,
                     const double* chi_packed, const double* g /*[W,2,N]*/, double* fock_sum /*[2,n,n]*/, void* ws, size_t ws_bytes);

/* ---- K3: ERI sweep ---------------------------------------------------------------------------
 * J[p,q] = sum_rt eri[p,q,r,t] P[r,t]  (coulomb_potential, grad_dft/molecule.py:811) and E_J = 1/2 <P,J>
 * (grad_dft/molecule.py:781-783).  With K != NULL the same ONE pass over the tensor also returns
 * K[p,r] = sum_qt eri[p,q,r,t] P[q,t] (the exchange pairing of BASELINE.json's "J/K"; not in the reference, SURVEY.md
 * section 0.3): 8 n^4 bytes for both.  K needs the gdft_workspace_bytes(GDFT_OP_ERI_J, ...) workspace (partial K per chunk
 * of q, summed in fixed order); J-only calls need none.  J of a J+K call and of a J-only call differ in summation order
 * (rounding level), each is run-to-run reproducible.  K and EJ may be NULL. */
int gdft_eri_jk(gdft_stream_t stream, int64_t n, const double* eri /*[n,n,n,n]*/, const double* P /*[n,n]*/,
                double* J /*[n,n]*/, double* K /*[n,n] or NULL*/, double* EJ /*[1] or NULL*/,
                void* ws, size_t ws_bytes);
/* cotangent of the sweep wrt P: Pbar[r,t] = sum_pq Jbar[p,q] eri[p,q,r,t] (exact for any eri) */
int gdft_eri_j_transpose(gdft_stream_t stream, int64_t n, const double* eri, const double* Jbar,
                         double* Pbar, void* ws, size_t ws_bytes);
/* cotangent of K wrt P: Pbar[q,t] = sum_pr Kbar[p,r] eri[p,q,r,t] (exact for any eri; same workspace) */
int gdft_eri_k_transpose(gdft_stream_t stream, int64_t n, const double* eri, const double* Kbar,
                         double* Pbar, void* ws, size_t ws_bytes);

/* Row-sharded sweep (at n = 400 the tensor is 205 GB and must be split over GPUs):
 * `eri_rows` is the contiguous block of `rows` (p,q)-rows [rows, n*n] a rank holds; J_rows[rows] are the matching
 * entries of J (row-major (p,q) order).  Per-row summation order is independent of the blocking, so the gathered J
 * is bitwise equal to gdft_eri_jk's.  The transpose returns this block's partial Pbar (sum over ranks = full). */
int gdft_eri_j_rows(gdft_stream_t stream, int64_t n, int64_t rows, const double* eri_rows, const double* P,
                    double* J_rows);
int gdft_eri_j_transpose_rows(gdft_stream_t stream, int64_t n, int64_t rows, const double* eri_rows,
                              const double* Jbar_rows, double* Pbar, void* ws, size_t ws_bytes);

/* E_nuc + E_1 + E_J = nuclear_repulsion + <P, h1e> + 1/2 <P, J> (grad_dft/molecule.py:697-733, 738-783) in one pass over the
 * n x n matrices; nuclear_repulsion and out are device scalars. */
int gdft_nonxc_energy(gdft_stream_t stream, int64_t n, const double* P /*[n,n]*/, const double* h1e, const double* J,
                      const double* nuclear_repulsion /*[1]*/, double* out /*[1]*/);

/* ---- packed rep_tensor ("shard or pack").  (pq|rt) = (qp|rt) = (pq|tr), so
 * J_pq = sum_{r>=t} (pq|rt) (P_rt + P_tr)(1 - delta_rt/2) needs the npair = n(n+1)/2 pair rows x pair columns only: a
 * quarter of the 8 n^4 bytes of grad_dft/molecule.py:811's sweep, laid out once per molecule as
 * packed[pair(p,q)][pair(r,t)], pair(i,j) = i(i+1)/2 + j (i >= j), row pitch = npair rounded up to even (padding 0).
 * gdft_eri_symmetry_defect writes {max |asymmetry|, max |value|} over the given (p,q) rows (device, 2 doubles) so that the
 * caller packs only tensors that ARE symmetric.  The packed matrix is symmetric, so the transpose of gdft_eri_j_packed
 * (the VJP of J w.r.t. P) is gdft_eri_j_packed itself.  Pair rows shard like (p,q) rows: each rank sweeps
 * [pair0, pair0 + pairs), writes zero elsewhere, and the Fock all-reduce assembles J. */
int64_t gdft_eri_npair(int64_t n);
size_t gdft_eri_packed_bytes(int64_t n, int64_t pairs);
size_t gdft_eri_packed_workspace(int64_t n);
int gdft_eri_symmetry_defect(gdft_stream_t stream, int64_t n, int64_t row0, int64_t rows, const double* eri_rows /*[rows,n,n]*/,
                             double* out2 /*[2]*/, void* ws, size_t ws_bytes);
int gdft_eri_pack(gdft_stream_t stream, int64_t n, int src_is_pair_rows, int64_t src_row0, int64_t src_rows,
                  const double* src /*[src_rows,n,n]: (p,q) rows src_row0.., or exactly the pair rows pair0..*/,
                  int64_t pair0, int64_t pairs, double* packed /*[pairs][npair_pad]*/);
int gdft_eri_j_packed(gdft_stream_t stream, int64_t n, int64_t pair0, int64_t pairs, const double* packed,
                      const double* P /*[n,n]*/, double* J /*[n,n]*/, double* EJ /*[1] or NULL*/, void* ws, size_t ws_bytes);

/* ---- K6: XC quadrature ------------------------------------------------------------------------
 * E = sum_r aclip(w_r) aclip(aclip(sum_f c[r,f] d[r,f]))  (grad_dft/functional.py:251-253,342;
 * aclip = abs_clip, grad_dft/molecule.py:687-689).  c_rows is 1 (constant functionals,
 * grad_dft/popular_functionals.py:347) or N. */
int gdft_xc_integrate_fwd(gdft_stream_t stream, int64_t N, int F, int64_t c_rows, const double* c,
                          const double* d, const double* w, double clip, double* E /*[1]*/, void* ws,
                          size_t ws_bytes);
int gdft_xc_integrate_bwd(gdft_stream_t stream, int64_t N, int F, int64_t c_rows, const double* c,
                          const double* d, const double* w, double clip, const double* E_bar /*[1]*/,
                          double* c_bar /*[c_rows,F] or NULL*/, double* d_bar /*[N,F] or NULL*/,
                          void* ws, size_t ws_bytes);

/* ---- K5: closed-form per-point features --------------------------------------------------------
 * out[N,F] for one of gdft_pointwise_id; inputs not used by the id may be NULL.  bwd returns the
 * cotangents of the inputs for cotangent out_bar[N,F] (forward-mode dual numbers inside the kernel,
 * with jnp.clip / jnp.where sub-gradient conventions). */
int gdft_pointwise_ncols(int id);
int gdft_pointwise_fwd(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho,
                       const double* grad_rho, const double* tau, const double* lapl, double* out);
int gdft_pointwise_bwd(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho,
                       const double* grad_rho, const double* tau, const double* lapl,
                       const double* out_bar, double* rho_bar, double* grad_rho_bar, double* tau_bar,
                       double* lapl_bar);

/* VJP of gdft_pointwise_bwd (second order: differentiating V_xc once more, i.e. training through the SCF loop,
 * grad_dft/evaluate.py:917-1038 under jax.grad).  u_* are the cotangents of gdft_pointwise_bwd's outputs (NULL = zero);
 * out_bar_bar[N,F] is the cotangent of out_bar, *_t those of the inputs.  Outputs may be NULL. */
int gdft_pointwise_bwd2(gdft_stream_t stream, int64_t N, int id, double clip, const double* rho,
                        const double* grad_rho, const double* tau, const double* lapl, const double* out_bar,
                        const double* u_rho, const double* u_grad_rho, const double* u_tau, const double* u_lapl,
                        double* out_bar_bar, double* rho_t, double* grad_rho_t, double* tau_t, double* lapl_t);

/* First-order XC build of a closed-form functional with a constant coefficient row in ONE pass per grid point: the features
 * of `id` (evaluated once, on dual numbers), the optional exact-exchange column h = sum_{w,s} ehf[w,s,r]
 * (popular_functionals.py:330-338), abs_clip of the densities (functional.py:160-185), e = sum_f coef_f d_f, abs_clip of e and of
 * the weights, E = sum_r w_r e_r (functional.py:219-253, 316-342), and the cotangents of the grid quantities and of ehf seeded by
 * dE/dd_f = w_r [|e_r| > clip] coef_f [|d_f| > clip].  coef holds the F feature coefficients and, when W > 0, the coefficient of
 * the exact-exchange column.  Same per-point arithmetic as gdft_pointwise_fwd/_bwd + gdft_xc_integrate_fwd/_bwd. */
size_t gdft_xc_point_workspace(int64_t N);
int gdft_xc_point_fused(gdft_stream_t stream, int64_t N, int id, double clip, const double* coef_host, int ncoef, const double* rho,
                        const double* grad_rho, const double* tau, const double* lapl, const double* ehf /*[W,2,N] or NULL*/, int W,
                        const double* weights, double* E /*[1]*/, double* rho_bar, double* grad_rho_bar, double* tau_bar,
                        double* lapl_bar, double* ehf_bar /*[W,2,N] or NULL*/, void* ws, size_t ws_bytes);

/* ---- coefficient-network residual block ----------------------------
 * out = elu(LayerNorm(y + res) * scale + bias) over the last axis of [N, W] (W even, <= 512), the loop body of DM21's
 * default_nn after its Dense layer (grad_dft/functional.py:809-819; flax LayerNorm: biased variance, eps inside the
 * square root).  res may be NULL.  stats[N,2] receives (mean, 1/sqrt(var+eps)) per row for the reverse pass.
 * bwd: z_bar[N,W] is the cotangent of z = y + res (hence of both y and res); scale_bar / bias_bar [W] may be NULL. */
int gdft_ln_elu_fwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* res,
                    const double* scale, const double* bias, double eps, double* out, double* stats);
int gdft_ln_elu_bwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* res,
                    const double* scale, const double* bias, const double* stats, const double* out_bar,
                    double* z_bar, double* scale_bar, double* bias_bar, void* ws, size_t ws_bytes);

/* The same block with the Dense bias folded in: z = y + ybias + res where y = x K is the bare library GEMM and ybias[W]
 * the Dense bias (flax Dense = x K + k, grad_dft/functional.py:811); the reverse pass also returns ybias_bar[W] = column
 * sums of z_bar, so neither the broadcast add nor its reduction is a pass of its own.  ybias / ybias_bar may be NULL. */
int gdft_dense_ln_elu_fwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* ybias,
                          const double* res, const double* scale, const double* bias, double eps, double* out,
                          double* stats);
int gdft_dense_ln_elu_bwd(gdft_stream_t stream, int64_t N, int64_t W, const double* y, const double* ybias,
                          const double* res, const double* scale, const double* bias, const double* stats,
                          const double* fwd_out /*the forward output, or NULL: elu' is then recomputed with exp*/,
                          const double* out_bar, double* z_bar, double* scale_bar, double* bias_bar,
                          double* ybias_bar, void* ws, size_t ws_bytes);

/* ---- coefficient-network Dense layers as FP64 tensor-core GEMMs (csrc/dense_gemm.cu; row f2) -------------------------
 * flax Dense (x K + k) of DM21's default_nn and its residual blocks, grad_dft/functional.py:793-822, with the elementwise
 * tail of the block fused into the GEMM epilogue.  Shapes: K even, Wd a multiple of 8 and <= 256 (gdft_dense_supported).
 * `kernel_t` is the TRANSPOSED Dense kernel [Wd, K] (so that both operands stream with the same tile layout).
 *   gdft_dense_fwd        out[N,Wd] = x[N,K] kernel (+ bias[Wd]) (+ res[N,Wd]).  Also the input cotangent of a Dense layer:
 *                         x_bar = y_bar kernel^T is gdft_dense_fwd(y_bar, kernel_t := kernel as stored, ...).
 *   gdft_dense_block_fwd  one residual block, out = elu(LayerNorm(x kernel + dense_bias + x) * scale + bias) over rows of width
 *                         W (flax LayerNorm: biased variance, eps inside the root); also writes xhat[N,W] (the normalised
 *                         rows) and rstd[N] for the reverse pass.
 *   gdft_dense_block_bwd  reverse pass across the boundary between two consecutive blocks: from z_bar[N,W] (cotangent of
 *                         z = x kernel + dense_bias + x of THIS block, whose kernel [W,W] is passed as stored) it forms the
 *                         cotangent of this block's input, x_bar = z_bar kernel^T + z_bar -- the PREVIOUS block's output
 *                         cotangent -- and undoes the previous block's ELU and LayerNorm in the epilogue: prev_z_bar[N,W],
 *                         and the previous block's parameter cotangents prev_scale_bar / prev_bias_bar (LayerNorm) and
 *                         prev_dense_bias_bar (= column sums of prev_z_bar), each [W] or NULL.  ws >= ceil(N/32)*3*W*8 bytes.
 *   gdft_dense_bwd_weight kernel_bar[K,Wd] = x[N,K]^T z_bar[N,Wd] (split-K over the rows, deterministic); K a multiple of 8;
 *                         ws >= 148*K*Wd*8 bytes.                                                                        */
int gdft_dense_supported(int64_t K, int64_t Wd);
int gdft_dense_fwd(gdft_stream_t stream, int64_t N, int64_t K, int64_t Wd, const double* x, const double* kernel_t,
                   const double* bias, const double* res, double* out);
int gdft_dense_block_fwd(gdft_stream_t stream, int64_t N, int64_t W, const double* x, const double* kernel_t,
                         const double* dense_bias, const double* scale, const double* bias, double eps, double* out,
                         double* xhat, double* rstd);
int gdft_dense_block_bwd(gdft_stream_t stream, int64_t N, int64_t W, const double* z_bar, const double* kernel,
                         const double* prev_out, const double* prev_xhat, const double* prev_rstd, const double* prev_scale,
                         double* prev_z_bar, double* prev_scale_bar, double* prev_bias_bar, double* prev_dense_bias_bar,
                         void* ws, size_t ws_bytes);
int gdft_dense_bwd_weight(gdft_stream_t stream, int64_t N, int64_t K, int64_t Wd, const double* x, const double* z_bar,
                          double* kernel_bar, void* ws, size_t ws_bytes);
/* The last block of a trunk (no GEMM follows whose epilogue could undo its ELU / LayerNorm): z_bar[N,W] and the block's
 * parameter cotangents from out_bar, in one streaming pass over (out_bar, out, xhat).  ws >= 592*3*W*8 bytes. */
int gdft_dense_block_bwd_last(gdft_stream_t stream, int64_t N, int64_t W, const double* out_bar, const double* out,
                              const double* xhat, const double* rstd, const double* scale, double* z_bar, double* scale_bar,
                              double* bias_bar, double* dense_bias_bar, void* ws, size_t ws_bytes);

/* ---- SCF harness: small symmetric eigenproblem -------------------------
 * evals[b, n] ascending and evecs[b, n, n] (columns) of the symmetric matrices A[b, n, n], n <= gdft_sym_eigh_max_n():
 * what jnp.linalg.eigh returns inside safe_eigh (grad_dft/utils/eigenproblem.py:26-106), as one CTA per matrix of
 * parallel-order cyclic Jacobi in shared memory.  Stream-ordered with no status word, hence capturable in a CUDA graph
 * together with the rest of the SCF iteration.  Larger matrices stay with the host framework's cuSOLVER path. */
int gdft_sym_eigh_max_n(void);
int gdft_sym_eigh(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, double* evals, double* evecs);
/* Warm start: V0[b, n, n] orthogonal (the eigenvectors of the previous SCF cycle; NULL = cold).  The sweeps run on
 * V0^T A V0, which is nearly diagonal once the SCF settles, and the eigenvectors returned are V0 V'.  Used for n <= 64;
 * larger n ignore V0.  The result is the eigen-decomposition of A either way; only the sweep count changes. */
int gdft_sym_eigh_warm(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, const double* V0, double* evals,
                       double* evecs);
/* The general entry: n <= 64 one CTA per matrix (two-sided Jacobi in shared memory, above); 64 < n <= gdft_sym_eigh_max_n()
 * (320: the benzene/def2-TZVP class) one 8-CTA thread-block cluster per matrix running one-sided (Hestenes) Jacobi on
 * (A + sigma I) V0 with the column pairs handed round the cluster through distributed shared memory (csrc/eigh_cluster.cu);
 * V0 warm-starts both.  A is taken as symmetric (its rows are read as columns).  `info[b]` (optional, DEVICE int per
 * matrix, written stream-ordered: poll it lazily or assert on it in tests) receives the number of sweeps, -1 if the sweep
 * bound was hit before convergence, -2 for a non-finite result.  No host synchronisation: capturable in a CUDA graph. */
int gdft_sym_eigh_ex(gdft_stream_t stream, int64_t batch, int64_t n, const double* A, const double* V0, double* evals,
                     double* evecs, int* info);

/* The two reductions of the CDIIS step (grad_dft/evaluate.py:1165 "iskl,jskl->sij" and :1198 "si,isjk->sjk") over the ring
 * buffers err_vec / fock_vec [m,2,n,n]: gram[2,m,m] (symmetric, deterministic summation order) and the extrapolated
 * out[2,n,n] = sum_i x[s,i] fock_vec[i,s].  As degenerate GEMMs they cost 21 us each inside the H2O-shaped iteration. */
int gdft_diis_gram(gdft_stream_t stream, int m, int64_t n, const double* err_vec, double* gram);
/* The bordered CDIIS matrix of evaluate.py:1167-1181 straight from the ring buffer: B[s,0,0] = 0, B[s,0,1+i] = B[s,1+i,0] =
 * (i <= cycle), B[s,1+i,1+j] = gram[s,i,j], with 1 on the diagonal of the slots that are not live yet. */
int gdft_diis_matrix(gdft_stream_t stream, int m, int64_t n, int cycle, const double* err_vec, double* B /*[2,m+1,m+1]*/);
int gdft_diis_combine(gdft_stream_t stream, int m, int64_t n, const double* x /*[2,m]*/, const double* fock_vec,
                      double* out /*[2,n,n]*/);

/* The n x n tail of one DIIS SCF iteration of a small molecule (n <= gdft_scf_stage_max_n() = 64, max_diis <= 16) as two
 * kernels around the eigensolver, one CTA per spin, no host-visible status (graph-capturable):
 * gdft_scf_diis_step: err = FDS - (FDS)^T, ring-buffer slot written in place (logical entry i lives in physical slot
 *   (head + i) % m, head = max(0, cycle - m) % m: the slot rotates where upstream shifts the whole buffer; cycle == m is
 *   dropped like upstream's out-of-bounds .at[].set()), the one new row/column of the persistent Gram matrix, the bordered
 *   CDIIS matrix, x = B^-1 e_0 (LU with partial pivoting), F' = sum_i x_i F_i, and the reduced matrix C = L^-1 F' L^-T
 *   (grad_dft/evaluate.py:1111-1205, grad_dft/utils/eigenproblem.py:125-127).
 * gdft_scf_occupy: mo_coeff = L^-T V, aufbau occupations by stable rank with nelec = round(sum occ_prev), rdm1 = C occ C^T
 *   (eigenproblem.py:129, grad_dft/molecule.py:815-889). */
/* Aufbau occupations of grad_dft/molecule.py:851-889 for any n (stable rank by counting, no sort): occ[2,n] from evals[2,n]
 * and nelec_s = round(sum of occ_prev[s,:]). */
int gdft_aufbau_occupations(gdft_stream_t stream, int64_t n, const double* evals /*[2,n]*/, const double* occ_prev /*[2,n]*/,
                            double* occ /*[2,n]*/);
int gdft_scf_stage_max_n(void);
int gdft_scf_diis_step(gdft_stream_t stream, int64_t n, int m, int cycle, const double* fock /*[2,n,n]*/, const double* rdm1 /*[2,n,n]*/,
                       const double* overlap /*[n,n]*/, const double* L_inv /*[n,n]*/, double* fock_vec /*[m,2,n,n]*/,
                       double* err_vec /*[m,2,n,n]*/, double* gram /*[2,m,m]*/, double* x_out /*[2,m]*/, double* fock_out /*[2,n,n]*/,
                       double* C_out /*[2,n,n]*/);
int gdft_scf_occupy(gdft_stream_t stream, int64_t n, const double* evals /*[2,n]*/, const double* V /*[2,n,n]*/, const double* L_inv,
                    const double* occ_prev /*[2,n]*/, double* mo_coeff /*[2,n,n]*/, double* mo_occ /*[2,n]*/, double* rdm1 /*[2,n,n]*/);

/* abs_clip (grad_dft/molecule.py:687-689) and its VJP in one elementwise pass: out[i] = |src[i]| > thr ? x[i] : 0
 * (x = src: the clip; x = a cotangent: the cotangent of the clip's input).  NaN in src gives 0, as jnp.where does. */
int gdft_abs_clip(gdft_stream_t stream, int64_t count, const double* x, const double* src, double thr, double* out);

/* ---- chi generation tail ----------------------------------------------
 * chi[r, s, a] = sum_{b,d} rdm1[s,b,d] ao[r,b] nu[r,d,a] for the Nc grid points of one nu chunk and one
 * range-separation parameter: the "...bd,b,da->...a" einsum that generate_chi_tensor vmaps over a chunk
 * (grad_dft/interface/pyscf.py:1110-1119); nu[Nc,n,n] are the per-point screened-Coulomb integrals that
 * _nu_chunk yields (grad_dft/external/_hf_density.py:69-103; libcint, out of path).  Row r of the chunk reads
 * ao + r*ao_ld (n values) and writes chi + r*chi_ld ([2, n] contiguous), so the caller can point `chi` at
 * chi_full[start, w, 0, 0] with chi_ld = W*2*n (the reference layout chi[N,W,2,n], grad_dft/molecule.py:92) and
 * no concatenate/stack pass is needed.  One HBM pass over nu (8 n^2 bytes per point).  n <= gdft_chi_contract_max_n(). */
int64_t gdft_chi_contract_max_n(void);
int gdft_chi_contract(gdft_stream_t stream, int64_t Nc, int64_t n, const double* ao, int64_t ao_ld,
                      const double* rdm1 /*[2,n,n]*/, const double* nu /*[Nc,n,n]*/, double* chi, int64_t chi_ld);

/* ---- predictor glue ----------------------------------------------------------------------------
 * fock = aclip(1/2 (X + X^T)), X = aclip(h1e + J + Dbar)   (grad_dft/train.py:148-163) */
int gdft_fock_assemble(gdft_stream_t stream, int64_t n, const double* h1e, const double* J,
                       const double* rdm1_bar /*[2,n,n]*/, double clip, double* fock /*[2,n,n]*/);
/* fock = aclip(fock + V + V^T)                              (grad_dft/train.py:205-206, 212-213) */
int gdft_fock_add_sym(gdft_stream_t stream, int64_t n, const double* V /*[2,n,n]*/, double clip,
                      double* fock /*[2,n,n]*/);

/* ---- the exchange step of the grid-sharded path (not in the reference, which is single-device) ----
 * One sum over ranks of the packed [V_xc | J | V_HF ... | E_xc] buffer closes a sharded Fock build
 * (the partial sums of grad_dft/train.py:147-213 computed on each rank's grid rows).
 *
 * gdft_allreduce_fock: in-place ncclAllReduce(sum, float64) on the caller's communicator and stream.  libnccl is resolved
 * with dlopen at first use; gdft_nccl_* create a communicator from a unique id the host side broadcasts. */
int gdft_nccl_available(void);
size_t gdft_nccl_unique_id_bytes(void);
int gdft_nccl_unique_id(void* id_host /*[gdft_nccl_unique_id_bytes()]*/);
int gdft_nccl_comm_create(const void* id_host, int rank, int world, void** nccl_comm_out);
int gdft_nccl_comm_destroy(void* nccl_comm);
int gdft_allreduce_fock(void* nccl_comm /*ncclComm_t*/, gdft_stream_t stream, double* packed /*[count], in place*/, size_t count);
/* gdft_allreduce_fock_p2p: the library's own exchange kernel over NVLink peer memory.  A communicator owns a payload of
 * `capacity` doubles in IPC-shareable device memory (gdft_comm_buffer: ordinary device memory, kernels may write their
 * results straight into it); one launch per rank announces, reduces slice `rank` over all peers in rank order
 * (reduce-scatter by peer loads), stores the sums into every peer's payload (all-gather by peer stores) and completes a
 * handshake.  Bitwise identical on all ranks and run to run; no host involvement (CUDA-graph capturable).  gdft_comm_create /
 * _connect / _destroy are SETUP calls (they allocate and synchronise); handles are exchanged by the host side
 * (one process per GPU: gdft_comm_handle + gdft_comm_connect; one process driving several GPUs: gdft_comm_connect_local).
 * Calls on one communicator must be issued in the same order on every rank and never concurrently from two streams. */
typedef struct gdft_comm gdft_comm;
size_t gdft_comm_handle_bytes(void);
int gdft_comm_create(int rank, int world, size_t capacity, gdft_comm** out);
double* gdft_comm_buffer(gdft_comm* comm);
size_t gdft_comm_capacity(gdft_comm* comm);
int gdft_comm_handle(gdft_comm* comm, void* handle_host /*[gdft_comm_handle_bytes()]*/);
int gdft_comm_connect(gdft_comm* comm, const void* handles_host /*[world][gdft_comm_handle_bytes()], rank order*/);
int gdft_comm_connect_local(gdft_comm* comm, gdft_comm* const* all /*[world], rank order*/);
int gdft_comm_status(gdft_comm* comm, int* status_host, unsigned long long* epoch_host); /* host-synchronous */
int gdft_comm_destroy(gdft_comm* comm);
int gdft_allreduce_fock_p2p(gdft_stream_t stream, gdft_comm* comm, size_t count);

/* ---- XLA custom-call adapters (graddft_b200/csrc/jax_ffi.cu) --------------------------------------
 * Legacy custom-call ABI void(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len): operands
 * then results in `buffers`, a packed dims struct (gdft_xla_dims_size() bytes; layout in jax_ffi.py) in `opaque`.
 * These let jax.ffi / xla_client register the kernels as JAX custom calls (north-star binding); they forward to
 * the entry points above and contain no arithmetic. */
int gdft_xla_last_status(void);
size_t gdft_xla_dims_size(void);
void gdft_pack_basis_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_pack_chi_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_density_fwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_density_bwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_hf_fock_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_j_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_j_transpose_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_jk_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_k_transpose_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_xc_integrate_fwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_xc_integrate_bwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_pointwise_fwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_pointwise_bwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_pointwise_bwd2_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_j_rows_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_eri_j_transpose_rows_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_ln_elu_fwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_ln_elu_bwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_dense_ln_elu_fwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_dense_ln_elu_bwd_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_sym_eigh_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_chi_contract_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_diis_gram_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);
void gdft_diis_combine_xla(gdft_stream_t stream, void** buffers, const char* opaque, size_t opaque_len);

#ifdef __cplusplus
}
#endif
#endif /* GDFT_B200_H */
