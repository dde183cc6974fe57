// Library-level entry points: version/status, tensor-map construction, workspace sizing, basis packing.
#include "common.cuh"

namespace gdft {

thread_local int g_last_cuda_error = 0;
std::atomic<unsigned long long> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { g_last_cuda_error = (int)cudaErrorNotSupported; return GDFT_CUDA_ERROR; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15) || (box0 * 8) % 16 || box0 > 256 ||
      box1 > 256)
    return GDFT_BAD_ALIGNMENT;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; attempt++) {
    r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    // a host thread that has not touched the runtime yet (the autograd engine's worker calling an entry point whose first
    // CUDA call is this driver function): bind the device's primary context to it and try again
    (void)cudaFree(nullptr);
  }
  if (r != CUDA_SUCCESS) { g_last_cuda_error = 100000 + (int)r; return GDFT_CUDA_ERROR; }
  return GDFT_OK;
}

size_t density_fwd_workspace(int64_t n);
size_t density_bwd_workspace(int64_t N, int64_t n, int ncoef_rows, int nout);
size_t eri_workspace(int64_t n);
size_t integrate_workspace(int64_t N);
size_t ln_elu_workspace(int64_t N, int64_t W);
size_t dense_workspace(int64_t N, int64_t K, int64_t Wd);

// ---- basis packing ------------------------------------------------------------------------------
// packed[c][r][b]: c=0 ao; c=1..3 grad_ao[r][b][c-1]; c=4 sum_i grad2_ao[r][b][i]; columns n..npad-1 zero.
__global__ void pack_basis_kernel(int64_t N, int n, int npad, const double* __restrict__ ao, const double* __restrict__ gao,
                                  const double* __restrict__ g2ao, double* __restrict__ packed, int nplanes) {
  const int64_t r = blockIdx.x;
  const size_t plane = (size_t)N * npad;
  for (int b = threadIdx.x; b < npad; b += blockDim.x) {
    const bool in = b < n;
    const size_t src = (size_t)r * n + b, dst = (size_t)r * npad + b;
    packed[dst] = in ? ao[src] : 0.0;
    if (nplanes >= 4) {
      double gx = 0, gy = 0, gz = 0;
      if (in) { gx = gao[src * 3]; gy = gao[src * 3 + 1]; gz = gao[src * 3 + 2]; }
      packed[plane + dst] = gx; packed[2 * plane + dst] = gy; packed[3 * plane + dst] = gz;
    }
    if (nplanes >= 5) packed[4 * plane + dst] = in ? (g2ao[src * 3] + g2ao[src * 3 + 1] + g2ao[src * 3 + 2]) : 0.0;
  }
}

// chi[r][w][s][c] -> chi_packed[w][s][r][c padded]
__global__ void pack_chi_kernel(int64_t N, int n, int npad, int W, const double* __restrict__ chi, double* __restrict__ out) {
  const int64_t r = blockIdx.x;
  const int ws = blockIdx.y;
  const double* src = chi + ((size_t)r * 2 * W + ws) * n;
  double* dst = out + ((size_t)ws * N + r) * npad;
  for (int c = threadIdx.x; c < npad; c += blockDim.x) dst[c] = c < n ? src[c] : 0.0;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_version(void) { return 100; }
extern "C" int gdft_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" unsigned long long gdft_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* gdft_status_string(int s) {
  switch (s) {
    case GDFT_OK: return "ok";
    case GDFT_BAD_SHAPE: return "bad shape";
    case GDFT_BAD_ALIGNMENT: return "bad alignment";
    case GDFT_WORKSPACE_TOO_SMALL: return "workspace too small";
    case GDFT_CUDA_ERROR: return "CUDA error";
    case GDFT_BAD_ARGUMENT: return "bad argument";
    default: return "unknown status";
  }
}
extern "C" int gdft_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
extern "C" int64_t gdft_npad(int64_t n) { return npad_of(n); }
extern "C" size_t gdft_packed_basis_bytes(int64_t N, int64_t n, int nplanes) {
  if (N <= 0 || n <= 0 || nplanes < 1) return 0;
  return (size_t)nplanes * (size_t)N * (size_t)npad_of(n) * 8;
}

extern "C" int gdft_pack_basis(gdft_stream_t stream_, int64_t N, int64_t n, const double* ao, const double* grad_ao,
                               const double* grad2_ao, double* packed, int nplanes) {
  if (N <= 0 || n <= 0 || N > (int64_t)2147483000) return GDFT_BAD_SHAPE;
  if (nplanes != 1 && nplanes != 4 && nplanes != 5) return GDFT_BAD_SHAPE;
  if (!ao || !packed || (nplanes >= 4 && !grad_ao) || (nplanes >= 5 && !grad2_ao)) return GDFT_BAD_ARGUMENT;
  if (!aligned16(packed)) return GDFT_BAD_ALIGNMENT;
  const int npad = (int)npad_of(n);
  pack_basis_kernel<<<(unsigned)N, 128, 0, static_cast<cudaStream_t>(stream_)>>>(N, (int)n, npad, ao, grad_ao, grad2_ao, packed,
                                                                                 nplanes);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_pack_chi(gdft_stream_t stream_, int64_t N, int64_t n, int W, const double* chi, double* chi_packed) {
  if (N <= 0 || n <= 0 || W <= 0 || W > 8 || N > (int64_t)2147483000) return GDFT_BAD_SHAPE;
  if (!chi || !chi_packed) return GDFT_BAD_ARGUMENT;
  if (!aligned16(chi_packed)) return GDFT_BAD_ALIGNMENT;
  const int npad = (int)npad_of(n);
  dim3 grid((unsigned)N, 2 * W);
  pack_chi_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream_)>>>(N, (int)n, npad, W, chi, chi_packed);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" size_t gdft_workspace_bytes(int op, int64_t N, int64_t n, int flags, int W) {
  switch (op) {
    case GDFT_OP_DENSITY_FWD: return density_fwd_workspace(n);
    case GDFT_OP_DENSITY_BWD: return density_bwd_workspace(N, n, 12, 2);
    case GDFT_OP_HF_FOCK: return density_bwd_workspace(N, n, 2 * (W > 0 ? W : 1), 2 * (W > 0 ? W : 1));
    case GDFT_OP_ERI_J: return eri_workspace(n);
    case GDFT_OP_XC_INTEGRATE: return integrate_workspace(N);
    case GDFT_OP_LN_ELU: return ln_elu_workspace(N, n);
    case GDFT_OP_DENSE: return dense_workspace(N, n, flags);
    default: (void)flags; return 0;
  }
}
