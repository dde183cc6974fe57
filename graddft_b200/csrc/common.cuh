// Shared device/host helpers for libgdft_b200: status codes, mbarrier/TMA PTX wrappers, the FP64
// DMMA wrapper, and host-side tensor-map creation through the driver entry point (no libcuda link).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/gdft_b200.h"

namespace gdft {

extern thread_local int g_last_cuda_error;
extern std::atomic<unsigned long long> g_launches;  // statistics only: number of kernels launched by this library

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return GDFT_CUDA_ERROR;
}
#define GDFT_CUDA_TRY(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::gdft::cuda_fail(_e);     \
  } while (0)
// every kernel launch in the library is followed by this; the counter backs gdft_launch_count()
#define GDFT_LAUNCH_CHECK()                                  \
  do {                                                       \
    ::gdft::g_launches.fetch_add(1, std::memory_order_relaxed); \
    GDFT_CUDA_TRY(cudaGetLastError());                       \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
inline int64_t npad_of(int64_t n) { return round_up(n, 8); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// bump allocator over the caller's workspace (256-byte aligned slices)
struct Workspace {
  char* base;
  size_t size, used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (base == nullptr || used + bytes > size) { used += bytes; return nullptr; }
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

// Builds a 3-D FP64 tiled tensor map {dim0 (fastest), dim1, dim2}; strides in bytes for dim1, dim2.
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box0, uint32_t box1);

// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared tile copy (SASS: UTMALDG), completion signalled on `bar` by transaction bytes
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// generic-proxy writes to smem that a later async-proxy (TMA) op will overwrite/read
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4,row) * B(4x8,col), FP64 tensor op (SASS: DMMA.8x8x4).
// lane = 4*g + t : a = A[g][t], b = B[t][g], c = {C[g][2t], C[g][2t+1]}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace gdft
