// The one exchange step of the grid-sharded path: the sum over ranks of the packed
// [V_xc | J | V_HF... | E_xc] buffer that closes a Fock build.  Two implementations behind the C-ABI:
//
//   gdft_allreduce_fock        -- SURVEY.md 8b's signature: in-place ncclAllReduce(sum, f64) on the caller's communicator
//                                 and stream.  libnccl is resolved with dlopen at first use (the library has no link-time
//                                 dependency on it and still loads on a box without NCCL).
//   gdft_allreduce_fock_p2p    -- our own kernel over NVLink peer memory: every rank's payload lives in a cudaMalloc'ed,
//                                 IPC-shared buffer (gdft_comm_*); ONE launch per rank does announce -> reduce-scatter (rank r
//                                 sums slice r of every peer's buffer in rank order with 128-bit peer loads) -> all-gather
//                                 (writes the sums straight into every peer's buffer) -> completion handshake.  No host
//                                 involvement, no NCCL proxy thread, device-resident epoch counter => the call can be
//                                 captured in a CUDA graph with the rest of the SCF iteration.  Every element is summed by
//                                 exactly one rank in a fixed order, so the result is bitwise identical on all ranks and
//                                 from run to run.  The density VJP's second-stage reduce can write straight into the
//                                 payload (it is ordinary device memory), so no pack/cat copies precede the exchange.
//
// Flags are 64-bit epochs written with st.release.sys and polled with ld.acquire.sys; a poll that exceeds
// GDFT_COMM_TIMEOUT_NS gives up, records status 1 in the communicator and lets the kernel finish (results undefined,
// gdft_comm_status() reports it) instead of hanging the GPU.
#include <dlfcn.h>
#include <string.h>
#include <new>
#include "common.cuh"

#define GDFT_COMM_MAX_RANKS 16
#define GDFT_COMM_TIMEOUT_NS 20000000000ull
#define GDFT_COMM_HEADER_BYTES 1024

namespace gdft {

struct CommHeader {  // at the head of every rank's allocation; ready/done are written by the peers
  unsigned long long ready[GDFT_COMM_MAX_RANKS];
  unsigned long long done[GDFT_COMM_MAX_RANKS];
  unsigned long long epoch;  // local: number of completed exchanges
  unsigned int counter;      // local: CTAs of the running launch that have finished their slice
  int status;                // local: 0 ok, 1 a poll timed out
};
static_assert(sizeof(CommHeader) <= GDFT_COMM_HEADER_BYTES, "header");

struct PeerTable {
  char* base[GDFT_COMM_MAX_RANKS];  // allocation base of every rank (header, then payload)
};

}  // namespace gdft

struct gdft_comm {
  int rank, world, device;
  size_t capacity;  // doubles
  char* local;      // cudaMalloc'ed: header + payload
  gdft::PeerTable peers;
  bool opened[GDFT_COMM_MAX_RANKS];
  bool connected;
};

namespace gdft {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// peer payload written by earlier kernels of the peer: bypass L1 (never cached here before, but keep it explicit)
__device__ __forceinline__ double2 ld_peer(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(double2* p, double2 v) {
  asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ bool poll_ge(const unsigned long long* flag, unsigned long long want, int* status) {
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flag) < want) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > GDFT_COMM_TIMEOUT_NS) {
      *status = 1;
      return false;
    }
  }
  return true;
}

template <int WORLD>
__global__ void __launch_bounds__(256) allreduce_fock_kernel(PeerTable peers, int rank, int world_rt, long long count2) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  CommHeader* me = reinterpret_cast<CommHeader*>(peers.base[rank]);
  __shared__ int s_last;
  const unsigned long long ep = *reinterpret_cast<volatile unsigned long long*>(&me->epoch) + 1;

  // 1. announce: my payload is complete (stream order) -- one flag per peer, written into the peer's header
  if (blockIdx.x == 0 && threadIdx.x < world)
    st_release_sys(&reinterpret_cast<CommHeader*>(peers.base[threadIdx.x])->ready[rank], ep);
  // 2. every CTA waits until every peer has announced (local polls)
  if (threadIdx.x < world) poll_ge(&me->ready[threadIdx.x], ep, &me->status);
  __syncthreads();

  // 3. reduce slice `rank` over the peers in rank order, write the sums into every peer's buffer
  const long long per = (count2 + world - 1) / world;
  const long long lo = per * rank, hi = lo + per < count2 ? lo + per : count2;
  double2* pay[GDFT_COMM_MAX_RANKS];
#pragma unroll
  for (int p = 0; p < GDFT_COMM_MAX_RANKS; ++p)
    if (p < world) pay[p] = reinterpret_cast<double2*>(peers.base[p] + GDFT_COMM_HEADER_BYTES);
  // U elements per thread in flight (U x world independent peer loads before the first add: the loop is NVLink-latency-bound)
  constexpr int U = (WORLD > 0 && WORLD <= 4) ? 4 : 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += U * stride) {
    double2 v[U][GDFT_COMM_MAX_RANKS];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int p = 0; p < GDFT_COMM_MAX_RANKS; ++p)
        if (p < world && i + u * stride < hi) v[u][p] = ld_peer(pay[p] + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= hi) break;
      double2 acc = v[u][0];
#pragma unroll
      for (int p = 1; p < GDFT_COMM_MAX_RANKS; ++p)
        if (p < world) { acc.x += v[u][p].x; acc.y += v[u][p].y; }
#pragma unroll
      for (int p = 0; p < GDFT_COMM_MAX_RANKS; ++p)
        if (p < world) st_peer(pay[p] + i + u * stride, acc);
    }
  }

  // 4. completion: the last CTA of this launch tells every peer that this rank's reads and writes are over and waits for
  //    the same from them; only then may the local payload be read (all slices have landed) or overwritten (nobody reads it)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&me->counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < world) {
    st_release_sys(&reinterpret_cast<CommHeader*>(peers.base[threadIdx.x])->done[rank], ep);
    poll_ge(&me->done[threadIdx.x], ep, &me->status);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    me->counter = 0;
    *reinterpret_cast<volatile unsigned long long*>(&me->epoch) = ep;
    __threadfence();
  }
}

// ---- NCCL through dlopen ----------------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
typedef int (*nccl_get_unique_id_t)(NcclUniqueId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclUniqueId, int);
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
struct NcclApi {
  void* handle = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  bool ok = false;
};
static const NcclApi& nccl_api() {
  static const NcclApi api = [] {  // C++11 magic static: initialised once, thread-safe, immutable afterwards
    NcclApi a;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the host framework has usually mapped it already
    if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW);
    if (!a.handle) return a;
    a.get_unique_id = (nccl_get_unique_id_t)dlsym(a.handle, "ncclGetUniqueId");
    a.comm_init_rank = (nccl_comm_init_rank_t)dlsym(a.handle, "ncclCommInitRank");
    a.comm_destroy = (nccl_comm_destroy_t)dlsym(a.handle, "ncclCommDestroy");
    a.all_reduce = (nccl_all_reduce_t)dlsym(a.handle, "ncclAllReduce");
    a.ok = a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce;
    return a;
  }();
  return api;
}

}  // namespace gdft
using namespace gdft;

// ---- gdft_allreduce_fock(ncclComm_t, cudaStream_t, double* packed_E_and_Dbar, size_t count) ---------------
extern "C" int gdft_nccl_available(void) { return nccl_api().ok ? 1 : 0; }
extern "C" size_t gdft_nccl_unique_id_bytes(void) { return sizeof(NcclUniqueId); }
extern "C" int gdft_nccl_unique_id(void* id_host) {
  if (!id_host) return GDFT_BAD_ARGUMENT;
  if (!nccl_api().ok) return GDFT_BAD_ARGUMENT;
  return nccl_api().get_unique_id(static_cast<NcclUniqueId*>(id_host)) == 0 ? GDFT_OK : GDFT_CUDA_ERROR;
}
extern "C" int gdft_nccl_comm_create(const void* id_host, int rank, int world, void** comm_out) {
  if (!id_host || !comm_out || world < 1 || rank < 0 || rank >= world) return GDFT_BAD_ARGUMENT;
  if (!nccl_api().ok) return GDFT_BAD_ARGUMENT;
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  return nccl_api().comm_init_rank(comm_out, world, id, rank) == 0 ? GDFT_OK : GDFT_CUDA_ERROR;
}
extern "C" int gdft_nccl_comm_destroy(void* comm) {
  if (!comm || !nccl_api().ok) return GDFT_BAD_ARGUMENT;
  return nccl_api().comm_destroy(comm) == 0 ? GDFT_OK : GDFT_CUDA_ERROR;
}
extern "C" int gdft_allreduce_fock(void* nccl_comm, gdft_stream_t stream, double* packed, size_t count) {
  if (!nccl_comm || !packed) return GDFT_BAD_ARGUMENT;
  if (count == 0) return GDFT_OK;
  if (!nccl_api().ok) return GDFT_BAD_ARGUMENT;
  const int nccl_float64 = 8, nccl_sum = 0;
  return nccl_api().all_reduce(packed, packed, count, nccl_float64, nccl_sum, nccl_comm, (cudaStream_t)stream) == 0 ? GDFT_OK : GDFT_CUDA_ERROR;
}

// ---- peer-memory communicator ----------------------------------------------------------------------------------------------
extern "C" size_t gdft_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

// Setup call (allocates): header + `capacity` doubles of payload on the current device, zero-initialised.
extern "C" int gdft_comm_create(int rank, int world, size_t capacity, gdft_comm** out) {
  if (!out || world < 1 || world > GDFT_COMM_MAX_RANKS || rank < 0 || rank >= world || capacity == 0) return GDFT_BAD_ARGUMENT;
  gdft_comm* c = new (std::nothrow) gdft_comm();
  if (!c) return GDFT_BAD_ARGUMENT;
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  c->capacity = (capacity + 1) & ~size_t(1);
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&c->local, GDFT_COMM_HEADER_BYTES + c->capacity * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, GDFT_COMM_HEADER_BYTES + c->capacity * sizeof(double));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    delete c;
    return cuda_fail(e);
  }
  c->peers.base[rank] = c->local;
  c->connected = (world == 1);
  *out = c;
  return GDFT_OK;
}
extern "C" double* gdft_comm_buffer(gdft_comm* c) { return c ? reinterpret_cast<double*>(c->local + GDFT_COMM_HEADER_BYTES) : nullptr; }
extern "C" size_t gdft_comm_capacity(gdft_comm* c) { return c ? c->capacity : 0; }
extern "C" int gdft_comm_handle(gdft_comm* c, void* handle_host) {
  if (!c || !handle_host) return GDFT_BAD_ARGUMENT;
  cudaIpcMemHandle_t h;
  GDFT_CUDA_TRY(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle_host, &h, sizeof(h));
  return GDFT_OK;
}
// handles_host: world handles in rank order (the local one is ignored): one process per GPU, exchanged by the host side
extern "C" int gdft_comm_connect(gdft_comm* c, const void* handles_host) {
  if (!c || !handles_host) return GDFT_BAD_ARGUMENT;
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank || c->opened[p]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles_host) + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    GDFT_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peers.base[p] = static_cast<char*>(ptr);
    c->opened[p] = true;
  }
  c->connected = true;
  return GDFT_OK;
}
// Same-process variant (several ranks driven by one process, e.g. one host thread per GPU, or the single-GPU test):
// `all` lists the communicators of every rank in rank order; peer access between their devices is enabled here.
extern "C" int gdft_comm_connect_local(gdft_comm* c, gdft_comm* const* all) {
  if (!c || !all) return GDFT_BAD_ARGUMENT;
  for (int p = 0; p < c->world; ++p) {
    if (!all[p] || all[p]->world != c->world || all[p]->rank != p || all[p]->capacity != c->capacity) return GDFT_BAD_ARGUMENT;
    if (all[p]->device != c->device) {
      int can = 0;
      GDFT_CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, all[p]->device));
      if (!can) return GDFT_BAD_ARGUMENT;
      int cur = 0;
      GDFT_CUDA_TRY(cudaGetDevice(&cur));
      GDFT_CUDA_TRY(cudaSetDevice(c->device));
      cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
      cudaSetDevice(cur);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e);
      (void)cudaGetLastError();
    }
    c->peers.base[p] = all[p]->local;
  }
  c->connected = true;
  return GDFT_OK;
}
// Host-synchronous (debug / end of run): 0 ok, 1 a flag poll timed out in some exchange since creation
extern "C" int gdft_comm_status(gdft_comm* c, int* status_host, unsigned long long* epoch_host) {
  if (!c) return GDFT_BAD_ARGUMENT;
  CommHeader h;
  GDFT_CUDA_TRY(cudaMemcpy(&h, c->local, sizeof(h), cudaMemcpyDeviceToHost));
  if (status_host) *status_host = h.status;
  if (epoch_host) *epoch_host = h.epoch;
  return GDFT_OK;
}
extern "C" int gdft_comm_destroy(gdft_comm* c) {
  if (!c) return GDFT_BAD_ARGUMENT;
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->peers.base[p]);
  cudaFree(c->local);
  delete c;
  return GDFT_OK;
}

// In place on the communicator's payload: payload[0:count] <- sum over ranks, stream-ordered, graph-capturable.
extern "C" int gdft_allreduce_fock_p2p(gdft_stream_t stream, gdft_comm* c, size_t count) {
  if (!c || !c->connected) return GDFT_BAD_ARGUMENT;
  if (count > c->capacity) return GDFT_BAD_SHAPE;
  if (count == 0 || c->world == 1) return GDFT_OK;
  const long long count2 = (long long)((count + 1) / 2);  // capacity is even and the tail is owned by the communicator
  const long long per = (count2 + c->world - 1) / c->world;
  int blocks = (int)imin64(imax64((per + 1023) / 1024, 1), 96);  // ~4 elements per thread, <= 96 CTAs per rank: the exchange is latency-bound
  cudaStream_t s = (cudaStream_t)stream;
  switch (c->world) {
    case 2: allreduce_fock_kernel<2><<<blocks, 256, 0, s>>>(c->peers, c->rank, c->world, count2); break;
    case 4: allreduce_fock_kernel<4><<<blocks, 256, 0, s>>>(c->peers, c->rank, c->world, count2); break;
    case 8: allreduce_fock_kernel<8><<<blocks, 256, 0, s>>>(c->peers, c->rank, c->world, count2); break;
    default: allreduce_fock_kernel<0><<<blocks, 256, 0, s>>>(c->peers, c->rank, c->world, count2); break;
  }
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
