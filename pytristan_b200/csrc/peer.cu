// Peer-memory plumbing of the multi-GPU path (SURVEY.md 8e ii/iii): one process per GPU on one
// NVSwitch box, every rank maps the exchange buffers of the others (CUDA IPC) and the kernels
// store into / load from them directly over NVLink.
//
//   pt_peer_alloc / pt_peer_free      exchange buffers (cudaMalloc: IPC-exportable, zero filled)
//   pt_ipc_export / open / close      64-byte handles, exchanged by the host through torch.distributed
//   pt_peer_barrier                   device-side barrier: release-store of an epoch into every
//                                     peer's flag array, acquire-spin on the own one (bounded)
//   pt_peer_scatter                   redistribution of a field through a pt_scatter_map with
//                                     16-byte peer stores (the operand that is not produced by a
//                                     dense kernel; the produced one leaves through the scatter
//                                     epilogue of pt_dense_apply_scatter)
#include "common.cuh"
#include <string.h>

namespace {

__device__ __forceinline__ void st_release_sys(uint64_t* ptr, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* ptr) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

// flags[r] points at rank r's flag array (nranks slots of rank r's own memory).  Thread t
// publishes `epoch` in slot `rank` of peer t, then waits until slot t of the own array reached
// `epoch`.  All earlier work of the stream is visible to the peers before the flag is (the kernel
// starts after it; the store is a system-scope release), and everything the peers wrote before
// their flag is visible to the kernels that follow this one.
__global__ void peer_barrier_kernel(uint64_t* const* flags, int rank, int nranks, uint64_t epoch,
                                    uint64_t timeout_ns, int* status) {
  const int t = threadIdx.x;
  if (t >= nranks) return;
  __threadfence_system();
  st_release_sys(flags[t] + rank, epoch);
  const uint64_t* mine = flags[rank] + t;
  const uint64_t t0 = global_timer_ns();
  while (ld_acquire_sys(mine) < epoch) {
    if (timeout_ns && global_timer_ns() - t0 > timeout_ns) {
      if (status) atomicExch(status, 1 + t);   // peer t never arrived
      break;
    }
    __nanosleep(100);
  }
  __threadfence_system();
}

struct ScatterCopyParams {
  const double* __restrict__ src;
  double* const* base;
  int n, chunk;
  int64_t after, before;
  int64_t a_inner, s_outer, s_inner, s_i, offset;
};

// One thread per V doubles of a contiguous run.  before > 1: a run is the `before` contiguous
// doubles of one (a, i); before == 1 (and s_i == 1): a run is the `chunk` rows that go to one rank.
// 128-thread CTAs at <= 40 registers: small enough to share an SM with two resident CTAs of the
// even-odd DMMA kernel (224 registers x 128 threads each), so the redistribution of the field
// really runs UNDER the operator kernel of the other stream instead of after it
template <int V>
__global__ void __launch_bounds__(128) peer_scatter_kernel(const ScatterCopyParams p, int64_t run_len,
                                                            int64_t total) {
  const int64_t per_run = run_len / V;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t run = idx / per_run;
    const int64_t pos = (idx - run * per_run) * V;
    int64_t a, i, b;
    if (p.before > 1) { a = run / p.n; i = run - a * p.n; b = pos; }
    else {
      const int64_t runs_per_a = (p.n + p.chunk - 1) / p.chunk;
      a = run / runs_per_a;
      i = (run - a * runs_per_a) * p.chunk + pos;
      b = 0;
      if (i >= p.n) continue;
    }
    const int q = int(i / p.chunk);
    const int64_t ao = a / p.a_inner;
    double* dst = p.base[q] + p.offset + ao * p.s_outer + (a - ao * p.a_inner) * p.s_inner +
                  (i - int64_t(q) * p.chunk) * p.s_i + b;
    const double* src = p.src + (a * p.n + i) * p.before + b;
    if (V == 2) *reinterpret_cast<double2*>(dst) = *reinterpret_cast<const double2*>(src);
    else *dst = *src;
  }
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

extern "C" int pt_peer_alloc(int64_t bytes, void** ptr) {
  PT_REQUIRE(ptr && bytes > 0, "pt_peer_alloc: bad argument");
  void* q = nullptr;
  PT_CUDA(cudaMalloc(&q, size_t(bytes)));
  cudaError_t e = cudaMemset(q, 0, size_t(bytes));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(q);
    return pt::fail(PT_ERR_CUDA, "pt_peer_alloc: %s", cudaGetErrorString(e));
  }
  *ptr = q;
  return PT_OK;
}

extern "C" int pt_peer_free(void* ptr) {
  if (ptr) PT_CUDA(cudaFree(ptr));
  return PT_OK;
}

extern "C" int pt_ipc_export(const void* ptr, void* handle64) {
  PT_REQUIRE(ptr && handle64, "pt_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == PT_IPC_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  PT_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return PT_OK;
}

extern "C" int pt_ipc_open(const void* handle64, void** ptr) {
  PT_REQUIRE(ptr && handle64, "pt_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* q = nullptr;
  PT_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = q;
  return PT_OK;
}

extern "C" int pt_ipc_close(void* ptr) {
  if (ptr) PT_CUDA(cudaIpcCloseMemHandle(ptr));
  return PT_OK;
}

extern "C" int pt_peer_barrier(uint64_t* const* flags, int rank, int nranks, uint64_t epoch,
                               int64_t timeout_ns, int* status, pt_stream_t stream) {
  PT_REQUIRE(flags && nranks >= 1 && nranks <= 64 && rank >= 0 && rank < nranks,
             "pt_peer_barrier: bad argument (rank %d of %d)", rank, nranks);
  PT_REQUIRE(timeout_ns >= 0, "pt_peer_barrier: negative timeout");
  peer_barrier_kernel<<<1, 64, 0, pt::as_stream(stream)>>>(flags, rank, nranks, epoch,
                                                           uint64_t(timeout_ns), status);
  return pt::check_launch("pt_peer_barrier");
}

extern "C" int pt_peer_scatter(const double* src, int n, int64_t after, int64_t before,
                               const pt_scatter_map* map, pt_stream_t stream) {
  PT_REQUIRE(src && map && map->base, "pt_peer_scatter: null pointer");
  PT_REQUIRE(n >= 1 && after >= 0 && before >= 1, "pt_peer_scatter: bad field shape");
  PT_REQUIRE(map->nranks >= 1 && map->chunk >= 1 && int64_t(map->chunk) * map->nranks >= n,
             "pt_peer_scatter: %d destinations x %d rows do not cover n=%d", map->nranks, map->chunk, n);
  PT_REQUIRE(map->a_inner >= 1, "pt_peer_scatter: a_inner must be >= 1");
  PT_REQUIRE(before > 1 || map->s_i == 1, "pt_peer_scatter: before == 1 needs s_i == 1");
  if (after == 0) return PT_OK;
  ScatterCopyParams p;
  p.src = src; p.base = map->base; p.n = n; p.chunk = map->chunk; p.after = after; p.before = before;
  p.a_inner = map->a_inner; p.s_outer = map->s_outer; p.s_inner = map->s_inner; p.s_i = map->s_i;
  p.offset = map->offset;
  const int64_t run_len = before > 1 ? before : map->chunk;
  const int64_t runs = before > 1 ? after * n : after * ((n + map->chunk - 1) / map->chunk);
  const bool even = !((map->s_outer | map->s_inner | map->offset | run_len) & 1) &&
                    (before > 1 ? !(map->s_i & 1) : (n % 2 == 0));
  const bool vec = even && aligned16(src);
  const int64_t total = runs * (run_len / (vec ? 2 : 1));
  int64_t blocks = (total + 127) / 128;
  const int64_t cap = int64_t(pt::sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (vec) peer_scatter_kernel<2><<<unsigned(blocks), 128, 0, pt::as_stream(stream)>>>(p, run_len, total);
  else peer_scatter_kernel<1><<<unsigned(blocks), 128, 0, pt::as_stream(stream)>>>(p, run_len, total);
  return pt::check_launch("pt_peer_scatter");
}
