// CPU emulation shim for the CUDA sources under co_snarks_b200/csrc (TEST INFRASTRUCTURE ONLY).
//
// Built by tests/emu/build_emu.py with g++ -DCS_EMU; lets the exact device algorithms (field
// arithmetic, point formulas, MSM / NTT kernels and their host drivers) run on a box without a GPU
// so they can be compared with the oracle before GPU minutes are spent.  It is never linked into
// the product library and is not a fallback: libcosnarks_gpu.so has no CPU path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

#define __shared__ static
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __ldg(p) (*(p))

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2

namespace cs { namespace emu {
void syncthreads();
void* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem, bool uses_sync, const std::function<void()>& body);
}}  // namespace cs::emu

static inline void __syncthreads() { cs::emu::syncthreads(); }
static inline void __syncwarp() {}
// warp shuffles (32-bit), emulated with a block-wide exchange: every thread of the block must make the call together,
// i.e. the kernel is launched with CS_LAUNCH_SYNC and has no early return ahead of it
namespace cs { namespace emu { uint32_t shfl_from(uint32_t v, int delta_kind, unsigned arg);
void shfl_bytes(const void* in, void* out, size_t bytes, int kind, unsigned arg); } }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) { return cs::emu::shfl_from(v, 0, d); }
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) { return cs::emu::shfl_from(v, 1, d); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, unsigned m) { return cs::emu::shfl_from(v, 2, m); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, unsigned lane) { return cs::emu::shfl_from(v, 3, lane); }
static inline void __threadfence() {}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
  uint32_t o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __brev(uint32_t x) {
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}

// ---- minimal fake runtime ------------------------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return cudaSuccess; }
// IPC within the emulation = the pointer itself (all "processes" share one address space)
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** out, cudaIpcMemHandle_t h, unsigned) {
  memcpy(out, h.reserved, sizeof(void*)); return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8

#define CS_LAUNCH(kernel, grid, block, smem, stream, ...) \
  do { cs::launch_counter()++; cs::emu::launch(dim3(grid), dim3(block), smem, false, [&]() { kernel(__VA_ARGS__); }); } while (0)
#define CS_LAUNCH_SYNC(kernel, grid, block, smem, stream, ...) \
  do { cs::launch_counter()++; cs::emu::launch(dim3(grid), dim3(block), smem, true, [&]() { kernel(__VA_ARGS__); }); } while (0)
#define CS_DYN_SMEM(type, name) type* name = (type*)cs::emu::dyn_smem()
