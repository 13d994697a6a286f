// CPU emulation shim runtime (TEST INFRASTRUCTURE ONLY) -- see cs_emu.h.
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <algorithm>
#include "cs_emu.h"

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace cs { namespace emu {

static std::vector<unsigned char> g_smem;
static bool g_threaded = false;

// reusable sense-reversing barrier; spins with yield (the emulated blocks have more threads than the box has cores)
struct Barrier {
  std::atomic<unsigned> waiting{0}, gen{0};
  unsigned count = 0;
  void reset(unsigned n) { count = n; waiting.store(0); }
  void wait() {
    const unsigned g = gen.load(std::memory_order_acquire);
    if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == count) {
      waiting.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    } else {
      unsigned spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 16) std::this_thread::yield();
    }
  }
};
static Barrier g_bar;

void syncthreads() {
  if (!g_threaded) { fprintf(stderr, "cs_emu: __syncthreads in a kernel launched without CS_LAUNCH_SYNC\n"); abort(); }
  g_bar.wait();
}
void* dyn_smem() { return g_smem.data(); }

static unsigned char g_shfl_bytes[1024][512];
void shfl_bytes(const void* in, void* out, size_t bytes, int kind, unsigned arg) {
  if (!g_threaded) { fprintf(stderr, "cs_emu: warp shuffle in a kernel launched without CS_LAUNCH_SYNC\n"); abort(); }
  if (bytes > 512) { fprintf(stderr, "cs_emu: shuffle payload too large\n"); abort(); }
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned lane = tid & 31, base = tid & ~31u;
  unsigned src = tid;
  if (kind == 0) src = lane >= arg ? tid - arg : tid;
  else if (kind == 1) src = lane + arg < 32 ? tid + arg : tid;
  else if (kind == 2) src = base | ((lane ^ arg) & 31);
  else src = base | (arg & 31);
  const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
  if (src >= nthreads) src = tid;
  memcpy(g_shfl_bytes[tid], in, bytes);
  g_bar.wait();
  memcpy(out, g_shfl_bytes[src], bytes);
  g_bar.wait();
}
static uint32_t g_shfl[1024];
uint32_t shfl_from(uint32_t v, int kind, unsigned arg) {
  if (!g_threaded) { fprintf(stderr, "cs_emu: warp shuffle in a kernel launched without CS_LAUNCH_SYNC\n"); abort(); }
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned lane = tid & 31, base = tid & ~31u;
  unsigned src = tid;
  if (kind == 0) src = lane >= arg ? tid - arg : tid;
  else if (kind == 1) src = lane + arg < 32 ? tid + arg : tid;
  else if (kind == 2) src = base | ((lane ^ arg) & 31);
  else src = base | (arg & 31);
  const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
  if (src >= nthreads) src = tid;
  g_shfl[tid] = v;
  g_bar.wait();
  const uint32_t r = g_shfl[src];
  g_bar.wait();
  return r;
}

static std::mutex g_launch_mutex;  // one emulated kernel at a time, whichever host thread launches it

void launch(dim3 grid, dim3 block, size_t smem, bool uses_sync, const std::function<void()>& body) {
  std::lock_guard<std::mutex> launch_lock(g_launch_mutex);
  g_smem.assign(smem + 16, 0);
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  auto set_ids = [&](unsigned long long bl, unsigned t) {
    blockIdx = uint3{(unsigned)(bl % grid.x), (unsigned)((bl / grid.x) % grid.y), (unsigned)(bl / ((unsigned long long)grid.x * grid.y))};
    blockDim = block;
    gridDim = grid;
    threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
  };
  if (!uses_sync) {
    g_threaded = false;
    for (unsigned long long bl = 0; bl < nblocks; bl++)
      for (unsigned t = 0; t < nthreads; t++) { set_ids(bl, t); body(); }
    return;
  }
  // kernels with barriers / shuffles: one host thread per CUDA thread, created once per launch; the blocks run one
  // after the other (static __shared__ storage is a single copy), separated by a barrier
  g_threaded = true;
  g_bar.reset(nthreads);
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      for (unsigned long long bl = 0; bl < nblocks; bl++) {
        set_ids(bl, t);
        body();
        g_bar.wait();
        if (t == 0 && smem) std::fill(g_smem.begin(), g_smem.end(), 0);
        if (smem) g_bar.wait();
      }
    });
  for (auto& x : th) x.join();
  g_threaded = false;
}

}}  // namespace cs::emu
