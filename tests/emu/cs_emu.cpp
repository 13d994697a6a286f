// CPU emulation shim runtime (TEST INFRASTRUCTURE ONLY) -- see cs_emu.h.
#include <stdio.h>
#include "cs_emu.h"

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace cs { namespace emu {

static std::vector<unsigned char> g_smem;
static bool g_threaded = false;

// simple reusable barrier
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  unsigned count = 0, waiting = 0, gen = 0;
  void reset(unsigned n) { count = n; waiting = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    unsigned g = gen;
    if (++waiting == count) { waiting = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
static Barrier g_bar;

void syncthreads() {
  if (!g_threaded) { fprintf(stderr, "cs_emu: __syncthreads in a kernel launched without CS_LAUNCH_SYNC\n"); abort(); }
  g_bar.wait();
}
void* dyn_smem() { return g_smem.data(); }

static std::mutex g_launch_mutex;  // one emulated kernel at a time, whichever host thread launches it

void launch(dim3 grid, dim3 block, size_t smem, bool uses_sync, const std::function<void()>& body) {
  std::lock_guard<std::mutex> launch_lock(g_launch_mutex);
  g_smem.assign(smem + 16, 0);
  unsigned nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        auto run_thread = [&](unsigned t) {
          blockIdx = uint3{bx, by, bz};
          blockDim = block;
          gridDim = grid;
          threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          body();
        };
        if (!uses_sync) {
          g_threaded = false;
          for (unsigned t = 0; t < nthreads; t++) run_thread(t);
        } else {
          g_threaded = true;
          g_bar.reset(nthreads);
          std::vector<std::thread> th;
          th.reserve(nthreads);
          for (unsigned t = 0; t < nthreads; t++) th.emplace_back(run_thread, t);
          for (auto& x : th) x.join();
          g_threaded = false;
        }
      }
}

}}  // namespace cs::emu
