// Integer-pipe microbenchmark for the MSM/NTT roofline: sustained IMAD (32x32+32) and IMAD.WIDE
// (32x32+64) issue rate on B200, and the throughput of the library's Montgomery multiplication.
// MEASURED_PEAKS.json only carries HBM and bf16 peaks; 256-bit modular arithmetic is bound by this
// pipe instead (DESIGN.md "Rooflines").  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../co_snarks_b200/csrc/cs_params.cuh"
#include "../co_snarks_b200/csrc/cs_field.cuh"

template <int CHAINS>
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint64_t acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a + c), "r"(b));
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a + c), "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// CH independent Montgomery-multiplication chains per thread
template <int CH>
__global__ void k_montmul(uint32_t* out, int iters) {
  typedef cs::Fp<cs::Bn254Fq> F;
  F x[CH], y;
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 8; i++) x[c].l[i] = threadIdx.x * 77 + i + c;
  for (int i = 0; i < 8; i++) y.l[i] = blockIdx.x + i * 3 + 1;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = x[c] * y;
  }
  uint32_t s = 0;
  for (int c = 0; c < CH; c++) s += x[c].l[0] ^ x[c].l[7];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_ms(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  void* buf;
  cudaMalloc(&buf, (size_t)sms * 16 * 1024 * 8);
  printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 4096;
  for (int threads : {256, 512, 1024}) {
    int blocks = sms * (2048 / threads);
    float ms = time_ms([&] { k_imad_wide<8><<<blocks, threads>>>((uint64_t*)buf, 3, 5, iters); });
    double ops = (double)blocks * threads * 8 * iters;
    printf(", \"imad_wide_tops_t%d\": %.3f", threads, ops / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_imad_lo<8><<<blocks, threads>>>((uint32_t*)buf, 3, 5, iters); });
    printf(", \"imad_lo_tops_t%d\": %.3f", threads, ops / (ms * 1e-3) / 1e12);
  }
  for (int wps : {1, 2, 3, 4, 8}) {  // warps per SM sub-partition
    int threads = 128, blocks = sms * wps;  // 128 threads = 1 warp per SMSP per block
    float ms = time_ms([&] { k_montmul<1><<<blocks, threads>>>((uint32_t*)buf, 2048); });
    double muls = (double)blocks * threads * 2048;
    printf(", \"montmul_gmuls_ch1_w%d\": %.2f", wps, muls / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_montmul<2><<<blocks, threads>>>((uint32_t*)buf, 2048); });
    printf(", \"montmul_gmuls_ch2_w%d\": %.2f", wps, 2 * muls / (ms * 1e-3) / 1e9);
  }
  printf("}\n");
  return 0;
}
