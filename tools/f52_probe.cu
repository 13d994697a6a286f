// FP64-pipe Montgomery product (cs_field52.cuh) against the integer-pipe product (cs_field.cuh) on the device:
// bit-exactness on random operands and sustained throughput.  Output: one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../co_snarks_b200/csrc/cs_params.cuh"
#include "../co_snarks_b200/csrc/cs_params52.cuh"
#include "../co_snarks_b200/csrc/cs_field52.cuh"

using namespace cs;
typedef Fp<Bn254Fq> F;

__device__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

__global__ void k_check(int* mismatches, int rounds) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int r = 0; r < rounds; r++) {
    F x, y;
    for (int i = 0; i < 8; i++) { x.l[i] = lcg(s); y.l[i] = lcg(s); }
    x.l[7] &= 0x1fffffff; y.l[7] &= 0x1fffffff;  // < 2^253 < p
    if (r == 0 && threadIdx.x < 4) {  // edge values
      for (int i = 0; i < 8; i++) x.l[i] = (threadIdx.x & 1) ? Bn254Fq::mod(i) : 0;
      if (threadIdx.x & 1) x.l[0] -= 1;               // p - 1
      if (threadIdx.x & 2) y = x;
    }
    F ref = x * y;
    I52 a = f52_from_fp<Bn254Fq52, Bn254Fq>(x), b = f52_from_fp<Bn254Fq52, Bn254Fq>(y);
    I52 m = f52_mul<Bn254Fq52>(f52_to_double(a), f52_to_double(b));
    F got = f52_to_fp<Bn254Fq52, Bn254Fq>(m);
    F ref2 = x.sqr();
    F got2 = f52_to_fp<Bn254Fq52, Bn254Fq>(f52_sqr<Bn254Fq52>(f52_to_double(a)));
    // lazy chain: ((a*b) - a + 4p) * (b + a) stays within the 8p operand bound
    I52 d = f52_sub<Bn254Fq52, 4>(m, a);
    I52 e = f52_add_i<Bn254Fq52>(b, a);
    F got3 = f52_to_fp<Bn254Fq52, Bn254Fq>(f52_mul<Bn254Fq52>(f52_to_double(d), f52_to_double(e)));
    F ref3 = (ref - x) * (y + x);
    if (!(ref == got) || !(ref2 == got2) || !(ref3 == got3)) atomicAdd(mismatches, 1);
  }
}

template <int CH>
__global__ void k_mul52(uint64_t* out, int iters) {
  I52 x[CH];
  D52 y;
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 5; i++) x[c].l[i] = ((uint64_t)(threadIdx.x * 77 + i + c) * 0x9E3779B97F4A7C15ull) & F52_MASK;
  for (int i = 0; i < 5; i++) y.l[i] = (double)(((uint64_t)(blockIdx.x + i * 3 + 1) * 0x9E3779B97F4A7C15ull) & F52_MASK);
  x[0].l[4] &= 0xffffffffffull; y.l[4] = 12345.0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = f52_mul<Bn254Fq52>(f52_to_double(x[c]), y);
  }
  uint64_t s = 0;
  for (int c = 0; c < CH; c++) s += x[c].l[0] ^ x[c].l[4];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_sqr52(uint64_t* out, int iters) {
  I52 x[CH];
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 5; i++) x[c].l[i] = ((uint64_t)(threadIdx.x * 77 + i + c) * 0x9E3779B97F4A7C15ull) & (i == 4 ? 0xffffffffffull : F52_MASK);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = f52_sqr<Bn254Fq52>(f52_to_double(x[c]));
  }
  uint64_t s = 0;
  for (int c = 0; c < CH; c++) s += x[c].l[0] ^ x[c].l[4];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_mul32(uint32_t* out, int iters) {
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
  const int sms = p.multiProcessorCount;
  void* buf;
  cudaMalloc(&buf, (size_t)sms * 2048 * 8 * 2);
  int* d_mis;
  cudaMalloc(&d_mis, 4);
  cudaMemset(d_mis, 0, 4);
  k_check<<<sms * 2, 128>>>(d_mis, 64);
  cudaDeviceSynchronize();
  int mis = -1;
  cudaMemcpy(&mis, d_mis, 4, cudaMemcpyDeviceToHost);
  printf("{\"device\": \"%s\", \"checked_products\": %d, \"mismatches\": %d, \"last_error\": \"%s\"", p.name, sms * 2 * 128 * 64 * 3, mis,
         cudaGetErrorString(cudaGetLastError()));
  const int iters = 1024;
  for (int wps : {1, 2, 3, 4, 6, 8}) {
    const int threads = 128, blocks = sms * wps;
    const double muls = (double)blocks * threads * iters;
    float ms = time_ms([&] { k_mul52<1><<<blocks, threads>>>((uint64_t*)buf, iters); });
    printf(", \"mul52_gmuls_ch1_w%d\": %.2f", wps, muls / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_mul52<2><<<blocks, threads>>>((uint64_t*)buf, iters); });
    printf(", \"mul52_gmuls_ch2_w%d\": %.2f", wps, 2 * muls / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_sqr52<1><<<blocks, threads>>>((uint64_t*)buf, iters); });
    printf(", \"sqr52_gmuls_ch1_w%d\": %.2f", wps, muls / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { k_mul32<1><<<blocks, threads>>>((uint32_t*)buf, iters); });
    printf(", \"mul32_gmuls_ch1_w%d\": %.2f", wps, muls / (ms * 1e-3) / 1e9);
  }
  printf("}\n");
  return 0;
}
