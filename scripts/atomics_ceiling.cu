// Ceiling experiment for the high-cardinality group-by: how fast can 60 M rows fold 7 accumulator words
// each into G groups with native 64-bit atomics, whatever else the kernel does?  Rows carry a random group id; variants:
//   soa   acc[w][G]      (the operator's layout: one array per accumulator word)
//   aos   acc[g][8]      (one 64-byte line per group)
//   none  no atomics: the loads + hash only (the streaming floor of this harness)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/atomics_ceiling.cu -o /tmp/atomics_ceiling
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(const double* __restrict__ v, long long n, u64 G, double* acc, u64* cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double sink = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = __ldcs(v + i);
    const u64 g = mix((u64)i) % G;
    if (MODE == 0) {  // soa
#pragma unroll
      for (int w = 0; w < 5; w++) atomicAdd(acc + (size_t)w * G + g, x);
      atomicAdd(cnt + g, 1ULL);
      atomicAdd(cnt + G + g, (u64)x);
    } else if (MODE == 1) {  // aos
      double* a = acc + g * 8;
#pragma unroll
      for (int w = 0; w < 5; w++) atomicAdd(a + w, x);
      atomicAdd((u64*)(a + 5), 1ULL);
      atomicAdd((u64*)(a + 6), (u64)x);
    } else {
      sink += x * (double)g;
    }
  }
  if (MODE == 2 && sink == 12345.678) acc[0] = sink;
}
int main() {
  const long long n = 60000003;
  double* v; cudaMalloc(&v, n * 8); cudaMemset(v, 0, n * 8);
  for (u64 G : {8ULL, 50ULL, 2500ULL, 125000ULL, 4000000ULL}) {
    double* acc; u64* cnt;
    cudaMalloc(&acc, G * 8 * 8); cudaMalloc(&cnt, G * 2 * 8);
    cudaMemset(acc, 0, G * 64); cudaMemset(cnt, 0, G * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"soa", "aos", "none"};
    for (int mode = 0; mode < 3; mode++) {
      float best = 1e9;
      for (int it = 0; it < 3; it++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(v, n, G, acc, cnt);
        if (mode == 1) k<1><<<148 * 8, 256>>>(v, n, G, acc, cnt);
        if (mode == 2) k<2><<<148 * 8, 256>>>(v, n, G, acc, cnt);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("G=%8llu %-4s %8.3f ms  %7.1f G atomics/s\n", G, names[mode], best, mode < 2 ? n * 7 / best / 1e6 : 0.0);
    }
    cudaFree(acc); cudaFree(cnt);
  }
  return 0;
}
