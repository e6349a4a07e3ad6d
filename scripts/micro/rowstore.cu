// micro-benchmark: per-round cost of streaming 32 rows x 128 B: loads, stores, both, on one warp / two warps
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(const float *in, float *out, float *sink, long long stride, int rounds) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s0 = blockIdx.x * 32;
  float acc = 0.f;
  for (int it = 0; it < rounds; it++) {
    float v[32];
    const bool do_ld = (MODE == 0 || MODE == 2 || (MODE == 3 && warp == 0));
    const bool do_st = (MODE == 1 || MODE == 2 || (MODE == 3 && warp == 1));
    if (do_ld) {
#pragma unroll
      for (int r = 0; r < 32; r++) v[r] = in[(long long)(s0 + r) * stride + (long long)it * 32 + lane];
#pragma unroll
      for (int r = 0; r < 32; r++) acc += v[r];
    }
    if (do_st) {
#pragma unroll
      for (int r = 0; r < 32; r++) out[(long long)(s0 + r) * stride + (long long)it * 32 + lane] = acc + r;
    }
    if (MODE == 3) __syncthreads();
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int n = 1024;
  const long long st = 12960;
  float *in, *out, *sink;
  cudaMalloc(&in, (size_t)n * st * 4);
  cudaMalloc(&out, (size_t)n * st * 4);
  cudaMalloc(&sink, n * 8);
  cudaMemset(in, 0, (size_t)n * st * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int rounds = 360;
  const char *names[] = {"loads only", "stores only", "loads+stores one warp", "loads warp0 / stores warp1 + barrier"};
  for (int mode = 0; mode < 4; mode++) {
    float ms;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<n / 32, 32>>>(in, out, sink, st, rounds);
      if (mode == 1) k<1><<<n / 32, 32>>>(in, out, sink, st, rounds);
      if (mode == 2) k<2><<<n / 32, 32>>>(in, out, sink, st, rounds);
      if (mode == 3) k<3><<<n / 32, 64>>>(in, out, sink, st, rounds);
      cudaEventRecord(b);
      cudaDeviceSynchronize();
      cudaEventElapsedTime(&ms, a, b);
    }
    printf("%-40s: %.3f ms total, %.2f us per round (%s)\n", names[mode], ms, ms * 1e3 / rounds, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
