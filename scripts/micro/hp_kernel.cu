// micro-benchmark of ns::highpass_body (K0) in isolation on B200
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../crispy_b200/csrc/ns_pipe.cuh"

__global__ void __launch_bounds__(ns::kHpThreads) kern(const __grid_constant__ ns::Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ns::highpass_body(p, *reinterpret_cast<ns::HpSmem *>(smem_raw));
}

int main(int argc, char **argv) {
  const int n = 1024, nf = argc > 1 ? atoi(argv[1]) : 24;
  ns::Params p;
  memset(&p, 0, sizeof(p));
  float *in, *hp, *state;
  cudaMalloc(&in, (size_t)n * nf * 480 * 4);
  cudaMemset(in, 0, (size_t)n * nf * 480 * 4);
  p.hp_stride = ns::kHist + nf * 480;
  cudaMalloc(&hp, (size_t)n * p.hp_stride * 4);
  cudaMalloc(&state, (size_t)n * ns::kStateFloats * 4);
  cudaMemset(state, 0, (size_t)n * ns::kStateFloats * 4);
  p.in = in; p.hp = hp; p.state = state; p.in_stride = nf * 480; p.n_streams = n; p.n_frames = nf; p.chunk_cap = nf;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ns::HpSmem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a); kern<<<n / 32, ns::kHpThreads, sizeof(ns::HpSmem)>>>(p); cudaEventRecord(b);
    cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, a, b);
    printf("highpass_body %d frames x %d streams: %.3f ms = %.1f ns/sample (%s)\n", nf, n, ms, ms * 1e6 / (nf * 480.0),
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
