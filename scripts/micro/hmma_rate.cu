// hmma_rate.cu -- how fast does sm_100a issue the legacy warp-level mma.sync.m16n8k16 (bf16 -> f32)?  K4 and the v7 pitch
// filter are built on it.  Reports cycles per HMMA per warp for dependent and independent accumulator chains and the
// aggregate rate per SM at 1..16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void k(float *out, int iters, unsigned long long *cycles) {
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f003f00u};
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
#pragma unroll
    for (int e = 0; e < 4; e++) d[c][e] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) mma(d[c], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += d[c][0] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
}

template <int CHAINS>
void run(int warps_per_sm, float *d_out, unsigned long long *d_cyc) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CHAINS><<<148, 32 * warps_per_sm>>>(d_out, 64, d_cyc);
  cudaEventRecord(e0);
  k<CHAINS><<<148, 32 * warps_per_sm>>>(d_out, iters, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cyc;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * CHAINS;
  printf("chains %d warps/SM %2d: %6.1f cycles per HMMA per warp, %6.2f HMMA/cycle/SM, %7.1f dense TFLOP/s\n", CHAINS, warps_per_sm,
         cyc / n, n * warps_per_sm / cyc, 148.0 * warps_per_sm * n * (16 * 8 * 16 * 2) / (ms * 1e-3) / 1e12);
}

int main() {
  float *d_out;
  unsigned long long *d_cyc;
  cudaMalloc(&d_out, 148 * 1024 * sizeof(float));
  cudaMalloc(&d_cyc, 8);
  for (int w : {1, 4, 8, 16}) run<1>(w, d_out, d_cyc);
  for (int w : {1, 4, 8, 16}) run<4>(w, d_out, d_cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
