// micro-benchmark: issue cost of the FP64-side instructions the exact biquad uses (F2F both ways, DFMA, DADD, DMUL)
// and of FADD for scale, per warp-instruction, with 1, 2, 4 and 8 warps resident on one SM: tells whether the unit
// behind each is private to an SM sub-partition (cycles per instruction stay flat up to 4 warps) or shared by the SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(long long *cycles, float *sink, int iters, float seed) {
  // eight independent chains per thread so that latency does not bound the loop
  float f[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = seed + i + threadIdx.x, d[i] = (double)seed * (i + 1) + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) d[i] = (double)f[i], f[i] = __double2float_rn(d[i]) + 1.0f;       // F2F.F64.F32 + F2F.F32.F64 + FADD
      if (OP == 1) d[i] = fma(d[i], 1.0000001, 0.5);                                // DFMA
      if (OP == 2) d[i] = d[i] + 1.25;                                              // DADD
      if (OP == 3) d[i] = d[i] * 1.0000001;                                         // DMUL
      if (OP == 4) f[i] = __fadd_rn(f[i], 1.25f);                                   // FADD
      if (OP == 5) d[i] = d[i] + (double)f[i], f[i] = __fadd_rn(f[i], 1.0f);        // F2F.F64.F32 + DADD + FADD
      if (OP == 6) f[i] = __fadd_rn(f[i], __double2float_rn(d[i])), d[i] = d[i] + 1.25;  // F2F.F32.F64 + FADD + DADD
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) acc += f[i] + (float)d[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  long long *c;
  float *s;
  cudaMalloc(&c, 8 * 16);
  cudaMalloc(&s, 4 * 1024);
  const int iters = 20000;
  const char *names[] = {"F2F.F64.F32 + F2F.F32.F64 + FADD", "DFMA", "DADD", "DMUL", "FADD", "F2F.F64.F32 + DADD + FADD", "F2F.F32.F64 + FADD + DADD"};
  for (int op = 0; op < 7; op++) {
    printf("%-36s", names[op]);
    for (int warps : {1, 2, 4, 8}) {
      for (int rep = 0; rep < 2; rep++) {
        switch (op) {
          case 0: k<0><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          case 1: k<1><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          case 2: k<2><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          case 3: k<3><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          case 4: k<4><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          case 5: k<5><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
          default: k<6><<<1, 32 * warps>>>(c, s, iters, 1.5f); break;
        }
      }
      cudaDeviceSynchronize();
      long long hc;
      cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
      printf("  %dw: %6.2f", warps, (double)hc / ((double)iters * 8));
    }
    printf("   cycles per group of instructions per warp (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
