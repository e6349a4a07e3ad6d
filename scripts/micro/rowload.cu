// micro-benchmark: a warp streams 32 rows (one per stream) W bytes at a time; how long per round?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VEC>  // floats per lane per row per round (1 -> 128 B per row, 4 -> 512 B per row)
__global__ void k(const float *in, float *out, long long stride, int rounds, int depth) {
  const int lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * 32;
  float acc = 0.f;
  for (int it = 0; it < rounds; it++) {
    float v[32 * VEC];
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const float *p = in + (long long)(s0 + r) * stride + (long long)it * 32 * VEC + lane * VEC;
      if (VEC == 1) v[r] = *p;
      else { float4 q = *reinterpret_cast<const float4 *>(p); v[4*r]=q.x; v[4*r+1]=q.y; v[4*r+2]=q.z; v[4*r+3]=q.w; }
    }
#pragma unroll
    for (int r = 0; r < 32 * VEC; r++) acc += v[r];
  }
  out[blockIdx.x * 32 + lane] = acc;
}

int main() {
  const int n = 1024;
  float *in, *out;
  const long long maxstride = 16 * 1024 * 1024 / 4 + 4096;
  cudaMalloc(&in, (size_t)n * maxstride * 4 / 8);  // enough for strides up to 2 MB
  cudaMalloc(&out, n * 4);
  cudaMemset(in, 0, (size_t)n * maxstride * 4 / 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const long long strides[] = {11520, 11520 + 32, 11520 + 32 * 17, 12960, 16384, 16384 + 32, 65536, 65536 + 32 * 5, 500000};
  for (long long st : strides) {
    for (int vec = 1; vec <= 4; vec += 3) {
      const int rounds = 360 / vec;
      float ms;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        if (vec == 1) k<1><<<n / 32, 32>>>(in, out, st, rounds, 0); else k<4><<<n / 32, 32>>>(in, out, st, rounds, 0);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, a, b);
      }
      printf("stride %8lld floats (%9lld B), %3d B per row per round: %.3f ms total, %.2f us per round, %.1f GB/s (%s)\n", st, st * 4,
             vec * 128, ms, ms * 1e3 / rounds, n * 11520.0 * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
