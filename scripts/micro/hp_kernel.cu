// micro-benchmark of ns::highpass_body pieces on B200
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../crispy_b200/csrc/ns_pipe.cuh"

template <int MODE>
__global__ void __launch_bounds__(32) kern(const __grid_constant__ ns::Params p) {
  __shared__ ns::HpSmem sm;
  if (MODE == 0) { ns::highpass_body(p, sm); return; }
  // MODE 1: recursion + tile I/O only (no history copies)
  const int lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * 32;
  const int nrows = 32;
  float m0 = 0.f, m1 = 0.f;
  const double a0 = (double)-1.99599f, a1 = (double)0.99600f;
  const int nsamp = p.n_frames * ns::kFrame;
  float nxt[32];
#pragma unroll
  for (int r = 0; r < 32; r++) nxt[r] = ns::load_sample(p, s0 + r, lane);
  for (int base = 0; base < nsamp; base += 32) {
#pragma unroll
    for (int r = 0; r < 32; r++) sm.tile[r][lane] = nxt[r];
    __syncwarp();
    if (base + 32 < nsamp) {
#pragma unroll
      for (int r = 0; r < 32; r++) nxt[r] = ns::load_sample(p, s0 + r, base + 32 + lane);
    }
    if (MODE != 3) {
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
      const float xi = sm.tile[lane][i];
      const float yi = xi + m0;
      const double xd = (double)xi, yd = (double)yi;
      m0 = (float)((double)m1 + (-2.0 * xd - a0 * yd));
      m1 = (float)(xd - a1 * yd);
      sm.tile[lane][i] = yi;
    }
    }
    __syncwarp();
    if (MODE != 2)
      for (int r = 0; r < nrows; r++) p.hp[(long long)(s0 + r) * p.hp_stride + ns::kHist + base + lane] = sm.tile[r][lane];
    __syncwarp();
  }
  p.state[(s0 + lane) * ns::kStateFloats + ns::kStHp] = m0 + m1;
}

int main() {
  const int n = 1024, nf = 24;
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
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
#define RUN(M, what)                                                 \
  kern<M><<<n / 32, 32>>>(p);                                          \
  cudaEventRecord(a); kern<M><<<n / 32, 32>>>(p); cudaEventRecord(b);  \
  cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, a, b);            \
  printf("mode %d (%s): %.3f ms  (%s)\n", M, what, ms, cudaGetErrorString(cudaGetLastError()));
  RUN(0, "highpass_body as shipped")
  RUN(1, "no history copies")
  RUN(2, "no history copies, no hp stores")
  RUN(3, "tile I/O only, no recursion")
  return 0;
}
