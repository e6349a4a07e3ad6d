// ffma2.cu -- does sm_100a's packed FP32 pipe (FFMA2) help an unfused multiply-add chain?
// K1's exactness contract needs  s = fl(s + fl(x*y))  (two roundings).  Scalar: FMUL + FADD = 2 issue slots per MAC.
// Packed: p = fma.rn.f32x2(x, y, -0.0) (exactly fl(x*y)), s = fma.rn.f32x2(p, 1.0, s) (exactly fl(p + s)):
// 2 issue slots per TWO MACs.  Measures MAC/s per variant and checks the packed form bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int NACC = 8;  // independent chains per thread (pairs for the packed variants)

template <int MODE>
__global__ void __launch_bounds__(256) burn(float *out, int iters, float seed, float one_rt, float nz_rt) {
  const float x0 = seed + threadIdx.x * 1e-3f, y0 = 1e-6f * threadIdx.x;
  if (MODE == 0) {  // FFMA
    float a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = i;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < NACC; i++) a[i] = fmaf(a[(i + 1) % NACC], y0, a[i]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) {  // FMUL + FADD (unfused MAC)
    float a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = i;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < NACC; i++) a[i] = __fadd_rn(a[i], __fmul_rn(a[(i + 1) % NACC], y0));
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 2) {  // FFMA2 (fused, packed)
    u64 a[NACC / 2];
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) a[i] = pk(i, i + 1);
    const u64 yy = pk(y0, y0), xx = pk(x0, x0);
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < NACC / 2; i++) a[i] = fma2(a[(i + 1) % (NACC / 2)], yy, a[i]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) { float u, v; upk(a[i], u, v); s += u + v; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {  // unfused MAC on the packed pipe: p = fma2(x, y, -0), s = fma2(p, 1, s)
    u64 a[NACC / 2];
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) a[i] = pk(i, i + 1);
    const u64 yy = pk(y0, y0), one = pk(one_rt, one_rt), nz = pk(nz_rt, nz_rt);
    u64 xs[NACC / 2];
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) xs[i] = pk(x0 + 2 * i, x0 + 2 * i + 1);
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < NACC / 2; i++) a[i] = fma2(fma2(a[(i + 1) % (NACC / 2)], yy, nz), one, a[i]);
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) { float u, v; upk(a[i], u, v); s += u + v; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

// exactness: n random (x, y, s) triples incl. signed zeros, denormals, infinities
__global__ void exact(const float *x, const float *y, const float *s, uint32_t *bad, int n, float one_rt, float nz_rt) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i + 1 >= n) return;
  const float r0 = __fadd_rn(s[i], __fmul_rn(x[i], y[i])), r1 = __fadd_rn(s[i + 1], __fmul_rn(x[i + 1], y[i + 1]));
  const u64 p = fma2(pk(x[i], x[i + 1]), pk(y[i], y[i + 1]), pk(nz_rt, nz_rt));
  const u64 q = fma2(p, pk(one_rt, one_rt), pk(s[i], s[i + 1]));
  float q0, q1;
  upk(q, q0, q1);
  const bool nan0 = r0 != r0, nan1 = r1 != r1;
  if ((nan0 ? (q0 == q0) : (__float_as_uint(q0) != __float_as_uint(r0))) || (nan1 ? (q1 == q1) : (__float_as_uint(q1) != __float_as_uint(r1))))
    atomicAdd(bad, 1u);
}

template <int MODE>
double run(const char *name, float *d_out, int iters) {
  const int blocks = 148 * 8, threads = 256;
  burn<MODE><<<blocks, threads>>>(d_out, 16, 1.f, 1.f, -0.f);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    burn<MODE><<<blocks, threads>>>(d_out, iters, 1.f, 1.f, -0.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double macs = (double)blocks * threads * NACC * iters;
  printf("%-34s %8.3f ms  %8.2f T MAC/s\n", name, best, macs / best / 1e9);
  return macs / best / 1e9;
}

int main() {
  float *d_out;
  cudaMalloc(&d_out, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000;
  const double f = run<0>("FFMA (fused, scalar)", d_out, iters);
  run<1>("FMUL+FADD (unfused, scalar)", d_out, iters);
  run<2>("FFMA2 (fused, packed)", d_out, iters);
  run<3>("FFMA2 x2 (unfused MAC, packed)", d_out, iters);
  printf("fp32_burst_tflops_ffma %.2f\n", 2 * f);
  // exactness
  const int n = 1 << 22;
  float *hx = (float *)malloc(3 * n * sizeof(float)), *hy = hx + n, *hs = hy + n;
  srand(1);
  auto rnd = [] { uint32_t u = ((uint32_t)rand() << 16) ^ (uint32_t)rand() ^ ((uint32_t)rand() << 31); float f; memcpy(&f, &u, 4); return f; };
  for (int i = 0; i < n; i++) {
    const int c = i & 7;
    hx[i] = rnd(), hy[i] = rnd(), hs[i] = rnd();
    if (c == 1) hx[i] = 0.f, hs[i] = -0.f;
    if (c == 2) hx[i] = -0.f, hs[i] = 0.f;
    if (c == 3) { hx[i] = (float)(rand() % 65536 - 32768); hy[i] = (float)(rand() % 65536 - 32768) * 0.37f; hs[i] = (float)rand(); }
    if (c == 4) { hx[i] = 1e-20f * (rand() % 100); hy[i] = 1e-20f * (rand() % 100); hs[i] = 1e-39f * (rand() % 100); }
  }
  float *dx;
  uint32_t *dbad, hbad = 0;
  cudaMalloc(&dx, 3 * n * sizeof(float));
  cudaMalloc(&dbad, 4);
  cudaMemset(dbad, 0, 4);
  cudaMemcpy(dx, hx, 3 * n * sizeof(float), cudaMemcpyHostToDevice);
  exact<<<n / 2 / 256, 256>>>(dx, dx + n, dx + 2 * n, dbad, n, 1.f, -0.f);
  cudaMemcpy(&hbad, dbad, 4, cudaMemcpyDeviceToHost);
  printf("packed unfused MAC vs FMUL+FADD: %u mismatching pairs of %d  (%s)\n", hbad, n / 2, cudaGetErrorString(cudaGetLastError()));
  return hbad != 0;
}
