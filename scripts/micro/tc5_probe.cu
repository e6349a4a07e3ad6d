// tc5_probe.cu -- one tcgen05.mma (A from TMEM, B from shared memory through a K-major no-swizzle descriptor, D in TMEM)
// checked against the host: pins the layouts the tcgen05 recurrent core relies on before it is written.
//   D[128 x N] (+)= A[128 x 16] . B[16 x N],  bf16 x bf16 -> f32, two K-steps (accumulate on the second)
//   A: lane = row m, 32-bit column c holds (A[m][2c], A[m][2c+1])                        (tcgen05.st 32x32b)
//   B: element (n, k) at (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2 bytes   (canonical K-major)
//   D: lane = row m, column n (f32)                                                          (tcgen05.ld 32x32b)
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc5_probe tc5_probe.cu && ./tc5_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

constexpr int N = 48, KSTEPS = 2;
constexpr int kSbo = 128, kLbo = N / 8 * 128;  // 8-row groups back to back; the two K halves of a step N/8 groups apart
constexpr int kBStep = 2 * kLbo;               // bytes per K-step of B

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t bf16_bits(float x) { return __float_as_uint(x) >> 16; }  // exact for small ints

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, int *status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B -> shared memory, canonical layout
  for (int i = tid; i < KSTEPS * 16 * N; i += 128) {
    const int k = i / N, n = i % N, ks = k / 16, kk = k % 16;
    const int off = ks * kBStep + (n / 8) * kSbo + (kk / 8) * kLbo + (n % 8) * 16 + (kk % 8) * 2;
    *reinterpret_cast<uint16_t *>(smem + off) = (uint16_t)bf16_bits(B[k * N + n]);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of B visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  // A -> TMEM columns 48 .. 48 + 8 * KSTEPS of this thread's lane (row m = tid)
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  for (int ks = 0; ks < KSTEPS; ks++) {
    uint32_t r[8];
    for (int c = 0; c < 8; c++)
      r[c] = bf16_bits(A[tid * (16 * KSTEPS) + ks * 16 + 2 * c]) | (bf16_bits(A[tid * (16 * KSTEPS) + ks * 16 + 2 * c + 1]) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + 48 + 8 * ks),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    // instruction descriptor: D f32 (bits 4-5 = 1), A bf16 (7-9 = 1), B bf16 (10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < KSTEPS; ks++) {
      const uint32_t b_addr = smem_u32(smem) + ks * kBStep;
      const uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kLbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) | (1ull << 46);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tbase),
          "r"(tbase + 48 + 8 * ks), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everyone waits for the MMAs (bounded spin: a mistake must not hang the box)
  {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; spin++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
    if (!done) {
      if (tid == 0) *status = -1;
      return;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(lane_addr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int c = 0; c < 16; c++) D[tid * N + c0 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64));
  if (tid == 0) *status = 1;
  (void)lane;
}

int main() {
  const int K = 16 * KSTEPS;
  static float hA[128 * 16 * KSTEPS], hB[16 * KSTEPS * N], hD[128 * N], ref[128 * N];
  for (int m = 0; m < 128; m++)
    for (int k = 0; k < K; k++) hA[m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
  for (int k = 0; k < K; k++)
    for (int n = 0; n < N; n++) hB[k * N + n] = (float)((k * 5 + n * 13) % 17 - 8);
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      float s = 0;
      for (int k = 0; k < K; k++) s += hA[m * K + k] * hB[k * N + n];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  int *dS, hS = 0;
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dB, sizeof(hB));
  cudaMalloc(&dD, sizeof(hD));
  cudaMalloc(&dS, 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, sizeof(hD));
  cudaMemset(dS, 0, 4);
  probe<<<1, 128, KSTEPS * kBStep>>>(dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&hS, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 128 * N; i++) bad += hD[i] != ref[i];
  printf("tc5_probe: %s, status %d, %d of %d mismatches; D[0][0..3] = %g %g %g %g (ref %g %g %g %g); D[77][5] = %g (ref %g)\n",
         cudaGetErrorString(e), hS, bad, 128 * N, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3], hD[77 * N + 5], ref[77 * N + 5]);
  return bad != 0 || hS != 1;
}
