// tc5_rate.cu -- what one tcgen05.mma (A from tensor memory, B from shared memory, K-major no-swizzle, M = 128, K = 16,
// bf16 -> f32) costs as a function of N, and what a dependent round trip (issue -> commit -> mbarrier wait) costs:
// the two numbers that size the tcgen05 recurrent core (ns_rnn_tc5.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc5_rate tc5_rate.cu && ./tc5_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, int NMMA, bool SS>
__device__ void run(uint32_t tbase, unsigned char *smem, uint64_t *mbar, unsigned &parity, long long *out, int slot) {
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo = N / 8 * 128;
  const uint64_t bdesc = (uint64_t)((smem_u32(smem) & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  // A from shared memory for the SS form: 128 rows x 16 K, LBO = 128 / 8 * 128
  const uint64_t adesc = (uint64_t)(((smem_u32(smem) + 16384) & 0x3FFFFu) >> 4) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  constexpr int REPS = 20;
  long long best = 1ll << 60;
  for (int rep = 0; rep < REPS; rep++) {
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < NMMA; i++) {
      const uint32_t acc = i > 0;
      if (SS)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tbase + 256),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
      else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tbase + 256),
            "r"(tbase + 8 * (i & 15)), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; spin++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");
    const long long t1 = clock64();
    if (t1 - t0 < best) best = t1 - t0;
  }
  out[slot] = best;
}

__global__ void __launch_bounds__(128) rate(long long *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3F803F80u;  // bf16 ones
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  {
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 128; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_addr + c), "r"(0x3F803F80u));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    unsigned parity = 0;
    int s = 0;
#define BOTH(N)                                              \
  run<N, 1, false>(tbase, smem, &mbar, parity, out, s++);    \
  run<N, 64, false>(tbase, smem, &mbar, parity, out, s++);   \
  run<N, 1, true>(tbase, smem, &mbar, parity, out, s++);     \
  run<N, 64, true>(tbase, smem, &mbar, parity, out, s++);
    BOTH(16) BOTH(32) BOTH(48) BOTH(64) BOTH(96) BOTH(128) BOTH(192) BOTH(256)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
  long long *d, h[32];
  cudaMalloc(&d, sizeof(h));
  cudaMemset(d, 0, sizeof(h));
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  rate<<<1, 128, 65536>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("tc5_rate: %s\n", cudaGetErrorString(e));
  const int ns[8] = {16, 32, 48, 64, 96, 128, 192, 256};
  printf("%6s | %28s | %28s\n", "N", "A in TMEM: 1 MMA, 64 MMAs", "A in smem: 1 MMA, 64 MMAs");
  for (int i = 0; i < 8; i++)
    printf("%6d | round trip %5lld, %6.1f / MMA | round trip %5lld, %6.1f / MMA   (floor 128 N / 256 = %d)\n", ns[i], h[4 * i],
           (h[4 * i + 1] - h[4 * i]) / 63.0, h[4 * i + 2], (h[4 * i + 3] - h[4 * i + 2]) / 63.0, ns[i] / 2);
  return e != cudaSuccess;
}
