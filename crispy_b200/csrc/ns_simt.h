// ns_simt.h -- the handful of SIMT primitives the stream kernel uses, in two spellings:
//  * device build (nvcc, sm_100a): thin wrappers over threadIdx, named barriers (bar.sync id, n)
//    and warp shuffles;
//  * NS_HOST_EMU build (g++): one OS thread per CUDA thread, pthread barriers for bar.sync and a
//    per-warp exchange buffer for shuffles.  The emulation exists so the CPU test-suite can run the
//    kernel's exact control flow against the oracle without a GPU.  It is test plumbing: the
//    product library never contains it.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(NS_HOST_EMU)

#define NS_DEV __device__ __forceinline__
#define NS_DEV_NOINLINE __device__ __noinline__
namespace ns {
NS_DEV uint32_t f2u(float v) { return __float_as_uint(v); }
NS_DEV float u2f(uint32_t v) { return __uint_as_float(v); }
}  // namespace ns

namespace ns {
struct Simt {
  static NS_DEV int tid() { return (int)threadIdx.x; }
  static NS_DEV int cta() { return (int)blockIdx.x; }
  static NS_DEV void cta_sync() { __syncthreads(); }
  // bar.sync with an explicit id and thread count: only the `n` threads of one stream group meet.
  static NS_DEV void group_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
  }
  static NS_DEV void warp_sync() { __syncwarp(); }
  static NS_DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static NS_DEV float shfl_down(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
  static NS_DEV float shfl_up(float v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
  static NS_DEV float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static NS_DEV int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static NS_DEV int shfl_down(int v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
  static NS_DEV int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static NS_DEV double shfl_up(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
  static NS_DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static NS_DEV unsigned ballot(bool pred) { return __ballot_sync(0xffffffffu, pred); }
  static NS_DEV int atomic_add_shared(int *p, int v) { return atomicAdd(p, v); }
  static NS_DEV int n_ctas() { return (int)gridDim.x; }
  // 16-byte asynchronous global -> shared copy (LDGSTS), grouped with commit / wait<N pending groups>
  static NS_DEV void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
  }
  static NS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  template <int N>
  static NS_DEV void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
  }
  static NS_DEV void prefetch_l2(const void *gmem) { asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem)); }
  // read-only global loads through L1 (LDG.CONSTANT): the small tables every CTA reads in order
  static NS_DEV float ldg(const float *p) { return __ldg(p); }
  static NS_DEV int ldg(const int *p) { return __ldg(p); }
  static NS_DEV float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
  // ---- TMA 1-D bulk copies (cp.async.bulk, UBLKCP) completing on a shared-memory mbarrier: one thread moves a
  // whole tile with one instruction; consumers wait on the barrier's phase parity.  dst, src and bytes are
  // multiples of 16.
  static NS_DEV void mbar_init(uint64_t *bar, int count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  static NS_DEV void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
  }
  static NS_DEV void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned long long g = (unsigned long long)__cvta_generic_to_global(gmem_src);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(g),
                 "r"(bytes), "r"(a)
                 : "memory");
  }
  static NS_DEV void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NS_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NS_MBAR_DONE;\n"
        "bra NS_MBAR_WAIT;\n"
        "NS_MBAR_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
  }
  // orders this thread's earlier generic-proxy accesses to shared memory before its later async-proxy copies
  static NS_DEV void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  // shared-memory flags between the specialised warps of one CTA (no bar.sync): values only grow
  // release / acquire ordering between the warps of a CTA: fence.acq_rel.cta (MEMBAR.ALL.CTA).  __threadfence_block()
  // is the sequentially consistent fence (MEMBAR.SC.CTA), which the flag hand-overs below do not need.
  static NS_DEV void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
  static NS_DEV void flag_set(int *f, int v) {
    fence_cta();
    *reinterpret_cast<volatile int *>(f) = v;
  }
  static NS_DEV void flag_wait(const int *f, int v, bool relaxed) {
    while (*reinterpret_cast<const volatile int *>(f) < v) {
      if (relaxed) __nanosleep(200);
    }
    fence_cta();
  }
  // the same two without their fence, for a warp that waits on / raises several flags around ONE fence_cta()
  static NS_DEV void flag_store(int *f, int v) { *reinterpret_cast<volatile int *>(f) = v; }
  static NS_DEV void flag_poll(const int *f, int v, bool relaxed) {
    while (*reinterpret_cast<const volatile int *>(f) < v) {
      if (relaxed) __nanosleep(200);
    }
  }
  // two floats -> packed bf16 pair, round to nearest even; `lo` lands in the low halfword
  static NS_DEV uint32_t bf16x2_rn(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
  }
  // warp-wide tensor-pipe MMA (HMMA): D[16x8] += A[16x16] . B[16x8], bf16 inputs, f32 accumulate.
  // Fragment layouts are the PTX m16n8k16 ones (ns_common.h restates them).
  static NS_DEV void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};
}  // namespace ns

#else  // ---------------------------------------------------------------- host emulation

#include <pthread.h>
#include <sched.h>
#include <string.h>

#define NS_DEV inline
#define NS_DEV_NOINLINE inline
namespace ns {
inline uint32_t f2u(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  return u;
}
inline float u2f(uint32_t u) {
  float v;
  memcpy(&v, &u, 4);
  return v;
}
}  // namespace ns

namespace ns {
struct EmuWarp {
  pthread_barrier_t bar;
  uint64_t xch[32];
  uint32_t frag[32][6];  // mma emulation: every lane's A (4 words) and B (2 words) fragments
};
struct EmuCta {
  pthread_barrier_t cta_bar;
  pthread_barrier_t group_bar[16];
  EmuWarp *warps;
  int cta_index;
  int n_ctas;
};
struct EmuThread {
  EmuCta *cta;
  int tid;
};
extern thread_local EmuThread g_emu;

struct Simt {
  static int tid() { return g_emu.tid; }
  static int cta() { return g_emu.cta->cta_index; }
  static void cta_sync() { pthread_barrier_wait(&g_emu.cta->cta_bar); }
  static void group_sync(int id, int) { pthread_barrier_wait(&g_emu.cta->group_bar[id]); }
  static void warp_sync() { pthread_barrier_wait(&g_emu.cta->warps[g_emu.tid >> 5].bar); }
  template <class T>
  static T xchg(T v, int src_lane) {
    EmuWarp &w = g_emu.cta->warps[g_emu.tid >> 5];
    int lane = g_emu.tid & 31;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.xch[lane] = bits;
    pthread_barrier_wait(&w.bar);
    T r = v;
    if (src_lane >= 0 && src_lane < 32) memcpy(&r, &w.xch[src_lane], sizeof(T));
    pthread_barrier_wait(&w.bar);
    return r;
  }
  static int lane() { return g_emu.tid & 31; }
  static float shfl_xor(float v, int m) { return xchg(v, lane() ^ m); }
  static float shfl_down(float v, int d) { return xchg(v, lane() + d); }
  static float shfl_up(float v, int d) { return xchg(v, lane() - d); }
  static float shfl(float v, int src) { return xchg(v, src); }
  static int shfl_xor(int v, int m) { return xchg(v, lane() ^ m); }
  static int shfl_down(int v, int d) { return xchg(v, lane() + d); }
  static int shfl(int v, int src) { return xchg(v, src); }
  static double shfl_up(double v, int d) { return xchg(v, lane() - d); }
  static double shfl(double v, int src) { return xchg(v, src); }
  static unsigned ballot(bool pred) {
    EmuWarp &w = g_emu.cta->warps[g_emu.tid >> 5];
    w.xch[lane()] = pred ? 1u : 0u;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(w.xch[i] & 1u) << i;
    pthread_barrier_wait(&w.bar);
    return m;
  }
  static int atomic_add_shared(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
  static int n_ctas() { return g_emu.cta->n_ctas; }
  static void cp_async16(void *smem_dst, const void *gmem_src) { memcpy(smem_dst, gmem_src, 16); }
  static void cp_async_commit() {}
  template <int N>
  static void cp_async_wait() {}
  static uint32_t bf16x2_rn(float lo, float hi) {
    auto rn = [](float x) {
      const uint32_t u = f2u(x);
      return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
    };
    return rn(lo) | (rn(hi) << 16);
  }
  static void prefetch_l2(const void *) {}
  static float ldg(const float *p) { return *p; }
  static int ldg(const int *p) { return *p; }
  // bulk copies complete at once in the emulation; callers follow every mbar_wait with a group barrier
  static void mbar_init(uint64_t *, int) {}
  static void mbar_expect_tx(uint64_t *, unsigned) {}
  static void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *) { memcpy(smem_dst, gmem_src, bytes); }
  static void mbar_wait(uint64_t *, unsigned) {}
  static void fence_async_proxy() {}
  static void fence_cta() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
  static void flag_set(int *f, int v) { __atomic_store_n(f, v, __ATOMIC_SEQ_CST); }
  static void flag_wait(const int *f, int v, bool) {
    while (__atomic_load_n(f, __ATOMIC_SEQ_CST) < v) sched_yield();
  }
  static void flag_store(int *f, int v) { __atomic_store_n(f, v, __ATOMIC_SEQ_CST); }
  static void flag_poll(const int *f, int v, bool relaxed) { flag_wait(f, v, relaxed); }
  // mma.sync.m16n8k16 (bf16 x bf16 -> f32) emulated from the lanes' fragments: A element (row, k)
  // sits in lane (row%8)*4 + (k%8)/2, word row/8 + 2*(k/8), halfword k%2; B element (k, col) in
  // lane col*4 + (k%8)/2, word k/8, halfword k%2; this lane owns D (lane/4 [+8], 2*(lane%4) [+1]).
  static void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    EmuWarp &w = g_emu.cta->warps[g_emu.tid >> 5];
    const int ln = lane();
    for (int i = 0; i < 4; i++) w.frag[ln][i] = a[i];
    for (int i = 0; i < 2; i++) w.frag[ln][4 + i] = b[i];
    pthread_barrier_wait(&w.bar);
    auto bf = [](uint32_t word, int half) { return u2f(((word >> (16 * half)) & 0xFFFFu) << 16); };
    for (int e = 0; e < 4; e++) {
      const int row = (ln >> 2) + 8 * (e >> 1), col = 2 * (ln & 3) + (e & 1);
      double sum = 0.0;
      for (int k = 0; k < 16; k++) {
        const float av = bf(w.frag[(row & 7) * 4 + ((k & 7) >> 1)][(row >> 3) + 2 * (k >> 3)], k & 1);
        const float bv = bf(w.frag[col * 4 + ((k & 7) >> 1)][4 + (k >> 3)], k & 1);
        sum += (double)av * (double)bv;
      }
      d[e] = (float)((double)d[e] + sum);
    }
    pthread_barrier_wait(&w.bar);
  }
};
}  // namespace ns

#endif
