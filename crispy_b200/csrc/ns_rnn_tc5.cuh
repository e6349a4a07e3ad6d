// ns_rnn_tc5.cuh -- K4 on the fifth-generation tensor cores: the recurrent core (a14: dense -> VAD GRU -> noise GRU ->
// denoise GRU -> dense) of 128 streams per CTA as tcgen05.mma (UTCHMMA) products with every operand where Blackwell
// wants it:
//   * A (activations, 128 streams x K): TENSOR MEMORY.  Lane = stream, 32-bit column c holds the bf16 pair of inputs
//     (2c, 2c + 1); two planes, hi at columns [0, 216) and lo = bf16(x - hi) at [216, 432), so x = hi + lo to 2^-17.
//     The epilogue threads write them with tcgen05.st and the MMA reads them in place (the .ts form: A from TMEM).
//   * B (weights, int8 exact in bf16): SHARED MEMORY, 187.5 KB resident for the whole launch, one block per product in
//     the canonical K-major no-swizzle layout a UMMA shared-memory descriptor addresses (8 x 16 B core matrices).
//   * D (f32 accumulators, 128 x N): TENSOR MEMORY columns [432, 512), read back with tcgen05.ld for the epilogue.
// One frame step is thirteen rounds (the eight products of the network, the wide ones cut into chunks of <= 64
// columns because A fills most of tensor memory): one thread issues the round's MMAs (hi and lo plane into the same
// accumulator) and commits them to an mbarrier; the 256 epilogue threads wait on it, load their half of the columns of
// their stream's row, apply bias / table tanh / GRU algebra in registers (the GRU states themselves stay in f32
// registers) and store the next products' operands straight back into tensor memory.
// 1,024 streams take 8 CTAs instead of the 64 the warp-level mma.sync core occupies (ns_pipe.cuh rnn_body).
//
// Device-only (no host emulation: tensor memory has no CPU stand-in); parity is checked on the GPU against the oracle
// and against the mma.sync core (tests/test_gpu_parity.py).
#pragma once
#include "ns_common.h"

namespace ns {
namespace tc5 {

constexpr int kStreams = 128;
constexpr int kThreads = 256;   // thread = (stream row m = tid & 127, column half hs = tid >> 7)
constexpr int kColLo = 216;     // first column of the lo plane
constexpr int kColD = 432;      // accumulator columns
constexpr int kTmemCols = 512;
constexpr int kNumRounds = 13;
constexpr int kMaxKt = 14;

// positions of the activation vectors in the A element space (k-tile v = elements 16 v .. 16 v + 15; the same tiling as
// the mma.sync core: ns_common.h kKt*)
constexpr int kPosDense = 0, kPosVadH = 24, kPosDense2 = 48, kPosVadR = 72, kPosNoiseH = 96, kPosNoiseR = 144,
              kPosDenH = 192, kPosDenR = 288, kPosFeat = 384;

struct Round {
  int n_kt;
  int kt[kMaxKt];
  int n;      // columns of the round (a multiple of 16): half 0 takes columns [0, n/2), half 1 the rest
  int boff;   // byte offset of the round's weight block
};
// weight block of a round: [k-tile][K half][n / 8][8 columns][8 inputs] bf16 = n * 32 bytes per k-tile
constexpr Round kRounds[kNumRounds] = {
    {3, {24, 25, 26}, 32, 0},                                              // 0 input_dense
    {3, {0, 1, 2}, 48, 3072},                                              // 1 vad z | r
    {3, {3, 4, 5}, 32, 7680},                                              // 2 vad candidate
    {9, {0, 1, 2, 24, 25, 26, 6, 7, 8}, 48, 10752},                        // 3 noise z
    {9, {0, 1, 2, 24, 25, 26, 6, 7, 8}, 64, 24576},                        // 4 noise r (+ vad_output)
    {9, {0, 1, 2, 24, 25, 26, 9, 10, 11}, 48, 43008},                      // 5 noise candidate
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 12, 13, 14, 15, 16, 17}, 48, 56832},  // 6 denoise z, neurons 48 h + 0..23
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 12, 13, 14, 15, 16, 17}, 48, 78336},  // 7 denoise z, neurons 48 h + 24..47
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 12, 13, 14, 15, 16, 17}, 48, 99840},  // 8 denoise r, neurons 48 h + 0..23
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 12, 13, 14, 15, 16, 17}, 48, 121344}, // 9 denoise r, neurons 48 h + 24..47
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 18, 19, 20, 21, 22, 23}, 48, 142848}, // 10 denoise candidate, 48 h + 0..23
    {14, {1, 2, 6, 7, 8, 24, 25, 26, 18, 19, 20, 21, 22, 23}, 48, 164352}, // 11 denoise candidate, 48 h + 24..47
    {6, {12, 13, 14, 15, 16, 17}, 32, 185856},                             // 12 denoise_output
};
constexpr int kWeightBytes = 185856 + 6 * 32 * 32;  // 192,000
constexpr int kBiasPerRound = 64;

// what column `col` of round `r` computes: layer 0 dense, 1 vad GRU, 2 noise GRU, 3 denoise GRU, 4 output, 5 vad_output;
// gate 0 z, 1 r, 2 candidate; unit index; layer -1 = padding.  Shared by the host packer and (implicitly) the epilogues.
struct ColInfo {
  int layer, gate, unit;
};
inline ColInfo col_info(int r, int col) {
  const int half_n = kRounds[r].n / 2, h = col / half_n, j = col % half_n;
  switch (r) {
    case 0: return j < 12 ? ColInfo{0, 0, 12 * h + j} : ColInfo{-1, 0, 0};
    case 1: return j < 12 ? ColInfo{1, 0, 12 * h + j} : ColInfo{1, 1, 12 * h + j - 12};
    case 2: return j < 12 ? ColInfo{1, 2, 12 * h + j} : ColInfo{-1, 0, 0};
    case 3: return ColInfo{2, 0, 24 * h + j};
    case 4: return j < 24 ? ColInfo{2, 1, 24 * h + j} : ((h == 0 && j == 24) ? ColInfo{5, 0, 0} : ColInfo{-1, 0, 0});
    case 5: return ColInfo{2, 2, 24 * h + j};
    case 6: return ColInfo{3, 0, 48 * h + j};
    case 7: return ColInfo{3, 0, 48 * h + 24 + j};
    case 8: return ColInfo{3, 1, 48 * h + j};
    case 9: return ColInfo{3, 1, 48 * h + 24 + j};
    case 10: return ColInfo{3, 2, 48 * h + j};
    case 11: return ColInfo{3, 2, 48 * h + 24 + j};
    default: return j < 11 ? ColInfo{4, 0, 11 * h + j} : ColInfo{-1, 0, 0};
  }
}

#if defined(__CUDACC__) && !defined(NS_HOST_EMU)

struct Smem {
  alignas(128) unsigned char w[kWeightBytes];
  float bias[kNumRounds * kBiasPerRound];
  float tansig[204];
  alignas(8) unsigned long long mbar;
  unsigned tmem_base;
};
static_assert(sizeof(Smem) <= 227 * 1024, "tcgen05 recurrent core exceeds the shared memory of a CTA");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int N>
struct TmemIo;
template <>
struct TmemIo<2> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v[0]), "r"(v[1]));
  }
};
template <>
struct TmemIo<4> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
  }
};
template <>
struct TmemIo<8> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
  }
};
// N consecutive columns in pieces of P columns (every piece starts at a multiple of P columns: callers choose P so)
template <int N, int P>
__device__ __forceinline__ void tmem_ld(unsigned a, unsigned *v) {
  static_assert(N % P == 0, "whole pieces");
#pragma unroll
  for (int i = 0; i < N; i += P) TmemIo<P>::ld(a + i, v + i);
}
template <int N, int P>
__device__ __forceinline__ void tmem_st(unsigned a, const unsigned *v) {
  static_assert(N % P == 0, "whole pieces");
#pragma unroll
  for (int i = 0; i < N; i += P) TmemIo<P>::st(a + i, v + i);
}

// NV values of this thread's stream -> hi / lo bf16 pairs at A position `pos` (pos / 2 a multiple of P) of both planes
template <int NV, int P>
__device__ __forceinline__ void store_act(unsigned lane_base, int pos, const float *v) {
  unsigned hi[NV / 2], lo[NV / 2];
#pragma unroll
  for (int i = 0; i < NV / 2; i++) bf16_split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  tmem_st<NV / 2, P>(lane_base + (unsigned)(pos >> 1), hi);
  tmem_st<NV / 2, P>(lane_base + (unsigned)(kColLo + (pos >> 1)), lo);
}

__device__ __forceinline__ void rnn_tc5_body(const Params &p, Smem &s) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hs = warp >> 2;        // TMEM lane quarter of this warp, column half
  const int m = 32 * q + lane;                   // stream row within the CTA
  const int stream = blockIdx.x * kStreams + m;
  const bool live = stream < p.n_streams;
  // ---- weights, biases, tables -> shared memory; tensor memory; the mbarrier
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.rnn_words);
    uint4 *dst = reinterpret_cast<uint4 *>(s.w);
    for (int i = tid; i < kWeightBytes / 16; i += kThreads) dst[i] = src[i];
    for (int i = tid; i < kNumRounds * kBiasPerRound; i += kThreads) s.bias[i] = p.rnn_bias[i];
    for (int i = tid; i < 204; i += kThreads) s.tansig[i] = p.tables->tansig[i];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weights, written through the generic proxy, for the MMAs
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tbase = s.tmem_base;
  const unsigned lane_base = tbase + ((unsigned)(32 * q) << 16);
  const float *tab = s.tansig;
  const RnnHeader &H = *p.rnn_hdr;
  const int act_dense = H.activation[kJDense], act_vad = H.activation[kJVadC], act_noise = H.activation[kJNoiseC],
            act_den = H.activation[kJDenC], act_out = H.activation[kJOut], act_vadout = H.activation[kJVadOut];
  // ---- recurrent state of this thread's units: vad 12 hs + 0..11, noise 24 hs + 0..23, denoise 48 hs + 0..47, lastg 11 hs + 0..10
  float hv[12], hn[24], hd[48], lastg[11];
  {
    const float *st = p.state + (long long)(live ? stream : 0) * kStateFloats;
#pragma unroll
    for (int j = 0; j < 12; j++) hv[j] = live ? st[kStHVad + 12 * hs + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 24; j++) hn[j] = live ? st[kStHNoise + 24 * hs + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 48; j++) hd[j] = live ? st[kStHDen + 48 * hs + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 11; j++) lastg[j] = live ? st[kStLastG + 11 * hs + j] : 0.f;
  }
  store_act<12, 2>(lane_base, kPosVadH + 12 * hs, hv);
  store_act<24, 4>(lane_base, kPosNoiseH + 24 * hs, hn);
  store_act<48, 8>(lane_base, kPosDenH + 48 * hs, hd);
  // feature words of (16-stream group, frame): K3b's layout (ns_pipe.cuh features_body): pair q of row r at word widx
  const int row16 = m & 15;
  const uint32_t *fq_src = p.featq + (long long)(stream >> 4) * p.chunk_cap * kFeatBlockWords;
  unsigned parity = 0;
  unsigned err = 0;
  auto round_mma = [&](auto ri_tag) {  // one thread: every k-tile of the round, hi then lo plane, into the accumulator columns
    constexpr Round R = kRounds[decltype(ri_tag)::value];
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(R.n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const unsigned lbo = (unsigned)(R.n / 8 * 128);
    const unsigned b0 = smem_u32(s.w) + (unsigned)R.boff;
    unsigned first = 1;
#pragma unroll
    for (int i = 0; i < R.n_kt; i++) {
      const unsigned long long bdesc = (unsigned long long)(((b0 + (unsigned)(i * R.n * 32)) & 0x3FFFFu) >> 4) |
                                       ((unsigned long long)(lbo >> 4) << 16) | ((unsigned long long)(128 >> 4) << 32) | (1ull << 46);
#pragma unroll
      for (int plane = 0; plane < 2; plane++) {
        const unsigned a = tbase + (unsigned)(plane * kColLo + 8 * R.kt[i]);
        const unsigned acc = first ? 0u : 1u;
        first = 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tbase + kColD),
            "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.mbar)) : "memory");
  };
#ifdef NS_TC5_CLOCKS
  long long tk_issue = 0, tk_wait = 0, tk_epi = 0, tk_mark = clock64();
#define NS_TC5_TICK(acc) do { const long long now_ = clock64(); acc += now_ - tk_mark; tk_mark = now_; } while (0)
#else
#define NS_TC5_TICK(acc) do {} while (0)
#endif
  auto wait_round = [&]() {
    NS_TC5_TICK(tk_issue);
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; spin++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&s.mbar)), "r"(parity)
                   : "memory");
    if (!done) err = 1;  // a wedged tensor pipe must not hang the device: results are garbage, the host sees the flag
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");
    NS_TC5_TICK(tk_wait);
  };
  // the accumulators of this thread's half of the round's columns, scaled: x = (acc + bias) / 256
  auto load_pre = [&](int r, auto nc_tag, float *x) {
    constexpr int NC = decltype(nc_tag)::value;
    unsigned v[NC];
    tmem_ld<NC, 8>(lane_base + (unsigned)(kColD + hs * NC), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float *b = s.bias + r * kBiasPerRound + hs * NC;
#pragma unroll
    for (int j = 0; j < NC; j++) x[j] = (__uint_as_float(v[j]) + b[j]) * (1.f / 256);
  };
  auto end_round = [&]() {  // operands stored, accumulators read: the next round may issue
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    NS_TC5_TICK(tk_epi);
  };

  for (int t = 0; t < p.n_frames; t++) {
    // features of this frame -> tensor memory (half 0 stores the hi plane, half 1 the lo plane)
    bool sil = true;
    {
      unsigned fw[24];
      const uint32_t *blk = fq_src + (long long)t * kFeatBlockWords + hs * (kFeatKt * kKtWords);
#pragma unroll
      for (int qq = 0; qq < 24; qq++) {
        const int widx = ((qq >> 3) * 32 + (row16 & 7) * 4 + (qq & 3)) * 4 + (row16 >> 3) + 2 * ((qq & 7) >> 2);
        fw[qq] = live ? __ldg(blk + widx) : 0u;
      }
      if (live) sil = __ldg(fq_src + (long long)t * kFeatBlockWords + 2 * kFeatKt * kKtWords + row16) != 0u;
      tmem_st<24, 8>(lane_base + (unsigned)(hs * kColLo + (kPosFeat >> 1)), fw);
    }
    end_round();
    float vad = 0.f;
    float x[32];
    // 0: input_dense -> dense (both copies)
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<0>{});
    }
    wait_round();
    load_pre(0, IntC<16>{}, x);
    {
      float y[12];
#pragma unroll
      for (int j = 0; j < 12; j++) y[j] = activate(tab, act_dense, x[j]);
      store_act<12, 2>(lane_base, kPosDense + 12 * hs, y);
      store_act<12, 2>(lane_base, kPosDense2 + 12 * hs, y);
    }
    end_round();
    // 1: vad z | r
    float zv[12];
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<1>{});
    }
    wait_round();
    load_pre(1, IntC<24>{}, x);
    {
      float rh[12];
#pragma unroll
      for (int j = 0; j < 12; j++) {
        zv[j] = sigmoid_approx(tab, x[j]);
        rh[j] = hv[j] * sigmoid_approx(tab, x[12 + j]);
      }
      store_act<12, 2>(lane_base, kPosVadR + 12 * hs, rh);
    }
    end_round();
    // 2: vad candidate -> vad state
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<2>{});
    }
    wait_round();
    load_pre(2, IntC<16>{}, x);
#pragma unroll
    for (int j = 0; j < 12; j++) {
      const float c = activate(tab, act_vad, x[j]);
      const float hnew = zv[j] * hv[j] + (1.f - zv[j]) * c;
      if (!sil) hv[j] = hnew;
    }
    store_act<12, 2>(lane_base, kPosVadH + 12 * hs, hv);
    end_round();
    // 3: noise z
    float zn[24];
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<3>{});
    }
    wait_round();
    load_pre(3, IntC<24>{}, x);
#pragma unroll
    for (int j = 0; j < 24; j++) zn[j] = sigmoid_approx(tab, x[j]);
    end_round();
    // 4: noise r (+ vad_output in column 24 of half 0)
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<4>{});
    }
    wait_round();
    load_pre(4, IntC<32>{}, x);
    {
      float rh[24];
#pragma unroll
      for (int j = 0; j < 24; j++) rh[j] = hn[j] * sigmoid_approx(tab, x[j]);
      store_act<24, 4>(lane_base, kPosNoiseR + 24 * hs, rh);
      if (hs == 0) vad = activate(tab, act_vadout, x[24]);
    }
    end_round();
    // 5: noise candidate -> noise state
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<5>{});
    }
    wait_round();
    load_pre(5, IntC<24>{}, x);
#pragma unroll
    for (int j = 0; j < 24; j++) {
      const float c = activate(tab, act_noise, x[j]);
      const float hnew = zn[j] * hn[j] + (1.f - zn[j]) * c;
      if (!sil) hn[j] = hnew;
    }
    store_act<24, 4>(lane_base, kPosNoiseH + 24 * hs, hn);
    end_round();
    // 6, 7: denoise z
    float zd[48];
    auto den_z = [&](auto c_tag) {
      constexpr int c = decltype(c_tag)::value;
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        round_mma(IntC<6 + c>{});
      }
      wait_round();
      load_pre(6 + c, IntC<24>{}, x);
#pragma unroll
      for (int j = 0; j < 24; j++) zd[24 * c + j] = sigmoid_approx(tab, x[j]);
      end_round();
    };
    den_z(IntC<0>{});
    den_z(IntC<1>{});
    // 8, 9: denoise r -> r * h
    auto den_r = [&](auto c_tag) {
      constexpr int c = decltype(c_tag)::value;
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        round_mma(IntC<8 + c>{});
      }
      wait_round();
      load_pre(8 + c, IntC<24>{}, x);
      float rh[24];
#pragma unroll
      for (int j = 0; j < 24; j++) rh[j] = hd[24 * c + j] * sigmoid_approx(tab, x[j]);
      store_act<24, 4>(lane_base, kPosDenR + 48 * hs + 24 * c, rh);
      end_round();
    };
    den_r(IntC<0>{});
    den_r(IntC<1>{});
    // 10, 11: denoise candidate -> denoise state (the r rounds above read the old state: it is replaced only here)
    auto den_c = [&](auto c_tag) {
      constexpr int c = decltype(c_tag)::value;
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        round_mma(IntC<10 + c>{});
      }
      wait_round();
      load_pre(10 + c, IntC<24>{}, x);
#pragma unroll
      for (int j = 0; j < 24; j++) {
        const float cc = activate(tab, act_den, x[j]);
        const float hnew = zd[24 * c + j] * hd[24 * c + j] + (1.f - zd[24 * c + j]) * cc;
        if (!sil) hd[24 * c + j] = hnew;
      }
      store_act<24, 4>(lane_base, kPosDenH + 48 * hs + 24 * c, hd + 24 * c);  // the second chunk's product reads r * h, not the state
      end_round();
    };
    den_c(IntC<0>{});
    den_c(IntC<1>{});
    // 12: denoise_output -> band gains; g = max(g, 0.6 lastg)
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      round_mma(IntC<12>{});
    }
    wait_round();
    load_pre(12, IntC<16>{}, x);
    if (live) {
      float *rec = p.rec + ((long long)stream * p.chunk_cap + t) * kRecFloats;
#pragma unroll
      for (int j = 0; j < 11; j++) {
        float graw = 0.f, g = 0.f;
        if (!sil) {
          graw = activate(tab, act_out, x[j]);
          g = fmaxf(graw, .6f * lastg[j]);
          lastg[j] = g;
        }
        rec[kRecGRaw + 11 * hs + j] = graw;
        rec[kRecG + 11 * hs + j] = g;
      }
      if (hs == 0) {
        const float v = sil ? 0.f : vad;
        rec[kRecVad] = v;
        if (p.vad) p.vad[(long long)stream * p.vad_stride + p.frame0 + t] = v;
      }
    }
    // the next frame's feature store and its end_round() order this round's accumulator reads before the next MMAs
  }
  if (live) {
    float *st = p.state + (long long)stream * kStateFloats;
#pragma unroll
    for (int j = 0; j < 12; j++) st[kStHVad + 12 * hs + j] = hv[j];
#pragma unroll
    for (int j = 0; j < 24; j++) st[kStHNoise + 24 * hs + j] = hn[j];
#pragma unroll
    for (int j = 0; j < 48; j++) st[kStHDen + 48 * hs + j] = hd[j];
#pragma unroll
    for (int j = 0; j < 11; j++) st[kStLastG + 11 * hs + j] = lastg[j];
    if (err && tid == 0) reinterpret_cast<int *>(st)[kStFrameCount] = -1;  // poison the informational frame counter
  }
#ifdef NS_TC5_CLOCKS
  if (tid == 0 && blockIdx.x == 0)
    printf("tc5 clocks per frame (thread 0): MMA issue %lld, wait for MMAs %lld, epilogue + stores + barrier %lld cycles\n",
           tk_issue / p.n_frames, tk_wait / p.n_frames, tk_epi / p.n_frames);
#endif
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kTmemCols));
}

#endif  // __CUDACC__

}  // namespace tc5
}  // namespace ns
