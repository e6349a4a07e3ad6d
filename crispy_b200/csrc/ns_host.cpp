// ns_host.cpp -- tables, model blobs, weight repacking (host only; see ns_host.h)
#include "ns_host.h"
#include "ns_rnn_tc5.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

namespace ns {

static const int kEband5ms[kBands] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  10, 12,
                                      14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100};

void make_tables(Tables &t) {
  memset(&t, 0, sizeof(t));
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int k = 0; k < 480; k++) {
    t.w480[k].x = (float)cosl(2 * pi * k / 480);
    t.w480[k].y = (float)-sinl(2 * pi * k / 480);
  }
  for (int k = 0; k <= 240; k++) {
    t.w960[k].x = (float)cosl(2 * pi * k / 960);
    t.w960[k].y = (float)-sinl(2 * pi * k / 960);
  }
  for (int i = 0; i < kFrame; i++) {
    const long double s = sinl(.5L * pi * (i + .5L) / kFrame);
    t.win[i] = (float)sinl(.5L * pi * s * s);
  }
  for (int i = 0; i < kBands; i++)
    for (int j = 0; j < kBands; j++) {
      long double v = cosl((i + .5L) * j * pi / kBands);
      if (j == 0) v *= sqrtl(.5L);
      t.dct[i * kBands + j] = (float)v;
    }
  for (int i = 0; i <= 200; i++) t.tansig[i] = (float)(floor(tanh(.04 * i) * 1e6 + .5) / 1e6);
  for (int i = 0; i < kBands; i++) t.eband[i] = kEband5ms[i] * 4;
  for (int i = 0; i < kBands - 1; i++) {
    const int lo = t.eband[i], n = t.eband[i + 1] - lo;
    for (int j = 0; j < n; j++) {
      t.bin_band[lo + j] = i;
      t.bin_frac[lo + j] = (float)j / (float)n;
    }
  }
}

// ---- synthetic weights: specified in DESIGN.md; an independent copy lives in the oracle ----------
static uint64_t splitmix64(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void fill_q(std::vector<int8_t> &dst, size_t n, uint64_t &s, int amp, int offset) {
  dst.resize(n);
  for (size_t i = 0; i < n; i++) {
    const uint64_t u = splitmix64(s);
    const int t = (int)(u & 0xFFFF) + (int)((u >> 16) & 0xFFFF) + (int)((u >> 32) & 0xFFFF) +
                  (int)((u >> 48) & 0xFFFF) - 131070;
    int w = (t * amp) / 131072 + offset;
    if (w > 127) w = 127;
    if (w < -127) w = -127;
    dst[i] = (int8_t)w;
  }
}

static void shape_default(Model &m) {
  m.input_dense.nb_inputs = 42, m.input_dense.nb_neurons = 24, m.input_dense.activation = 0;
  m.vad_gru.nb_inputs = 24, m.vad_gru.nb_neurons = 24, m.vad_gru.activation = 2;
  m.vad_output.nb_inputs = 24, m.vad_output.nb_neurons = 1, m.vad_output.activation = 1;
  m.noise_gru.nb_inputs = 90, m.noise_gru.nb_neurons = 48, m.noise_gru.activation = 2;
  m.denoise_gru.nb_inputs = 114, m.denoise_gru.nb_neurons = 96, m.denoise_gru.activation = 2;
  m.denoise_output.nb_inputs = 96, m.denoise_output.nb_neurons = 22, m.denoise_output.activation = 1;
}

void model_synthetic(Model &m, uint64_t seed) {
  shape_default(m);
  uint64_t s = seed ^ 0xC215B200C215B200ull;
  fill_q(m.input_dense.weights, 42 * 24, s, 80, 0);
  fill_q(m.input_dense.bias, 24, s, 80, 0);
  fill_q(m.vad_gru.input_weights, 24 * 72, s, 250, 0);
  fill_q(m.vad_gru.recurrent_weights, 24 * 72, s, 110, 0);
  fill_q(m.vad_gru.bias, 72, s, 100, 0);
  fill_q(m.vad_output.weights, 24, s, 500, 0);
  fill_q(m.vad_output.bias, 1, s, 40, 0);
  fill_q(m.noise_gru.input_weights, 90 * 144, s, 160, 0);
  fill_q(m.noise_gru.recurrent_weights, 48 * 144, s, 80, 0);
  fill_q(m.noise_gru.bias, 144, s, 100, 0);
  fill_q(m.denoise_gru.input_weights, 114 * 288, s, 160, 0);
  fill_q(m.denoise_gru.recurrent_weights, 96 * 288, s, 60, 0);
  fill_q(m.denoise_gru.bias, 288, s, 100, 0);
  fill_q(m.denoise_output.weights, 96 * 22, s, 400, 0);
  fill_q(m.denoise_output.bias, 22, s, 120, 40);
}

// ---- blobs ---------------------------------------------------------------------------------------
static const char kMagic[8] = {'C', 'R', 'N', 'S', 'M', 'D', 'L', '1'};
static const char kTextMagic[] = "rnnoise-nu model file version 1";

static void put_u32(std::vector<uint8_t> &b, uint32_t v) {
  for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
static void put_arr(std::vector<uint8_t> &b, const std::vector<int8_t> &a) {
  b.insert(b.end(), reinterpret_cast<const uint8_t *>(a.data()), reinterpret_cast<const uint8_t *>(a.data()) + a.size());
}
static void put_dense(std::vector<uint8_t> &b, const DenseLayer &l) {
  put_u32(b, 0), put_u32(b, (uint32_t)l.nb_inputs), put_u32(b, (uint32_t)l.nb_neurons), put_u32(b, (uint32_t)l.activation);
  put_arr(b, l.weights), put_arr(b, l.bias);
}
static void put_gru(std::vector<uint8_t> &b, const GruLayer &l) {
  put_u32(b, 1), put_u32(b, (uint32_t)l.nb_inputs), put_u32(b, (uint32_t)l.nb_neurons), put_u32(b, (uint32_t)l.activation);
  put_arr(b, l.input_weights), put_arr(b, l.recurrent_weights), put_arr(b, l.bias);
}
std::vector<uint8_t> model_to_bytes(const Model &m) {
  std::vector<uint8_t> b(kMagic, kMagic + 8);
  put_dense(b, m.input_dense);
  put_gru(b, m.vad_gru);
  put_dense(b, m.vad_output);
  put_gru(b, m.noise_gru);
  put_gru(b, m.denoise_gru);
  put_dense(b, m.denoise_output);
  return b;
}

struct Reader {
  const uint8_t *p;
  size_t len, off;
  bool ok;
  uint32_t u32() {
    if (off + 4 > len) {
      ok = false;
      return 0;
    }
    uint32_t v = (uint32_t)p[off] | ((uint32_t)p[off + 1] << 8) | ((uint32_t)p[off + 2] << 16) | ((uint32_t)p[off + 3] << 24);
    off += 4;
    return v;
  }
  void arr(std::vector<int8_t> &a, size_t n) {
    if (off + n > len) {
      ok = false;
      return;
    }
    a.assign(reinterpret_cast<const int8_t *>(p + off), reinterpret_cast<const int8_t *>(p + off) + n);
    off += n;
  }
};
static void get_dense(Reader &r, DenseLayer &l) {
  const uint32_t kind = r.u32(), in = r.u32(), out = r.u32(), act = r.u32();
  if (!r.ok || kind != 0 || (int)in != l.nb_inputs || (int)out != l.nb_neurons || act > 2) {
    r.ok = false;
    return;
  }
  l.activation = (int)act;
  r.arr(l.weights, (size_t)in * out);
  r.arr(l.bias, out);
}
static void get_gru(Reader &r, GruLayer &l) {
  const uint32_t kind = r.u32(), in = r.u32(), n = r.u32(), act = r.u32();
  if (!r.ok || kind != 1 || (int)in != l.nb_inputs || (int)n != l.nb_neurons || act > 2) {
    r.ok = false;
    return;
  }
  l.activation = (int)act;
  r.arr(l.input_weights, (size_t)in * 3 * n);
  r.arr(l.recurrent_weights, (size_t)n * 3 * n);
  r.arr(l.bias, (size_t)3 * n);
}

struct TextReader {
  const char *p, *end;
  bool ok;
  long next() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
    if (p >= end) {
      ok = false;
      return 0;
    }
    char *e = nullptr;
    const long v = strtol(p, &e, 10);
    if (e == p) {
      ok = false;
      return 0;
    }
    p = e;
    return v;
  }
  void arr(std::vector<int8_t> &a, size_t n) {
    a.resize(n);
    for (size_t i = 0; i < n && ok; i++) {
      long v = next();
      if (v > 127) v = 127;
      if (v < -128) v = -128;
      a[i] = (int8_t)v;
    }
  }
};
static void text_dense(TextReader &t, DenseLayer &l) {
  const long in = t.next(), out = t.next(), act = t.next();
  if (!t.ok || in != l.nb_inputs || out != l.nb_neurons || act < 0 || act > 2) {
    t.ok = false;
    return;
  }
  l.activation = (int)act;
  t.arr(l.weights, (size_t)in * out);
  t.arr(l.bias, (size_t)out);
}
static void text_gru(TextReader &t, GruLayer &l) {
  const long in = t.next(), n = t.next(), act = t.next();
  if (!t.ok || in != l.nb_inputs || n != l.nb_neurons || act < 0 || act > 2) {
    t.ok = false;
    return;
  }
  l.activation = (int)act;
  t.arr(l.input_weights, (size_t)in * 3 * n);
  t.arr(l.recurrent_weights, (size_t)n * 3 * n);
  t.arr(l.bias, (size_t)3 * n);
}

bool model_from_bytes(Model &m, const void *blob, size_t len, std::string &err) {
  shape_default(m);
  if (!blob) {
    err = "null model blob";
    return false;
  }
  if (len >= 8 && memcmp(blob, kMagic, 8) == 0) {
    Reader r{reinterpret_cast<const uint8_t *>(blob), len, 8, true};
    get_dense(r, m.input_dense);
    get_gru(r, m.vad_gru);
    get_dense(r, m.vad_output);
    get_gru(r, m.noise_gru);
    get_gru(r, m.denoise_gru);
    get_dense(r, m.denoise_output);
    if (!r.ok || r.off != len) {
      err = "malformed CRNSMDL1 model blob (layer shape/length mismatch)";
      return false;
    }
    return true;
  }
  const size_t tl = sizeof(kTextMagic) - 1;
  if (len >= tl && memcmp(blob, kTextMagic, tl) == 0) {
    TextReader t{reinterpret_cast<const char *>(blob) + tl, reinterpret_cast<const char *>(blob) + len, true};
    text_dense(t, m.input_dense);
    text_gru(t, m.vad_gru);
    text_gru(t, m.noise_gru);
    text_gru(t, m.denoise_gru);
    text_dense(t, m.denoise_output);
    text_dense(t, m.vad_output);
    if (!t.ok) {
      err = "malformed rnnoise-nu text model";
      return false;
    }
    return true;
  }
  err = "unknown model blob format (expected CRNSMDL1 or rnnoise-nu text)";
  return false;
}

// ---- repacking for the tensor-pipe recurrent core (ns_common.h "recurrent core (K4)") ---------------
static uint32_t bf16_bits(int v) {  // int8 value -> bf16 bit pattern (exact)
  float f = (float)v;
  uint32_t u;
  memcpy(&u, &f, 4);
  return u >> 16;
}

namespace {
enum Src { kSrcNone = 0, kSrcDense, kSrcVadH, kSrcFeat, kSrcNoiseH, kSrcDenH };
struct SrcIdx {
  Src src;
  int idx;
};
// which activation sits at input kk of virtual k-tile vkt (ns_common.h kKt*)
SrcIdx seg_src(int vkt, int kk) {
  if (vkt < kKtNH) {  // DV and DVR: [dense 24 | vad state (or r*h) 24]
    const int pos = (vkt % 3) * 16 + kk;
    return pos < 24 ? SrcIdx{kSrcDense, pos} : SrcIdx{kSrcVadH, pos - 24};
  }
  if (vkt < kKtDH) return SrcIdx{kSrcNoiseH, ((vkt - kKtNH) % 3) * 16 + kk};
  if (vkt < kKtResident) return SrcIdx{kSrcDenH, ((vkt - kKtDH) % 6) * 16 + kk};
  const int f = (vkt - kKtF) * 16 + kk;
  return f < kFeatures ? SrcIdx{kSrcFeat, f} : SrcIdx{kSrcNone, 0};
}
}  // namespace

// weight(src, idx, col) of one job; 0 where the job does not read that input / column
using WeightFn = std::function<int(SrcIdx, int col)>;

template <int... KT>
static std::vector<int> kt_vec(KtList<KT...>) {
  return std::vector<int>{KT...};
}

// B fragments + biases of product J, in the compile-time shape the kernel unrolls over (ns_common.h MmaShape)
template <int J>
static void add_mma_job(PackedRnn &out, int n_out, int act, const WeightFn &w, const std::function<int(int col)> &bias) {
  const std::vector<int> kts = kt_vec(typename MmaShape<J>::Kt{});
  const int nnt = MmaShape<J>::nnt;
  if ((int)out.words.size() != MmaOff<J>::w || (int)out.bias.size() != MmaOff<J>::b || (n_out + 7) / 8 != nnt) abort();
  out.hdr.activation[J] = act;
  for (size_t i = 0; i < kts.size(); i++)
    for (int nt = 0; nt < nnt; nt++)
      for (int lane = 0; lane < 32; lane++)
        for (int r = 0; r < 2; r++) {
          const int n = nt * 8 + lane / 4;
          uint32_t word = 0;
          for (int h = 0; h < 2; h++) {
            const int kk = (lane % 4) * 2 + 8 * r + h;
            const int v = (n < n_out) ? w(seg_src(kts[i], kk), n) : 0;
            word |= bf16_bits(v) << (16 * h);
          }
          out.words.push_back(word);
        }
  for (int col = 0; col < nnt * 8; col++) out.bias.push_back(col < n_out ? (float)bias(col) : 0.f);
}

void pack_rnn(const Model &m, PackedRnn &out) {
  out.words.clear();
  out.bias.clear();
  memset(&out.hdr, 0, sizeof(out.hdr));
  const DenseLayer &d0 = m.input_dense, &dv = m.vad_output, &dg = m.denoise_output;
  const GruLayer &gv = m.vad_gru, &gn = m.noise_gru, &gd = m.denoise_gru;
  auto gin = [](const GruLayer &g, int row, int col) { return (int)g.input_weights[(size_t)row * 3 * g.nb_neurons + col]; };
  auto grec = [](const GruLayer &g, int row, int col) { return (int)g.recurrent_weights[(size_t)row * 3 * g.nb_neurons + col]; };
  // input_dense: features -> 24
  add_mma_job<kJDense>(out, 24, d0.activation,
              [&](SrcIdx s, int col) { return s.src == kSrcFeat ? (int)d0.weights[(size_t)s.idx * 24 + col] : 0; },
              [&](int col) { return (int)d0.bias[col]; });
  // vad_gru: input dense(24), state 24
  auto vad_w = [&](int col0) {
    return [&, col0](SrcIdx s, int col) {
      if (s.src == kSrcDense) return gin(gv, s.idx, col0 + col);
      if (s.src == kSrcVadH) return grec(gv, s.idx, col0 + col);
      return 0;
    };
  };
  add_mma_job<kJVadZR>(out, 48, 1, vad_w(0), [&](int col) { return (int)gv.bias[col]; });
  add_mma_job<kJVadC>(out, 24, gv.activation, vad_w(48), [&](int col) { return (int)gv.bias[48 + col]; });
  // noise_gru: input [dense 24 | vad state 24 | features 42], state 48
  auto noise_w = [&](int col0) {
    return [&, col0](SrcIdx s, int col) {
      if (s.src == kSrcDense) return gin(gn, s.idx, col0 + col);
      if (s.src == kSrcVadH) return gin(gn, 24 + s.idx, col0 + col);
      if (s.src == kSrcFeat) return gin(gn, 48 + s.idx, col0 + col);
      if (s.src == kSrcNoiseH) return grec(gn, s.idx, col0 + col);
      return 0;
    };
  };
  add_mma_job<kJNoiseZR>(out, 96, 1, noise_w(0), [&](int col) { return (int)gn.bias[col]; });
  add_mma_job<kJNoiseC>(out, 48, gn.activation, noise_w(96),
              [&](int col) { return (int)gn.bias[96 + col]; });
  // denoise_gru: input [vad state 24 | noise state 48 | features 42], state 96.  DV k-tile 1 also
  // holds dense[16..24): zero weights.
  auto den_w = [&](int col0) {
    return [&, col0](SrcIdx s, int col) {
      if (s.src == kSrcVadH) return gin(gd, s.idx, col0 + col);
      if (s.src == kSrcNoiseH) return gin(gd, 24 + s.idx, col0 + col);
      if (s.src == kSrcFeat) return gin(gd, 72 + s.idx, col0 + col);
      if (s.src == kSrcDenH) return grec(gd, s.idx, col0 + col);
      return 0;
    };
  };
  add_mma_job<kJDenZR>(out, 192, 1, den_w(0),
              [&](int col) { return (int)gd.bias[col]; });
  add_mma_job<kJDenC>(out, 96, gd.activation, den_w(192),
              [&](int col) { return (int)gd.bias[192 + col]; });
  // denoise_output: denoise state -> 22 band gains
  add_mma_job<kJOut>(out, 22, dg.activation,
              [&](SrcIdx s, int col) { return s.src == kSrcDenH ? (int)dg.weights[(size_t)s.idx * 22 + col] : 0; },
              [&](int col) { return (int)dg.bias[col]; });
  // vad_output: vad state -> 1 (column 0 of one n-tile)
  add_mma_job<kJVadOut>(out, 1, dv.activation,
              [&](SrcIdx s, int col) { return (s.src == kSrcVadH && col == 0) ? (int)dv.weights[s.idx] : 0; },
              [&](int col) { return (int)dv.bias[col]; });
  out.hdr.n_words = (int32_t)out.words.size();
  out.hdr.n_bias = (int32_t)out.bias.size();
}

// ---- repacking for the tcgen05 recurrent core (ns_rnn_tc5.cuh): per round one weight block in the canonical K-major
// no-swizzle layout of a UMMA shared-memory descriptor, element (column n, input kk of k-tile i) at
//   i * n_cols * 32 + (kk / 8) * (n_cols / 8) * 128 + (n / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2   bytes,
// and the biases per (round, column).
void pack_rnn_tc5(const Model &m, std::vector<uint8_t> &w, std::vector<float> &bias) {
  using namespace tc5;
  const DenseLayer &d0 = m.input_dense, &dv = m.vad_output, &dg = m.denoise_output;
  const GruLayer &gv = m.vad_gru, &gn = m.noise_gru, &gd = m.denoise_gru;
  auto gin = [](const GruLayer &g, int row, int col) { return (int)g.input_weights[(size_t)row * 3 * g.nb_neurons + col]; };
  auto grec = [](const GruLayer &g, int row, int col) { return (int)g.recurrent_weights[(size_t)row * 3 * g.nb_neurons + col]; };
  auto in_range = [](int p, int lo, int n) { return p >= lo && p < lo + n; };
  // weight of the unit (layer, gate, unit) on the activation at A position p
  auto weight = [&](const ColInfo &c, int p) -> int {
    const int feat = p - kPosFeat;
    switch (c.layer) {
      case 0: return (feat >= 0 && feat < kFeatures) ? (int)d0.weights[(size_t)feat * 24 + c.unit] : 0;
      case 1: {
        const int col = 24 * c.gate + c.unit;
        if (c.gate < 2) {
          if (in_range(p, kPosDense, 24)) return gin(gv, p - kPosDense, col);
          if (in_range(p, kPosVadH, 24)) return grec(gv, p - kPosVadH, col);
        } else {
          if (in_range(p, kPosDense, 24)) return gin(gv, p - kPosDense, col);
          if (in_range(p, kPosVadR, 24)) return grec(gv, p - kPosVadR, col);
        }
        return 0;
      }
      case 2: {
        const int col = 48 * c.gate + c.unit;
        if (in_range(p, kPosDense, 24)) return gin(gn, p - kPosDense, col);
        if (in_range(p, kPosVadH, 24)) return gin(gn, 24 + p - kPosVadH, col);
        if (feat >= 0 && feat < kFeatures) return gin(gn, 48 + feat, col);
        if (c.gate < 2 && in_range(p, kPosNoiseH, 48)) return grec(gn, p - kPosNoiseH, col);
        if (c.gate == 2 && in_range(p, kPosNoiseR, 48)) return grec(gn, p - kPosNoiseR, col);
        return 0;
      }
      case 3: {
        const int col = 96 * c.gate + c.unit;
        if (in_range(p, kPosVadH, 24)) return gin(gd, p - kPosVadH, col);
        if (in_range(p, kPosNoiseH, 48)) return gin(gd, 24 + p - kPosNoiseH, col);
        if (feat >= 0 && feat < kFeatures) return gin(gd, 72 + feat, col);
        if (c.gate < 2 && in_range(p, kPosDenH, 96)) return grec(gd, p - kPosDenH, col);
        if (c.gate == 2 && in_range(p, kPosDenR, 96)) return grec(gd, p - kPosDenR, col);
        return 0;
      }
      case 4: return in_range(p, kPosDenH, 96) ? (int)dg.weights[(size_t)(p - kPosDenH) * 22 + c.unit] : 0;
      case 5: return in_range(p, kPosVadH, 24) ? (int)dv.weights[p - kPosVadH] : 0;
      default: return 0;
    }
  };
  auto unit_bias = [&](const ColInfo &c) -> int {
    switch (c.layer) {
      case 0: return d0.bias[c.unit];
      case 1: return gv.bias[24 * c.gate + c.unit];
      case 2: return gn.bias[48 * c.gate + c.unit];
      case 3: return gd.bias[96 * c.gate + c.unit];
      case 4: return dg.bias[c.unit];
      case 5: return dv.bias[0];
      default: return 0;
    }
  };
  w.assign(kWeightBytes, 0);
  bias.assign((size_t)kNumRounds * kBiasPerRound, 0.f);
  for (int r = 0; r < kNumRounds; r++) {
    const Round &R = kRounds[r];
    for (int n = 0; n < R.n; n++) {
      const ColInfo c = col_info(r, n);
      bias[(size_t)r * kBiasPerRound + n] = (float)unit_bias(c);
      for (int i = 0; i < R.n_kt; i++)
        for (int kk = 0; kk < 16; kk++) {
          const int v = weight(c, 16 * R.kt[i] + kk);
          const size_t off = (size_t)R.boff + (size_t)i * R.n * 32 + (size_t)(kk / 8) * (R.n / 8) * 128 + (size_t)(n / 8) * 128 +
                             (size_t)(n % 8) * 16 + (size_t)(kk % 8) * 2;
          const uint32_t bits = bf16_bits(v);
          w[off] = (uint8_t)bits;
          w[off + 1] = (uint8_t)(bits >> 8);
        }
    }
  }
}

}  // namespace ns
