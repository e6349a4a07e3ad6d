// ns_common.h -- constants, table and parameter structs shared by the sm_100a kernels, the host
// library and the host-side SIMT emulation used by the CPU tests.
//
// Domain vocabulary follows the reference path (nnnoiseless::DenoiseState behind
// /root/reference/src-tauri/src/audio.rs:268): frames of 480 samples at 48 kHz, a 960-sample
// analysis window, 481 frequency bins, 22 bands, 42 features, a 1728-sample pitch buffer.
#pragma once
#include <stdint.h>

namespace ns {

constexpr int kFrame = 480;
constexpr int kWindow = 960;
constexpr int kFreq = 481;
constexpr int kBands = 22;
constexpr int kFeatures = 42;
constexpr int kCepsMem = 8;
constexpr int kDeltaCeps = 6;
constexpr int kPitchMin = 60;
constexpr int kPitchMax = 768;
constexpr int kPitchBuf = 1728;
constexpr int kGroupThreads = 128;   // threads cooperating on one frame's spectra

// ---- pipeline geometry ---------------------------------------------------------------------------
// A call is cut into chunks of <= chunk_cap frames.  Per chunk and stream the engine keeps, in HBM:
//   hp  : the high-passed signal, kHist samples of history + chunk*480 new samples (f32)
//   tab : per frame, the pitch candidate table remove_doubling's serial decision walks (kTabWords)
//   rec : per frame, the small record passed between phases (kRecFloats)
//   spec: per frame, the spectra X and P (K3 -> K5)
constexpr int kHist = 1440;          // >= 1248 (pitch_buf history) and a multiple of 480
constexpr int kLpLen = 864;          // pitch_buf downsampled by 2
constexpr int kLpStride = 872;       // padded row of the per-frame downsampled buffers
constexpr int kMaxK = 15;            // remove_doubling examines T0/k for k = 1..15
constexpr int kTabWords = 48;
//   word 0      : T0 | n_k << 16      (n_k = candidates present, k = 1..n_k)
//   word 1      : unused
//   word 2+3(k-1): T_k | pitch_index_k << 16 ; g_k (f32 bits) ; pitch_gain_k (f32 bits)
constexpr int kRecFloats = 144;
constexpr int kRecPitchIndex = 0;    // int bits
constexpr int kRecSilence = 1;       // int bits
constexpr int kRecPitchGain = 2;
constexpr int kRecVad = 3;
constexpr int kRecExp = 4;           // 22 (normalised band correlation)
constexpr int kRecCeps = 26;         // 22 (DCT of log band energy, offsets applied)
constexpr int kRecTail = 48;         // 7: features[34..40]
constexpr int kRecGRaw = 56;         // 22: band gains as the RNN emits them (the pitch filter uses these)
constexpr int kRecG = 78;            // 22: band gains after the 0.6*lastg smoothing
constexpr int kRecEx = 100;          // 22: band energy of the frame
constexpr int kRecEp = 122;          // 22: band energy of the pitch-lagged window
static_assert(kRecEp + kBands <= kRecFloats, "record overflow");
constexpr int kSpecStride = 482;     // complex bins per stored spectrum (481 used); per frame: X then P

struct alignas(8) cf {
  float x, y;
};
struct alignas(16) f4 {
  float x, y, z, w;
};

// ---- per-stream persistent state (the fields of nnnoiseless::DenoiseState), one block of
// kStateFloats f32 per stream in HBM.
constexpr int kStHist = 0;                        // 1440: last high-passed samples (pitch_buf + analysis_mem)
constexpr int kStSynth = kStHist + kHist;         // 2x480: synthesis_mem, double buffered: a chunk reads copy
                                                  //        synth_sel and writes the other (its runs are concurrent)
constexpr int kStCeps = kStSynth + 2 * kFrame;    // 176 : cepstral_mem[8][22]
constexpr int kStLastG = kStCeps + 176;           // 22  : lastg
constexpr int kStHVad = kStLastG + 22;            // 24  : vad_gru_state
constexpr int kStHNoise = kStHVad + 24;           // 48  : noise_gru_state
constexpr int kStHDen = kStHNoise + 48;           // 96  : denoise_gru_state
constexpr int kStHp = kStHDen + 96;               // 2   : mem_hp_x (f32, as upstream)
constexpr int kStLastGain = kStHp + 2;            // 1
constexpr int kStLastPeriod = kStLastGain + 1;    // 1 (int bits)
constexpr int kStMemId = kStLastPeriod + 1;       // 1 (int bits)
constexpr int kStFrameCount = kStMemId + 1;       // 1 (int bits, informational)
constexpr int kStateFloats = 2784;                // padded to a multiple of 32 floats
static_assert(kStFrameCount + 1 <= kStateFloats, "state layout overflow");

// ---- per-frame debug taps (optional), mirrors oracle rno_debug
constexpr int kDbgFeatures = 0;
constexpr int kDbgGains = 42;
constexpr int kDbgEx = 64;
constexpr int kDbgEp = 86;
constexpr int kDbgExp = 108;
constexpr int kDbgPitchGain = 130;
constexpr int kDbgVad = 131;
constexpr int kDbgPitchIndex = 132;  // stored as float
constexpr int kDbgSilence = 133;
constexpr int kDbgFloats = 136;

// ---- constant tables, generated on the host in f64 and copied to shared memory per CTA
struct Tables {
  cf w480[480];          // exp(-2 pi i k/480)
  cf w960[244];          // exp(-2 pi i k/960), k <= 240
  float win[480];        // Vorbis power-complementary half window
  float dct[484];        // dct_table[j*22+i] = cos((j+.5) i pi/22) (* sqrt(.5) for i == 0)
  float tansig[204];     // tanh(0.04 i) rounded to 6 decimals, i <= 200
  float bin_frac[400];   // j / band_size for bin k inside band interval bin_band[k]
  int32_t bin_band[400]; // band interval index (0..20) of bin k
  int32_t eband[24];     // band edges in bins (eband5ms * 4), 22 used
};

// ---- RNN weights repacked for the recurrent-core kernel (K4).
// Activations of the 8 streams of a CTA live in shared memory as two f32 arrays of rows [k][8]:
//   A: dense(24) | vad_gru_state(24) | features(42) | noise_gru_state(48) | denoise_gru_state(96)
//   R: r*h of the three GRUs: vad(24) | noise(48) | denoise(96)
// so that every matrix-vector job reads one contiguous range of A plus, for the candidate-gate jobs,
// one contiguous range of R.  A job with N outputs is worked by cp = ceil(N/2) column pairs x ksplit
// slices of its K inputs; thread tj = ks*cp + pair keeps 2 x 8 accumulators and reads its weights
// as packed bf16 pairs (int8 values are exact in bf16) from words[w_off + i*(cp*ksplit) + tj].
constexpr int kRnnThreads = 256;
constexpr int kActDense = 0, kActHVad = 24, kActFeat = 48, kActHNoise = 90, kActHDen = 138, kActRows = 234;
constexpr int kRhVad = 0, kRhNoise = 24, kRhDen = 72, kRhRows = 168;
constexpr int kNumJobs = 8;
enum JobKind : int32_t { kJobDense = 0, kJobZR = 1, kJobC = 2 };
struct JobDesc {
  int32_t n_out;       // N
  int32_t activation;  // 0 tanh, 1 sigmoid, 2 relu (of the layer; ZR jobs are always sigmoid)
  int32_t kind;        // JobKind
  int32_t cp, ksplit, len;  // column pairs, K slices, slice length
  int32_t k_total;     // K = len1 + len2
  int32_t off1, len1;  // rows of A
  int32_t off2, len2;  // rows of R (candidate-gate jobs), len2 == 0 otherwise
  int32_t w_off;       // offset (uint32 words) into words
  int32_t b_off;       // offset into bias
  int32_t out_off;     // Dense: row of A (or -1: gains); ZR / C: row of A holding this GRU's state
  int32_t rh_off;      // ZR: row of R receiving r*h
  int32_t pad;
};
struct RnnHeader {
  JobDesc jobs[kNumJobs];
  int32_t vad_w_off;   // vad_output: 24 f32 weights + bias at bias[vad_w_off .. +25)
  int32_t vad_activation;
  int32_t n_words;
  int32_t n_bias;
};

// ---- launch parameters
enum Flags : uint32_t {
  kFlagInI16 = 1u << 0,       // input samples are int16 (16-bit scale)
  kFlagOutI16 = 1u << 1,      // output samples are int16 (round to nearest, saturate)
  kFlagUnitScale = 1u << 2,   // audio.rs:261-273 wrapper arithmetic: x32768 in, /32768 + clamp + *volume out
  kFlagMixStereoI16 = 1u << 3 // f1: out = interleaved stereo i16 of clamp(dn + app) * 32767 (trunc)
};

// One chunk of one call.  Caller pointers (in/out/vad/app/dbg) address frame 0 of the CALL; the
// kernels add frame0.  Engine pointers (hp/tab/rec) address the chunk's workspace slot.
struct Params {
  const void *in;
  void *out;
  float *vad;             // [n_streams][vad_stride] or null
  const float *app;       // f1: app audio, unit scale, same geometry as out frames; may be null
  float *dbg;             // [n_streams][n_frames_call][kDbgFloats] or null
  float *state;           // [n_streams][kStateFloats]
  float *hp;              // [n_streams][hp_stride]
  uint32_t *tab;          // [n_streams][chunk_cap][kTabWords]
  float *rec;             // [n_streams][chunk_cap][kRecFloats]
  cf *spec;               // [n_streams][chunk_cap][2][kSpecStride]: X and P of every frame
  const Tables *tables;
  const RnnHeader *rnn_hdr;
  const uint32_t *rnn_words;
  const float *rnn_bias;
  long long in_stride;    // samples between streams
  long long out_stride;   // samples (mono) or stereo pairs between streams
  long long vad_stride;
  long long app_stride;
  long long hp_stride;    // kHist + chunk_cap*480
  int n_streams;
  int n_frames;           // frames in this chunk
  int frame0;             // index of the chunk's first frame within the call
  int n_frames_call;      // frames in the whole call (debug tap geometry)
  int chunk_cap;
  int synth_sel;          // which synthesis_mem copy this chunk reads (chunk counter & 1)
  int out_frame_offset;   // output frame t is stored at frame slot t + out_frame_offset (skipped if < 0)
  uint32_t flags;
  float volume;
};

}  // namespace ns
