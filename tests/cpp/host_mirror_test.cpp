// host_mirror_test.cpp -- the C++ host layer (include/crispy_ns.hpp) exercised the way the reference exercises its
// own Rust code: the first two groups are the reference's unit tests restated against the mirror
// (src-tauri/src/audio.rs:1040-1096 LinearResampler, src-tauri/src/recording.rs:406-520 WavWriter), the third drives
// RnnNoiseProcessor / DenoiseState / BatchDenoiser on the GPU and leaves its outputs in files that
// tests/test_cpp_host.py compares with the oracle (the oracle is test infrastructure; this program never links it).
//
//   host_mirror_test cpu <tmpdir>            no CUDA device needed (WAV I/O is host code inside the library)
//   host_mirror_test gpu <tmpdir>            needs <tmpdir>/clip48k.f32 and <tmpdir>/clip441.f32 (unit-scale f32)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "crispy_ns.hpp"

using namespace crispy;

static int g_failed = 0;
#define EXPECT(cond, ...)                                          \
  do {                                                             \
    if (!(cond)) {                                                 \
      std::fprintf(stderr, "FAIL %s:%d: %s -- ", __FILE__, __LINE__, #cond); \
      std::fprintf(stderr, __VA_ARGS__);                           \
      std::fprintf(stderr, "\n");                                  \
      g_failed++;                                                  \
    }                                                              \
  } while (0)

// ---- audio.rs:1040-1096 ------------------------------------------------------------------------------------------------
static void linear_resampler_same_rate_passthrough() {
  auto resampler = LinearResampler::new_(48000.0f, 48000.0f);
  std::vector<float> output;
  for (int i = 0; i < 10; i++) resampler.process_sample((float)i * 0.1f, [&](float s) { output.push_back(s); });
  EXPECT(output.size() == 10, "same rate: each input produces exactly one output, got %zu", output.size());
  for (std::size_t i = 0; i < output.size(); i++) EXPECT(std::fabs(output[i] - (float)i * 0.1f) < 0.001f, "Sample %zu mismatch", i);
}
static void linear_resampler_downsample_produces_fewer() {
  auto resampler = LinearResampler::new_(48000.0f, 16000.0f);
  std::vector<float> output;
  for (int i = 0; i < 300; i++) resampler.process_sample(0.5f, [&](float s) { output.push_back(s); });
  EXPECT(output.size() > 80 && output.size() < 120, "Expected ~100 output samples, got %zu", output.size());
}
static void linear_resampler_upsample_produces_more() {
  auto resampler = LinearResampler::new_(16000.0f, 48000.0f);
  std::vector<float> output;
  for (int i = 0; i < 100; i++) resampler.process_sample(0.5f, [&](float s) { output.push_back(s); });
  EXPECT(output.size() > 250 && output.size() < 350, "Expected ~300 output samples, got %zu", output.size());
}
static void linear_resampler_rates_preserved() {
  auto resampler = LinearResampler::new_(44100.0f, 48000.0f);
  auto [input, output] = resampler.rates();
  EXPECT(std::fabs(input - 44100.0f) < 0.1f && std::fabs(output - 48000.0f) < 0.1f, "rates");
}
static void linear_resampler_set_rates_updates() {
  auto resampler = LinearResampler::new_(48000.0f, 48000.0f);
  resampler.set_rates(44100.0f, 16000.0f);
  auto [input, output] = resampler.rates();
  EXPECT(std::fabs(input - 44100.0f) < 0.1f && std::fabs(output - 16000.0f) < 0.1f, "set_rates");
}
// the library's batched count (crispy_ns_linear_resample_count replays the same f64 positions) agrees with the streaming
// mirror for the rate pairs the reference meets
static void linear_resampler_count_matches_the_library() {
  const float rates[][2] = {{44100.f, 48000.f}, {16000.f, 48000.f}, {96000.f, 48000.f}, {48000.4f, 48000.f}, {22050.f, 48000.f}};
  for (auto &r : rates)
    for (long n : {0L, 1L, 2L, 441L, 4410L, 48001L}) {
      auto rs = LinearResampler::new_(r[0], r[1]);
      std::size_t cnt = 0;
      for (long i = 0; i < n; i++) rs.process_sample(0.25f, [&](float) { cnt++; });
      EXPECT((long long)cnt == (long long)crispy_ns_linear_resample_count(r[0], r[1], n), "count %g -> %g, n = %ld", r[0], r[1], n);
    }
}

// ---- recording.rs:406-520 ----------------------------------------------------------------------------------------------
static bool file_exists(const std::string &p) { return std::ifstream(p).good(); }

static void wav_writer_creates_file(const std::string &dir) {
  const std::string path = dir + "/test_create.wav";
  auto writer = WavWriter::new_(path);
  EXPECT(writer.output_path() == path, "output_path");
  const std::string finalized_path = std::move(writer).finalize();
  EXPECT(finalized_path == path, "finalize returns the path");
  EXPECT(file_exists(path), "file exists");
}
static void wav_writer_writes_silence(const std::string &dir) {
  const std::string path = dir + "/test_silence.wav";
  auto writer = WavWriter::new_(path);
  std::vector<float> left(48000, 0.0f), right(48000, 0.0f);
  writer.write_samples(left, right);
  std::move(writer).finalize();
  const WavData w = wav_read_pcm16(path);
  EXPECT(w.channels == (int)CHANNELS && w.sample_rate == (int)SAMPLE_RATE, "spec %d ch %d Hz", w.channels, w.sample_rate);
  EXPECT(w.interleaved.size() == 48000 * 2, "48000 samples * 2 channels, got %zu", w.interleaved.size());
  bool all_zero = true;
  for (auto s : w.interleaved) all_zero = all_zero && s == 0;
  EXPECT(all_zero, "all silence");
  // the 44-byte header hound writes for 16-bit stereo: RIFF size = 36 + data, fmt chunk of 16 bytes, PCM tag 1
  std::ifstream f(path, std::ios::binary);
  unsigned char h[44] = {};
  f.read((char *)h, 44);
  const unsigned data_bytes = 48000u * 2u * 2u;
  auto le32 = [&](int o) { return (unsigned)h[o] | ((unsigned)h[o + 1] << 8) | ((unsigned)h[o + 2] << 16) | ((unsigned)h[o + 3] << 24); };
  EXPECT(std::memcmp(h, "RIFF", 4) == 0 && std::memcmp(h + 8, "WAVEfmt ", 8) == 0 && std::memcmp(h + 36, "data", 4) == 0, "chunk ids");
  EXPECT(le32(4) == 36 + data_bytes && le32(16) == 16 && le32(24) == 48000 && le32(28) == 48000u * 4u && le32(40) == data_bytes, "header sizes");
  EXPECT(h[20] == 1 && h[22] == 2 && h[32] == 4 && h[34] == 16, "format fields");
}
static void wav_writer_writes_audio_data(const std::string &dir) {
  const std::string path = dir + "/test_data.wav";
  auto writer = WavWriter::new_(path);
  std::vector<float> left(100, 0.5f), right(100, -0.5f);
  writer.write_samples(left, right);
  std::move(writer).finalize();
  const WavData w = wav_read_pcm16(path);
  EXPECT(w.interleaved.size() == 200, "100 * 2 channels");
  const std::int16_t expected_left = (std::int16_t)(0.5f * 32767.0f), expected_right = (std::int16_t)(-0.5f * 32767.0f);
  for (int i = 0; i < 100 && w.interleaved.size() == 200; i++)
    EXPECT(w.interleaved[i * 2] == expected_left && w.interleaved[i * 2 + 1] == expected_right, "interleave at %d", i);
}
static void wav_writer_clamps_samples(const std::string &dir) {
  const std::string path = dir + "/test_clamp.wav";
  auto writer = WavWriter::new_(path);
  writer.write_samples({2.0f, -3.0f}, {1.5f, -1.5f});
  std::move(writer).finalize();
  const WavData w = wav_read_pcm16(path);
  EXPECT(w.interleaved.size() == 4, "two frames");
  if (w.interleaved.size() == 4) {
    EXPECT(w.interleaved[0] == 32767 && w.interleaved[1] == 32767, "2.0 / 1.5 clamped to 1.0");
    EXPECT(w.interleaved[2] == -32767 && w.interleaved[3] == -32767, "-3.0 / -1.5 clamped to -1.0 (x 32767, truncated)");
  }
}
static void wav_writer_rejects_mismatched_channels(const std::string &dir) {
  const std::string path = dir + "/test_mismatch.wav";
  auto writer = WavWriter::new_(path);
  bool threw = false;
  try {
    writer.write_samples({0.0f, 0.0f, 0.0f}, {0.0f, 0.0f});
  } catch (const Error &e) {
    threw = std::string(e.what()).find("mismatch") != std::string::npos;
  }
  EXPECT(threw, "mismatched channel lengths are an error whose text says so (recording.rs:103)");
}
static void wav_writer_rejects_an_unwritable_path(const std::string &dir) {
  bool threw = false;
  try {
    WavWriter::new_(dir + "/no/such/dir/x.wav");
  } catch (const Error &e) {
    threw = std::string(e.what()).find("Failed to create WAV writer") == 0;
  }
  EXPECT(threw, "hound::WavWriter::create fails on an unwritable path; so does new_");
}

// the streaming interpolator's samples for a seeded clip, left for the pytest side to compare bit for bit with the
// oracle's restatement of audio.rs:108-133 (run for the plain build and for one with -march=native -ffp-contract=fast:
// the products feeding an addition must not turn into fused multiply-adds)
static void linear_resampler_dump(const std::string &dir) {
  std::ifstream probe(dir + "/clip441.f32", std::ios::binary);
  if (!probe) return;
  probe.close();
  std::ifstream f(dir + "/clip441.f32", std::ios::binary | std::ios::ate);
  std::vector<float> x((std::size_t)f.tellg() / 4), y;
  f.seekg(0);
  f.read((char *)x.data(), (std::streamsize)(x.size() * 4));
  auto rs = LinearResampler::new_(44100.0f, 48000.0f);
  for (float s : x) rs.process_sample(s, [&](float o) { y.push_back(o); });
  std::ofstream o(dir + "/out_linres.f32", std::ios::binary);
  o.write((const char *)y.data(), (std::streamsize)(y.size() * 4));
}

// ---- no CPU fallback -----------------------------------------------------------------------------------------------------
static void no_device_is_an_error_not_a_fallback() {
  int code = 0;
  try {
    auto st = DenoiseState::new_();
  } catch (const Error &e) {
    code = e.code();
  }
  EXPECT(code == CRISPY_NS_ENODEV, "DenoiseState::new_ without a CUDA device throws ENODEV (got %d)", code);
  code = 0;
  try {
    BatchDenoiser b(4);
  } catch (const Error &e) {
    code = e.code();
  }
  EXPECT(code == CRISPY_NS_ENODEV, "BatchDenoiser without a CUDA device throws ENODEV (got %d)", code);
}

// ---- GPU: the operator as audio.rs drives it ----------------------------------------------------------------------------
static std::vector<float> read_f32(const std::string &p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) {
    std::fprintf(stderr, "cannot read %s\n", p.c_str());
    std::exit(2);
  }
  std::vector<float> v((std::size_t)f.tellg() / 4);
  f.seekg(0);
  f.read((char *)v.data(), (std::streamsize)(v.size() * 4));
  return v;
}
static void write_f32(const std::string &p, const std::vector<float> &v) {
  std::ofstream f(p, std::ios::binary);
  f.write((const char *)v.data(), (std::streamsize)(v.size() * 4));
}

static void gpu_operator(const std::string &dir) {
  const std::vector<float> clip48 = read_f32(dir + "/clip48k.f32"), clip441 = read_f32(dir + "/clip441.f32");
  const Model model = Model::synthetic(0);

  // (1) 48 kHz microphone, volume 0.5, playback device at 44.1 kHz: push_sample's outputs, and what next_sample
  //     (audio.rs:297-314) plays when the output callback pulls 441 samples per 10 ms
  {
    RnnNoiseProcessor p(48000.0f, 44100.0f, 0.5f, &model);
    std::vector<float> pushed, played;
    std::size_t n_some = 0;
    EXPECT(p.next_sample() == 0.0f, "nothing buffered: silence");
    for (float s : clip48) {
      if (auto out = p.push_sample(s)) {
        EXPECT(out->size() == FRAME_SIZE, "a push yields nothing or one frame");
        pushed.insert(pushed.end(), out->begin(), out->end());
        n_some++;
        for (int i = 0; i < 441; i++) played.push_back(p.next_sample());
      }
    }
    EXPECT(n_some == clip48.size() / FRAME_SIZE - 1, "first frame dropped (audio.rs:275-278): %zu frames out", n_some);
    write_f32(dir + "/out_push48k.f32", pushed);
    write_f32(dir + "/out_played441.f32", played);
  }
  // (2) a 44.1 kHz microphone: LinearResampler in front (audio.rs:217-221), full volume, volume clamped from 3.0
  {
    RnnNoiseProcessor p(44100.0f, 48000.0f, 3.0f, &model);
    EXPECT(p.volume() == 1.0f, "volume.clamp(0, 1)");
    std::vector<float> pushed;
    for (float s : clip441)
      if (auto out = p.push_sample(s)) pushed.insert(pushed.end(), out->begin(), out->end());
    write_f32(dir + "/out_push441.f32", pushed);
  }
  // (3) DenoiseState alone, 16-bit scale, VAD returned; and the same frames through BatchDenoiser in one call:
  //     per-frame calls and one batched call give the same bits (the state is carried the same way)
  {
    auto st = DenoiseState::new_(&model);
    const std::size_t nf = clip48.size() / FRAME_SIZE;
    std::vector<float> in16(nf * FRAME_SIZE), out16(nf * FRAME_SIZE), vad(nf);
    for (std::size_t i = 0; i < in16.size(); i++) in16[i] = clip48[i] * 32768.0f;
    for (std::size_t f = 0; f < nf; f++)
      vad[f] = st->process_frame(out16.data() + f * FRAME_SIZE, FRAME_SIZE, in16.data() + f * FRAME_SIZE, FRAME_SIZE);
    write_f32(dir + "/out_frames16.f32", out16);
    write_f32(dir + "/out_vad.f32", vad);
    bool threw = false;
    try {
      std::vector<float> small(479), out(480);
      st->process_frame(out, small);
    } catch (const std::invalid_argument &) {
      threw = true;
    }
    EXPECT(threw, "a frame of the wrong length is a programming error (upstream asserts)");

    BatchDenoiser b(1, &model);
    PinnedBuffer<float> hin(nf * FRAME_SIZE), hout(nf * FRAME_SIZE), hvad(nf);
    for (std::size_t i = 0; i < nf * FRAME_SIZE; i++) hin[i] = clip48[i];
    b.process_streams(hin.data(), hout.data(), hvad.data(), (int)nf, (std::int64_t)(nf * FRAME_SIZE));
    std::size_t diff = 0, vdiff = 0;
    for (std::size_t i = 0; i < nf * FRAME_SIZE; i++) {
      const float o = detail::clampf(out16[i] / 32768.0f, -1.0f, 1.0f);
      diff += o != hout[i];
    }
    for (std::size_t f = 0; f < nf; f++) vdiff += vad[f] != hvad[f];
    EXPECT(diff == 0 && vdiff == 0, "frame-by-frame and batched calls agree bit for bit (%zu samples, %zu VADs differ)", diff, vdiff);

    // checkpoint / resume: the second half after load_state equals the second half of the uninterrupted run
    BatchDenoiser c(1, &model);
    const std::size_t h = nf / 2;
    std::vector<float> o2(nf * FRAME_SIZE);
    c.process_streams(hin.data(), hout.data(), nullptr, (int)h, (std::int64_t)(nf * FRAME_SIZE));
    const auto blob = c.save_state();
    BatchDenoiser d(1, &model);
    d.load_state(blob);
    PinnedBuffer<float> hout2(nf * FRAME_SIZE);
    d.process_streams(hin.data() + h * FRAME_SIZE, hout2.data(), nullptr, (int)(nf - h), (std::int64_t)(nf * FRAME_SIZE));
    std::size_t rdiff = 0;
    for (std::size_t i = 0; i < (nf - h) * FRAME_SIZE; i++) {
      const float o = detail::clampf(out16[h * FRAME_SIZE + i] / 32768.0f, -1.0f, 1.0f);
      rdiff += o != hout2[i];
    }
    EXPECT(rdiff == 0, "save_state / load_state resume bit for bit (%zu differ)", rdiff);
  }
  // (4) the recorder's path: microphone denoised + app audio -> dual-mono PCM16 -> WavWriter, and the app-audio
  //     resampler (recording.rs:13-39) bringing a 44.1 kHz source to the recorder's rate first
  {
    const std::size_t nf = clip441.size() / 441;  // 441 samples at 44.1 kHz are one 480-sample frame
    std::vector<float> app441(clip441.begin(), clip441.begin() + (std::ptrdiff_t)(nf * 441));
    for (float &v : app441) v *= 0.25f;
    std::vector<float> app48 = resample_audio(app441, 44100, 48000);
    EXPECT(app48.size() == nf * FRAME_SIZE, "resample_audio: %zu samples for %zu frames", app48.size(), nf);
    write_f32(dir + "/out_app48.f32", app48);
    const std::size_t n = std::min(nf, clip48.size() / FRAME_SIZE);
    BatchDenoiser b(1, &model);
    std::vector<std::int16_t> lr(2 * n * FRAME_SIZE);
    b.process_and_mix(clip48.data(), app48.data(), lr.data(), (int)n, (std::int64_t)(n * FRAME_SIZE));
    auto w = WavWriter::new_(dir + "/out_meeting.wav");
    w.write_interleaved(lr.data(), n * FRAME_SIZE);
    std::move(w).finalize();
  }
}

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s cpu|gpu <tmpdir>\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[1], dir = argv[2];
  try {
    linear_resampler_same_rate_passthrough();
    linear_resampler_downsample_produces_fewer();
    linear_resampler_upsample_produces_more();
    linear_resampler_rates_preserved();
    linear_resampler_set_rates_updates();
    linear_resampler_count_matches_the_library();
    wav_writer_creates_file(dir);
    wav_writer_writes_silence(dir);
    wav_writer_writes_audio_data(dir);
    wav_writer_clamps_samples(dir);
    wav_writer_rejects_mismatched_channels(dir);
    wav_writer_rejects_an_unwritable_path(dir);
    linear_resampler_dump(dir);
    if (mode == "gpu")
      gpu_operator(dir);
    else if (device_count() == 0)
      no_device_is_an_error_not_a_fallback();
  } catch (const std::exception &e) {
    std::fprintf(stderr, "FAIL: unexpected exception: %s\n", e.what());
    return 1;
  }
  if (g_failed) return 1;
  std::printf("host_mirror_test: ok (%s)\n", mode.c_str());
  return 0;
}
