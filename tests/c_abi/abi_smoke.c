/* abi_smoke.c -- a plain C consumer of include/crispy_ns.h, linked against libcrispy_ns.so.
 *
 * What the Rust shim (bindings/rust/ns_gpu.rs) does through `extern "C"`, done from C so that the header itself --
 * not the Python ctypes table -- is what gets compiled and exercised:
 *   DenoiseState::new / process_frame (audio.rs:229, :268), the batched host call, the multi-device host call,
 *   the WAV-level entry point, state save/load, error reporting.
 * Without a CUDA device it checks the no-fallback contract (CRISPY_NS_ENODEV everywhere) and exits 0.
 * Build: gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude tests/c_abi/abi_smoke.c -Lcrispy_b200 -lcrispy_ns -lm
 */
#include "crispy_ns.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(expr)                                                                               \
  do {                                                                                            \
    int rc_ = (expr);                                                                             \
    if (rc_ != CRISPY_NS_OK) {                                                                    \
      fprintf(stderr, "FAIL %s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #expr, rc_, crispy_ns_last_error()); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) {                                                        \
      fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      return 1;                                                           \
    }                                                                     \
  } while (0)

/* speech-like test signal in [-1, 1]: harmonics of a gliding f0 plus LCG noise */
static void make_signal(float *x, int n, unsigned seed) {
  unsigned r = seed * 2654435761u + 12345u;
  double ph = 0.0;
  int i;
  for (i = 0; i < n; i++) {
    const double f0 = 120.0 + 40.0 * sin(2.0 * 3.14159265358979 * 0.7 * i / 48000.0 + seed);
    double v;
    ph += 2.0 * 3.14159265358979 * f0 / 48000.0;
    r = r * 1664525u + 1013904223u;
    v = 0.2 * sin(ph) + 0.1 * sin(2 * ph) + 0.05 * sin(3 * ph) + 0.04 * (((int)(r >> 16) & 0xFFFF) / 32768.0 - 1.0);
    x[i] = (float)v;
  }
}

int main(int argc, char **argv) {
  const char *tmpdir = argc > 1 ? argv[1] : "/tmp";
  crispy_ns_model *model = NULL;
  crispy_ns_state *st = NULL;
  crispy_ns_batch *batch = NULL;
  crispy_ns_multi *multi = NULL;
  enum { NS = 3, NF = 40, N = NF * CRISPY_NS_FRAME_SIZE };
  int i, t;

  EXPECT(crispy_ns_frame_size() == CRISPY_NS_FRAME_SIZE);
  CHECK(crispy_ns_model_synthetic(0, &model));
  {
    size_t need = 0;
    CHECK(crispy_ns_model_to_bytes(model, NULL, 0, &need));
    EXPECT(need > 87503);
  }
  if (crispy_ns_device_count() == 0) { /* the no-fallback contract */
    int dev = 0;
    EXPECT(crispy_ns_create(model, 0, &st) == CRISPY_NS_ENODEV);
    EXPECT(strstr(crispy_ns_last_error(), "no CPU fallback") != NULL);
    EXPECT(crispy_ns_batch_create(model, 0, 4, &batch) == CRISPY_NS_ENODEV);
    EXPECT(crispy_ns_multi_create(model, &dev, 1, 4, &multi) == CRISPY_NS_ENODEV);
    EXPECT(crispy_ns_measure_fp32(0, NULL, NULL) == CRISPY_NS_ENODEV);
    crispy_ns_model_destroy(model);
    printf("abi_smoke: no CUDA device, ENODEV contract ok\n");
    return 0;
  }

  {
    float *x, *y_frame, *y_batch, *y_multi, *vad_b, *vad_m, vad_f[NF];
    void *px = NULL, *py = NULL, *pv = NULL;
    size_t state_bytes;
    void *state_blob;
    int dev0 = 0;
    /* pinned host buffers, as the Rust side would allocate for recordings in RAM */
    CHECK(crispy_ns_host_alloc(&px, sizeof(float) * NS * N));
    CHECK(crispy_ns_host_alloc(&py, sizeof(float) * NS * N));
    CHECK(crispy_ns_host_alloc(&pv, sizeof(float) * NS * NF));
    x = (float *)px, y_batch = (float *)py, vad_b = (float *)pv;
    y_frame = (float *)malloc(sizeof(float) * N);
    y_multi = (float *)malloc(sizeof(float) * NS * N);
    vad_m = (float *)malloc(sizeof(float) * NS * NF);
    for (i = 0; i < NS; i++) make_signal(x + (size_t)i * N, N, (unsigned)i);

    /* DenoiseState::new + process_frame, 16-bit scale (audio.rs:261-268) */
    CHECK(crispy_ns_create(model, 0, &st));
    for (t = 0; t < NF; t++) {
      float in[CRISPY_NS_FRAME_SIZE];
      for (i = 0; i < CRISPY_NS_FRAME_SIZE; i++) in[i] = x[t * CRISPY_NS_FRAME_SIZE + i] * 32768.0f;
      CHECK(crispy_ns_process_frame(st, y_frame + t * CRISPY_NS_FRAME_SIZE, in, &vad_f[t]));
    }
    /* batched host call, unit scale (the wrapper arithmetic fused) */
    CHECK(crispy_ns_batch_create(model, 0, NS, &batch));
    EXPECT(crispy_ns_batch_n_streams(batch) == NS);
    CHECK(crispy_ns_process_streams_host(batch, x, y_batch, vad_b, NULL, NF, N, N, NF, 0, CRISPY_NS_UNIT_SCALE, 1.0f));
    /* frame-at-a-time and batched results are the same computation */
    {
      double worst = 0.0;
      for (i = 0; i < N; i++) {
        float a = y_frame[i] / 32768.0f, d;
        a = a > 1.f ? 1.f : (a < -1.f ? -1.f : a);
        d = (float)fabs(a - y_batch[i]);
        if (d > worst) worst = d;
      }
      for (t = 0; t < NF; t++) EXPECT(fabs(vad_f[t] - vad_b[t]) <= 1e-6);
      EXPECT(worst <= 1e-6);
    }
    /* state save -> load round trip; a bad length is an error, not a crash */
    state_bytes = crispy_ns_batch_state_size(batch);
    state_blob = malloc(state_bytes);
    CHECK(crispy_ns_batch_save_state(batch, state_blob, state_bytes));
    EXPECT(crispy_ns_batch_load_state(batch, state_blob, state_bytes - 1) == CRISPY_NS_EINVAL);
    CHECK(crispy_ns_batch_load_state(batch, state_blob, state_bytes));
    free(state_blob);
    /* OUT_I16 takes no volume (header) */
    EXPECT(crispy_ns_process_streams_host(batch, x, y_batch, NULL, NULL, NF, N, N, NF, 0,
                                          CRISPY_NS_UNIT_SCALE | CRISPY_NS_OUT_I16, 0.5f) == CRISPY_NS_EINVAL);
    /* multi-device handle on device 0 alone (and twice device 0: two blocks): same bits as the batch */
    {
      int devs[2] = {0, 0}, nd, d, f, n;
      CHECK(crispy_ns_batch_reset(batch));
      CHECK(crispy_ns_process_streams_host(batch, x, y_batch, vad_b, NULL, NF, N, N, NF, 0, CRISPY_NS_UNIT_SCALE, 1.0f));
      CHECK(crispy_ns_multi_create(model, devs, 2, NS, &multi));
      nd = crispy_ns_multi_n_devices(multi);
      EXPECT(nd == 2);
      CHECK(crispy_ns_multi_stream_range(multi, 1, &d, &f, &n));
      EXPECT(d == 0 && f == 2 && n == 1);
      CHECK(crispy_ns_multi_process_streams_host(multi, x, y_multi, vad_m, NULL, NF, N, N, NF, 0, CRISPY_NS_UNIT_SCALE, 1.0f));
      EXPECT(memcmp(y_multi, y_batch, sizeof(float) * NS * N) == 0);
      EXPECT(memcmp(vad_m, vad_b, sizeof(float) * NS * NF) == 0);
      crispy_ns_multi_destroy(multi);
      (void)dev0;
    }
    /* WAV level: two stereo recordings of different length in, dual-mono out */
    {
      char p_in0[512], p_in1[512], p_out0[512], p_out1[512];
      const char *pin[2], *pout[2];
      int16_t *pcm = (int16_t *)malloc(sizeof(int16_t) * 2 * N);
      int16_t *back = (int16_t *)malloc(sizeof(int16_t) * 2 * N);
      float mv[2];
      int64_t nfr = 0;
      int ch = 0, sr = 0, k, len1 = N - 700;
      snprintf(p_in0, sizeof p_in0, "%s/abi_in0.wav", tmpdir);
      snprintf(p_in1, sizeof p_in1, "%s/abi_in1.wav", tmpdir);
      snprintf(p_out0, sizeof p_out0, "%s/abi_out0.wav", tmpdir);
      snprintf(p_out1, sizeof p_out1, "%s/abi_out1.wav", tmpdir);
      for (k = 0; k < 2; k++) {
        const int len = k ? len1 : N;
        for (i = 0; i < len; i++) {
          float v = x[(size_t)k * N + i];
          v = v > 1.f ? 1.f : (v < -1.f ? -1.f : v);
          pcm[2 * i] = pcm[2 * i + 1] = (int16_t)(v * 32767.0f); /* recording.rs:108-110 */
        }
        CHECK(crispy_ns_wav_write_pcm16(k ? p_in1 : p_in0, pcm, len, 2, 48000));
      }
      pin[0] = p_in0, pin[1] = p_in1, pout[0] = p_out0, pout[1] = p_out1;
      CHECK(crispy_ns_denoise_wav_files(model, 0, pin, pout, 2, 0, 1.0f, mv));
      CHECK(crispy_ns_wav_read_pcm16(p_out1, back, 2 * N, &nfr, &ch, &sr));
      EXPECT(nfr == len1 && ch == 2 && sr == 48000);
      for (i = 0; i < len1; i++) EXPECT(back[2 * i] == back[2 * i + 1]); /* dual mono */
      EXPECT(mv[0] >= 0.f && mv[0] <= 1.f && mv[1] >= 0.f && mv[1] <= 1.f);
      EXPECT(crispy_ns_denoise_wav_files(model, 0, pin, pout, 0, 0, 1.0f, NULL) == CRISPY_NS_EINVAL);
      pin[1] = "/nonexistent/dir/x.wav";
      EXPECT(crispy_ns_denoise_wav_files(model, 0, pin, pout, 2, 0, 1.0f, NULL) == CRISPY_NS_EIO);
      free(pcm);
      free(back);
      remove(p_in0), remove(p_in1), remove(p_out0), remove(p_out1);
    }
    {
      double tf = 0.0, tm = 0.0;
      CHECK(crispy_ns_measure_fp32(0, &tf, &tm));
      EXPECT(tf > 1.0 && tm > 0.5);
      printf("abi_smoke: fp32 burst %.1f TFLOP/s fused, %.2f T MAC/s unfused\n", tf, tm);
    }
    crispy_ns_batch_destroy(batch);
    crispy_ns_destroy(st);
    crispy_ns_host_free(px);
    crispy_ns_host_free(py);
    crispy_ns_host_free(pv);
    free(y_frame);
    free(y_multi);
    free(vad_m);
  }
  crispy_ns_model_destroy(model);
  printf("abi_smoke: ok\n");
  return 0;
}
