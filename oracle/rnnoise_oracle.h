/*
 * oracle/rnnoise_oracle.h -- CPU restatement of the RNNoise denoiser that sleep3r/crispy calls
 * through `nnnoiseless::DenoiseState` (reference call sites: src-tauri/src/audio.rs:4, :203, :229,
 * :268; crate pin nnnoiseless 0.5.2 at Cargo.lock:2825-2838).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (crispy_b200/, libcrispy_ns.so) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED: the nnnoiseless crate (source + its embedded ~87.5K int8 weights) is not
 * vendored under /root/reference, no Rust toolchain exists in this image, and the reference tree
 * holds no golden vectors for this path.  This file restates the
 * published algorithm of nnnoiseless 0.5.2 == xiph/rnnoise (denoise.c, pitch.c, celt_lpc.c, rnn.c)
 * in scalar f32 with the upstream summation order.  It is anchored on (a) the reference's call
 * sites and wrapper arithmetic (audio.rs:242-295), (b) the reference's own unit tests for the
 * neighbouring rows (LinearResampler audio.rs:1040-1096, WavWriter recording.rs:454-504), and
 * (c) algorithm-level known answers (window power-complementarity, DCT orthonormality, tanh
 * table, analysis/synthesis perfect reconstruction on the silence path), and (d) two NumPy transliterations of the
 * published algorithm written independently of this file (tests/np_pitch.py: pitch decisions and gain bit for bit;
 * tests/np_denoise.py: the whole frame in float64, agreement to float32 accuracy).
 */
#ifndef RNNOISE_ORACLE_H
#define RNNOISE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNO_FRAME_SIZE 480
#define RNO_WINDOW_SIZE 960
#define RNO_FREQ_SIZE 481
#define RNO_NB_BANDS 22
#define RNO_NB_FEATURES 42
#define RNO_PITCH_BUF_SIZE 1728

typedef struct rno_model rno_model;
typedef struct rno_state rno_state;

/* Intermediate taps of the most recent process_frame call (for stage-by-stage parity tests). */
typedef struct rno_debug {
  float features[RNO_NB_FEATURES];
  float gains[RNO_NB_BANDS];
  float Ex[RNO_NB_BANDS];
  float Ep[RNO_NB_BANDS];
  float Exp[RNO_NB_BANDS];
  float pitch_gain;
  float vad;
  int32_t pitch_index;
  int32_t silence;
} rno_debug;

/* ---- model (the six layers of SURVEY.md Appendix A.6) ---- */
rno_model *rno_model_synthetic(uint64_t seed);
rno_model *rno_model_from_bytes(const void *blob, size_t len); /* "CRNSMDL1" binary or rnnoise-nu text */
size_t rno_model_to_bytes(const rno_model *m, void *buf, size_t cap);
void rno_model_free(rno_model *m);

/* ---- DenoiseState::new / process_frame (audio.rs:229, :268) ---- */
rno_state *rno_create(const rno_model *m);
void rno_destroy(rno_state *st);
void rno_reset(rno_state *st);
/* in/out: 480 f32 in 16-bit scale (audio.rs:264 multiplies by 32768 first). Returns VAD prob. */
float rno_process_frame(rno_state *st, float *out, const float *in);
void rno_get_debug(const rno_state *st, rno_debug *dbg);

/* ---- batch helper: n_streams independent DenoiseStates over n_frames frames, n_threads pthreads
 *      over streams (the CPU baseline of SURVEY.md 8d).  flags: bit0 = unit-scale wrapper
 *      arithmetic of audio.rs:261-273 (x32768 in, /32768 + clamp + *volume out).  Output frame t
 *      is written at out + s*out_stride + t*480 (no first-frame drop here). ---- */
int rno_process_streams(const rno_model *m, const float *in, float *out, float *vad, int n_streams,
                        int n_frames, long in_stride, long out_stride, unsigned flags, float volume,
                        int n_threads);

/* the same, additionally recording every frame's discrete decisions ([n_streams][n_frames] each, any may be
 * NULL): pitch_index and pitch gain out of remove_doubling, the silence gate, and branch_margin = how far the pitch
 * filter's discontinuous `Exp > g ? 1 : ...` branch is from flipping: the smallest |Exp[b] - g[b]| over the bands whose
 * r reaches the output, in units of the rounding noise two float32 implementations carry on Exp and g, scaled so
 * that 1e-4 means "within the noise" (rnnoise_oracle.c rno_process_frame); 1e30 on silent frames */
int rno_process_streams_trace(const rno_model *m, const float *in, float *out, float *vad, int n_streams,
                              int n_frames, long in_stride, long out_stride, unsigned flags, float volume,
                              int n_threads, int32_t *pitch_index, float *pitch_gain, int32_t *silence,
                              float *branch_margin);

/* Summation order of the pitch path's inner products: 0 = sequential (xiph C order; default; what the CUDA kernels
 * reproduce), 1 = four interleaved partial sums in celt_inner_prod / dual_inner_prod, 2 = also inside the
 * cross-correlation kernels.  Process-wide; set before any thread runs.  A measurement aid: see rnnoise_oracle.c. */
void rno_set_sum_policy(int policy);
int rno_get_sum_policy(void);
/* Measurement aid: Exp (1 + rel_exp) + abs_exp and g + logit_g g (1 - g) inside the pitch_filter call only (0, 0, 0 = off). */
void rno_set_pf_perturb(float rel_exp, float logit_g, float abs_exp);
/* the band gains of the last frame as the RNN produced them (before g = max(g, 0.6 lastg)): what pitch_filter compares Exp with */
void rno_get_raw_gains(const rno_state *st, float *graw);

/* ---- neighbouring rows ---- */
/* capture callbacks' downmix to mono (audio.rs:754-755, :816-818, :879-884); fmt 0 f32, 1 i16, 2 u16 */
void rno_downmix_mono(const void *in, int fmt, int n_channels, size_t n_frames, float *out);

/* f2, app audio: resample_audio (recording.rs:13-39); returns the number of samples produced */
size_t rno_resample_audio(const float *samples, size_t len, size_t from_rate, size_t to_rate, float *out, size_t out_cap);

/* a4: LinearResampler (audio.rs:73-134). Streaming; returns number of samples emitted. */
typedef struct rno_linres {
  float input_rate, output_rate, last_sample;
  int has_last;
  double input_pos, next_output_pos;
} rno_linres;
void rno_linres_init(rno_linres *r, float input_rate, float output_rate);
size_t rno_linres_process(rno_linres *r, const float *in, size_t n_in, float *out, size_t out_cap);

/* a3: RnnNoiseProcessor::push_sample semantics over a whole buffer (audio.rs:242-295):
 * optional linear resample to 48k, frame assembly, x32768, process_frame, /32768, clamp, *volume,
 * first frame dropped.  Returns samples written to out. */
size_t rno_processor_run(const rno_model *m, float input_rate, float volume, const float *in,
                         size_t n_in, float *out, size_t out_cap);

/* f1: dual-mono mix + PCM16 quantiser (commands/recording.rs:260-264, recording.rs:101-121).
 * out is interleaved stereo i16, both channels = trunc(clamp(mic+app,-1,1)*32767). */
void rno_mix_dual_mono_i16(const float *mic, const float *app, size_t n, int16_t *out_interleaved);

/* f2 (north_star item 4): windowed-sinc polyphase resampler, "rubato-equivalent" front end.  The
 * reference itself only uses a linear interpolator on this path (audio.rs:108-133); rubato 0.16.2
 * (Cargo.lock:4166) appears in the tree as FftFixedIn ahead of transcription
 * (commands/transcription.rs:201-207).  This restates rubato's published synchronous sinc design
 * (SincFixedIn: make_sincs + BlackmanHarris2 window, sinc_len taps, f_cutoff relative to Nyquist,
 * cutoff scaled by the ratio when downsampling) with the oversampling factor set to L of the
 * reduced ratio L/M = output_rate/input_rate, so every output lands exactly on a tabulated phase
 * and no inter-phase interpolation is needed; delay-compensated (zero-phase), zeros outside
 * [0, n_in).  PARITY UNPINNED like the rest of this file: no rubato source or vectors are available.
 *   out[n] = sum_{k<sinc_len} h[(n*M) % L][k] * in[floor(n*M/L) - sinc_len/2 + 1 + k]
 * accumulated in f32 with fmaf in ascending k.  Returns samples written (<= out_cap); count only
 * when out == NULL.  Returns 0 for unsupported ratios (L > 1024) or bad sinc_len. */
size_t rno_sinc_resample_count(int input_rate, int output_rate, size_t n_in);
size_t rno_sinc_resample(const float *in, size_t n_in, float *out, size_t out_cap, int input_rate,
                         int output_rate, int sinc_len, float f_cutoff);
/* the polyphase table itself ([L][sinc_len] f32) for known-answer tests; returns L, or 0 */
int rno_sinc_table(int input_rate, int output_rate, int sinc_len, float f_cutoff, float *table,
                   size_t cap_floats, int *M_out);

/* tables exposed for known-answer tests */
const float *rno_half_window(void);  /* 480 */
const float *rno_dct_table(void);    /* 22*22 */
const float *rno_tansig_table(void); /* 201 */
float rno_tansig_approx(float x);
float rno_sigmoid_approx(float x);
/* forward (scaled 1/960) and inverse (unscaled) 960-point real transforms as used on the path */
void rno_forward_transform(float *out_re_im_481x2, const float *in960);
void rno_inverse_transform(float *out960, const float *in_re_im_481x2);

#ifdef __cplusplus
}
#endif
#endif
