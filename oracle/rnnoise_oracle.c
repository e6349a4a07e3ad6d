/*
 * oracle/rnnoise_oracle.c -- scalar f32 restatement of nnnoiseless 0.5.2 / xiph rnnoise.
 * TEST INFRASTRUCTURE ONLY (see rnnoise_oracle.h).  PARITY UNPINNED (see header).
 *
 * Each function names the reference interface it stands behind and the upstream routine it
 * restates.  The only in-tree citations possible are the call site `denoise.process_frame`
 * at /root/reference/src-tauri/src/audio.rs:268 and the crate pin Cargo.lock:2825-2838; the
 * upstream file names (nnnoiseless src/{denoise,features,pitch,rnn,util}.rs == rnnoise
 * src/{denoise,pitch,celt_lpc,rnn}.c) are given per function.
 *
 * Build: gcc -O2 -ffp-contract=off (Rust never contracts a*b+c into an FMA, so neither do we).
 *
 * Provenance of the algorithm: this file restates, routine by routine and under the upstream identifiers,
 * xiph/rnnoise (src/denoise.c, pitch.c, celt_lpc.c, rnn.c) -- the code nnnoiseless 0.5.2 ports to Rust.  It was
 * written from the published algorithm, not copied from /root/reference (which does not contain it).  The upstream
 * sources carry this notice, reproduced here because the routines below follow them closely:
 *
 *   Copyright (c) 2017-2018 Mozilla; Copyright (c) 2007-2009 Xiph.Org Foundation; Copyright (c) 2003-2008
 *   Jean-Marc Valin; Copyright (c) 2007-2008 CSIRO; Copyright (c) 2008-2011 Octasic Inc.
 *   Redistribution and use in source and binary forms, with or without modification, are permitted provided
 *   that the following conditions are met:
 *   - Redistributions of source code must retain the above copyright notice, this list of conditions and the
 *     following disclaimer.
 *   - Redistributions in binary form must reproduce the above copyright notice, this list of conditions and
 *     the following disclaimer in the documentation and/or other materials provided with the distribution.
 *   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR IMPLIED
 *   WARRANTIES, INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A
 *   PARTICULAR PURPOSE ARE DISCLAIMED.  IN NO EVENT SHALL THE FOUNDATION OR CONTRIBUTORS BE LIABLE FOR ANY
 *   DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES (INCLUDING, BUT NOT LIMITED TO,
 *   PROCUREMENT OF SUBSTITUTE GOODS OR SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION)
 *   HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING
 *   NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE
 *   POSSIBILITY OF SUCH DAMAGE.
 *
 * OPEN POINTS against nnnoiseless itself (SURVEY.md Appendix A [verify]): (1) whether the
 * biquad widens to f64 as the C does (implemented: yes); (2) the GRU activations of the shipped model (carried by
 * the model blob, so the real weights decide); (3) the model blob format (CRNSMDL1 is this repo's own container);
 * (4) the summation order of the inner products (rno_set_sum_policy below).
 */
#include "rnnoise_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FRAME_SIZE 480
#define WINDOW_SIZE 960
#define FREQ_SIZE 481
#define PITCH_MIN_PERIOD 60
#define PITCH_MAX_PERIOD 768
#define PITCH_FRAME_SIZE 960
#define PITCH_BUF_SIZE (PITCH_MAX_PERIOD + PITCH_FRAME_SIZE)
#define NB_BANDS 22
#define CEPS_MEM 8
#define NB_DELTA_CEPS 6
#define NB_FEATURES (NB_BANDS + 3 * NB_DELTA_CEPS + 2)
#define FRAME_SIZE_SHIFT 2
#define WEIGHTS_SCALE (1.f / 256)

#define ACT_TANH 0
#define ACT_SIGMOID 1
#define ACT_RELU 2

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
  float r, i;
} cpx;

/* upstream: denoise.c `eband5ms` / nnnoiseless lib.rs EBAND_5MS */
static const int eband5ms[NB_BANDS] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  10, 12,
                                       14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100};

/* ------------------------------------------------------------------------------------------ */
/* tables                                                                                     */
/* ------------------------------------------------------------------------------------------ */
static float g_half_window[FRAME_SIZE];
static float g_dct_table[NB_BANDS * NB_BANDS];
static float g_tansig[201];
static cpx g_tw960[481];  /* exp(-2 pi i k / 960), k <= 480 */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_fft_plan(void);
static void init_tables(void) {
  int i, j;
  /* upstream: denoise.c check_init(): half_window, dct_table */
  for (i = 0; i < FRAME_SIZE; i++) {
    double s = sin(.5 * M_PI * (i + .5) / FRAME_SIZE);
    g_half_window[i] = (float)sin(.5 * M_PI * s * s);
  }
  for (i = 0; i < NB_BANDS; i++)
    for (j = 0; j < NB_BANDS; j++) {
      double v = cos((i + .5) * j * M_PI / NB_BANDS);
      if (j == 0) v *= sqrt(.5);
      g_dct_table[i * NB_BANDS + j] = (float)v;
    }
  /* upstream: tansig_table.h -- tanh(0.04 i) printed with 6 decimals */
  for (i = 0; i <= 200; i++) g_tansig[i] = (float)(floor(tanh(.04 * i) * 1e6 + .5) / 1e6);
  for (i = 0; i <= 480; i++) {
    g_tw960[i].r = (float)cos(2 * M_PI * i / 960);
    g_tw960[i].i = (float)-sin(2 * M_PI * i / 960);
  }
  init_fft_plan();
}
static void ensure_tables(void) { pthread_once(&g_once, init_tables); }

const float *rno_half_window(void) {
  ensure_tables();
  return g_half_window;
}
const float *rno_dct_table(void) {
  ensure_tables();
  return g_dct_table;
}
const float *rno_tansig_table(void) {
  ensure_tables();
  return g_tansig;
}

/* ------------------------------------------------------------------------------------------ */
/* FFT: 480-point complex mixed radix (Stockham, 5.3.4.4.2) in f32 + real packing.  Stands in for */
/* easyfft/realfft/rustfft (Cargo.lock:1200-1210, :3927-3933, :4199-4210), which compute the   */
/* same DFT in f32; only rounding differs.                                                    */
/* ------------------------------------------------------------------------------------------ */
static cpx cmul(cpx a, cpx b) {
  cpx c;
  c.r = a.r * b.r - a.i * b.i;
  c.i = a.r * b.i + a.i * b.r;
  return c;
}

/* out = DFT480(in): Stockham autosort (decimation in time), radices 5, 3, 4, 4, 2 with the twiddles of every stage
 * tabulated in stage order (g_stage_tw, init_tables).  Stage with radix p after Ns = product of the earlier radices:
 * butterfly j = b*Ns + k reads in[j + r*N/p] * W_{Ns p}^{r k} and writes out[b*Ns*p + k + r*Ns]. */
#define FFT_STAGES 5
static const int g_radix[FFT_STAGES] = {5, 3, 4, 4, 2};
static cpx g_stage_tw[480 * 2]; /* sum over stages of Ns * (p - 1) <= 480 + ... */
static int g_stage_off[FFT_STAGES];

static void init_fft_plan(void) {
  int s, k, r, ns = 1, off = 0;
  for (s = 0; s < FFT_STAGES; s++) {
    const int p = g_radix[s];
    g_stage_off[s] = off;
    for (k = 0; k < ns; k++)
      for (r = 1; r < p; r++) {
        const double a = -2.0 * M_PI * (double)(r * k) / (double)(ns * p);
        g_stage_tw[off].r = (float)cos(a);
        g_stage_tw[off].i = (float)sin(a);
        off++;
      }
    ns *= p;
  }
}

static inline cpx cadd(cpx a, cpx b) {
  cpx c;
  c.r = a.r + b.r;
  c.i = a.i + b.i;
  return c;
}
static inline cpx csub(cpx a, cpx b) {
  cpx c;
  c.r = a.r - b.r;
  c.i = a.i - b.i;
  return c;
}
static inline cpx mul_mi(cpx a) { /* a * (-i) */
  cpx c;
  c.r = a.i;
  c.i = -a.r;
  return c;
}

static void butterfly(cpx *v, int p) {
  if (p == 2) {
    cpx a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
  } else if (p == 4) {
    cpx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3);
    v[3] = csub(t1, t3);
  } else if (p == 3) {
    const float s3 = 0.86602540378443864676f;
    cpx a = cadd(v[1], v[2]), d = csub(v[1], v[2]), m, n;
    m.r = v[0].r - .5f * a.r;
    m.i = v[0].i - .5f * a.i;
    n.r = s3 * d.i; /* -i * s3 * d = (s3 d.i, -s3 d.r) */
    n.i = -s3 * d.r;
    v[0] = cadd(v[0], a);
    v[1] = cadd(m, n);
    v[2] = csub(m, n);
  } else { /* 5 */
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    cpx a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    cpx m1, m2, n1, n2, x0 = v[0];
    m1.r = x0.r + c1 * a1.r + c2 * a2.r;
    m1.i = x0.i + c1 * a1.i + c2 * a2.i;
    m2.r = x0.r + c2 * a1.r + c1 * a2.r;
    m2.i = x0.i + c2 * a1.i + c1 * a2.i;
    n1.r = s1 * b1.i + s2 * b2.i; /* -i * (s1 b1 + s2 b2) */
    n1.i = -(s1 * b1.r + s2 * b2.r);
    n2.r = s2 * b1.i - s1 * b2.i; /* -i * (s2 b1 - s1 b2) */
    n2.i = -(s2 * b1.r - s1 * b2.r);
    v[0].r = x0.r + a1.r + a2.r;
    v[0].i = x0.i + a1.i + a2.i;
    v[1] = cadd(m1, n1);
    v[4] = csub(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
  }
}

static void fft480(cpx *out, const cpx *in) {
  cpx buf[2][480];
  const cpx *src = in;
  int s, ns = 1;
  for (s = 0; s < FFT_STAGES; s++) {
    const int p = g_radix[s], m = 480 / p;
    cpx *dst = (s == FFT_STAGES - 1) ? out : buf[s & 1];
    const cpx *tw = g_stage_tw + g_stage_off[s];
    int b, k, r;
    for (b = 0; b < m / ns; b++)
      for (k = 0; k < ns; k++) {
        cpx v[5];
        const int j = b * ns + k;
        v[0] = src[j];
        for (r = 1; r < p; r++) v[r] = (ns == 1) ? src[j + r * m] : cmul(src[j + r * m], tw[k * (p - 1) + r - 1]);
        butterfly(v, p);
        for (r = 0; r < p; r++) dst[b * ns * p + k + r * ns] = v[r];
      }
    src = dst;
    ns *= p;
  }
}

/* upstream: denoise.c forward_transform (kiss_fft scales by 1/N); nnnoiseless scales by
 * 1/WINDOW_SIZE after an unscaled real FFT. */
static void forward_transform(cpx *out, const float *in) {
  cpx z[480], Z[480];
  int k;
  const float norm = 1.0f / WINDOW_SIZE;
  for (k = 0; k < 480; k++) {
    z[k].r = in[2 * k];
    z[k].i = in[2 * k + 1];
  }
  fft480(Z, z);
  for (k = 0; k <= 480; k++) {
    cpx a = Z[k % 480], b = Z[(480 - k) % 480], e, o, t;
    b.i = -b.i; /* conj(Z[N-k]) */
    e.r = .5f * (a.r + b.r);
    e.i = .5f * (a.i + b.i);
    o.r = .5f * (a.r - b.r);
    o.i = .5f * (a.i - b.i);
    /* X[k] = E[k] - i W960^k O[k] */
    t = cmul(o, g_tw960[k]);
    out[k].r = (e.r + t.i) * norm;
    out[k].i = (e.i - t.r) * norm;
  }
}

/* upstream: denoise.c inverse_transform -- Hermitian extension, UNSCALED inverse DFT. */
static void inverse_transform(float *out, const cpx *in) {
  cpx z[480], Z[480];
  int k;
  /* x[n] = sum_{k<960} X[k] e^{+2 pi i k n/960};  with z[m] = x[2m] + i x[2m+1]:
   * z = IDFT480( (X[k] + conj(X[480-k])) + i e^{+2 pi i k/960} (X[k] - conj(X[480-k])) ).
   * IDFT480(Y)[m] = conj(DFT480(conj(Y)))[m]. */
  for (k = 0; k < 480; k++) {
    cpx a = in[k], b = in[480 - k], e, o, w, t;
    b.i = -b.i;
    e.r = a.r + b.r;
    e.i = a.i + b.i;
    o.r = a.r - b.r;
    o.i = a.i - b.i;
    w = g_tw960[k];
    w.i = -w.i; /* e^{+...} */
    t = cmul(o, w);
    /* e + i t, then conjugate for the forward-FFT trick */
    Z[k].r = e.r - t.i;
    Z[k].i = -(e.i + t.r);
  }
  fft480(z, Z);
  for (k = 0; k < 480; k++) {
    out[2 * k] = z[k].r;
    out[2 * k + 1] = -z[k].i;
  }
}

void rno_forward_transform(float *out, const float *in) {
  ensure_tables();
  forward_transform((cpx *)out, in);
}
void rno_inverse_transform(float *out, const float *in) {
  ensure_tables();
  inverse_transform(out, (const cpx *)in);
}

/* ------------------------------------------------------------------------------------------ */
/* model                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int nb_inputs, nb_neurons, activation;
  int8_t *weights; /* [in][out] */
  int8_t *bias;    /* [out] */
} dense_layer;
typedef struct {
  int nb_inputs, nb_neurons, activation;
  int8_t *input_weights;     /* [in][3N] */
  int8_t *recurrent_weights; /* [N][3N] */
  int8_t *bias;              /* [3N] */
} gru_layer;

struct rno_model {
  dense_layer input_dense;
  gru_layer vad_gru;
  dense_layer vad_output;
  gru_layer noise_gru;
  gru_layer denoise_gru;
  dense_layer denoise_output;
};

static void dense_alloc(dense_layer *l, int in, int out, int act) {
  l->nb_inputs = in;
  l->nb_neurons = out;
  l->activation = act;
  l->weights = (int8_t *)calloc((size_t)in * out, 1);
  l->bias = (int8_t *)calloc((size_t)out, 1);
}
static void gru_alloc(gru_layer *l, int in, int n, int act) {
  l->nb_inputs = in;
  l->nb_neurons = n;
  l->activation = act;
  l->input_weights = (int8_t *)calloc((size_t)in * 3 * n, 1);
  l->recurrent_weights = (int8_t *)calloc((size_t)n * 3 * n, 1);
  l->bias = (int8_t *)calloc((size_t)3 * n, 1);
}

void rno_model_free(rno_model *m) {
  if (!m) return;
  free(m->input_dense.weights);
  free(m->input_dense.bias);
  free(m->vad_output.weights);
  free(m->vad_output.bias);
  free(m->denoise_output.weights);
  free(m->denoise_output.bias);
  free(m->vad_gru.input_weights);
  free(m->vad_gru.recurrent_weights);
  free(m->vad_gru.bias);
  free(m->noise_gru.input_weights);
  free(m->noise_gru.recurrent_weights);
  free(m->noise_gru.bias);
  free(m->denoise_gru.input_weights);
  free(m->denoise_gru.recurrent_weights);
  free(m->denoise_gru.bias);
  free(m);
}

static rno_model *model_alloc_default_shapes(void) {
  /* SURVEY.md Appendix A.6 topology (rnnoise rnn_data.c) */
  rno_model *m = (rno_model *)calloc(1, sizeof(*m));
  dense_alloc(&m->input_dense, 42, 24, ACT_TANH);
  gru_alloc(&m->vad_gru, 24, 24, ACT_RELU);
  dense_alloc(&m->vad_output, 24, 1, ACT_SIGMOID);
  gru_alloc(&m->noise_gru, 90, 48, ACT_RELU);
  gru_alloc(&m->denoise_gru, 114, 96, ACT_RELU);
  dense_alloc(&m->denoise_output, 96, 22, ACT_SIGMOID);
  return m;
}

/* Seeded synthetic int8 weights (the real nnnoiseless weights are not available offline).
 * The generator is specified in DESIGN.md "Synthetic model" and implemented twice on purpose:
 * here and in crispy_b200/csrc/model.cpp; tests assert both emit identical bytes. */
static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void fill_q(int8_t *dst, size_t n, uint64_t *s, int amp, int offset) {
  size_t i;
  for (i = 0; i < n; i++) {
    uint64_t u = splitmix64(s);
    int a = (int)(u & 0xFFFF), b = (int)((u >> 16) & 0xFFFF), c = (int)((u >> 32) & 0xFFFF),
        d = (int)((u >> 48) & 0xFFFF);
    int t = a + b + c + d - 131070; /* ~N(0, 37837^2), Irwin-Hall n=4 */
    int w = (t * amp) / 131072 + offset;
    if (w > 127) w = 127;
    if (w < -127) w = -127;
    dst[i] = (int8_t)w;
  }
}

rno_model *rno_model_synthetic(uint64_t seed) {
  rno_model *m = model_alloc_default_shapes();
  uint64_t s = seed ^ 0xC215B200C215B200ull;
  fill_q(m->input_dense.weights, 42 * 24, &s, 80, 0);
  fill_q(m->input_dense.bias, 24, &s, 80, 0);
  fill_q(m->vad_gru.input_weights, 24 * 72, &s, 250, 0);
  fill_q(m->vad_gru.recurrent_weights, 24 * 72, &s, 110, 0);
  fill_q(m->vad_gru.bias, 72, &s, 100, 0);
  fill_q(m->vad_output.weights, 24, &s, 500, 0);
  fill_q(m->vad_output.bias, 1, &s, 40, 0);
  fill_q(m->noise_gru.input_weights, 90 * 144, &s, 160, 0);
  fill_q(m->noise_gru.recurrent_weights, 48 * 144, &s, 80, 0);
  fill_q(m->noise_gru.bias, 144, &s, 100, 0);
  fill_q(m->denoise_gru.input_weights, 114 * 288, &s, 160, 0);
  fill_q(m->denoise_gru.recurrent_weights, 96 * 288, &s, 60, 0);
  fill_q(m->denoise_gru.bias, 288, &s, 100, 0);
  fill_q(m->denoise_output.weights, 96 * 22, &s, 400, 0);
  fill_q(m->denoise_output.bias, 22, &s, 120, 40);
  return m;
}

/* binary blob "CRNSMDL1": magic[8]; then 6 layers in A.6 order, each
 * u32 kind(0 dense,1 gru) | u32 nb_inputs | u32 nb_neurons | u32 activation | int8 arrays */
static const char kMagic[8] = {'C', 'R', 'N', 'S', 'M', 'D', 'L', '1'};

static size_t put_u32(uint8_t *p, size_t off, size_t cap, uint32_t v) {
  if (p && off + 4 <= cap) {
    p[off] = (uint8_t)v;
    p[off + 1] = (uint8_t)(v >> 8);
    p[off + 2] = (uint8_t)(v >> 16);
    p[off + 3] = (uint8_t)(v >> 24);
  }
  return off + 4;
}
static size_t put_bytes(uint8_t *p, size_t off, size_t cap, const void *src, size_t n) {
  if (p && off + n <= cap) memcpy(p + off, src, n);
  return off + n;
}
static size_t put_dense(uint8_t *p, size_t off, size_t cap, const dense_layer *l) {
  off = put_u32(p, off, cap, 0);
  off = put_u32(p, off, cap, (uint32_t)l->nb_inputs);
  off = put_u32(p, off, cap, (uint32_t)l->nb_neurons);
  off = put_u32(p, off, cap, (uint32_t)l->activation);
  off = put_bytes(p, off, cap, l->weights, (size_t)l->nb_inputs * l->nb_neurons);
  off = put_bytes(p, off, cap, l->bias, (size_t)l->nb_neurons);
  return off;
}
static size_t put_gru(uint8_t *p, size_t off, size_t cap, const gru_layer *l) {
  size_t n3 = 3 * (size_t)l->nb_neurons;
  off = put_u32(p, off, cap, 1);
  off = put_u32(p, off, cap, (uint32_t)l->nb_inputs);
  off = put_u32(p, off, cap, (uint32_t)l->nb_neurons);
  off = put_u32(p, off, cap, (uint32_t)l->activation);
  off = put_bytes(p, off, cap, l->input_weights, (size_t)l->nb_inputs * n3);
  off = put_bytes(p, off, cap, l->recurrent_weights, (size_t)l->nb_neurons * n3);
  off = put_bytes(p, off, cap, l->bias, n3);
  return off;
}

size_t rno_model_to_bytes(const rno_model *m, void *buf, size_t cap) {
  uint8_t *p = (uint8_t *)buf;
  size_t off = put_bytes(p, 0, cap, kMagic, 8);
  off = put_dense(p, off, cap, &m->input_dense);
  off = put_gru(p, off, cap, &m->vad_gru);
  off = put_dense(p, off, cap, &m->vad_output);
  off = put_gru(p, off, cap, &m->noise_gru);
  off = put_gru(p, off, cap, &m->denoise_gru);
  off = put_dense(p, off, cap, &m->denoise_output);
  return off;
}

typedef struct {
  const uint8_t *p;
  size_t len, off;
  int err;
} rd;
static uint32_t get_u32(rd *r) {
  uint32_t v;
  if (r->off + 4 > r->len) {
    r->err = 1;
    return 0;
  }
  v = (uint32_t)r->p[r->off] | ((uint32_t)r->p[r->off + 1] << 8) |
      ((uint32_t)r->p[r->off + 2] << 16) | ((uint32_t)r->p[r->off + 3] << 24);
  r->off += 4;
  return v;
}
static void get_bytes(rd *r, void *dst, size_t n) {
  if (r->off + n > r->len) {
    r->err = 1;
    return;
  }
  memcpy(dst, r->p + r->off, n);
  r->off += n;
}
static void get_dense(rd *r, dense_layer *l) {
  uint32_t kind = get_u32(r), in = get_u32(r), out = get_u32(r), act = get_u32(r);
  if (r->err || kind != 0 || (int)in != l->nb_inputs || (int)out != l->nb_neurons || act > 2) {
    r->err = 1;
    return;
  }
  l->activation = (int)act;
  get_bytes(r, l->weights, (size_t)in * out);
  get_bytes(r, l->bias, out);
}
static void get_gru(rd *r, gru_layer *l) {
  uint32_t kind = get_u32(r), in = get_u32(r), n = get_u32(r), act = get_u32(r);
  if (r->err || kind != 1 || (int)in != l->nb_inputs || (int)n != l->nb_neurons || act > 2) {
    r->err = 1;
    return;
  }
  l->activation = (int)act;
  get_bytes(r, l->input_weights, (size_t)in * 3 * n);
  get_bytes(r, l->recurrent_weights, (size_t)n * 3 * n);
  get_bytes(r, l->bias, (size_t)3 * n);
}

/* rnnoise-nu text model: "rnnoise-nu model file version 1", then per layer a header line
 * (dense: in out act; gru: in out act) followed by whitespace-separated integers.  Layer order in
 * that format: input_dense, vad_gru, noise_gru, denoise_gru, denoise_output, vad_output. */
typedef struct {
  const char *p, *end;
  int err;
} trd;
static long text_int(trd *t) {
  char *e;
  long v;
  while (t->p < t->end && (*t->p == ' ' || *t->p == '\n' || *t->p == '\r' || *t->p == '\t')) t->p++;
  if (t->p >= t->end) {
    t->err = 1;
    return 0;
  }
  v = strtol(t->p, &e, 10);
  if (e == t->p) {
    t->err = 1;
    return 0;
  }
  t->p = e;
  return v;
}
static void text_arr(trd *t, int8_t *dst, size_t n) {
  size_t i;
  for (i = 0; i < n && !t->err; i++) {
    long v = text_int(t);
    if (v > 127) v = 127;
    if (v < -128) v = -128;
    dst[i] = (int8_t)v;
  }
}
static void text_dense(trd *t, dense_layer *l) {
  long in = text_int(t), out = text_int(t), act = text_int(t);
  if (t->err || in != l->nb_inputs || out != l->nb_neurons || act < 0 || act > 2) {
    t->err = 1;
    return;
  }
  l->activation = (int)act;
  text_arr(t, l->weights, (size_t)in * out);
  text_arr(t, l->bias, (size_t)out);
}
static void text_gru(trd *t, gru_layer *l) {
  long in = text_int(t), n = text_int(t), act = text_int(t);
  if (t->err || in != l->nb_inputs || n != l->nb_neurons || act < 0 || act > 2) {
    t->err = 1;
    return;
  }
  l->activation = (int)act;
  text_arr(t, l->input_weights, (size_t)in * 3 * n);
  text_arr(t, l->recurrent_weights, (size_t)n * 3 * n);
  text_arr(t, l->bias, (size_t)3 * n);
}

rno_model *rno_model_from_bytes(const void *blob, size_t len) {
  static const char kText[] = "rnnoise-nu model file version 1";
  rno_model *m;
  if (!blob) return NULL;
  m = model_alloc_default_shapes();
  if (len >= 8 && memcmp(blob, kMagic, 8) == 0) {
    rd r;
    r.p = (const uint8_t *)blob;
    r.len = len;
    r.off = 8;
    r.err = 0;
    get_dense(&r, &m->input_dense);
    get_gru(&r, &m->vad_gru);
    get_dense(&r, &m->vad_output);
    get_gru(&r, &m->noise_gru);
    get_gru(&r, &m->denoise_gru);
    get_dense(&r, &m->denoise_output);
    if (r.err || r.off != len) {
      rno_model_free(m);
      return NULL;
    }
    return m;
  }
  if (len >= sizeof(kText) - 1 && memcmp(blob, kText, sizeof(kText) - 1) == 0) {
    trd t;
    t.p = (const char *)blob + sizeof(kText) - 1;
    t.end = (const char *)blob + len;
    t.err = 0;
    text_dense(&t, &m->input_dense);
    text_gru(&t, &m->vad_gru);
    text_gru(&t, &m->noise_gru);
    text_gru(&t, &m->denoise_gru);
    text_dense(&t, &m->denoise_output);
    text_dense(&t, &m->vad_output);
    if (t.err) {
      rno_model_free(m);
      return NULL;
    }
    return m;
  }
  rno_model_free(m);
  return NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* state (nnnoiseless DenoiseState; reference type at audio.rs:203)                            */
/* ------------------------------------------------------------------------------------------ */
struct rno_state {
  const rno_model *model;
  float analysis_mem[FRAME_SIZE];
  float cepstral_mem[CEPS_MEM][NB_BANDS];
  int memid;
  float synthesis_mem[FRAME_SIZE];
  float pitch_buf[PITCH_BUF_SIZE];
  float last_gain;
  int last_period;
  float mem_hp_x[2];
  float lastg[NB_BANDS];
  float vad_gru_state[24];
  float noise_gru_state[48];
  float denoise_gru_state[96];
  rno_debug dbg;
  float graw[NB_BANDS]; /* test taps: the band gains as the RNN emitted them, and ... */
  float branch_margin;  /* ... min over bands of |Exp - g| of the last frame (1e30 on silent frames) */
};

rno_state *rno_create(const rno_model *m) {
  rno_state *st;
  ensure_tables();
  st = (rno_state *)calloc(1, sizeof(*st));
  st->model = m;
  return st;
}
void rno_reset(rno_state *st) {
  const rno_model *m = st->model;
  memset(st, 0, sizeof(*st));
  st->model = m;
}
void rno_destroy(rno_state *st) { free(st); }
void rno_get_debug(const rno_state *st, rno_debug *dbg) { *dbg = st->dbg; }

/* ------------------------------------------------------------------------------------------ */
/* a6: biquad high-pass.  upstream denoise.c biquad() (double intermediates, f32 memory).      */
/* ------------------------------------------------------------------------------------------ */
static void biquad(float *y, float mem[2], const float *x, const float *b, const float *a, int N) {
  int i;
  for (i = 0; i < N; i++) {
    float xi = x[i];
    float yi = x[i] + mem[0];
    mem[0] = (float)(mem[1] + (b[0] * (double)xi - a[0] * (double)yi));
    mem[1] = (float)(b[1] * (double)xi - a[1] * (double)yi);
    y[i] = yi;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* a7/a8: window, band energy / correlation, interpolation, DCT (denoise.c)                   */
/* ------------------------------------------------------------------------------------------ */
static void apply_window(float *x) {
  int i;
  for (i = 0; i < FRAME_SIZE; i++) {
    x[i] *= g_half_window[i];
    x[WINDOW_SIZE - 1 - i] *= g_half_window[i];
  }
}

static void compute_band_energy(float *bandE, const cpx *X) {
  int i, j;
  float sum[NB_BANDS] = {0};
  for (i = 0; i < NB_BANDS - 1; i++) {
    int band_size = (eband5ms[i + 1] - eband5ms[i]) << FRAME_SIZE_SHIFT;
    for (j = 0; j < band_size; j++) {
      float frac = (float)j / band_size;
      int k = (eband5ms[i] << FRAME_SIZE_SHIFT) + j;
      float tmp = X[k].r * X[k].r + X[k].i * X[k].i;
      sum[i] += (1 - frac) * tmp;
      sum[i + 1] += frac * tmp;
    }
  }
  sum[0] *= 2;
  sum[NB_BANDS - 1] *= 2;
  for (i = 0; i < NB_BANDS; i++) bandE[i] = sum[i];
}

static void compute_band_corr(float *bandE, const cpx *X, const cpx *P) {
  int i, j;
  float sum[NB_BANDS] = {0};
  for (i = 0; i < NB_BANDS - 1; i++) {
    int band_size = (eband5ms[i + 1] - eband5ms[i]) << FRAME_SIZE_SHIFT;
    for (j = 0; j < band_size; j++) {
      float frac = (float)j / band_size;
      int k = (eband5ms[i] << FRAME_SIZE_SHIFT) + j;
      float tmp = X[k].r * P[k].r + X[k].i * P[k].i;
      sum[i] += (1 - frac) * tmp;
      sum[i + 1] += frac * tmp;
    }
  }
  sum[0] *= 2;
  sum[NB_BANDS - 1] *= 2;
  for (i = 0; i < NB_BANDS; i++) bandE[i] = sum[i];
}

static void interp_band_gain(float *g, const float *bandE) {
  int i, j;
  memset(g, 0, FREQ_SIZE * sizeof(float));
  for (i = 0; i < NB_BANDS - 1; i++) {
    int band_size = (eband5ms[i + 1] - eband5ms[i]) << FRAME_SIZE_SHIFT;
    for (j = 0; j < band_size; j++) {
      float frac = (float)j / band_size;
      g[(eband5ms[i] << FRAME_SIZE_SHIFT) + j] = (1 - frac) * bandE[i] + frac * bandE[i + 1];
    }
  }
}

static void dct(float *out, const float *in) {
  int i, j;
  const float scale = (float)sqrt(2. / 22);
  for (i = 0; i < NB_BANDS; i++) {
    float sum = 0;
    for (j = 0; j < NB_BANDS; j++) sum += in[j] * g_dct_table[j * NB_BANDS + i];
    out[i] = sum * scale;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* a9: pitch_downsample (pitch.c) + _celt_autocorr/_celt_lpc (celt_lpc.c) + celt_fir5          */
/* ------------------------------------------------------------------------------------------ */
/* Summation-order policy of the pitch path's inner products (process-wide; set it before any thread runs).
 *   0  sequential, ascending index: xiph/rnnoise's C (celt_inner_prod, xcorr_kernel per lag).  THE DEFAULT, and the
 *      order the CUDA kernels reproduce bit for bit.
 *   1  four interleaved partial sums (s[j & 3] += x[j] y[j], then ((s0 + s1) + s2) + s3, tail sequential) in
 *      celt_inner_prod / dual_inner_prod only -- the shape of a hand-unrolled Rust inner product such as a port may
 *      use; the cross-correlation kernels stay sequential per lag.
 *   2  policy 1, and the same four-way split inside celt_pitch_xcorr (hence _celt_autocorr) as well.
 * Nothing here claims to be what nnnoiseless does: the crate source is absent.  The switch exists to MEASURE how
 * often a different but equally legitimate float32 order flips a discrete pitch decision (tests/test_oracle.py,
 * profiles/r2_parity.json): the decisions are bit-exact against the oracle's order only. */
static int g_sum_policy = 0;
/* Measurement aid (tests/tools only): perturb the pitch filter's inputs the way another float32 implementation
 * would -- Exp absolutely and relatively (Exp (1 + rel_exp) + abs_exp: a normalised correlation carries the rounding
 * of a sum of positive and negative terms), the band gains in the logit domain (g + logit_g g (1 - g): what an error
 * of logit_g in the output layer's pre-activation does) -- inside the pitch_filter call only, nothing that feeds the
 * state.  Exposes the frames on which RNNoise's pitch filter is discontinuous (Exp > g ? 1 : ...). */
static float g_pf_dexp = 0.f, g_pf_dg = 0.f, g_pf_aexp = 0.f;
void rno_get_raw_gains(const rno_state *st, float *graw) { memcpy(graw, st->graw, sizeof(st->graw)); }
void rno_set_pf_perturb(float rel_exp, float logit_g, float abs_exp) {
  g_pf_dexp = rel_exp;
  g_pf_dg = logit_g;
  g_pf_aexp = abs_exp;
}
void rno_set_sum_policy(int policy) { g_sum_policy = policy < 0 ? 0 : (policy > 2 ? 2 : policy); }
int rno_get_sum_policy(void) { return g_sum_policy; }

static float dot4(const float *x, const float *y, int N) {
  float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s;
  int i, n4 = N & ~3;
  for (i = 0; i < n4; i += 4) {
    s0 += x[i] * y[i];
    s1 += x[i + 1] * y[i + 1];
    s2 += x[i + 2] * y[i + 2];
    s3 += x[i + 3] * y[i + 3];
  }
  s = ((s0 + s1) + s2) + s3;
  for (; i < N; i++) s += x[i] * y[i];
  return s;
}

static void celt_pitch_xcorr(const float *x, const float *y, float *xcorr, int len, int max_pitch) {
  int i, j;
  if (g_sum_policy >= 2) {
    for (i = 0; i < max_pitch; i++) xcorr[i] = dot4(x, y + i, len);
    return;
  }
  /* eight lags at a time, each lag's sum still accumulated in ascending j with a rounded product and a rounded
   * add: bit-identical to the lag-by-lag loop, but the compiler can keep the eight sums in one vector register */
  for (i = 0; i + 8 <= max_pitch; i += 8) {
    float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float *yy = y + i;
    for (j = 0; j < len; j++) {
      const float xj = x[j];
      int l;
      for (l = 0; l < 8; l++) sum[l] += xj * yy[j + l];
    }
    for (j = 0; j < 8; j++) xcorr[i + j] = sum[j];
  }
  for (; i < max_pitch; i++) {
    float sum = 0;
    for (j = 0; j < len; j++) sum += x[j] * y[i + j];
    xcorr[i] = sum;
  }
}
static float celt_inner_prod(const float *x, const float *y, int N) {
  int i;
  float xy = 0;
  if (g_sum_policy >= 1) return dot4(x, y, N);
  for (i = 0; i < N; i++) xy += x[i] * y[i];
  return xy;
}
static void dual_inner_prod(const float *x, const float *y01, const float *y02, int N, float *xy1,
                            float *xy2) {
  int i;
  float a = 0, b = 0;
  if (g_sum_policy >= 1) {
    *xy1 = dot4(x, y01, N);
    *xy2 = dot4(x, y02, N);
    return;
  }
  for (i = 0; i < N; i++) {
    a += x[i] * y01[i];
    b += x[i] * y02[i];
  }
  *xy1 = a;
  *xy2 = b;
}

static void celt_autocorr(const float *x, float *ac, int lag, int n) {
  int i, k;
  int fastN = n - lag;
  celt_pitch_xcorr(x, x, ac, fastN, lag + 1);
  for (k = 0; k <= lag; k++) {
    float d = 0;
    for (i = k + fastN; i < n; i++) d += x[i] * x[i - k];
    ac[k] += d;
  }
}

static void celt_lpc(float *lpc, const float *ac, int p) {
  int i, j;
  float error = ac[0];
  for (i = 0; i < p; i++) lpc[i] = 0;
  if (ac[0] != 0) {
    for (i = 0; i < p; i++) {
      float rr = 0, r;
      for (j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
      rr += ac[i + 1];
      r = -rr / error;
      lpc[i] = r;
      for (j = 0; j < (i + 1) >> 1; j++) {
        float tmp1 = lpc[j], tmp2 = lpc[i - 1 - j];
        lpc[j] = tmp1 + r * tmp2;
        lpc[i - 1 - j] = tmp2 + r * tmp1;
      }
      error = error - r * r * error;
      if (error < .001f * ac[0]) break;
    }
  }
}

static void celt_fir5(const float *x, const float *num, float *y, int N, float *mem) {
  int i;
  float num0 = num[0], num1 = num[1], num2 = num[2], num3 = num[3], num4 = num[4];
  float mem0 = mem[0], mem1 = mem[1], mem2 = mem[2], mem3 = mem[3], mem4 = mem[4];
  for (i = 0; i < N; i++) {
    float sum = x[i];
    sum += num0 * mem0;
    sum += num1 * mem1;
    sum += num2 * mem2;
    sum += num3 * mem3;
    sum += num4 * mem4;
    mem4 = mem3;
    mem3 = mem2;
    mem2 = mem1;
    mem1 = mem0;
    mem0 = x[i];
    y[i] = sum;
  }
  mem[0] = mem0;
  mem[1] = mem1;
  mem[2] = mem2;
  mem[3] = mem3;
  mem[4] = mem4;
}

static void pitch_downsample(const float *x, float *x_lp, int len) {
  int i;
  float ac[5];
  float tmp = 1.f;
  float lpc[4], mem[5] = {0, 0, 0, 0, 0};
  float lpc2[5];
  float c1 = .8f;
  for (i = 1; i < len >> 1; i++)
    x_lp[i] = .5f * (.5f * (x[2 * i - 1] + x[2 * i + 1]) + x[2 * i]);
  x_lp[0] = .5f * (.5f * x[1] + x[0]);
  celt_autocorr(x_lp, ac, 4, len >> 1);
  ac[0] *= 1.0001f;
  for (i = 1; i <= 4; i++) ac[i] -= ac[i] * (.008f * i) * (.008f * i);
  celt_lpc(lpc, ac, 4);
  for (i = 0; i < 4; i++) {
    tmp = .9f * tmp;
    lpc[i] = lpc[i] * tmp;
  }
  lpc2[0] = lpc[0] + .8f;
  lpc2[1] = lpc[1] + c1 * lpc[0];
  lpc2[2] = lpc[2] + c1 * lpc[1];
  lpc2[3] = lpc[3] + c1 * lpc[2];
  lpc2[4] = c1 * lpc[3];
  celt_fir5(x_lp, lpc2, x_lp, len >> 1, mem);
}

/* ------------------------------------------------------------------------------------------ */
/* a10: pitch_search / find_best_pitch (pitch.c)                                              */
/* ------------------------------------------------------------------------------------------ */
static void find_best_pitch(const float *xcorr, const float *y, int len, int max_pitch,
                            int *best_pitch) {
  int i, j;
  float Syy = 1;
  float best_num[2] = {-1, -1};
  float best_den[2] = {0, 0};
  best_pitch[0] = 0;
  best_pitch[1] = 1;
  for (j = 0; j < len; j++) Syy += y[j] * y[j];
  for (i = 0; i < max_pitch; i++) {
    if (xcorr[i] > 0) {
      float num;
      float xcorr16 = xcorr[i];
      xcorr16 *= 1e-12f;
      num = xcorr16 * xcorr16;
      if (num * best_den[1] > best_num[1] * Syy) {
        if (num * best_den[0] > best_num[0] * Syy) {
          best_num[1] = best_num[0];
          best_den[1] = best_den[0];
          best_pitch[1] = best_pitch[0];
          best_num[0] = num;
          best_den[0] = Syy;
          best_pitch[0] = i;
        } else {
          best_num[1] = num;
          best_den[1] = Syy;
          best_pitch[1] = i;
        }
      }
    }
    Syy += y[i + len] * y[i + len] - y[i] * y[i];
    if (Syy < 1) Syy = 1;
  }
}

static void pitch_search(const float *x_lp, const float *y, int len, int max_pitch, int *pitch) {
  int i, j;
  int lag = len + max_pitch;
  int best_pitch[2] = {0, 0};
  int offset;
  float x_lp4[PITCH_FRAME_SIZE >> 2];
  float y_lp4[(PITCH_FRAME_SIZE + PITCH_MAX_PERIOD) >> 2];
  float xcorr[PITCH_MAX_PERIOD >> 1];
  for (j = 0; j < len >> 2; j++) x_lp4[j] = x_lp[2 * j];
  for (j = 0; j < lag >> 2; j++) y_lp4[j] = y[2 * j];
  celt_pitch_xcorr(x_lp4, y_lp4, xcorr, len >> 2, max_pitch >> 2);
  find_best_pitch(xcorr, y_lp4, len >> 2, max_pitch >> 2, best_pitch);
  for (i = 0; i < max_pitch >> 1; i++) {
    float sum;
    xcorr[i] = 0;
    if (abs(i - 2 * best_pitch[0]) > 2 && abs(i - 2 * best_pitch[1]) > 2) continue;
    sum = celt_inner_prod(x_lp, y + i, len >> 1);
    xcorr[i] = sum < -1 ? -1 : sum;
  }
  find_best_pitch(xcorr, y, len >> 1, max_pitch >> 1, best_pitch);
  if (best_pitch[0] > 0 && best_pitch[0] < (max_pitch >> 1) - 1) {
    float a = xcorr[best_pitch[0] - 1], b = xcorr[best_pitch[0]], c = xcorr[best_pitch[0] + 1];
    if ((c - a) > .7f * (b - a))
      offset = 1;
    else if ((a - c) > .7f * (b - c))
      offset = -1;
    else
      offset = 0;
  } else {
    offset = 0;
  }
  *pitch = 2 * best_pitch[0] - offset;
}

/* ------------------------------------------------------------------------------------------ */
/* a11: remove_doubling (pitch.c)                                                             */
/* ------------------------------------------------------------------------------------------ */
static float compute_pitch_gain(float xy, float xx, float yy) {
  return xy / (float)sqrt(1 + xx * yy);
}
static const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};

static float remove_doubling(const float *x, int maxperiod, int minperiod, int N, int *T0_,
                             int prev_period, float prev_gain) {
  int k, i, T, T0;
  float g, g0, pg;
  float xy, xx, yy, xy2;
  float xcorr[3];
  float best_xy, best_yy;
  int offset;
  int minperiod0 = minperiod;
  float yy_lookup[(PITCH_MAX_PERIOD >> 1) + 1];
  maxperiod /= 2;
  minperiod /= 2;
  *T0_ /= 2;
  prev_period /= 2;
  N /= 2;
  x += maxperiod;
  if (*T0_ >= maxperiod) *T0_ = maxperiod - 1;
  T = T0 = *T0_;
  dual_inner_prod(x, x, x - T0, N, &xx, &xy);
  yy_lookup[0] = xx;
  yy = xx;
  for (i = 1; i <= maxperiod; i++) {
    yy = yy + x[-i] * x[-i] - x[N - i] * x[N - i];
    yy_lookup[i] = yy < 0 ? 0 : yy;
  }
  yy = yy_lookup[T0];
  best_xy = xy;
  best_yy = yy;
  g = g0 = compute_pitch_gain(xy, xx, yy);
  for (k = 2; k <= 15; k++) {
    int T1, T1b;
    float g1, cont = 0, thresh;
    T1 = (2 * T0 + k) / (2 * k);
    if (T1 < minperiod) break;
    if (k == 2) {
      if (T1 + T0 > maxperiod)
        T1b = T0;
      else
        T1b = T0 + T1;
    } else {
      T1b = (2 * second_check[k] * T0 + k) / (2 * k);
    }
    dual_inner_prod(x, &x[-T1], &x[-T1b], N, &xy, &xy2);
    xy = .5f * (xy + xy2);
    yy = .5f * (yy_lookup[T1] + yy_lookup[T1b]);
    g1 = compute_pitch_gain(xy, xx, yy);
    if (abs(T1 - prev_period) <= 1)
      cont = prev_gain;
    else if (abs(T1 - prev_period) <= 2 && 5 * k * k < T0)
      cont = .5f * prev_gain;
    else
      cont = 0;
    thresh = .7f * g0 - cont;
    if (thresh < .3f) thresh = .3f;
    if (T1 < 3 * minperiod) {
      thresh = .85f * g0 - cont;
      if (thresh < .4f) thresh = .4f;
    } else if (T1 < 2 * minperiod) {
      thresh = .9f * g0 - cont;
      if (thresh < .5f) thresh = .5f;
    }
    if (g1 > thresh) {
      best_xy = xy;
      best_yy = yy;
      T = T1;
      g = g1;
    }
  }
  if (best_xy < 0) best_xy = 0;
  if (best_yy <= best_xy)
    pg = 1.f;
  else
    pg = best_xy / (best_yy + 1);
  for (k = 0; k < 3; k++) xcorr[k] = celt_inner_prod(x, x - (T + k - 1), N);
  if ((xcorr[2] - xcorr[0]) > .7f * (xcorr[1] - xcorr[0]))
    offset = 1;
  else if ((xcorr[0] - xcorr[2]) > .7f * (xcorr[1] - xcorr[2]))
    offset = -1;
  else
    offset = 0;
  if (pg > g) pg = g;
  *T0_ = 2 * T + offset;
  if (*T0_ < minperiod0) *T0_ = minperiod0;
  return pg;
}

/* ------------------------------------------------------------------------------------------ */
/* a7, a12, a13: frame_analysis + compute_frame_features (denoise.c / nnnoiseless features.rs) */
/* ------------------------------------------------------------------------------------------ */
static void frame_analysis(rno_state *st, cpx *X, float *Ex, const float *in) {
  float x[WINDOW_SIZE];
  memcpy(x, st->analysis_mem, FRAME_SIZE * sizeof(float));
  memcpy(x + FRAME_SIZE, in, FRAME_SIZE * sizeof(float));
  memcpy(st->analysis_mem, in, FRAME_SIZE * sizeof(float));
  apply_window(x);
  forward_transform(X, x);
  compute_band_energy(Ex, X);
}

static int compute_frame_features(rno_state *st, cpx *X, cpx *P, float *Ex, float *Ep, float *Exp,
                                  float *features, const float *in) {
  int i, j, k;
  float E = 0;
  float *ceps_0, *ceps_1, *ceps_2;
  float spec_variability = 0;
  float Ly[NB_BANDS];
  float p[WINDOW_SIZE];
  float pitch_buf[PITCH_BUF_SIZE >> 1];
  int pitch_index;
  float gain;
  float tmp[NB_BANDS];
  float follow, logMax;
  frame_analysis(st, X, Ex, in);
  memmove(st->pitch_buf, &st->pitch_buf[FRAME_SIZE], (PITCH_BUF_SIZE - FRAME_SIZE) * sizeof(float));
  memcpy(&st->pitch_buf[PITCH_BUF_SIZE - FRAME_SIZE], in, FRAME_SIZE * sizeof(float));
  pitch_downsample(st->pitch_buf, pitch_buf, PITCH_BUF_SIZE);
  pitch_search(pitch_buf + (PITCH_MAX_PERIOD >> 1), pitch_buf, PITCH_FRAME_SIZE,
               PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD, &pitch_index);
  pitch_index = PITCH_MAX_PERIOD - pitch_index;
  gain = remove_doubling(pitch_buf, PITCH_MAX_PERIOD, PITCH_MIN_PERIOD, PITCH_FRAME_SIZE,
                         &pitch_index, st->last_period, st->last_gain);
  st->last_period = pitch_index;
  st->last_gain = gain;
  st->dbg.pitch_index = pitch_index;
  st->dbg.pitch_gain = gain;
  for (i = 0; i < WINDOW_SIZE; i++)
    p[i] = st->pitch_buf[PITCH_BUF_SIZE - WINDOW_SIZE - pitch_index + i];
  apply_window(p);
  forward_transform(P, p);
  compute_band_energy(Ep, P);
  compute_band_corr(Exp, X, P);
  for (i = 0; i < NB_BANDS; i++) Exp[i] = Exp[i] / (float)sqrt(.001 + Ex[i] * Ep[i]);
  dct(tmp, Exp);
  for (i = 0; i < NB_DELTA_CEPS; i++) features[NB_BANDS + 2 * NB_DELTA_CEPS + i] = tmp[i];
  features[NB_BANDS + 2 * NB_DELTA_CEPS] -= 1.3f;
  features[NB_BANDS + 2 * NB_DELTA_CEPS + 1] -= 0.9f;
  features[NB_BANDS + 3 * NB_DELTA_CEPS] = .01f * (pitch_index - 300);
  logMax = -2;
  follow = -2;
  for (i = 0; i < NB_BANDS; i++) {
    Ly[i] = (float)log10(1e-2 + Ex[i]);
    Ly[i] = fmaxf(logMax - 7, fmaxf(follow - 1.5f, Ly[i]));
    logMax = fmaxf(logMax, Ly[i]);
    follow = fmaxf(follow - 1.5f, Ly[i]);
    E += Ex[i];
  }
  if (E < 0.04f) {
    /* silence: features zeroed, cepstral history / RNN state untouched */
    memset(features, 0, NB_FEATURES * sizeof(float));
    return 1;
  }
  dct(features, Ly);
  features[0] -= 12;
  features[1] -= 4;
  ceps_0 = st->cepstral_mem[st->memid];
  ceps_1 = (st->memid < 1) ? st->cepstral_mem[CEPS_MEM + st->memid - 1] : st->cepstral_mem[st->memid - 1];
  ceps_2 = (st->memid < 2) ? st->cepstral_mem[CEPS_MEM + st->memid - 2] : st->cepstral_mem[st->memid - 2];
  for (i = 0; i < NB_BANDS; i++) ceps_0[i] = features[i];
  st->memid++;
  for (i = 0; i < NB_DELTA_CEPS; i++) {
    features[i] = ceps_0[i] + ceps_1[i] + ceps_2[i];
    features[NB_BANDS + i] = ceps_0[i] - ceps_2[i];
    features[NB_BANDS + NB_DELTA_CEPS + i] = ceps_0[i] - 2 * ceps_1[i] + ceps_2[i];
  }
  if (st->memid == CEPS_MEM) st->memid = 0;
  for (i = 0; i < CEPS_MEM; i++) {
    float mindist = 1e15f;
    for (j = 0; j < CEPS_MEM; j++) {
      float dist = 0;
      for (k = 0; k < NB_BANDS; k++) {
        float t = st->cepstral_mem[i][k] - st->cepstral_mem[j][k];
        dist += t * t;
      }
      if (j != i) mindist = fminf(mindist, dist);
    }
    spec_variability += mindist;
  }
  features[NB_BANDS + 3 * NB_DELTA_CEPS + 1] = spec_variability / CEPS_MEM - 2.1f;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a14: RNN (rnn.c / nnnoiseless rnn.rs)                                                      */
/* ------------------------------------------------------------------------------------------ */
float rno_tansig_approx(float x) {
  int i;
  float y, dy;
  float sign = 1;
  ensure_tables();
  if (!(x < 8)) return 1;
  if (!(x > -8)) return -1;
  if (x < 0) {
    x = -x;
    sign = -1;
  }
  i = (int)floor(.5f + 25 * x);
  x -= .04f * i;
  y = g_tansig[i];
  dy = 1 - y * y;
  y = y + x * dy * (1 - y * x);
  return sign * y;
}
float rno_sigmoid_approx(float x) { return .5f + .5f * rno_tansig_approx(.5f * x); }
static float relu(float x) { return x < 0 ? 0 : x; }
static float activate(int act, float x) {
  if (act == ACT_SIGMOID) return rno_sigmoid_approx(x);
  if (act == ACT_TANH) return rno_tansig_approx(x);
  return relu(x);
}

/* sum[i] = bias[i] + sum_j W[j*stride + i] * in[j], j ascending: upstream walks j innermost per neuron i; here j is
 * the outer loop and every neuron keeps its own running sum, so each sum sees the same additions in the same order
 * (bit-identical) while the weight rows are read contiguously and the i loop vectorises. */
static void accum_rows(float *sum, const int8_t *w, int stride, const float *in, int M, int N) {
  int i, j;
  for (j = 0; j < M; j++) {
    const int8_t *row = w + (size_t)j * stride;
    const float v = in[j];
    for (i = 0; i < N; i++) sum[i] += row[i] * v;
  }
}

static void compute_dense(const dense_layer *layer, float *output, const float *input) {
  int i;
  int M = layer->nb_inputs, N = layer->nb_neurons;
  float sum[96];
  for (i = 0; i < N; i++) sum[i] = layer->bias[i];
  accum_rows(sum, layer->weights, N, input, M, N);
  for (i = 0; i < N; i++) output[i] = activate(layer->activation, WEIGHTS_SCALE * sum[i]);
}

static void compute_gru(const gru_layer *gru, float *state, const float *input) {
  int i, j;
  float z[96], r[96], h[96];
  int M = gru->nb_inputs, N = gru->nb_neurons, stride = 3 * N;
  float sum[2 * 96];
  /* update and reset gates share their inputs: columns 0..2N of the weight rows in one sweep */
  for (i = 0; i < 2 * N; i++) sum[i] = gru->bias[i];
  accum_rows(sum, gru->input_weights, stride, input, M, 2 * N);
  accum_rows(sum, gru->recurrent_weights, stride, state, N, 2 * N);
  for (i = 0; i < N; i++) z[i] = rno_sigmoid_approx(WEIGHTS_SCALE * sum[i]);
  for (i = 0; i < N; i++) r[i] = rno_sigmoid_approx(WEIGHTS_SCALE * sum[N + i]);
  /* candidate: upstream multiplies weight * state[j] * r[j] left to right, i.e. (w * state[j]) * r[j] */
  for (i = 0; i < N; i++) sum[i] = gru->bias[2 * N + i];
  accum_rows(sum, gru->input_weights + 2 * N, stride, input, M, N);
  for (j = 0; j < N; j++) {
    const int8_t *row = gru->recurrent_weights + 2 * N + (size_t)j * stride;
    const float sj = state[j], rj = r[j];
    for (i = 0; i < N; i++) sum[i] += row[i] * sj * rj;
  }
  for (i = 0; i < N; i++) {
    const float c = activate(gru->activation, WEIGHTS_SCALE * sum[i]);
    h[i] = z[i] * state[i] + (1 - z[i]) * c;
  }
  for (i = 0; i < N; i++) state[i] = h[i];
}

static void compute_rnn(rno_state *st, float *gains, float *vad, const float *input) {
  int i;
  const rno_model *m = st->model;
  float dense_out[24];
  float noise_input[90];
  float denoise_input[114];
  compute_dense(&m->input_dense, dense_out, input);
  compute_gru(&m->vad_gru, st->vad_gru_state, dense_out);
  compute_dense(&m->vad_output, vad, st->vad_gru_state);
  for (i = 0; i < 24; i++) noise_input[i] = dense_out[i];
  for (i = 0; i < 24; i++) noise_input[i + 24] = st->vad_gru_state[i];
  for (i = 0; i < 42; i++) noise_input[i + 48] = input[i];
  compute_gru(&m->noise_gru, st->noise_gru_state, noise_input);
  for (i = 0; i < 24; i++) denoise_input[i] = st->vad_gru_state[i];
  for (i = 0; i < 48; i++) denoise_input[i + 24] = st->noise_gru_state[i];
  for (i = 0; i < 42; i++) denoise_input[i + 72] = input[i];
  compute_gru(&m->denoise_gru, st->denoise_gru_state, denoise_input);
  compute_dense(&m->denoise_output, gains, st->denoise_gru_state);
}

/* ------------------------------------------------------------------------------------------ */
/* a15/a16: pitch_filter, frame_synthesis, process_frame (denoise.c / nnnoiseless denoise.rs)  */
/* ------------------------------------------------------------------------------------------ */
static void pitch_filter(cpx *X, const cpx *P, const float *Ex, const float *Ep, const float *Exp,
                         const float *g) {
  int i;
  float r[NB_BANDS];
  float rf[FREQ_SIZE];
  float newE[NB_BANDS];
  float norm[NB_BANDS];
  float normf[FREQ_SIZE];
  for (i = 0; i < NB_BANDS; i++) {
    if (Exp[i] > g[i])
      r[i] = 1;
    else
      r[i] = (Exp[i] * Exp[i]) * (1 - g[i] * g[i]) / (.001f + (g[i] * g[i]) * (1 - Exp[i] * Exp[i]));
    r[i] = (float)sqrt(fminf(1, fmaxf(0, r[i])));
    r[i] *= (float)sqrt(Ex[i] / (1e-8 + Ep[i]));
  }
  interp_band_gain(rf, r);
  for (i = 0; i < FREQ_SIZE; i++) {
    X[i].r += rf[i] * P[i].r;
    X[i].i += rf[i] * P[i].i;
  }
  compute_band_energy(newE, X);
  for (i = 0; i < NB_BANDS; i++) norm[i] = (float)sqrt(Ex[i] / (1e-8 + newE[i]));
  interp_band_gain(normf, norm);
  for (i = 0; i < FREQ_SIZE; i++) {
    X[i].r *= normf[i];
    X[i].i *= normf[i];
  }
}

static void frame_synthesis(rno_state *st, float *out, const cpx *y) {
  float x[WINDOW_SIZE];
  int i;
  inverse_transform(x, y);
  apply_window(x);
  for (i = 0; i < FRAME_SIZE; i++) out[i] = x[i] + st->synthesis_mem[i];
  memcpy(st->synthesis_mem, &x[FRAME_SIZE], FRAME_SIZE * sizeof(float));
}

float rno_process_frame(rno_state *st, float *out, const float *in) {
  int i;
  cpx X[FREQ_SIZE];
  cpx P[WINDOW_SIZE];
  float x[FRAME_SIZE];
  float Ex[NB_BANDS], Ep[NB_BANDS];
  float Exp[NB_BANDS];
  float features[NB_FEATURES];
  float g[NB_BANDS];
  float gf[FREQ_SIZE];
  float vad_prob = 0;
  int silence;
  static const float a_hp[2] = {-1.99599f, 0.99600f};
  static const float b_hp[2] = {-2, 1};
  for (i = 0; i < NB_BANDS; i++) g[i] = 0;
  biquad(x, st->mem_hp_x, in, b_hp, a_hp, FRAME_SIZE);
  silence = compute_frame_features(st, X, P, Ex, Ep, Exp, features, x);
  if (!silence) {
    compute_rnn(st, g, &vad_prob, features);
    memcpy(st->graw, g, sizeof(g));
    if (g_pf_dexp != 0.f || g_pf_dg != 0.f || g_pf_aexp != 0.f) {
      float Exp2[NB_BANDS], g2[NB_BANDS];
      for (i = 0; i < NB_BANDS; i++) {
        Exp2[i] = Exp[i] * (1.f + g_pf_dexp) + g_pf_aexp;
        g2[i] = g[i] + g_pf_dg * g[i] * (1.f - g[i]);
      }
      pitch_filter(X, P, Ex, Ep, Exp2, g2);
    } else {
      pitch_filter(X, P, Ex, Ep, Exp, g);
    }
    for (i = 0; i < NB_BANDS; i++) {
      float alpha = .6f;
      g[i] = fmaxf(g[i], alpha * st->lastg[i]);
      st->lastg[i] = g[i];
    }
    interp_band_gain(gf, g);
    for (i = 0; i < FREQ_SIZE; i++) {
      X[i].r *= gf[i];
      X[i].i *= gf[i];
    }
  }
  frame_synthesis(st, out, X);
  /* Distance of the pitch filter's `Exp > g ? 1 : ...` branch from flipping.  RNNoise is discontinuous there (r jumps
   * to 1 from a value that is ~0 when g is small), at any magnitude: Exp = 3e-7 against g = 0 takes the branch just
   * like 0.8 against 0.7.  Two float32 implementations agree on the two sides only up to their rounding noise:
   *   Exp (a normalised sum of positive and negative terms): ~2e-6 absolutely + 1e-5 relatively;
   *   g (a table sigmoid): the error of its pre-activation, ~4e-4, times g (1 - g) -- at most 1e-4, nothing where
   *   the sigmoid saturates.
   * margin of band b = |Exp - g| / noise_b, reported on the scale where 1e-4 means "within the noise".  Only bands
   * that reach the output count: r is interpolated over the neighbouring bands' bins (interp_band_gain), so band b
   * counts if the APPLIED gain max(g, 0.6 lastg) of b - 1, b or b + 1 exceeds 1e-3 (a band the RNN has just switched
   * off still plays at 0.6 of its previous gain, and a muted band between two audible ones shapes their bins). */
  st->branch_margin = 1e30f;
  if (!silence)
    for (i = 0; i < NB_BANDS; i++) {
      const float gr = st->graw[i];
      const float noise = 2e-6f + 1e-5f * (float)fabs(Exp[i]) + 4e-4f * gr * (1.f - gr);
      const float d = (float)fabs(Exp[i] - gr) * (1e-4f / noise);
      float audible = g[i];
      if (i > 0) audible = fmaxf(audible, g[i - 1]);
      if (i + 1 < NB_BANDS) audible = fmaxf(audible, g[i + 1]);
      if (audible > 1e-3f && d < st->branch_margin) st->branch_margin = d;
    }
  memcpy(st->dbg.features, features, sizeof(features));
  memcpy(st->dbg.gains, g, sizeof(g));
  memcpy(st->dbg.Ex, Ex, sizeof(Ex));
  memcpy(st->dbg.Ep, Ep, sizeof(Ep));
  memcpy(st->dbg.Exp, Exp, sizeof(Exp));
  st->dbg.vad = vad_prob;
  st->dbg.silence = silence;
  return vad_prob;
}

/* ------------------------------------------------------------------------------------------ */
/* batch helper (CPU baseline): pthreads over streams                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const rno_model *m;
  const float *in;
  float *out, *vad;
  int n_streams, n_frames, tid, n_threads;
  long in_stride, out_stride;
  unsigned flags;
  float volume;
  int32_t *tr_pitch, *tr_silence; /* optional per-frame decision traces, [n_streams][n_frames] */
  float *tr_pgain, *tr_margin;
} job;

static void *worker(void *arg) {
  job *j = (job *)arg;
  int s, t, i;
  float fin[FRAME_SIZE], fout[FRAME_SIZE];
  for (s = j->tid; s < j->n_streams; s += j->n_threads) {
    rno_state *st = rno_create(j->m);
    const float *src = j->in + (size_t)s * j->in_stride;
    float *dst = j->out + (size_t)s * j->out_stride;
    for (t = 0; t < j->n_frames; t++) {
      float v;
      if (j->flags & 1u) {
        /* wrapper arithmetic: audio.rs:261-273 */
        for (i = 0; i < FRAME_SIZE; i++) fin[i] = src[(size_t)t * FRAME_SIZE + i] * 32768.0f;
        v = rno_process_frame(st, fout, fin);
        for (i = 0; i < FRAME_SIZE; i++) {
          float o = fout[i] / 32768.0f;
          o = o < -1.f ? -1.f : (o > 1.f ? 1.f : o);
          dst[(size_t)t * FRAME_SIZE + i] = o * j->volume;
        }
      } else {
        v = rno_process_frame(st, dst + (size_t)t * FRAME_SIZE, src + (size_t)t * FRAME_SIZE);
      }
      if (j->vad) j->vad[(size_t)s * j->n_frames + t] = v;
      if (j->tr_pitch) j->tr_pitch[(size_t)s * j->n_frames + t] = st->dbg.pitch_index;
      if (j->tr_pgain) j->tr_pgain[(size_t)s * j->n_frames + t] = st->dbg.pitch_gain;
      if (j->tr_silence) j->tr_silence[(size_t)s * j->n_frames + t] = st->dbg.silence;
      if (j->tr_margin) j->tr_margin[(size_t)s * j->n_frames + t] = st->branch_margin;
    }
    rno_destroy(st);
  }
  return NULL;
}

int rno_process_streams(const rno_model *m, const float *in, float *out, float *vad, int n_streams,
                        int n_frames, long in_stride, long out_stride, unsigned flags, float volume,
                        int n_threads) {
  return rno_process_streams_trace(m, in, out, vad, n_streams, n_frames, in_stride, out_stride, flags, volume,
                                   n_threads, NULL, NULL, NULL, NULL);
}

int rno_process_streams_trace(const rno_model *m, const float *in, float *out, float *vad, int n_streams,
                              int n_frames, long in_stride, long out_stride, unsigned flags, float volume,
                              int n_threads, int32_t *pitch_index, float *pitch_gain, int32_t *silence,
                              float *branch_margin) {
  int t;
  pthread_t *th;
  job *jobs;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = n_streams > 0 ? n_streams : 1;
  ensure_tables();
  th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  jobs = (job *)malloc(sizeof(job) * (size_t)n_threads);
  for (t = 0; t < n_threads; t++) {
    job j;
    j.m = m;
    j.in = in;
    j.out = out;
    j.vad = vad;
    j.n_streams = n_streams;
    j.n_frames = n_frames;
    j.tid = t;
    j.n_threads = n_threads;
    j.in_stride = in_stride;
    j.out_stride = out_stride;
    j.flags = flags;
    j.volume = volume;
    j.tr_pitch = pitch_index;
    j.tr_pgain = pitch_gain;
    j.tr_silence = silence;
    j.tr_margin = branch_margin;
    jobs[t] = j;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a4: LinearResampler::process_sample (audio.rs:108-133)                                      */
/* ------------------------------------------------------------------------------------------ */
void rno_linres_init(rno_linres *r, float input_rate, float output_rate) {
  r->input_rate = input_rate;
  r->output_rate = output_rate;
  r->last_sample = 0.f;
  r->has_last = 0;
  r->input_pos = 0.0;
  r->next_output_pos = 0.0;
}

/* The capture callbacks' downmix to mono (audio.rs:754-755 f32, :816-818 i16, :879-884 u16): per interleaved frame,
 * the f32 sum in channel order of the samples brought to unit scale, divided by the channel count.  (Rust's
 * `iter().sum::<f32>()` folds from zero; whether that zero is +0.0 or -0.0 depends on the toolchain and only shows in
 * the sign of an all-negative-zero frame.)  fmt: 0 f32, 1 i16, 2 u16. */
void rno_downmix_mono(const void *in, int fmt, int n_channels, size_t n_frames, float *out) {
  size_t n;
  int c;
  for (n = 0; n < n_frames; n++) {
    float sum = 0.f;
    for (c = 0; c < n_channels; c++) {
      size_t k = n * (size_t)n_channels + (size_t)c;
      float v;
      if (fmt == 0)
        v = ((const float *)in)[k];
      else if (fmt == 1)
        v = (float)((const int16_t *)in)[k] / 32768.0f;
      else
        v = ((float)((const uint16_t *)in)[k] - 32768.0f) / 32768.0f;
      sum = sum + v;
    }
    out[n] = sum / (float)n_channels;
  }
}

/* f2, app audio: resample_audio (recording.rs:13-39), the recorder's whole-buffer linear interpolator for captured
 * app audio (recording.rs:356-360).  Line by line: ratio = from / to in f64; output_len = ceil(len / ratio); for
 * every i: src_pos = i * ratio, src_index = floor, frac = src_pos - src_index; two-sample interpolation with
 * `frac as f32`, the last sample alone where src_index + 1 runs off the end, nothing beyond.  Returns the number of
 * samples produced (all of them are written if out_cap allows). */
size_t rno_resample_audio(const float *samples, size_t len, size_t from_rate, size_t to_rate, float *out, size_t out_cap) {
  size_t i, k = 0, output_len;
  double ratio;
  if (from_rate == to_rate) { /* recording.rs:14-16 */
    for (i = 0; i < len; i++)
      if (i < out_cap) out[i] = samples[i];
    return len;
  }
  ratio = (double)from_rate / (double)to_rate;          /* recording.rs:18 */
  output_len = (size_t)ceil((double)len / ratio);       /* recording.rs:19 */
  for (i = 0; i < output_len; i++) {
    double src_pos = (double)i * ratio;                 /* recording.rs:23 */
    size_t src_index = (size_t)floor(src_pos);
    double frac = src_pos - (double)src_index;
    if (src_index + 1 < len) {                          /* recording.rs:27-31 */
      float sample1 = samples[src_index], sample2 = samples[src_index + 1];
      float o = sample1 + (sample2 - sample1) * (float)frac;
      if (k < out_cap) out[k] = o;
      k++;
    } else if (src_index < len) {                       /* recording.rs:32-35 */
      if (k < out_cap) out[k] = samples[src_index];
      k++;
    }
  }
  return k;
}

size_t rno_linres_process(rno_linres *r, const float *in, size_t n_in, float *out, size_t out_cap) {
  size_t n, k = 0;
  for (n = 0; n < n_in; n++) {
    float sample = in[n];
    double step;
    if (fabsf(r->input_rate - r->output_rate) < 1.0f) { /* audio.rs:109-112 */
      if (k < out_cap) out[k] = sample;
      k++;
      continue;
    }
    if (!r->has_last) { /* audio.rs:114-120: the first sample only primes the state */
      r->last_sample = sample;
      r->has_last = 1;
      r->input_pos = 0.0;
      r->next_output_pos = 0.0;
      continue;
    }
    r->input_pos += 1.0;
    step = (double)(r->input_rate / r->output_rate); /* f32 division, then widened (audio.rs:123) */
    while (r->next_output_pos <= r->input_pos) {
      float t = (float)(r->next_output_pos - (r->input_pos - 1.0));
      float o;
      t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
      o = r->last_sample + (sample - r->last_sample) * t;
      if (k < out_cap) out[k] = o;
      k++;
      r->next_output_pos += step;
    }
    r->last_sample = sample;
  }
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* a3: RnnNoiseProcessor::new + push_sample over a whole buffer (audio.rs:216-295)             */
/* ------------------------------------------------------------------------------------------ */
size_t rno_processor_run(const rno_model *m, float input_rate, float volume, const float *in,
                         size_t n_in, float *out, size_t out_cap) {
  rno_state *st = rno_create(m);
  rno_linres rs;
  int use_rs = fabsf(input_rate - 48000.0f) >= 1.0f; /* audio.rs:217 */
  float frame_in[FRAME_SIZE], frame_out[FRAME_SIZE];
  float tmp[8];
  int fill = 0, first_frame = 1;
  size_t n, k = 0;
  if (volume < 0.f) volume = 0.f; /* audio.rs:236 */
  if (volume > 1.f) volume = 1.f;
  rno_linres_init(&rs, input_rate, 48000.0f);
  for (n = 0; n < n_in; n++) {
    size_t cnt, c;
    if (use_rs) {
      cnt = rno_linres_process(&rs, in + n, 1, tmp, 8);
    } else {
      tmp[0] = in[n];
      cnt = 1;
    }
    for (c = 0; c < cnt; c++) {
      frame_in[fill++] = tmp[c] * 32768.0f; /* audio.rs:264 */
      if (fill == FRAME_SIZE) {
        int i;
        fill = 0;
        rno_process_frame(st, frame_out, frame_in); /* audio.rs:268 */
        if (first_frame) {                          /* audio.rs:275-278 */
          first_frame = 0;
          continue;
        }
        for (i = 0; i < FRAME_SIZE; i++) {
          float o = frame_out[i] / 32768.0f;
          o = o < -1.f ? -1.f : (o > 1.f ? 1.f : o);
          if (k < out_cap) out[k] = o * volume; /* audio.rs:272 */
          k++;
        }
      }
    }
  }
  rno_destroy(st);
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* f1: mixer + PCM16 quantiser (commands/recording.rs:260-264; recording.rs:108-110)           */
/* ------------------------------------------------------------------------------------------ */
void rno_mix_dual_mono_i16(const float *mic, const float *app, size_t n, int16_t *out) {
  size_t i;
  for (i = 0; i < n; i++) {
    float mixed = mic[i] + (app ? app[i] : 0.f);
    float c = mixed < -1.f ? -1.f : (mixed > 1.f ? 1.f : mixed);
    int16_t q = (int16_t)(c * 32767.0f); /* Rust `as i16`: truncation toward zero */
    out[2 * i] = q;
    out[2 * i + 1] = q;
  }
}

/* ---- f2 (north_star item 4): windowed-sinc polyphase resampler ----------------------------------
 * Follows rubato 0.16.2 (Cargo.lock:4166; source not vendored, restated from its documentation):
 * sinc.rs make_sincs -- y[x] = w[x] * sinc((x - tot/2) * f_cutoff / factor), tot = sinc_len * factor,
 * normalised so that sum(y) == factor; windows.rs BlackmanHarris2 -- the square of the 4-term
 * Blackman-Harris window; SincFixedIn -- f_cutoff scaled by the ratio when downsampling.
 * factor = L, so sub-filter p holds y[L*k + (L-1-p)]... here indexed directly by the fractional
 * position: h[p][k] = y at tau = (k - half + 1) - p/L input samples from the interpolation point. */
static int gcd_i(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}
static int sinc_ratio(int input_rate, int output_rate, int *L, int *M) {
  if (input_rate < 1 || output_rate < 1) return 0;
  int g = gcd_i(input_rate, output_rate);
  *L = output_rate / g;
  *M = input_rate / g;
  return *L <= 1024;
}
static double bh2_window(double x, double tot) { /* x in [0, tot) */
  const double PI = 3.14159265358979323846;
  double a = 2.0 * PI * x / tot;
  double w = 0.35875 - 0.48829 * cos(a) + 0.14128 * cos(2.0 * a) - 0.01168 * cos(3.0 * a);
  return w * w;
}
static int sinc_build(int L, int M, int sinc_len, float f_cutoff, float *table) {
  const double PI = 3.14159265358979323846;
  double fc = (double)f_cutoff;
  if (L < M) fc = fc * (double)L / (double)M;
  const long tot = (long)sinc_len * L;
  double *y = (double *)malloc(sizeof(double) * (size_t)tot);
  if (!y) return 0;
  double sum = 0.0;
  for (long x = 0; x < tot; x++) {
    double t = ((double)x - (double)(tot / 2)) * fc / (double)L;
    double s = t == 0.0 ? 1.0 : sin(PI * t) / (PI * t);
    y[x] = bh2_window((double)x, (double)tot) * s;
    sum += y[x];
  }
  sum /= (double)L;
  /* tau = (k - half + 1) - p/L  <=>  x = tot/2 + tau*L = L*(k + 1) - p   (x == tot -> tap is 0) */
  for (int p = 0; p < L; p++)
    for (int k = 0; k < sinc_len; k++) {
      long x = (long)L * (k + 1) - p;
      table[(size_t)p * sinc_len + k] = x < tot ? (float)(y[x] / sum) : 0.0f;
    }
  free(y);
  return 1;
}
int rno_sinc_table(int input_rate, int output_rate, int sinc_len, float f_cutoff, float *table,
                   size_t cap_floats, int *M_out) {
  int L, M;
  if (!sinc_ratio(input_rate, output_rate, &L, &M) || sinc_len < 2 || (sinc_len & 1)) return 0;
  if (cap_floats < (size_t)L * sinc_len) return 0;
  if (!sinc_build(L, M, sinc_len, f_cutoff, table)) return 0;
  if (M_out) *M_out = M;
  return L;
}
size_t rno_sinc_resample_count(int input_rate, int output_rate, size_t n_in) {
  int L, M;
  if (!sinc_ratio(input_rate, output_rate, &L, &M)) return 0;
  /* every n with n*M/L < n_in */
  return (size_t)(((unsigned long long)n_in * L + M - 1) / M);
}
size_t rno_sinc_resample(const float *in, size_t n_in, float *out, size_t out_cap, int input_rate,
                         int output_rate, int sinc_len, float f_cutoff) {
  int L, M;
  if (!sinc_ratio(input_rate, output_rate, &L, &M) || sinc_len < 2 || (sinc_len & 1)) return 0;
  size_t n_out = rno_sinc_resample_count(input_rate, output_rate, n_in);
  if (!out) return n_out;
  if (n_out > out_cap) n_out = out_cap;
  float *h = (float *)malloc(sizeof(float) * (size_t)L * sinc_len);
  if (!h || !sinc_build(L, M, sinc_len, f_cutoff, h)) {
    free(h);
    return 0;
  }
  const long half = sinc_len / 2;
  for (size_t n = 0; n < n_out; n++) {
    unsigned long long pos = (unsigned long long)n * M;
    long base = (long)(pos / L);
    int p = (int)(pos % L);
    const float *hp = h + (size_t)p * sinc_len;
    float acc = 0.f;
    for (int k = 0; k < sinc_len; k++) {
      long i = base - half + 1 + k;
      float x = (i >= 0 && (size_t)i < n_in) ? in[i] : 0.f;
      acc = fmaf(hp[k], x, acc);
    }
    out[n] = acc;
  }
  free(h);
  return n_out;
}
