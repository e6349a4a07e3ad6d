// ns_kernel.cuh -- the RNNoise stream kernel: device code for the whole denoise path behind
// nnnoiseless::DenoiseState::process_frame (/root/reference/src-tauri/src/audio.rs:268).
//
// One CTA owns S independent streams and walks them through time, one 480-sample frame per step,
// with every piece of DenoiseState resident in shared memory for the whole launch:
//   analysis  : biquad high-pass (f64 warp scan), pitch_downsample (LPC whitening), pitch_search,
//               remove_doubling, windowed 960-point real FFTs of the frame and of the pitch-lagged
//               window, 22-band energies / correlation, DCT cepstrum + delta features
//   recurrent : dense -> VAD GRU -> noise GRU -> denoise GRU -> dense, all S streams batched per
//               weight fetch (FP32 FMA; int8 weights stay int8 until the multiply)
//   synthesis : pitch filter, gain smoothing + interpolation, inverse FFT, window, overlap-add
// Each stream is worked by a 128-thread group that synchronises on its own named barrier
// (bar.sync id,128); only the recurrent phase meets CTA-wide.  Streams are independent, so there
// is no inter-CTA or inter-GPU exchange anywhere on the path.
//
// The same source compiles for sm_100a (nvcc) and for the host SIMT emulation (NS_HOST_EMU, tests).
#pragma once
#include "ns_common.h"
#include "ns_simt.h"

namespace ns {

// ------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------
struct StreamSmem {
  float ring[kRing];   // biquad output history: 4 frame slots (pitch_buf[1728] + analysis_mem)
  float synth[kFrame]; // synthesis_mem
  cf X[482];           // spectrum of the frame; scratch during pitch analysis
  cf P[482];           // spectrum of the pitch-lagged window; holds x_lp[864] during pitch analysis
  float ceps[kCepsMem * kBands];
  float lastg[24];
  float Ex[24], Ep[24], Exp[24], Ly[24], g[24], r[24], nrm[24], newE[24];
  float feat[44];
  float red[32];
  float dots[32];
  float dist[64];
  float fx[12], fy[12];
  int fi[12];
  float lpc2[8];
  double hp[2];
  float last_gain, pitch_gain, vad, pad0;
  int last_period, memid, ring_slot, pitch_index, T0, T, best0, best1, silence, pad1;
  long long frame_count;
};

struct RnnSmem {
  float feat[44 * 8];   // [row][stream]
  float dense[24 * 8];
  float hvad[24 * 8];
  float hnoise[48 * 8];
  float hden[96 * 8];
  float rh[96 * 8];
  float z[96 * 8];
  float gains[24 * 8];
  float vad[8];
  int silent[8];
};

template <int S>
struct CtaSmem {
  Tables tab;
  RnnSmem rnn;
  StreamSmem st[S];
};

struct Grp {
  int tid, lane, warp, bar;
};
NS_DEV void gsync(const Grp &g) { Simt::group_sync(g.bar, kGroupThreads); }
NS_DEV float warp_sum(float v) {
  v += Simt::shfl_xor(v, 16);
  v += Simt::shfl_xor(v, 8);
  v += Simt::shfl_xor(v, 4);
  v += Simt::shfl_xor(v, 2);
  v += Simt::shfl_xor(v, 1);
  return v;
}
NS_DEV f4 ld4(const float *p) { return *reinterpret_cast<const f4 *>(p); }
NS_DEV cf cmul(cf a, cf b) {
  cf c;
  c.x = a.x * b.x - a.y * b.y;
  c.y = a.x * b.y + a.y * b.x;
  return c;
}
NS_DEV cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
NS_DEV cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
NS_DEV cf mul_neg_i(cf a) { return cf{a.y, -a.x}; }  // a * (-i)
NS_DEV cf mul_pos_i(cf a) { return cf{-a.y, a.x}; }  // a * (+i)

// ------------------------------------------------------------------------------------------------
// 480-point complex FFT (forward, e^{-2 pi i nk/N}), Stockham radices 4,4,5,6, in place through
// registers: every active thread loads its butterfly, the group meets, then everyone stores.
// ------------------------------------------------------------------------------------------------
template <int R>
struct Dft;
template <>
struct Dft<3> {
  static NS_DEV void run(cf *v) {
    const float s = 0.86602540378443864676f;
    cf a = cadd(v[1], v[2]), b = csub(v[1], v[2]);
    cf m = cf{v[0].x - 0.5f * a.x, v[0].y - 0.5f * a.y};
    cf n = cf{s * b.x, s * b.y};
    v[0] = cadd(v[0], a);
    v[1] = cadd(m, mul_neg_i(n));
    v[2] = cadd(m, mul_pos_i(n));
  }
};
template <>
struct Dft<4> {
  static NS_DEV void run(cf *v) {
    cf t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    cf t2 = cadd(v[1], v[3]), t3 = mul_neg_i(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3);
    v[3] = csub(t1, t3);
  }
};
template <>
struct Dft<5> {
  static NS_DEV void run(cf *v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    cf a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    cf b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    cf m1 = cf{v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y};
    cf m2 = cf{v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y};
    cf n1 = cf{s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y};
    cf n2 = cf{s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y};
    v[0] = cf{v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y};
    v[1] = cadd(m1, mul_neg_i(n1));
    v[4] = cadd(m1, mul_pos_i(n1));
    v[2] = cadd(m2, mul_neg_i(n2));
    v[3] = cadd(m2, mul_pos_i(n2));
  }
};
template <>
struct Dft<6> {
  static NS_DEV void run(cf *v) {
    cf e[3] = {v[0], v[2], v[4]};
    cf o[3] = {v[1], v[3], v[5]};
    Dft<3>::run(e);
    Dft<3>::run(o);
    const cf w1 = cf{0.5f, -0.86602540378443864676f};   // W6
    const cf w2 = cf{-0.5f, -0.86602540378443864676f};  // W6^2
    o[1] = cmul(o[1], w1);
    o[2] = cmul(o[2], w2);
    v[0] = cadd(e[0], o[0]);
    v[1] = cadd(e[1], o[1]);
    v[2] = cadd(e[2], o[2]);
    v[3] = csub(e[0], o[0]);
    v[4] = csub(e[1], o[1]);
    v[5] = csub(e[2], o[2]);
  }
};

template <int R, int NS_, class Load>
NS_DEV void fft_stage(const Grp &g, const Tables &T, cf *buf, Load load) {
  constexpr int M = 480 / R;
  constexpr int TSTEP = 480 / (NS_ * R);
  cf v[R];
  const int j = g.tid;
  const bool act = j < M;
  int k = 0;
  if (act) {
    k = j % NS_;
    v[0] = load(j);
#pragma unroll
    for (int r = 1; r < R; r++) {
      cf x = load(j + r * M);
      v[r] = (NS_ == 1) ? x : cmul(x, T.w480[r * k * TSTEP]);
    }
    Dft<R>::run(v);
  }
  gsync(g);
  if (act) {
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) buf[j0 + r * NS_] = v[r];
  }
  gsync(g);
}

template <class LoadFirst>
NS_DEV void fft480(const Grp &g, const Tables &T, cf *buf, LoadFirst load_first) {
  auto from_buf = [&](int n) -> cf { return buf[n]; };
  fft_stage<4, 1>(g, T, buf, load_first);
  fft_stage<4, 4>(g, T, buf, from_buf);
  fft_stage<5, 16>(g, T, buf, from_buf);
  fft_stage<6, 80>(g, T, buf, from_buf);
}

// a7 / a12: X <- rFFT960(window . ring[(base+i) mod 1920]) / 960   (bins 0..480)
NS_DEV void rfft960_windowed(const Grp &g, const Tables &T, const float *ring, int base, cf *X) {
  auto load = [&](int n) -> cf {
    const int i0 = 2 * n;
    const float w0 = (i0 < kFrame) ? T.win[i0] : T.win[kWindow - 1 - i0];
    const float w1 = (i0 + 1 < kFrame) ? T.win[i0 + 1] : T.win[kWindow - 2 - i0];
    int a = base + i0;
    if (a >= kRing) a -= kRing;
    int b = a + 1;
    if (b >= kRing) b -= kRing;
    return cf{ring[a] * w0, ring[b] * w1};
  };
  fft480(g, T, X, load);
  const float norm = 1.0f / kWindow;
  for (int k = g.tid; k <= 240; k += kGroupThreads) {
    const cf a = X[k], b = X[k == 0 ? 0 : 480 - k], w = T.w960[k];
    const float er = .5f * (a.x + b.x), ei = .5f * (a.y - b.y);
    const float orr = .5f * (a.x - b.x), oi = .5f * (a.y + b.y);
    const float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
    X[k] = cf{(er + ti) * norm, (ei - tr) * norm};
    X[480 - k] = cf{(er - ti) * norm, (-ei - tr) * norm};
  }
  gsync(g);
}

// a16: unscaled inverse of the Hermitian spectrum X[0..480]; result left in X as 480 complex
// z[m] with x[2m] = z[m].x and x[2m+1] = -z[m].y (the conjugate of a forward FFT).
NS_DEV void irfft960_inplace(const Grp &g, const Tables &T, cf *X) {
  for (int k = g.tid; k <= 240; k += kGroupThreads) {
    const cf a = X[k], b = X[480 - k], w = T.w960[k];
    const float ex = a.x + b.x, ey = a.y - b.y;
    const float ox = a.x - b.x, oy = a.y + b.y;
    const float tr = ox * w.x + oy * w.y, ti = -ox * w.y + oy * w.x;
    X[k] = cf{ex - ti, -(ey + tr)};
    if (k != 0) X[480 - k] = cf{ex + ti, ey - tr};
  }
  gsync(g);
  auto from_buf = [&](int n) -> cf { return X[n]; };
  fft480(g, T, X, from_buf);
}

// ------------------------------------------------------------------------------------------------
// a8: 22 triangular bands over bins 0..400.  88 threads: band = tid/4, four lanes split the bins.
// ------------------------------------------------------------------------------------------------
template <class BinVal>
NS_DEV float band_accumulate(const Grp &g, const Tables &T, BinVal val) {
  float acc = 0.f;
  const int b = g.tid >> 2, sub = g.tid & 3;
  if (g.tid < 4 * kBands) {
    if (b >= 1) {
      const int lo = T.eband[b - 1], n = T.eband[b] - lo;
      for (int j = sub; j < n; j += 4) acc += ((float)j / (float)n) * val(lo + j);
    }
    if (b <= kBands - 2) {
      const int lo = T.eband[b], n = T.eband[b + 1] - lo;
      for (int j = sub; j < n; j += 4) acc += (1.f - (float)j / (float)n) * val(lo + j);
    }
  }
  acc += Simt::shfl_xor(acc, 1);
  acc += Simt::shfl_xor(acc, 2);
  if (b == 0 || b == kBands - 1) acc *= 2.f;
  return acc;  // valid in lanes with sub == 0 and tid < 88
}

// ------------------------------------------------------------------------------------------------
// a6: biquad high-pass over one frame (in place in its ring slot).  The recursion
// s' = A s + B x is linear, so warp 0 runs 15 samples per lane from a zero state, composes the
// lane end states with a Kogge-Stone scan of A^15 powers, and re-runs each lane from its true
// start state.  All in f64 (the reference widens to f64 per sample as well).
// ------------------------------------------------------------------------------------------------
NS_DEV void biquad_frame(const Grp &g, const Tables &T, StreamSmem &s, float *slot) {
  if (g.warp == 0) {
    const int lane = g.lane;
    const double a00 = T.hp_a[0], a01 = T.hp_a[1], a10 = T.hp_a[2], a11 = T.hp_a[3];
    const double b0 = T.hp_b[0], b1 = T.hp_b[1];
    float xs[15];
#pragma unroll
    for (int m = 0; m < 15; m++) xs[m] = slot[15 * lane + m];
    double m0 = lane == 0 ? s.hp[0] : 0.0, m1 = lane == 0 ? s.hp[1] : 0.0;
#pragma unroll
    for (int m = 0; m < 15; m++) {
      const double x = (double)xs[m];
      const double n0 = a00 * m0 + a01 * m1 + b0 * x;
      const double n1 = a10 * m0 + a11 * m1 + b1 * x;
      m0 = n0;
      m1 = n1;
    }
#pragma unroll
    for (int d = 0; d < 5; d++) {
      const double u0 = Simt::shfl_up(m0, 1 << d), u1 = Simt::shfl_up(m1, 1 << d);
      if (lane >= (1 << d)) {
        m0 += T.hp_pow[d][0] * u0 + T.hp_pow[d][1] * u1;
        m1 += T.hp_pow[d][2] * u0 + T.hp_pow[d][3] * u1;
      }
    }
    double s0 = Simt::shfl_up(m0, 1), s1 = Simt::shfl_up(m1, 1);
    if (lane == 0) {
      s0 = s.hp[0];
      s1 = s.hp[1];
    }
    // upstream form: y = x + m0; m0 = m1 + (b0 x - a0 y); m1 = b1 x - a1 y, with b = {-2, 1}
    const double a_0 = -a00, a_1 = -a10;
#pragma unroll
    for (int m = 0; m < 15; m++) {
      const double x = (double)xs[m];
      const double y = x + s0;
      s0 = s1 + (-2.0 * x - a_0 * y);
      s1 = x - a_1 * y;
      slot[15 * lane + m] = (float)y;
    }
    const double e0 = Simt::shfl(s0, 31), e1 = Simt::shfl(s1, 31);
    if (lane == 0) {
      s.hp[0] = e0;
      s.hp[1] = e1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a9: pitch_downsample.  pitch_buf[j] = ring[(base + j) mod 1920].  Result x_lp[864] in s.P.
// ------------------------------------------------------------------------------------------------
NS_DEV void pitch_downsample(const Grp &g, StreamSmem &s, int base) {
  float *lpraw = reinterpret_cast<float *>(s.X);
  float *lp = reinterpret_cast<float *>(s.P);
  auto pb = [&](int j) -> float {
    int a = base + j;
    if (a >= kRing) a -= kRing;
    return s.ring[a];
  };
  for (int i = g.tid; i < 864; i += kGroupThreads) {
    float v;
    if (i == 0)
      v = .5f * (.5f * pb(1) + pb(0));
    else
      v = .5f * (.5f * (pb(2 * i - 1) + pb(2 * i + 1)) + pb(2 * i));
    lpraw[i] = v;
  }
  gsync(g);
  float ac[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = g.tid; i < 864; i += kGroupThreads) {
    const float xi = lpraw[i];
#pragma unroll
    for (int k = 0; k < 5; k++)
      if (i >= k) ac[k] = fmaf(xi, lpraw[i - k], ac[k]);
  }
#pragma unroll
  for (int k = 0; k < 5; k++) {
    ac[k] = warp_sum(ac[k]);
    if (g.lane == 0) s.red[g.warp * 8 + k] = ac[k];
  }
  gsync(g);
  if (g.tid == 0) {
    float lpc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 5; k++) ac[k] = (s.red[k] + s.red[8 + k]) + (s.red[16 + k] + s.red[24 + k]);
    ac[0] *= 1.0001f;
#pragma unroll
    for (int i = 1; i <= 4; i++) ac[i] -= ac[i] * (.008f * i) * (.008f * i);
    float error = ac[0];
    if (ac[0] != 0.f) {
      for (int i = 0; i < 4; i++) {
        float rr = 0.f;
        for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
        rr += ac[i + 1];
        const float r = -rr / error;
        lpc[i] = r;
        for (int j = 0; j < ((i + 1) >> 1); j++) {
          const float t1 = lpc[j], t2 = lpc[i - 1 - j];
          lpc[j] = t1 + r * t2;
          lpc[i - 1 - j] = t2 + r * t1;
        }
        error = error - r * r * error;
        if (error < .001f * ac[0]) break;
      }
    }
    float tmp = 1.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      tmp = .9f * tmp;
      lpc[i] = lpc[i] * tmp;
    }
    s.lpc2[0] = lpc[0] + .8f;
    s.lpc2[1] = lpc[1] + .8f * lpc[0];
    s.lpc2[2] = lpc[2] + .8f * lpc[1];
    s.lpc2[3] = lpc[3] + .8f * lpc[2];
    s.lpc2[4] = .8f * lpc[3];
  }
  gsync(g);
  const float n0 = s.lpc2[0], n1 = s.lpc2[1], n2 = s.lpc2[2], n3 = s.lpc2[3], n4 = s.lpc2[4];
  for (int i = g.tid; i < 864; i += kGroupThreads) {
    float sum = lpraw[i];
    if (i >= 1) sum = fmaf(n0, lpraw[i - 1], sum);
    if (i >= 2) sum = fmaf(n1, lpraw[i - 2], sum);
    if (i >= 3) sum = fmaf(n2, lpraw[i - 3], sum);
    if (i >= 4) sum = fmaf(n3, lpraw[i - 4], sum);
    if (i >= 5) sum = fmaf(n4, lpraw[i - 5], sum);
    lp[i] = sum;
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// a10: pitch_search.  Coarse 4x-decimated cross-correlation (147 lags x 240 taps) register-tiled
// four lags per thread; top-2 by a warp merge that reproduces find_best_pitch's insertion order;
// fine 2x-decimated search on at most ten lags, one warp per dot product.
// ------------------------------------------------------------------------------------------------
struct Best2 {
  float n0, d0, n1, d1;
  int p0, p1;
};
NS_DEV void best_init(Best2 &b) {
  b.n0 = b.n1 = -1.f;
  b.d0 = b.d1 = 0.f;
  b.p0 = 0;
  b.p1 = 1;
}
NS_DEV void best_insert(Best2 &b, float num, float syy, int i) {
  if (num * b.d1 > b.n1 * syy) {
    if (num * b.d0 > b.n0 * syy) {
      b.n1 = b.n0;
      b.d1 = b.d0;
      b.p1 = b.p0;
      b.n0 = num;
      b.d0 = syy;
      b.p0 = i;
    } else {
      b.n1 = num;
      b.d1 = syy;
      b.p1 = i;
    }
  }
}

NS_DEV void pitch_search(const Grp &g, StreamSmem &s) {
  const float *lp = reinterpret_cast<const float *>(s.P);
  float *y4 = reinterpret_cast<float *>(s.X);  // 392 (387 valid, zero padded)
  float *x4 = y4 + 392;                        // 240
  float *xcp = y4 + 632;                       // 2 x 148 partial correlations
  for (int j = g.tid; j < 392; j += kGroupThreads) y4[j] = j < 387 ? lp[2 * j] : 0.f;
  for (int j = g.tid; j < 240; j += kGroupThreads) x4[j] = lp[384 + 2 * j];
  gsync(g);
  if (g.tid < 74) {
    const int q = g.tid % 37, h = g.tid / 37;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *yb = y4 + 4 * q;
    for (int j = 120 * h; j < 120 * h + 120; j += 4) {
      const f4 xv = ld4(x4 + j), ya = ld4(yb + j), yc = ld4(yb + j + 4);
      a0 = fmaf(xv.x, ya.x, a0); a1 = fmaf(xv.x, ya.y, a1); a2 = fmaf(xv.x, ya.z, a2); a3 = fmaf(xv.x, ya.w, a3);
      a0 = fmaf(xv.y, ya.y, a0); a1 = fmaf(xv.y, ya.z, a1); a2 = fmaf(xv.y, ya.w, a2); a3 = fmaf(xv.y, yc.x, a3);
      a0 = fmaf(xv.z, ya.z, a0); a1 = fmaf(xv.z, ya.w, a1); a2 = fmaf(xv.z, yc.x, a2); a3 = fmaf(xv.z, yc.y, a3);
      a0 = fmaf(xv.w, ya.w, a0); a1 = fmaf(xv.w, yc.x, a1); a2 = fmaf(xv.w, yc.y, a2); a3 = fmaf(xv.w, yc.z, a3);
    }
    float *dst = xcp + h * 148 + 4 * q;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  }
  gsync(g);
  if (g.warp == 0) {
    Best2 b;
    best_init(b);
    const int i0 = 5 * g.lane;
    float syy = 1.f;
    if (i0 < 147) {
      for (int j = 0; j < 240; j++) syy = fmaf(y4[i0 + j], y4[i0 + j], syy);
      for (int c = 0; c < 5; c++) {
        const int i = i0 + c;
        if (i < 147) {
          const float xc = xcp[i] + xcp[148 + i];
          if (xc > 0.f) {
            const float x16 = xc * 1e-12f;
            best_insert(b, x16 * x16, syy, i);
          }
          syy += y4[i + 240] * y4[i + 240] - y4[i] * y4[i];
          syy = fmaxf(1.f, syy);
        }
      }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Best2 o;
      o.n0 = Simt::shfl_down(b.n0, d); o.d0 = Simt::shfl_down(b.d0, d); o.p0 = Simt::shfl_down(b.p0, d);
      o.n1 = Simt::shfl_down(b.n1, d); o.d1 = Simt::shfl_down(b.d1, d); o.p1 = Simt::shfl_down(b.p1, d);
      if (g.lane + d < 32) {
        const bool second_first = (o.n1 >= 0.f) && (o.p1 < o.p0);
        if (second_first) best_insert(b, o.n1, o.d1, o.p1);
        if (o.n0 >= 0.f) best_insert(b, o.n0, o.d0, o.p0);
        if (!second_first && o.n1 >= 0.f) best_insert(b, o.n1, o.d1, o.p1);
      }
    }
    if (g.lane == 0) {
      s.best0 = b.p0;
      s.best1 = b.p1;
    }
  }
  gsync(g);
  // fine search: slots 0..4 around 2*best0, 5..9 around 2*best1 (minus duplicates)
  {
    const int c0 = 2 * s.best0, c1 = 2 * s.best1;
    for (int c = g.warp; c < 10; c += 4) {
      const int i = (c < 5) ? (c0 - 2 + c) : (c1 - 2 + (c - 5));
      const int dd = i - c0;
      const bool valid = (i >= 0) && (i < 294) && (c < 5 || dd > 2 || dd < -2);
      float sxy = 0.f, syy = 0.f;
      if (valid) {
        for (int j = g.lane; j < 480; j += 32) {
          const float yv = lp[i + j];
          sxy = fmaf(lp[384 + j], yv, sxy);
          syy = fmaf(yv, yv, syy);
        }
      }
      sxy = warp_sum(sxy);
      syy = warp_sum(syy);
      if (g.lane == 0) {
        s.fi[c] = valid ? i : -1;
        s.fx[c] = sxy < -1.f ? -1.f : sxy;
        s.fy[c] = 1.f + syy;
      }
    }
  }
  gsync(g);
  if (g.tid == 0) {
    Best2 b;
    best_init(b);
    const int first = (s.best1 < s.best0) ? 5 : 0;  // visit lags in ascending order
    for (int pass = 0; pass < 2; pass++) {
      const int off = pass == 0 ? first : 5 - first;
      for (int c = off; c < off + 5; c++) {
        if (s.fi[c] >= 0 && s.fx[c] > 0.f) {
          const float x16 = s.fx[c] * 1e-12f;
          best_insert(b, x16 * x16, s.fy[c], s.fi[c]);
        }
      }
    }
    const int bp = b.p0;
    int offset = 0;
    if (bp > 0 && bp < 293) {
      float a = 0.f, bb = 0.f, cc = 0.f;
      for (int c = 0; c < 10; c++) {
        if (s.fi[c] == bp - 1) a = s.fx[c];
        if (s.fi[c] == bp) bb = s.fx[c];
        if (s.fi[c] == bp + 1) cc = s.fx[c];
      }
      if ((cc - a) > .7f * (bb - a))
        offset = 1;
      else if ((a - cc) > .7f * (bb - cc))
        offset = -1;
    }
    const int pitch = 2 * bp - offset;
    const int pitch_index = kPitchMax - pitch;
    int T0 = pitch_index / 2;
    if (T0 >= 384) T0 = 383;
    s.T0 = T0;
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// a11: remove_doubling.  Every inner product depends only on T0, so all 30 run in parallel (one
// warp per dot); the running yy_lookup becomes a block prefix sum; only the <=14-step threshold
// walk that consults last_period / last_gain is serial (thread 0).
// ------------------------------------------------------------------------------------------------
NS_DEV float pitch_gain(float xy, float xx, float yy) { return xy / sqrtf(1.f + xx * yy); }

NS_DEV int rd_lag(int d, int T0) {  // lag of dot d: 0 -> xx, 1 -> T0, then (T1, T1b) for k = 2..15
  if (d == 0) return 0;
  if (d == 1) return T0;
  const int k = 2 + ((d - 2) >> 1);
  const int T1 = (2 * T0 + k) / (2 * k);
  if (((d - 2) & 1) == 0) return T1;
  if (k == 2) return (T1 + T0 > 384) ? T0 : T0 + T1;
  const int sc = (k == 6 || k == 12) ? 5 : ((k & 1) ? 2 : 3);  // second_check[k]
  return (2 * sc * T0 + k) / (2 * k);
}

NS_DEV void remove_doubling(const Grp &g, StreamSmem &s) {
  const float *lp = reinterpret_cast<const float *>(s.P);
  const float *x = lp + 384;
  float *D = reinterpret_cast<float *>(s.X);  // D[i] = sum_{m=1..i} x[-m]^2 - x[480-m]^2
  const int T0 = s.T0;
  {
    const int m = 3 * g.tid + 1;
    const float d1 = x[-m] * x[-m] - x[480 - m] * x[480 - m];
    const float d2 = x[-m - 1] * x[-m - 1] - x[479 - m] * x[479 - m];
    const float d3 = x[-m - 2] * x[-m - 2] - x[478 - m] * x[478 - m];
    const float l1 = d1, l2 = d1 + d2, l3 = l2 + d3;
    float inc = l3;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float u = Simt::shfl_up(inc, d);
      if (g.lane >= d) inc += u;
    }
    if (g.lane == 31) s.red[g.warp] = inc;
    gsync(g);
    float off = 0.f;
    for (int w = 0; w < g.warp; w++) off += s.red[w];
    const float excl = off + (inc - l3);
    D[m] = excl + l1;
    D[m + 1] = excl + l2;
    D[m + 2] = excl + l3;
    if (g.tid == 0) D[0] = 0.f;
  }
  for (int d = g.warp; d < 30; d += 4) {
    const int T = rd_lag(d, T0);
    float acc = 0.f;
    for (int j = g.lane; j < 480; j += 32) acc = fmaf(x[j], x[j - T], acc);
    acc = warp_sum(acc);
    if (g.lane == 0) s.dots[d] = acc;
  }
  gsync(g);
  if (g.tid == 0) {
    const float xx = s.dots[0];
    float xy = s.dots[1];
    float yy = fmaxf(0.f, xx + D[T0]);
    float best_xy = xy, best_yy = yy;
    const float g0 = pitch_gain(xy, xx, yy);
    float gg = g0;
    int T = T0;
    const int prev_period = s.last_period / 2;
    const float prev_gain = s.last_gain;
    for (int k = 2; k <= 15; k++) {
      const int T1 = (2 * T0 + k) / (2 * k);
      if (T1 < 30) break;
      const int T1b = rd_lag(3 + 2 * (k - 2), T0);
      xy = .5f * (s.dots[2 + 2 * (k - 2)] + s.dots[3 + 2 * (k - 2)]);
      yy = .5f * (fmaxf(0.f, xx + D[T1]) + fmaxf(0.f, xx + D[T1b]));
      const float g1 = pitch_gain(xy, xx, yy);
      float cont;
      const int dT = T1 > prev_period ? T1 - prev_period : prev_period - T1;
      if (dT <= 1)
        cont = prev_gain;
      else if (dT <= 2 && 5 * k * k < T0)
        cont = .5f * prev_gain;
      else
        cont = 0.f;
      float thresh = fmaxf(.3f, .7f * g0 - cont);
      if (T1 < 90)
        thresh = fmaxf(.4f, .85f * g0 - cont);
      else if (T1 < 60)
        thresh = fmaxf(.5f, .9f * g0 - cont);
      if (g1 > thresh) {
        best_xy = xy;
        best_yy = yy;
        T = T1;
        gg = g1;
      }
    }
    best_xy = fmaxf(0.f, best_xy);
    float pg = (best_yy <= best_xy) ? 1.f : best_xy / (best_yy + 1.f);
    if (pg > gg) pg = gg;
    s.T = T;
    s.pitch_gain = pg;
  }
  gsync(g);
  if (g.warp < 3) {
    const int T = s.T + g.warp - 1;
    float acc = 0.f;
    for (int j = g.lane; j < 480; j += 32) acc = fmaf(x[j], x[j - T], acc);
    acc = warp_sum(acc);
    if (g.lane == 0) s.dots[g.warp] = acc;
  }
  gsync(g);
  if (g.tid == 0) {
    const float c0 = s.dots[0], c1 = s.dots[1], c2 = s.dots[2];
    int offset = 0;
    if ((c2 - c0) > .7f * (c1 - c0))
      offset = 1;
    else if ((c0 - c2) > .7f * (c1 - c2))
      offset = -1;
    int pi = 2 * s.T + offset;
    if (pi < kPitchMin) pi = kPitchMin;
    s.pitch_index = pi;
    s.last_period = pi;
    s.last_gain = s.pitch_gain;
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// a13: features.  Needs Ex, Ep, Exp (raw band correlation) in shared memory.
// ------------------------------------------------------------------------------------------------
NS_DEV void compute_features(const Grp &g, const Tables &T, StreamSmem &s) {
  const float dct_scale = 0.30151134457776363f;  // sqrt(2/22)
  if (g.tid < kBands) {
    s.Exp[g.tid] = s.Exp[g.tid] / sqrtf(.001f + s.Ex[g.tid] * s.Ep[g.tid]);
  }
  if (g.tid == 32) {
    float logMax = -2.f, follow = -2.f, E = 0.f;
    for (int i = 0; i < kBands; i++) {
      float ly = log10f(1e-2f + s.Ex[i]);
      ly = fmaxf(logMax - 7.f, fmaxf(follow - 1.5f, ly));
      logMax = fmaxf(logMax, ly);
      follow = fmaxf(follow - 1.5f, ly);
      s.Ly[i] = ly;
      E += s.Ex[i];
    }
    s.silence = (E < 0.04f) ? 1 : 0;
  }
  gsync(g);
  const bool silent = s.silence != 0;
  if (g.tid < kBands) {
    float sum = 0.f;
    for (int j = 0; j < kBands; j++) sum += s.Ly[j] * T.dct[j * kBands + g.tid];
    float c = sum * dct_scale;
    if (g.tid == 0) c -= 12.f;
    if (g.tid == 1) c -= 4.f;
    if (!silent) {
      s.ceps[s.memid * kBands + g.tid] = c;
      s.feat[g.tid] = c;
    }
  } else if (g.tid >= 32 && g.tid < 32 + kDeltaCeps) {
    const int i = g.tid - 32;
    float sum = 0.f;
    for (int j = 0; j < kBands; j++) sum += s.Exp[j] * T.dct[j * kBands + i];
    float c = sum * dct_scale;
    if (i == 0) c -= 1.3f;
    if (i == 1) c -= 0.9f;
    s.feat[kBands + 2 * kDeltaCeps + i] = c;
  } else if (g.tid == 40) {
    s.feat[kBands + 3 * kDeltaCeps] = .01f * (float)(s.pitch_index - 300);
  }
  gsync(g);
  if (!silent) {
    const int m0 = s.memid, m1 = (s.memid + 7) & 7, m2 = (s.memid + 6) & 7;
    if (g.tid < kDeltaCeps) {
      const float c0 = s.ceps[m0 * kBands + g.tid], c1 = s.ceps[m1 * kBands + g.tid],
                  c2 = s.ceps[m2 * kBands + g.tid];
      s.feat[g.tid] = c0 + c1 + c2;
      s.feat[kBands + g.tid] = c0 - c2;
      s.feat[kBands + kDeltaCeps + g.tid] = c0 - 2.f * c1 + c2;
    } else if (g.tid >= 32 && g.tid < 96) {
      const int a = (g.tid - 32) >> 3, b = (g.tid - 32) & 7;
      float dist = 0.f;
      for (int k = 0; k < kBands; k++) {
        const float t = s.ceps[a * kBands + k] - s.ceps[b * kBands + k];
        dist += t * t;
      }
      s.dist[(a << 3) + b] = dist;
    }
  }
  gsync(g);
  if (!silent && g.tid == 0) {
    float sv = 0.f;
    for (int a = 0; a < kCepsMem; a++) {
      float mind = 1e15f;
      for (int b = 0; b < kCepsMem; b++)
        if (b != a) mind = fminf(mind, s.dist[(a << 3) + b]);
      sv += mind;
    }
    s.feat[kBands + 3 * kDeltaCeps + 1] = sv / kCepsMem - 2.1f;
    s.memid = (s.memid + 1) & 7;
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// a14: the recurrent core, all S streams of the CTA per weight fetch.
// ------------------------------------------------------------------------------------------------
NS_DEV float tansig_approx(const Tables &T, float x) {
  if (!(x < 8.f)) return 1.f;
  if (!(x > -8.f)) return -1.f;
  float sign = 1.f;
  if (x < 0.f) {
    x = -x;
    sign = -1.f;
  }
  const int i = (int)floorf(.5f + 25.f * x);
  x -= .04f * i;
  float y = T.tansig[i];
  const float dy = 1.f - y * y;
  y = y + x * dy * (1.f - y * x);
  return sign * y;
}
NS_DEV float sigmoid_approx(const Tables &T, float x) { return .5f + .5f * tansig_approx(T, .5f * x); }
NS_DEV float activate(const Tables &T, int act, float x) {
  if (act == 1) return sigmoid_approx(T, x);
  if (act == 0) return tansig_approx(T, x);
  return x < 0.f ? 0.f : x;
}

NS_DEV float *rnn_seg_ptr(RnnSmem &r, int id) {
  switch (id) {
    case kSegFeat: return r.feat;
    case kSegDense: return r.dense;
    case kSegHVad: return r.hvad;
    case kSegHNoise: return r.hnoise;
    case kSegHDen: return r.hden;
    default: return r.rh;
  }
}

// acc[s] = bias[col] + sum_rows W[row][col] * act[row][s]   for one output column `col`
NS_DEV void rnn_matvec(const JobDesc &jd, const uint32_t *__restrict__ words,
                       const float *__restrict__ bias, RnnSmem &r, int col, float (&acc)[8]) {
  const float b = bias[jd.b_off + col];
#pragma unroll
  for (int s = 0; s < 8; s++) acc[s] = b;
  const uint32_t *w = words + jd.w_off + col;
  const int n_out = jd.n_out;
  for (int sg = 0; sg < jd.n_segs; sg++) {
    const float *a = rnn_seg_ptr(r, jd.seg_id[sg]);
    const int k4 = jd.seg_k4[sg];
    for (int kk = 0; kk < k4; kk++) {
      const uint32_t wv = *w;
      w += n_out;
#pragma unroll
      for (int bb = 0; bb < 4; bb++) {
        const float wf = (float)(int)(int8_t)((wv >> (8 * bb)) & 0xFFu);
        const f4 lo = ld4(a), hi = ld4(a + 4);
        a += 8;
        acc[0] = fmaf(wf, lo.x, acc[0]);
        acc[1] = fmaf(wf, lo.y, acc[1]);
        acc[2] = fmaf(wf, lo.z, acc[2]);
        acc[3] = fmaf(wf, lo.w, acc[3]);
        acc[4] = fmaf(wf, hi.x, acc[4]);
        acc[5] = fmaf(wf, hi.y, acc[5]);
        acc[6] = fmaf(wf, hi.z, acc[6]);
        acc[7] = fmaf(wf, hi.w, acc[7]);
      }
    }
  }
}

NS_DEV void rnn_dense(const RnnHeader &H, int job, const Params &p, const Tables &T, RnnSmem &r,
                      int tid, int nthr, float *dst) {
  const JobDesc &jd = H.jobs[job];
  for (int col = tid; col < jd.n_out; col += nthr) {
    float acc[8];
    rnn_matvec(jd, p.rnn_words, p.rnn_bias, r, col, acc);
#pragma unroll
    for (int s = 0; s < 8; s++) dst[col * 8 + s] = activate(T, jd.activation, acc[s] * (1.f / 256));
  }
}
// z and r gates of a GRU with N neurons: columns [0,N) -> z, [N,2N) -> r*h into rh
NS_DEV void rnn_gru_zr(const RnnHeader &H, int job, const Params &p, const Tables &T, RnnSmem &r,
                       int tid, int nthr, const float *h) {
  const JobDesc &jd = H.jobs[job];
  const int N = jd.n_out >> 1;
  for (int col = tid; col < jd.n_out; col += nthr) {
    float acc[8];
    rnn_matvec(jd, p.rnn_words, p.rnn_bias, r, col, acc);
    if (col < N) {
#pragma unroll
      for (int s = 0; s < 8; s++) r.z[col * 8 + s] = sigmoid_approx(T, acc[s] * (1.f / 256));
    } else {
      const int i = col - N;
#pragma unroll
      for (int s = 0; s < 8; s++) r.rh[i * 8 + s] = h[i * 8 + s] * sigmoid_approx(T, acc[s] * (1.f / 256));
    }
  }
}
NS_DEV void rnn_gru_c(const RnnHeader &H, int job, const Params &p, const Tables &T, RnnSmem &r,
                      int tid, int nthr, float *h) {
  const JobDesc &jd = H.jobs[job];
  for (int col = tid; col < jd.n_out; col += nthr) {
    float acc[8];
    rnn_matvec(jd, p.rnn_words, p.rnn_bias, r, col, acc);
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const float c = activate(T, jd.activation, acc[s] * (1.f / 256));
      const float z = r.z[col * 8 + s], ho = h[col * 8 + s];
      const float hn = z * ho + (1.f - z) * c;
      h[col * 8 + s] = r.silent[s] ? ho : hn;
    }
  }
}

NS_DEV void rnn_phase(const Params &p, const Tables &T, RnnSmem &r, int tid, int nthr) {
  const RnnHeader &H = *p.rnn_hdr;
  rnn_dense(H, 0, p, T, r, tid, nthr, r.dense);
  Simt::cta_sync();
  rnn_gru_zr(H, 1, p, T, r, tid, nthr, r.hvad);
  Simt::cta_sync();
  rnn_gru_c(H, 2, p, T, r, tid, nthr, r.hvad);
  Simt::cta_sync();
  rnn_gru_zr(H, 4, p, T, r, tid, nthr, r.hnoise);
  if (tid == nthr - 1) {  // vad_output rides along on the last thread
    float acc[8];
    rnn_matvec(H.jobs[3], p.rnn_words, p.rnn_bias, r, 0, acc);
#pragma unroll
    for (int s = 0; s < 8; s++) r.vad[s] = activate(T, H.jobs[3].activation, acc[s] * (1.f / 256));
  }
  Simt::cta_sync();
  rnn_gru_c(H, 5, p, T, r, tid, nthr, r.hnoise);
  Simt::cta_sync();
  rnn_gru_zr(H, 6, p, T, r, tid, nthr, r.hden);
  Simt::cta_sync();
  rnn_gru_c(H, 7, p, T, r, tid, nthr, r.hden);
  Simt::cta_sync();
  rnn_dense(H, 8, p, T, r, tid, nthr, r.gains);
  Simt::cta_sync();
}

// ------------------------------------------------------------------------------------------------
// a15: pitch filter + gain smoothing + interpolation, applied to X in place
// ------------------------------------------------------------------------------------------------
NS_DEV float interp_band(const Tables &T, const float *v, int k) {  // k < 400
  const int b = T.bin_band[k];
  const float f = T.bin_frac[k];
  return (1.f - f) * v[b] + f * v[b + 1];
}

NS_DEV void pitch_filter_and_gains(const Grp &g, const Tables &T, StreamSmem &s) {
  if (g.tid < kBands) {
    const int i = g.tid;
    const float e = s.Exp[i], gi = s.g[i];
    float r;
    if (e > gi)
      r = 1.f;
    else
      r = (e * e) * (1.f - gi * gi) / (.001f + (gi * gi) * (1.f - e * e));
    r = sqrtf(fminf(1.f, fmaxf(0.f, r)));
    r *= sqrtf(s.Ex[i] / (1e-8f + s.Ep[i]));
    s.r[i] = r;
  }
  gsync(g);
  for (int k = g.tid; k < 400; k += kGroupThreads) {
    const float rf = interp_band(T, s.r, k);
    s.X[k].x += rf * s.P[k].x;
    s.X[k].y += rf * s.P[k].y;
  }
  gsync(g);
  {
    const float e = band_accumulate(g, T, [&](int k) { return s.X[k].x * s.X[k].x + s.X[k].y * s.X[k].y; });
    if (g.tid < 4 * kBands && (g.tid & 3) == 0) s.newE[g.tid >> 2] = e;
  }
  gsync(g);
  if (g.tid < kBands) {
    const int i = g.tid;
    s.nrm[i] = sqrtf(s.Ex[i] / (1e-8f + s.newE[i]));
    const float gi = fmaxf(s.g[i], .6f * s.lastg[i]);
    s.g[i] = gi;
    s.lastg[i] = gi;
  }
  gsync(g);
  for (int k = g.tid; k < kFreq; k += kGroupThreads) {
    if (k < 400) {
      const float nf = interp_band(T, s.nrm, k), gf = interp_band(T, s.g, k);
      s.X[k].x = (s.X[k].x * nf) * gf;
      s.X[k].y = (s.X[k].y * nf) * gf;
    } else {
      s.X[k] = cf{0.f, 0.f};
    }
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// frame I/O
// ------------------------------------------------------------------------------------------------
NS_DEV void load_frame(const Grp &g, const Params &p, int stream, int t, float *slot) {
  const long long off = (long long)stream * p.in_stride + (long long)t * kFrame;
  if (p.flags & kFlagInI16) {
    const int16_t *src = reinterpret_cast<const int16_t *>(p.in) + off;
    for (int i = g.tid; i < kFrame; i += kGroupThreads) slot[i] = (float)src[i];
  } else {
    const float *src = reinterpret_cast<const float *>(p.in) + off;
    const float sc = (p.flags & kFlagUnitScale) ? 32768.0f : 1.0f;
    for (int i = g.tid; i < kFrame; i += kGroupThreads) slot[i] = src[i] * sc;
  }
}

NS_DEV void store_frame(const Grp &g, const Tables &T, const Params &p, StreamSmem &s, int stream, int t) {
  // X holds z[m] with x[2m] = z.x, x[2m+1] = -z.y.  out[i] = x[i] w[i] + synth[i]; synth = x[480+i] w[479-i]
  const int slot = t + p.out_frame_offset;
  const float *zb = reinterpret_cast<const float *>(s.X);
  for (int i = g.tid; i < kFrame; i += kGroupThreads) {
    const float x0 = (i & 1) ? -zb[i] : zb[i];  // zb[2m] = z[m].x, zb[2m+1] = z[m].y
    const float x1 = (i & 1) ? -zb[kFrame + i] : zb[kFrame + i];
    const float o = x0 * T.win[i] + s.synth[i];
    s.synth[i] = x1 * T.win[kFrame - 1 - i];
    if (slot >= 0) {
      if (p.flags & kFlagMixStereoI16) {
        float dn = o / 32768.0f;
        dn = fminf(1.f, fmaxf(-1.f, dn)) * p.volume;
        const long long o_off = (long long)stream * p.out_stride + (long long)slot * kFrame + i;
        float mixed = dn;
        if (p.app) mixed += p.app[(long long)stream * p.app_stride + (long long)slot * kFrame + i];
        mixed = fminf(1.f, fmaxf(-1.f, mixed));
        const int16_t q = (int16_t)(int)(mixed * 32767.0f);  // truncation toward zero, as Rust `as i16`
        int16_t *dst = reinterpret_cast<int16_t *>(p.out) + 2 * o_off;
        dst[0] = q;
        dst[1] = q;
      } else if (p.flags & kFlagOutI16) {
        const long long o_off = (long long)stream * p.out_stride + (long long)slot * kFrame + i;
        float v = rintf(o);
        v = fminf(32767.f, fmaxf(-32768.f, v));
        reinterpret_cast<int16_t *>(p.out)[o_off] = (int16_t)(int)v;
      } else {
        const long long o_off = (long long)stream * p.out_stride + (long long)slot * kFrame + i;
        float v = o;
        if (p.flags & kFlagUnitScale) v = fminf(1.f, fmaxf(-1.f, o / 32768.0f)) * p.volume;
        reinterpret_cast<float *>(p.out)[o_off] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// state load / store (HBM <-> shared memory), once per launch
// ------------------------------------------------------------------------------------------------
NS_DEV void load_state(const Grp &g, const float *st, StreamSmem &s, RnnSmem &r, int sidx) {
  for (int i = g.tid; i < kRing; i += kGroupThreads) s.ring[i] = st[kStRing + i];
  for (int i = g.tid; i < kFrame; i += kGroupThreads) s.synth[i] = st[kStSynth + i];
  for (int i = g.tid; i < kCepsMem * kBands; i += kGroupThreads) s.ceps[i] = st[kStCeps + i];
  if (g.tid < kBands) s.lastg[g.tid] = st[kStLastG + g.tid];
  if (g.tid < 24) r.hvad[g.tid * 8 + sidx] = st[kStHVad + g.tid];
  if (g.tid < 48) r.hnoise[g.tid * 8 + sidx] = st[kStHNoise + g.tid];
  if (g.tid < 96) r.hden[g.tid * 8 + sidx] = st[kStHDen + g.tid];
  if (g.tid == 0) {
    const double *hp = reinterpret_cast<const double *>(st + kStHp);
    s.hp[0] = hp[0];
    s.hp[1] = hp[1];
    s.last_gain = st[kStLastGain];
    const int *ip = reinterpret_cast<const int *>(st);
    s.last_period = ip[kStLastPeriod];
    s.memid = ip[kStMemId];
    s.ring_slot = ip[kStRingSlot];
    s.frame_count = *reinterpret_cast<const long long *>(st + kStFrameCount);
  }
}
NS_DEV void store_state(const Grp &g, float *st, const StreamSmem &s, const RnnSmem &r, int sidx) {
  for (int i = g.tid; i < kRing; i += kGroupThreads) st[kStRing + i] = s.ring[i];
  for (int i = g.tid; i < kFrame; i += kGroupThreads) st[kStSynth + i] = s.synth[i];
  for (int i = g.tid; i < kCepsMem * kBands; i += kGroupThreads) st[kStCeps + i] = s.ceps[i];
  if (g.tid < kBands) st[kStLastG + g.tid] = s.lastg[g.tid];
  if (g.tid < 24) st[kStHVad + g.tid] = r.hvad[g.tid * 8 + sidx];
  if (g.tid < 48) st[kStHNoise + g.tid] = r.hnoise[g.tid * 8 + sidx];
  if (g.tid < 96) st[kStHDen + g.tid] = r.hden[g.tid * 8 + sidx];
  if (g.tid == 0) {
    double *hp = reinterpret_cast<double *>(st + kStHp);
    hp[0] = s.hp[0];
    hp[1] = s.hp[1];
    st[kStLastGain] = s.last_gain;
    int *ip = reinterpret_cast<int *>(st);
    ip[kStLastPeriod] = s.last_period;
    ip[kStMemId] = s.memid;
    ip[kStRingSlot] = s.ring_slot;
    *reinterpret_cast<long long *>(st + kStFrameCount) = s.frame_count;
  }
}

// ------------------------------------------------------------------------------------------------
// the stream kernel body: S streams per CTA, S*128 threads
// ------------------------------------------------------------------------------------------------
template <int S>
NS_DEV void stream_kernel_body(const Params &p, CtaSmem<S> &sm) {
  const int tid_cta = Simt::tid();
  const int n_threads = S * kGroupThreads;
  Grp g;
  const int sidx = tid_cta / kGroupThreads;
  g.tid = tid_cta % kGroupThreads;
  g.lane = g.tid & 31;
  g.warp = g.tid >> 5;
  g.bar = 1 + sidx;
  const int stream = Simt::cta() * S + sidx;
  const bool active = stream < p.n_streams;
  StreamSmem &s = sm.st[sidx];
  RnnSmem &r = sm.rnn;
  const Tables &T = sm.tab;

  {  // tables -> shared memory (word copy), RNN staging cleared
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.tab);
    for (int i = tid_cta; i < (int)(sizeof(Tables) / 4); i += n_threads) dst[i] = src[i];
    float *rz = reinterpret_cast<float *>(&sm.rnn);
    for (int i = tid_cta; i < (int)(sizeof(RnnSmem) / 4); i += n_threads) rz[i] = 0.f;
  }
  Simt::cta_sync();
  float *st_g = p.state + (long long)(active ? stream : 0) * kStateFloats;
  if (active) load_state(g, st_g, s, r, sidx);
  Simt::cta_sync();

  for (int t = 0; t < p.n_frames; t++) {
    if (active) {
      const int w = s.ring_slot;
      float *slot = s.ring + w * kFrame;
      load_frame(g, p, stream, t, slot);
      gsync(g);
      biquad_frame(g, T, s, slot);
      gsync(g);
      const int base_pb = (w * kFrame + 672) % kRing;  // pitch_buf[0]
      pitch_downsample(g, s, base_pb);
      pitch_search(g, s);
      remove_doubling(g, s);
      // spectra: X of [prev | cur], P of the window lagged by pitch_index
      int base_an = base_pb + 768;
      if (base_an >= kRing) base_an -= kRing;
      int base_p = base_pb + 768 - s.pitch_index;
      if (base_p >= kRing) base_p -= kRing;
      rfft960_windowed(g, T, s.ring, base_an, s.X);
      rfft960_windowed(g, T, s.ring, base_p, s.P);
      {
        const float ex = band_accumulate(g, T, [&](int k) { return s.X[k].x * s.X[k].x + s.X[k].y * s.X[k].y; });
        const float ep = band_accumulate(g, T, [&](int k) { return s.P[k].x * s.P[k].x + s.P[k].y * s.P[k].y; });
        const float exp_ = band_accumulate(g, T, [&](int k) { return s.X[k].x * s.P[k].x + s.X[k].y * s.P[k].y; });
        if (g.tid < 4 * kBands && (g.tid & 3) == 0) {
          s.Ex[g.tid >> 2] = ex;
          s.Ep[g.tid >> 2] = ep;
          s.Exp[g.tid >> 2] = exp_;
        }
      }
      gsync(g);
      compute_features(g, T, s);
      const bool silent = s.silence != 0;
      if (g.tid < 44) r.feat[g.tid * 8 + sidx] = (silent || g.tid >= kFeatures) ? 0.f : s.feat[g.tid];
      if (g.tid == 0) r.silent[sidx] = silent ? 1 : 0;
    } else if (g.tid == 0) {
      r.silent[sidx] = 1;
    }
    Simt::cta_sync();
    rnn_phase(p, T, r, tid_cta, n_threads);
    if (active) {
      const bool silent = s.silence != 0;
      if (g.tid < kBands) s.g[g.tid] = silent ? 0.f : r.gains[g.tid * 8 + sidx];
      if (g.tid == 0) s.vad = silent ? 0.f : r.vad[sidx];
      gsync(g);
      if (!silent) pitch_filter_and_gains(g, T, s);
      if (p.dbg) {
        float *d = p.dbg + ((long long)stream * p.n_frames + t) * kDbgFloats;
        if (g.tid < kFeatures) d[kDbgFeatures + g.tid] = silent ? 0.f : s.feat[g.tid];
        if (g.tid < kBands) {
          d[kDbgGains + g.tid] = s.g[g.tid];
          d[kDbgEx + g.tid] = s.Ex[g.tid];
          d[kDbgEp + g.tid] = s.Ep[g.tid];
          d[kDbgExp + g.tid] = s.Exp[g.tid];
        }
        if (g.tid == 0) {
          d[kDbgPitchGain] = s.pitch_gain;
          d[kDbgVad] = s.vad;
          d[kDbgPitchIndex] = (float)s.pitch_index;
          d[kDbgSilence] = (float)s.silence;
        }
      }
      irfft960_inplace(g, T, s.X);
      store_frame(g, T, p, s, stream, t);
      if (g.tid == 0) {
        if (p.vad) p.vad[(long long)stream * p.vad_stride + t] = s.vad;
        s.ring_slot = (s.ring_slot + 1) & 3;
        s.frame_count += 1;
      }
      gsync(g);
    }
  }
  Simt::cta_sync();
  if (active) store_state(g, st_g, s, r, sidx);
}

}  // namespace ns
