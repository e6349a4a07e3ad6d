// ns_pitch7.cuh -- K1, second generation: the coarse pitch search as a FILTERED-EXACT decision.
//
// pitch_search's coarse pass (pitch.c) correlates the 4x-decimated frame with 147 lags (147 x 240 multiply-adds,
// half of K1's arithmetic) only to name the two best lags: nothing else of the 147 sums survives.  So the sums are
// first computed APPROXIMATELY on the tensor pipe -- the correlation is a Toeplitz product,
//     xc[16 n + m] = sum_k A[m][k] B[k][n],  A[m][k] = x4[k - m],  B[k][n] = y4[k + 16 n],
// i.e. one 16 x 16 x 256 bf16 MMA per frame with hi + lo split operands (three products, error ~1e-5 of
// sqrt(Sxx Syy)) -- and bracketed by a rigorous error bound.  Only the lags whose score bracket reaches the second
// best lower bound ("K'", two or three lags on speech and on noise alike) can end up among find_best_pitch's two
// winners; for those the sum is recomputed in the oracle's order (ascending index, rounded product + rounded add)
// and find_best_pitch's insertion runs over them in lag order with the exact running energy.  Lags outside K' lose
// every comparison against the two winners by a margin far above float32 rounding, and a lag that cannot win does
// not change which lags do, so the result equals the full sequential scan bit for bit.  Frames where the bracket
// does not separate (more than 32 candidates, or correlations so small that num = (xc 1e-12)^2 nears the float32
// underflow range, where relative margins mean nothing) take the exact path for all 147 lags.
//
// Everything downstream (fine search, remove_doubling's candidate table) is the first generation's code.
#pragma once

namespace ns {

constexpr int kP7Threads = 352;      // warps 0..7: one per frame; 8, 9, 10: the serial chains S, B, C
constexpr int kP7MaxK = 32;          // candidates per frame the filtered path accepts
constexpr float kP7ErrC = 1.0f / 4096.0f;  // |xc_approx - xc_exact| <= kP7ErrC * sqrt(Sxx * Syy): hi+lo bf16 products
                                           // (3 * 2^-18), f32 accumulation in the tensor core and in the exact sum
                                           // (< 2^-13 together), with a factor of two to spare
constexpr float kP7MinXc = 1e-2f;    // below this the threshold element's num = (xc 1e-12)^2 < 1e-28: exact path

template <int R>
struct PitchSmem7 {
  static constexpr int kHLen = R * kFrame + 1248;
  static constexpr int kXlpFloats = (R * kLpStride > kHLen) ? R * kLpStride : kHLen;
  // raw downsampled rows; after the FIR each row holds YE hi[216] | YE lo[216] | yy_lookup[388]
  // (YE word p = bf16 pair (y4[2p], y4[2p+1]) of the 4x-decimated whitened signal, hi and lo planes)
  alignas(16) float xr[R * kLpStride];
  alignas(16) float xlp[kXlpFloats];    // first the high-passed window, then the whitened rows x_lp
  alignas(16) float sb6[R][148];        // exact Syy before every coarse lag (chain S)
  alignas(16) float pbx[R][52];         // exclusive prefix of the sums of squares of y4 in blocks of eight
  alignas(16) float ac[R][8];
  alignas(16) float lpc2[R][8];
  alignas(16) float fx[R][12];
  alignas(16) int fi[R][12];
  alignas(16) float xx[R];
  alignas(16) float s10[R][12];
  alignas(16) float kxc[R][kP7MaxK];    // exact coarse correlation of the frame's candidates
  int klag[R][kP7MaxK];
  int kcnt[R], kflag[R];                // kflag: 0 = filtered, 1 = no lag can be positive, 2 = exact path
  int best0[R], best1[R], T0[R], nk[R];
  int n_tri[4], n_sgl[4];
  // from P10 on: the work lists and the inner products of remove_doubling; before that (P6) the exact correlation
  // of all 147 lags of the frames that take the exact path
  union {
    struct {
      uint32_t tri[4][R * 16], sgl[4][R * 16];
      float dots[R][64];
    } w;
    float xcf[R][152];
  } u;
};

NS_DEV int p7_ctz(unsigned v) {
#if defined(__CUDACC__) && !defined(NS_HOST_EMU)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

struct Top2 {  // two largest (value, tag) by value
  float v0, v1, t0, t1;
};
NS_DEV void top2_push(Top2 &s, float v, float t) {
  const bool a = v > s.v0, b = v > s.v1;
  s.v1 = a ? s.v0 : (b ? v : s.v1);
  s.t1 = a ? s.t0 : (b ? t : s.t1);
  s.v0 = a ? v : s.v0;
  s.t0 = a ? t : s.t0;
}

#ifdef NS_P7_STATS
extern long g_p7_stats[4];
#endif
template <int R, int NT>
NS_DEV void pitch_body7(const Params &p, PitchSmem7<R> &sm) {
  static_assert(NT == kP7Threads && R == 8, "warp w < R owns frame w; three chain warps follow");
  const int tid = Simt::tid();
  NS_PHASE_BEGIN();
  const int lane = tid & 31, warp = tid >> 5;
  const int runs_per_stream = (p.n_frames + R - 1) / R;
  const int stream = Simt::cta() / runs_per_stream;
  const int t0 = (Simt::cta() % runs_per_stream) * R;
  const int nfr = (p.n_frames - t0) < R ? (p.n_frames - t0) : R;
  const float *row = p.hp + (long long)stream * p.hp_stride + 192 + (long long)t0 * kFrame;
  float *h = sm.xlp;

  // P0: the window of high-passed samples these frames' pitch buffers cover
  if (tid < 4) sm.n_tri[tid] = sm.n_sgl[tid] = 0;
  {
    const int n4 = (nfr * kFrame + 1248) / 4;
    const f4 *src = reinterpret_cast<const f4 *>(row);
    f4 *dst = reinterpret_cast<f4 *>(h);
    for (int i = tid; i < n4; i += NT) dst[i] = src[i];
  }
  Simt::cta_sync();
  NS_PHASE_MARK(1);
  // P1: a9 2x downsample, pitch_buf[j] of frame f = h[480 f + j]
  for (int it = tid; it < nfr * (kLpLen / 4); it += NT) {
    const int f = it / (kLpLen / 4), i0 = 4 * (it - f * (kLpLen / 4));
    const float *x = h + f * kFrame + 2 * i0;
    const f4 b4 = ld4(x), c4 = ld4(x + 4);
    f4 v;
    if (i0 == 0)
      v.x = .5f * (.5f * b4.y + b4.x);
    else
      v.x = .5f * (.5f * (x[-1] + b4.y) + b4.x);
    v.y = .5f * (.5f * (b4.y + b4.w) + b4.z);
    v.z = .5f * (.5f * (b4.w + c4.y) + c4.x);
    v.w = .5f * (.5f * (c4.y + c4.w) + c4.z);
    *reinterpret_cast<f4 *>(sm.xr + f * kLpStride + i0) = v;
  }
  Simt::cta_sync();
  NS_PHASE_MARK(2);
  // P2: _celt_autocorr, lags 0..4: one lane per (frame, lag) chain, twenty chains per warp
  if (tid < 64) {
    const int l = tid & 31, fl = l / 5, k = l - 5 * fl, f = 4 * (tid >> 5) + fl;
    if (l < 20 && f < nfr) {
      const float *x = sm.xr + f * kLpStride;
      const float *y = x + k;
      f4 xv = ld4(x);
      float y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3];
      float sum = 0.f;
      NS_UNROLL(NS_DOT_UNROLL)
      for (int i = 0; i < kLpLen - 4; i += 4) {
        const f4 xn = ld4(x + i + 4);
        const float n0 = y[i + 4], n1 = y[i + 5], n2 = y[i + 6], n3 = y[i + 7];
        sum += xv.x * y0;
        sum += xv.y * y1;
        sum += xv.z * y2;
        sum += xv.w * y3;
        xv = xn;
        y0 = n0, y1 = n1, y2 = n2, y3 = n3;
      }
      float d = 0.f;
      for (int i = k + kLpLen - 4; i < kLpLen; i++) d += x[i] * x[i - k];
      sm.ac[f][k] = sum + d;
    }
  }
  Simt::cta_sync();
  NS_PHASE_MARK(3);
  // P3: lag window, _celt_lpc (order 4), bandwidth expansion, the extra zero
  if (tid < nfr) {
    const int f = tid;
    float ac[5], lpc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 5; k++) ac[k] = sm.ac[f][k];
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
    sm.lpc2[f][0] = lpc[0] + .8f;
    sm.lpc2[f][1] = lpc[1] + .8f * lpc[0];
    sm.lpc2[f][2] = lpc[2] + .8f * lpc[1];
    sm.lpc2[f][3] = lpc[3] + .8f * lpc[2];
    sm.lpc2[f][4] = .8f * lpc[3];
  }
  Simt::cta_sync();
  NS_PHASE_MARK(4);
  // P4: celt_fir5 with zero initial memory -> x_lp (overwrites the window h, which is dead now)
  for (int it = tid; it < nfr * (kLpLen / 4); it += NT) {
    const int f = it / (kLpLen / 4), i0 = 4 * (it - f * (kLpLen / 4));
    const float *x = sm.xr + f * kLpStride + i0;
    const float *n = sm.lpc2[f];
    const f4 z4 = f4{0.f, 0.f, 0.f, 0.f};
    const f4 p4 = (i0 >= 8) ? ld4(x - 8) : z4, q4 = (i0 >= 4) ? ld4(x - 4) : z4, r4 = ld4(x);
    const float w[9] = {p4.w, q4.x, q4.y, q4.z, q4.w, r4.x, r4.y, r4.z, r4.w};
    const float n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3], n4 = n[4];
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float sum = w[5 + u];
      sum += n0 * w[4 + u];
      sum += n1 * w[3 + u];
      sum += n2 * w[2 + u];
      sum += n3 * w[1 + u];
      sum += n4 * w[u];
      o[u] = sum;
    }
    *reinterpret_cast<f4 *>(sm.xlp + f * kLpStride + i0) = f4{o[0], o[1], o[2], o[3]};
  }
  Simt::cta_sync();
  NS_PHASE_MARK(5);
  // P4b: the 4x-decimated signal y4[m] = x_lp[2m] as bf16 hi + lo pairs for the tensor pipe (x4[j] = y4[192 + j])
  for (int it = tid; it < nfr * 216; it += NT) {
    const int f = it / 216, w = it - f * 216;
    const f4 v = ld4(sm.xlp + f * kLpStride + 4 * w);  // y4[2w] = v.x, y4[2w+1] = v.z
    uint32_t hi, lo;
    bf16_split2(v.x, v.z, hi, lo);
    uint32_t *ye = reinterpret_cast<uint32_t *>(sm.xr + f * kLpStride);
    ye[w] = hi;
    ye[216 + w] = lo;
  }
  Simt::cta_sync();
  NS_PHASE_MARK(6);

  // ---- serial chains on their own warps (lane = frame), a budgeted number of 16-element blocks per phase, state in
  // registers across the barriers:
  //   S (warp 8): find_best_pitch's running energy of the coarse pass: Syy = 1 + sum_{j<240} y4[j]^2, then
  //     Syy <- max(1, Syy + y4[i+240]^2 - y4[i]^2), the value before every lag kept in sb6;
  //   B (warp 9): the same for the fine pass (480 taps of x_lp), kept before each of the ten candidate lags (s10);
  //   C (warp 10): remove_doubling's xx = sum x[j]^2 and its yy_lookup recurrence.
  const bool chain_s = warp == R && lane < nfr, chain_b = warp == R + 1 && lane < nfr, chain_c = warp == R + 2 && lane < nfr;
  float ch_acc = (chain_b || chain_s) ? 1.f : 0.f;
  int ch_blk = 0;
  auto chain_s_run = [&]() {  // the whole chain in one go (387 steps): done before the first barrier after P5
    if (!chain_s) return;
    const int f = lane;
    const float *x = sm.xlp + f * kLpStride;  // y4[m] = x[2m]
    float syy = 1.f;
    {
      f4 a = ld4(x), b = ld4(x + 4);
      for (int j = 0; j < 240; j += 4) {  // the next trip's operands are requested before this trip's chain of adds
        const f4 an = ld4(x + 2 * j + 8), bn = ld4(x + 2 * j + 12);
        const float p0 = a.x * a.x, p1 = a.z * a.z, p2 = b.x * b.x, p3 = b.z * b.z;
        syy += p0;
        syy += p1;
        syy += p2;
        syy += p3;
        a = an, b = bn;
      }
    }
    f4 a0 = ld4(x + 480), a1 = ld4(x + 484), b0 = ld4(x), b1 = ld4(x + 4);
    for (int i0 = 0; i0 < 148; i0 += 4) {
      const f4 a0n = ld4(x + 2 * (i0 + 244)), a1n = ld4(x + 2 * (i0 + 244) + 4), b0n = ld4(x + 2 * i0 + 8), b1n = ld4(x + 2 * i0 + 12);
      const float d0 = a0.x * a0.x - b0.x * b0.x, d1 = a0.z * a0.z - b0.z * b0.z, d2 = a1.x * a1.x - b1.x * b1.x,
                  d3 = a1.z * a1.z - b1.z * b1.z;
      f4 o;
      o.x = syy;
      syy += d0;
      syy = syy < 1.f ? 1.f : syy;
      o.y = syy;
      syy += d1;
      syy = syy < 1.f ? 1.f : syy;
      o.z = syy;
      syy += d2;
      syy = syy < 1.f ? 1.f : syy;
      o.w = syy;
      syy += d3;
      syy = syy < 1.f ? 1.f : syy;
      *reinterpret_cast<f4 *>(sm.sb6[f] + i0) = o;
      a0 = a0n, a1 = a1n, b0 = b0n, b1 = b1n;
    }
  };
  auto chain_run = [&](int budget, bool recur_b) {
    if (!chain_b && !chain_c) return;
    const int f = lane;
    const float *y = sm.xlp + f * kLpStride + (chain_c ? 384 : 0);
    for (; budget > 0 && ch_blk < 30; budget--, ch_blk++) {
      const float *q = y + 16 * ch_blk;
      const f4 v0 = ld4(q), v1 = ld4(q + 4), v2 = ld4(q + 8), v3 = ld4(q + 12);
      ch_acc += v0.x * v0.x, ch_acc += v0.y * v0.y, ch_acc += v0.z * v0.z, ch_acc += v0.w * v0.w;
      ch_acc += v1.x * v1.x, ch_acc += v1.y * v1.y, ch_acc += v1.z * v1.z, ch_acc += v1.w * v1.w;
      ch_acc += v2.x * v2.x, ch_acc += v2.y * v2.y, ch_acc += v2.z * v2.z, ch_acc += v2.w * v2.w;
      ch_acc += v3.x * v3.x, ch_acc += v3.y * v3.y, ch_acc += v3.z * v3.z, ch_acc += v3.w * v3.w;
      if (ch_blk == 29 && chain_c) {
        sm.xx[f] = ch_acc;
        sm.xr[f * kLpStride + 432] = ch_acc;  // yy_lookup[0]
      }
    }
    if (chain_b) {
      if (ch_blk < 30 || !recur_b) return;  // the recurrence waits for the coarse winners
      const int lo0 = 2 * sm.best0[f] - 2, lo1 = 2 * sm.best1[f] - 2;
      const int r1 = lo0 < lo1 ? lo0 : lo1, r2 = lo0 < lo1 ? lo1 : lo0;
      auto clampi = [](int v) { return v < 0 ? 0 : (v > 294 ? 294 : v); };
      const int a0 = clampi(r1), b0 = clampi(r2 > r1 + 5 ? r2 : r1 + 5), last = clampi(r2 + 5);
      float syy = ch_acc;
      for (int i0 = 0; i0 < last; i0 += 4) {
        const f4 ya = ld4(y + i0 + 480), yb = ld4(y + i0);
        const float d[4] = {ya.x * ya.x - yb.x * yb.x, ya.y * ya.y - yb.y * yb.y, ya.z * ya.z - yb.z * yb.z,
                            ya.w * ya.w - yb.w * yb.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          if ((unsigned)(i - a0) < 5u) sm.s10[f][i - a0] = syy;
          if ((unsigned)(i - b0) < 5u) sm.s10[f][5 + i - b0] = syy;
          syy += d[u];
          syy = syy < 1.f ? 1.f : syy;
        }
      }
      ch_blk = 1000;
    } else {
      float *yyl = sm.xr + f * kLpStride + 432;  // yy_lookup[0..384]
      for (; budget > 0 && ch_blk < 30 + 24; budget--, ch_blk++) {
        const int i1 = 1 + 16 * (ch_blk - 30);
        float yy = ch_acc;
#pragma unroll
        for (int hh = 0; hh < 4; hh++) {
          const int i0 = i1 + 4 * hh;
          const f4 a = ld4(y - i0 - 3), c = ld4(y + 477 - i0);
          const float av[4] = {a.w * a.w, a.z * a.z, a.y * a.y, a.x * a.x};
          const float cv[4] = {c.w * c.w, c.z * c.z, c.y * c.y, c.x * c.x};
#pragma unroll
          for (int u = 0; u < 4; u++) {
            yy = yy + av[u] - cv[u];
            yyl[i0 + u] = yy < 0.f ? 0.f : yy;
          }
        }
        ch_acc = yy;
      }
    }
  };

  // P5: the coarse correlation on the tensor pipe and its filter; warp f < R owns frame f
  if (warp < R) {
    const int f = warp;
    if (f < nfr) {
      const uint32_t *YEh = reinterpret_cast<const uint32_t *>(sm.xr + f * kLpStride), *YEl = YEh + 216;
      const float *xl = sm.xlp + f * kLpStride;  // y4[m] = xl[2m], x4[j] = xl[384 + 2j]
      const int g = lane >> 2, t = lane & 3;
      // sums of squares of y4 in blocks of eight -> exclusive prefix pbx[0..49]; largest magnitude of y4 (and so of x4)
      float amax = 0.f;
      {
        float bs0 = 0.f, bs1 = 0.f;  // blocks `lane` and `lane + 32` (< 54: y4[0 .. 432), which includes x4)
        {
          const float *q = xl + 16 * lane;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const f4 v = ld4(q + 4 * u);
            bs0 = fmaf(v.x, v.x, bs0);
            bs0 = fmaf(v.z, v.z, bs0);
            amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.z)));
          }
        }
        if (lane + 32 < 54) {
          const float *q = xl + 16 * (lane + 32);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const f4 v = ld4(q + 4 * u);
            if (lane + 32 < 49) {
              bs1 = fmaf(v.x, v.x, bs1);
              bs1 = fmaf(v.z, v.z, bs1);
            }
            amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.z)));
          }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) amax = fmaxf(amax, Simt::shfl_xor(amax, d));
        float s0 = bs0, s1 = bs1;  // inclusive scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float u0 = Simt::shfl_up(s0, d), u1 = Simt::shfl_up(s1, d);
          if (lane >= d) s0 += u0, s1 += u1;
        }
        const float tot0 = Simt::shfl(s0, 31);
        sm.pbx[f][lane] = s0 - bs0;
        if (lane + 32 < 52) sm.pbx[f][lane + 32] = tot0 + s1 - bs1;
      }
      Simt::warp_sync();
      NS_PHASE_MARK(20);
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[nt][e] = 0.f;
      // A[m][k] = x4[k - m] = y4[192 + k - m] (zero outside 0 <= k - m < 240).  Row g / g + 8, columns 2t, 2t + 1 (+ 8):
      // with e0 = k0 + 2t - g the four A registers are the pairs at x4 index e0, e0 - 8, e0 + 8, e0 (Toeplitz), and the
      // pair at e0 + 8 of one k-step is the pair at e0 - 8 of the next.  The pair's alignment is the parity of g, a
      // lane constant: an even pair is one word of YE, an odd one the high half of a word and the low half of the next.
      const bool oddg = (g & 1) != 0;
      auto pair_at = [&](const uint32_t *YE, int e) -> uint32_t {  // (x4[e], x4[e + 1]), e + 192 >= 0
        const int w = (192 + e) >> 1;
        const uint32_t a = YE[w];
        if (!oddg) return a;
        const uint32_t b = YE[w + 1];
        return (a >> 16) | (b << 16);
      };
      auto edge_mask = [](int e) -> uint32_t {
        uint32_t m = 0u;
        if (e >= 0 && e < 240) m |= 0x0000FFFFu;
        if (e + 1 >= 0 && e + 1 < 240) m |= 0xFFFF0000u;
        return m;
      };
      const int e00 = 2 * t - g;
      uint32_t ph_lo = pair_at(YEh, e00 - 8) & edge_mask(e00 - 8), pl_lo = pair_at(YEl, e00 - 8) & edge_mask(e00 - 8);
      const uint32_t *bh0 = YEh + t + 8 * g, *bl0 = YEl + t + 8 * g;  // B[k][n] = y4[k + 16 n]: word k0/2 + t + 8 n
#pragma unroll 4
      for (int ks = 0; ks < 16; ks++) {
        const int k0 = 16 * ks, e0 = k0 + e00;
        uint32_t ah[4], al[4];
        ah[1] = ph_lo, al[1] = pl_lo;
        ah[0] = pair_at(YEh, e0), al[0] = pair_at(YEl, e0);
        ah[2] = pair_at(YEh, e0 + 8), al[2] = pair_at(YEl, e0 + 8);
        if (ks == 0) {  // the band's upper-left corner: columns before the row's first tap
          const uint32_t m = edge_mask(e0);
          ah[0] &= m, al[0] &= m;
        }
        if (ks == 15) {  // and its lower-right corner
          const uint32_t m0 = edge_mask(e0), m2 = edge_mask(e0 + 8);
          ah[0] &= m0, al[0] &= m0, ah[2] &= m2, al[2] &= m2;
        }
        ah[3] = ah[0], al[3] = al[0];
        ph_lo = ah[2], pl_lo = al[2];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          const int w = (k0 >> 1) + 64 * nt;
          const uint32_t bh[2] = {bh0[w], bh0[w + 4]}, bl[2] = {bl0[w], bl0[w + 4]};
          Simt::mma_bf16_16816(acc[nt], ah, bh);
          Simt::mma_bf16_16816(acc[nt], ah, bl);
          Simt::mma_bf16_16816(acc[nt], al, bh);
        }
      }
      NS_PHASE_MARK(21);
      // this lane's eight lags: tile nt, element e -> m = g + 8 (e >> 1), n = 8 nt + 2 t + (e & 1), lag 16 n + m
      const float sxx_w = [&]() {  // Sxx = sum x4^2 (x4 = y4[192 .. 432))
        float s = 0.f;
        const float *q = xl + 384 + 16 * lane;  // 480 floats = 30 lanes x 16
        if (lane < 30) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const f4 v = ld4(q + 4 * u);
            s = fmaf(v.x, v.x, s);
            s = fmaf(v.z, v.z, s);
          }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += Simt::shfl_xor(s, d);
        return s;
      }();
      const float e_run = 3e-5f * (1.f + sm.pbx[f][49]);
      float up[8], lo[8], xm[8];
      int lagv[8];
      Top2 tp{-1.f, -1.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int nt = q >> 2, e = q & 3;
        const int m = g + 8 * (e >> 1), n = 8 * nt + 2 * t + (e & 1);
        const int i = 16 * n + m;
        lagv[q] = i;
        const bool valid = i < 147;
        const int c = (valid ? i : 0) >> 3;
        // window [i, i + 240) of y4 contains blocks c+1 .. c+29 and lies inside blocks c .. c+30.  The oracle's Syy is a
        // float32 running sum over everything from y4[0] on, so it may be off the window's true energy by the rounding
        // of ~400 operations on values up to the total energy: e_run
        const float w_lo = sm.pbx[f][c + 30] - sm.pbx[f][c + 1], w_hi = sm.pbx[f][c + 31] - sm.pbx[f][c];
        const float syy_lo = fmaxf(1.f, 1.f + w_lo * 0.9999f - e_run), syy_hi = 1.f + w_hi * 1.0001f + e_run;
        const float xa = acc[nt][e];
        const float dl = kP7ErrC * sqrtf(sxx_w * w_hi) * 1.001f;
        const float hi_x = fmaxf(xa + dl, 0.f), lo_x = fmaxf(xa - dl, 0.f);
        up[q] = valid ? (hi_x * hi_x) / syy_lo * 1.0002f : -1.f;
        lo[q] = valid ? (lo_x * lo_x) / syy_hi * 0.9998f : -1.f;
        xm[q] = lo_x;
        if (!(xa == xa) || !(dl == dl)) up[q] = valid ? 3.0e38f : -1.f;  // NaN / overflow: keep the lag (forces the exact path)
        top2_push(tp, lo[q], xm[q]);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {  // warp-wide two largest lower bounds
        const float ov0 = Simt::shfl_xor(tp.v0, d), ov1 = Simt::shfl_xor(tp.v1, d);
        const float ot0 = Simt::shfl_xor(tp.t0, d), ot1 = Simt::shfl_xor(tp.t1, d);
        top2_push(tp, ov0, ot0);
        top2_push(tp, ov1, ot1);
      }
      const float L2 = tp.v1;        // second best lower bound of the score (-1: fewer than two valid lags)
      const float L2_xc = tp.t1;     // xc - delta of that lag
      unsigned mine = 0u;
      int cnt = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const bool in = up[q] > 0.f && up[q] >= L2;  // up > 0 <=> the lag's correlation may be positive
        if (in) mine |= 1u << q, cnt++;
      }
      int incl = cnt;  // inclusive scan of the counts -> slots
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = Simt::shfl(incl, lane >= d ? lane - d : lane);
        if (lane >= d) incl += u;
      }
      const int total = Simt::shfl(incl, 31);
      // "no lag can be positive" is only certain when the decimated signal is exactly zero: on a signal decaying
      // through the denormal range the tensor pipe and the squares above flush to zero while the oracle's sums do not
      int flag = 0;
      if (total == 0)
        flag = (amax == 0.f) ? 1 : 2;
      else if (total > kP7MaxK || !(L2 > 0.f) || !(L2_xc >= kP7MinXc))
        flag = 2;
      if (flag == 0) {
        int slot = incl - cnt;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (mine & (1u << q)) sm.klag[f][slot++] = lagv[q];
      }
      NS_PHASE_MARK(22);
      if (lane == 0) {
        sm.kcnt[f] = flag == 0 ? total : 0;
        sm.kflag[f] = flag;
#ifdef NS_P7_STATS  // measurement builds of the host emulation: how often each path is taken, candidates per frame
        __atomic_fetch_add(&g_p7_stats[flag], 1, __ATOMIC_RELAXED);
        if (flag == 0) __atomic_fetch_add(&g_p7_stats[3], total, __ATOMIC_RELAXED);
#endif
#if defined(NS_PHASE_CLOCKS) && defined(__CUDACC__) && !defined(NS_HOST_EMU)
        atomicAdd(&g_pitch_phase_cycles[16 + flag], 1ull);  // frames per path; [19]: candidates on the filtered path
        if (flag == 0) atomicAdd(&g_pitch_phase_cycles[19], (unsigned long long)total);
#endif
      }
    }
  }
  chain_s_run();
  chain_run(15, false);
  Simt::cta_sync();
  NS_PHASE_MARK(7);
  // P6b: the candidates' correlations in the oracle's order, one lane per (frame, candidate); the frames on the exact
  // path get all 147 lags
  {
    int cum[R + 1];
    cum[0] = 0;
#pragma unroll
    for (int f = 0; f < R; f++) cum[f + 1] = cum[f] + (f < nfr ? (sm.kflag[f] == 2 ? 147 : sm.kcnt[f]) : 0);
    for (int q = tid; q < cum[R]; q += NT) {
      int f = 0;
#pragma unroll
      for (int ff = 1; ff < R; ff++)
        if (q >= cum[ff]) f = ff;
      const int it = q - cum[f];
      const bool all = sm.kflag[f] == 2;
      const int lag = all ? it : sm.klag[f][it];
      const float *xl = sm.xlp + f * kLpStride;
      const float *x4 = xl + 384;
      const bool odd = (lag & 1) != 0;
      const float *yq = xl + 2 * lag - (odd ? 2 : 0);  // 16-byte aligned; y4[lag + j] = xl[2 (lag + j)]
      float sum = 0.f;
      f4 xa = ld4(x4), ya = ld4(yq);
#pragma unroll 4
      for (int j = 0; j < 240; j += 2) {
        const f4 xn = ld4(x4 + 2 * j + 4), yn = ld4(yq + 2 * j + 4);
        // taps j, j+1: x4[j] = xa.x, x4[j+1] = xa.z; y4[lag+j], y4[lag+j+1]: even lag -> ya.x, ya.z; odd lag -> ya.z, yn.x
        const float y0 = odd ? ya.z : ya.x, y1 = odd ? yn.x : ya.z;
        sum += xa.x * y0;
        sum += xa.z * y1;
        xa = xn;
        ya = yn;
      }
      if (all)
        sm.u.xcf[f][lag] = sum;
      else
        sm.kxc[f][it] = sum;
    }
  }
  chain_run(8, false);
  Simt::cta_sync();
  NS_PHASE_MARK(8);
  // P6c: find_best_pitch's insertion over the candidates in lag order (frame warp), or over all lags (exact path)
  if (warp < R && warp < nfr) {
    const int f = warp;
    const int flag = sm.kflag[f];
    Best2 b;
    best_init(b);
    if (flag == 0) {
      const int n = sm.kcnt[f];
      int mylag = lane < n ? sm.klag[f][lane] : 0x7FFFFFFF;
      const float myxc = lane < n ? sm.kxc[f][lane] : 0.f;
      const float mysyy = lane < n ? sm.sb6[f][mylag] : 1.f;
      for (int s = 0; s < n; s++) {
        int mn = mylag;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const int o = Simt::shfl_xor(mn, d);
          mn = o < mn ? o : mn;
        }
        const unsigned who = Simt::ballot(mylag == mn);
        const int src = p7_ctz(who);
        const float xc = Simt::shfl(myxc, src), syy = Simt::shfl(mysyy, src);
        const float x16 = xc * 1e-12f;
        best_insert_sel(b, xc > 0.f, x16 * x16, syy, mn);
        if (lane == src) mylag = 0x7FFFFFFF;
      }
    } else if (flag == 2) {
      if (lane == 0) {
        for (int i = 0; i < 147; i++) {
          const float xc = sm.u.xcf[f][i];
          const float x16 = xc * 1e-12f;
          best_insert_sel(b, xc > 0.f, x16 * x16, sm.sb6[f][i], i);
        }
      }
      b.p0 = Simt::shfl(b.p0, 0);
      b.p1 = Simt::shfl(b.p1, 0);
    }
    if (lane == 0) {
      sm.best0[f] = b.p0;
      sm.best1[f] = b.p1;
    }
  }
  chain_run(7, false);  // B's sum of squares completes here
  Simt::cta_sync();
  NS_PHASE_MARK(9);
  // P7: fine search, at most ten lags around 2*best0 and 2*best1 (chains B and C finish in its shadow)
  for (int it = tid; it < nfr * 10; it += NT) {
    const int f = it / 10, c = it - f * 10;
    const float *lp = sm.xlp + f * kLpStride;
    const int c0 = 2 * sm.best0[f], c1 = 2 * sm.best1[f];
    const int i = (c < 5) ? (c0 - 2 + c) : (c1 - 2 + (c - 5));
    const int dd = i - c0;
    const bool ok = (i >= 0) && (i < 294) && (c < 5 || dd > 2 || dd < -2);
    float sum = 0.f;
    if (ok) sum = dot480_shifted(lp, i);
    sm.fi[f][c] = ok ? i : -1;
    sm.fx[f][c] = sum < -1.f ? -1.f : sum;
  }
  chain_run(1000, true);
  Simt::cta_sync();
  NS_PHASE_MARK(10);
  // P8: find_best_pitch on the fine correlation (zero outside the candidates), pseudo-interpolation
  if (tid < nfr) {
    const int f = tid;
    const int lo0 = 2 * sm.best0[f] - 2, lo1 = 2 * sm.best1[f] - 2;
    auto xcorr_at = [&](int i) -> float {
      const int d0 = i - lo0, d1 = i - lo1;
      if (d0 >= 0 && d0 < 5 && sm.fi[f][d0] == i) return sm.fx[f][d0];
      if (d1 >= 0 && d1 < 5 && sm.fi[f][5 + d1] == i) return sm.fx[f][5 + d1];
      return 0.f;
    };
    Best2 b;
    best_init(b);
    const int r1 = lo0 < lo1 ? lo0 : lo1, r2 = lo0 < lo1 ? lo1 : lo0;
    auto clampi = [](int v) { return v < 0 ? 0 : (v > 294 ? 294 : v); };
    const int a0 = clampi(r1), a1 = clampi(r1 + 5), b0 = clampi(r2 > r1 + 5 ? r2 : r1 + 5), b1 = clampi(r2 + 5);
    auto checked = [&](int from, int to, const float *syy) {
      for (int i = from; i < to; i++) {
        const float xc = xcorr_at(i);
        const float x16 = xc * 1e-12f;
        best_insert_sel(b, xc > 0.f, x16 * x16, syy[i - from], i);
      }
    };
    checked(a0, a1, sm.s10[f]);
    checked(b0, b1, sm.s10[f] + 5);
    const int bp = b.p0;
    int offset = 0;
    if (bp > 0 && bp < 293) {
      const float a = xcorr_at(bp - 1), bb = xcorr_at(bp), cc = xcorr_at(bp + 1);
      if ((cc - a) > .7f * (bb - a))
        offset = 1;
      else if ((a - cc) > .7f * (bb - cc))
        offset = -1;
    }
    const int pitch_index = kPitchMax - (2 * bp - offset);
    int T0 = pitch_index / 2;
    if (T0 >= 384) T0 = 383;
    sm.T0[f] = T0;
  }
  Simt::cta_sync();
  NS_PHASE_MARK(11);
  // P10: the candidate work list of remove_doubling, bucketed by window alignment (as the first generation)
  for (int it = tid; it < nfr * 16; it += NT) {
    const int f = it >> 4, k = it & 15;
    if (k == 0) continue;
    const int T0 = sm.T0[f];
    if (k > 1 && rd_T1(k, T0) < 30) continue;
    if (k == kMaxK || rd_T1(k + 1, T0) < 30) sm.nk[f] = k;
    const int Tc = (k == 1) ? T0 : rd_T1(k, T0);
    {
      const int bkt = (384 - Tc - 1) & 3;
      const int idx = Simt::atomic_add_shared(&sm.n_tri[bkt], 1);
      sm.u.w.tri[bkt][idx] = (uint32_t)(f | (Tc << 5) | (k << 14));
    }
    if (k > 1) {
      const int T1b = rd_T1b(k, T0, Tc);
      const int bkt = (384 - T1b) & 3;
      const int idx = Simt::atomic_add_shared(&sm.n_sgl[bkt], 1);
      sm.u.w.sgl[bkt][idx] = (uint32_t)(f | (T1b << 5) | (k << 14));
    }
  }
  Simt::cta_sync();
  NS_PHASE_MARK(12);
  // P11: the inner products.  Jobs 0..7: triples of bucket j % 4, half j / 4; jobs 8..11: singles of bucket j - 8
  {
    constexpr int NW = NT / 32;
    const int nwork = NW, w = warp;
    for (int job = w; job < 12; job += nwork) {
      if (job < 8) {
        const int bkt = job & 3, n = sm.n_tri[bkt];
        for (int it = (job >> 2) * 32 + lane; it < n; it += 64) {
          const int e = sm.u.w.tri[bkt][it];
          const int f = e & 31, Tc = (e >> 5) & 0x1FF, k = e >> 14;
          const float *rowp = sm.xlp + f * kLpStride;
          const float *ya = rowp + ((384 - Tc - 1) & ~3);
          float sp, sc, sm1;
          switch (bkt) {
            case 0: dot3_fixed_shift<480, 0>(rowp + 384, ya, sp, sc, sm1); break;
            case 1: dot3_fixed_shift<480, 1>(rowp + 384, ya, sp, sc, sm1); break;
            case 2: dot3_fixed_shift<480, 2>(rowp + 384, ya, sp, sc, sm1); break;
            default: dot3_fixed_shift<480, 3>(rowp + 384, ya, sp, sc, sm1); break;
          }
          const int d = 2 + 4 * (k - 2);
          sm.u.w.dots[f][k == 1 ? 0 : d] = sm1;
          sm.u.w.dots[f][k == 1 ? kDotXy0 : d + 1] = sc;
          sm.u.w.dots[f][k == 1 ? 1 : d + 2] = sp;
        }
      } else {
        const int bkt = job - 8, n = sm.n_sgl[bkt];
        for (int it = lane; it < n; it += 32) {
          const int e = sm.u.w.sgl[bkt][it];
          const int f = e & 31, lag = (e >> 5) & 0x1FF, k = e >> 14;
          const float *rowp = sm.xlp + f * kLpStride;
          const float *ya = rowp + ((384 - lag) & ~3);
          float sum;
          switch (bkt) {
            case 0: sum = dot_fixed_shift<480, 0>(rowp + 384, ya); break;
            case 1: sum = dot_fixed_shift<480, 1>(rowp + 384, ya); break;
            case 2: sum = dot_fixed_shift<480, 2>(rowp + 384, ya); break;
            default: sum = dot_fixed_shift<480, 3>(rowp + 384, ya); break;
          }
          sm.u.w.dots[f][2 + 4 * (k - 2) + 3] = sum;
        }
      }
    }
  }
  Simt::cta_sync();
  NS_PHASE_MARK(13);
  // P12: per candidate k: gain, the pitch gain it would report, refined pitch index -> table
  for (int it = tid; it < nfr * 16; it += NT) {
    const int f = it >> 4, k = it & 15;
    uint32_t *tab = p.tab + ((long long)stream * p.chunk_cap + (t0 + f)) * kTabWords;
    const int T0 = sm.T0[f], nk = sm.nk[f];
    if (k == 0) {
      tab[0] = (uint32_t)T0 | ((uint32_t)nk << 16);
      tab[1] = 0u;
      continue;
    }
    if (k > nk) continue;
    const float *yyl = sm.xr + f * kLpStride + 432;
    const float xx = sm.xx[f];
    float xy, yy, c0, c1, c2;
    int T;
    if (k == 1) {
      T = T0;
      xy = sm.u.w.dots[f][kDotXy0];
      yy = yyl[T0];
      c0 = sm.u.w.dots[f][0];
      c1 = xy;
      c2 = sm.u.w.dots[f][1];
    } else {
      const int d = 2 + 4 * (k - 2);
      T = rd_T1(k, T0);
      const int T1b = rd_T1b(k, T0, T);
      c0 = sm.u.w.dots[f][d];
      c1 = sm.u.w.dots[f][d + 1];
      c2 = sm.u.w.dots[f][d + 2];
      xy = .5f * (c1 + sm.u.w.dots[f][d + 3]);
      yy = .5f * (yyl[T] + yyl[T1b]);
    }
    const float g = pitch_gain_f(xy, xx, yy);
    const float bxy = xy < 0.f ? 0.f : xy;
    float pg = (yy <= bxy) ? 1.f : bxy / (yy + 1.f);
    if (pg > g) pg = g;
    int offset = 0;
    if ((c2 - c0) > .7f * (c1 - c0))
      offset = 1;
    else if ((c0 - c2) > .7f * (c1 - c2))
      offset = -1;
    int pi = 2 * T + offset;
    if (pi < kPitchMin) pi = kPitchMin;
    uint32_t *e = tab + 2 + 3 * (k - 1);
    e[0] = (uint32_t)T | ((uint32_t)pi << 16);
    e[1] = f2u(g);
    e[2] = f2u(pg);
  }
  NS_PHASE_MARK(14);
}

}  // namespace ns
