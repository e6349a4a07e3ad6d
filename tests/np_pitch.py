"""Independent NumPy transliteration of RNNoise's pitch analysis (xiph/rnnoise denoise.c biquad + pitch.c + celt_lpc.c,
the code nnnoiseless 0.5.2 ports), written against the published algorithm and not against
oracle/rnnoise_oracle.c, to cross-check the C oracle's pitch decisions.  float32 arithmetic in the upstream order:
a sum over products is np.cumsum over the rounded products (cumsum accumulates sequentially, unlike np.sum)."""
import numpy as np

F32 = np.float32
FRAME, PITCH_BUF, MINP, MAXP, PFRAME = 480, 1728, 60, 768, 960
SECOND_CHECK = [0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2]


def seqdot(a, b):
    a = np.asarray(a, F32)
    b = np.asarray(b, F32)
    if len(a) == 0:
        return F32(0)
    return np.cumsum(a * b, dtype=F32)[-1]


def biquad(x, mem):
    """denoise.c biquad(): f32 state, f64 intermediates; b = [-2, 1], a = [-1.99599, 0.99600]"""
    b0, b1, a0, a1 = -2.0, 1.0, float(F32(-1.99599)), float(F32(0.99600))
    y = np.empty(len(x), F32)
    m0, m1 = F32(mem[0]), F32(mem[1])
    for i, xi in enumerate(np.asarray(x, F32)):
        yi = F32(xi + m0)
        m0 = F32(float(m1) + (b0 * float(xi) - a0 * float(yi)))
        m1 = F32(b1 * float(xi) - a1 * float(yi))
        y[i] = yi
    mem[0], mem[1] = m0, m1
    return y


def celt_lpc(ac, p):
    lpc = np.zeros(p, F32)
    error = F32(ac[0])
    if ac[0] != 0:
        for i in range(p):
            rr = F32(0)
            for j in range(i):
                rr = F32(rr + F32(lpc[j] * ac[i - j]))
            rr = F32(rr + ac[i + 1])
            r = F32(-rr / error)
            lpc[i] = r
            for j in range((i + 1) >> 1):
                t1, t2 = lpc[j], lpc[i - 1 - j]
                lpc[j] = F32(t1 + F32(r * t2))
                lpc[i - 1 - j] = F32(t2 + F32(r * t1))
            error = F32(error - F32(F32(r * r) * error))
            if error < F32(F32(.001) * ac[0]):
                break
    return lpc


def pitch_downsample(buf):
    n = len(buf) >> 1
    b = np.asarray(buf, F32)
    lp = np.empty(n, F32)
    i = np.arange(1, n)
    lp[1:] = F32(.5) * (F32(.5) * (b[2 * i - 1] + b[2 * i + 1]) + b[2 * i])
    lp[0] = F32(.5) * (F32(.5) * b[1] + b[0])
    # _celt_autocorr(lp, ac, NULL, 0, 4, n): fastN = n - lag
    lag, fast_n = 4, n - 4
    ac = np.zeros(5, F32)
    for k in range(lag + 1):
        s = seqdot(lp[:fast_n], lp[k:k + fast_n])
        d = seqdot(lp[k + fast_n:n], lp[fast_n:n - k]) if k + fast_n < n else F32(0)
        ac[k] = F32(s + d)
    ac[0] = F32(ac[0] * F32(1.0001))
    for k in range(1, 5):
        ac[k] = F32(ac[k] - F32(F32(ac[k] * F32(F32(.008) * k)) * F32(F32(.008) * k)))
    lpc = celt_lpc(ac, 4)
    tmp = F32(1)
    for k in range(4):
        tmp = F32(F32(.9) * tmp)
        lpc[k] = F32(lpc[k] * tmp)
    c1 = F32(.8)
    num = [F32(lpc[0] + c1), F32(lpc[1] + F32(c1 * lpc[0])), F32(lpc[2] + F32(c1 * lpc[1])),
           F32(lpc[3] + F32(c1 * lpc[2])), F32(c1 * lpc[3])]
    # celt_fir5 with zero memory: sum = x[i]; sum += num0*x[i-1]; ... ; sum += num4*x[i-5]
    xp = np.concatenate([np.zeros(5, F32), lp])
    out = lp.copy()
    for m in range(5):
        out = (out + num[m] * xp[4 - m:4 - m + n]).astype(F32)
    return out


def find_best_pitch(xcorr, y, length, max_pitch):
    syy = F32(1)
    syy = F32(np.cumsum(np.concatenate([[syy], (y[:length] * y[:length]).astype(F32)]), dtype=F32)[-1])
    best_num, best_den, best_pitch = [F32(-1), F32(-1)], [F32(0), F32(0)], [0, 1]
    for i in range(max_pitch):
        if xcorr[i] > 0:
            x16 = F32(xcorr[i] * F32(1e-12))
            num = F32(x16 * x16)
            if F32(num * best_den[1]) > F32(best_num[1] * syy):
                if F32(num * best_den[0]) > F32(best_num[0] * syy):
                    best_num[1], best_den[1], best_pitch[1] = best_num[0], best_den[0], best_pitch[0]
                    best_num[0], best_den[0], best_pitch[0] = num, syy, i
                else:
                    best_num[1], best_den[1], best_pitch[1] = num, syy, i
        syy = F32(syy + F32(F32(y[i + length] * y[i + length]) - F32(y[i] * y[i])))
        syy = max(F32(1), syy)
    return best_pitch


def pitch_search(x_lp, y, length, max_pitch):
    x4 = x_lp[0:length >> 1:2][:length >> 2]
    y4 = y[0:(length + max_pitch) >> 1:2][:(length + max_pitch) >> 2]
    n4, m4 = length >> 2, max_pitch >> 2
    xcorr = np.array([seqdot(x4, y4[i:i + n4]) for i in range(m4)], F32)
    bp = find_best_pitch(xcorr, y4, n4, m4)
    m2, n2 = max_pitch >> 1, length >> 1
    xc = np.zeros(m2, F32)
    for i in range(m2):
        if abs(i - 2 * bp[0]) > 2 and abs(i - 2 * bp[1]) > 2:
            continue
        xc[i] = max(F32(-1), seqdot(x_lp[:n2], y[i:i + n2]))
    bp = find_best_pitch(xc, y, n2, m2)
    offset = 0
    if 0 < bp[0] < m2 - 1:
        a, b, c = xc[bp[0] - 1], xc[bp[0]], xc[bp[0] + 1]
        if F32(c - a) > F32(F32(.7) * F32(b - a)):
            offset = 1
        elif F32(a - c) > F32(F32(.7) * F32(b - c)):
            offset = -1
    return 2 * bp[0] - offset


def pitch_gain(xy, xx, yy):
    return F32(xy / np.sqrt(F32(F32(1) + F32(xx * yy)), dtype=F32))


def remove_doubling(lp, maxperiod, minperiod, n, t0, prev_period, prev_gain):
    minperiod0 = minperiod
    maxperiod, minperiod, t0, prev_period, n = maxperiod // 2, minperiod // 2, t0 // 2, prev_period // 2, n // 2
    x = maxperiod  # x[j] = lp[x + j]
    if t0 >= maxperiod:
        t0 = maxperiod - 1
    t = t0
    xs = lp[x:x + n]
    xx, xy = seqdot(xs, xs), seqdot(xs, lp[x - t0:x - t0 + n])
    yy_lookup = np.zeros(maxperiod + 1, F32)
    yy_lookup[0] = xx
    yy = xx
    for i in range(1, maxperiod + 1):
        yy = F32(F32(yy + F32(lp[x - i] * lp[x - i])) - F32(lp[x + n - i] * lp[x + n - i]))
        yy_lookup[i] = max(F32(0), yy)
    yy = yy_lookup[t0]
    best_xy, best_yy = xy, yy
    g = g0 = pitch_gain(xy, xx, yy)
    for k in range(2, 16):
        t1 = (2 * t0 + k) // (2 * k)
        if t1 < minperiod:
            break
        if k == 2:
            t1b = t0 if t1 + t0 > maxperiod else t0 + t1
        else:
            t1b = (2 * SECOND_CHECK[k] * t0 + k) // (2 * k)
        xy = seqdot(xs, lp[x - t1:x - t1 + n])
        xy2 = seqdot(xs, lp[x - t1b:x - t1b + n])
        xy = F32(F32(.5) * F32(xy + xy2))
        yy = F32(F32(.5) * F32(yy_lookup[t1] + yy_lookup[t1b]))
        g1 = pitch_gain(xy, xx, yy)
        if abs(t1 - prev_period) <= 1:
            cont = F32(prev_gain)
        elif abs(t1 - prev_period) <= 2 and 5 * k * k < t0:
            cont = F32(F32(.5) * F32(prev_gain))
        else:
            cont = F32(0)
        thresh = max(F32(.3), F32(F32(F32(.7) * g0) - cont))
        if t1 < 3 * minperiod:
            thresh = max(F32(.4), F32(F32(F32(.85) * g0) - cont))
        elif t1 < 2 * minperiod:
            thresh = max(F32(.5), F32(F32(F32(.9) * g0) - cont))
        if g1 > thresh:
            best_xy, best_yy, t, g = xy, yy, t1, g1
    best_xy = max(F32(0), best_xy)
    pg = F32(1) if best_yy <= best_xy else F32(best_xy / F32(best_yy + F32(1)))
    xcorr = [seqdot(xs, lp[x - (t + k - 1):x - (t + k - 1) + n]) for k in range(3)]
    if F32(xcorr[2] - xcorr[0]) > F32(F32(.7) * F32(xcorr[1] - xcorr[0])):
        offset = 1
    elif F32(xcorr[0] - xcorr[2]) > F32(F32(.7) * F32(xcorr[1] - xcorr[2])):
        offset = -1
    else:
        offset = 0
    if pg > g:
        pg = g
    t0_out = 2 * t + offset
    if t0_out < minperiod0:
        t0_out = minperiod0
    return t0_out, pg


class PitchTracker:
    """State of the pitch path of DenoiseState: biquad memory, pitch_buf, last_period, last_gain."""

    def __init__(self):
        self.mem = [F32(0), F32(0)]
        self.pitch_buf = np.zeros(PITCH_BUF, F32)
        self.last_period, self.last_gain = 0, F32(0)

    def frame(self, x480):
        x = biquad(x480, self.mem)
        self.x_hp = x
        self.pitch_buf = np.concatenate([self.pitch_buf[FRAME:], x])
        lp = pitch_downsample(self.pitch_buf)
        pitch = pitch_search(lp[MAXP >> 1:], lp, PFRAME, MAXP - 3 * MINP)
        pitch_index = MAXP - pitch
        pitch_index, gain = remove_doubling(lp, MAXP, MINP, PFRAME, pitch_index, self.last_period, self.last_gain)
        self.last_period, self.last_gain = pitch_index, gain
        return pitch_index, gain
