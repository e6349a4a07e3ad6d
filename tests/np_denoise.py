"""Independent NumPy transliteration of the rest of RNNoise's process_frame (xiph/rnnoise denoise.c, rnn.c: the code
nnnoiseless 0.5.2 ports), in float64 with NumPy's FFT: frame analysis, band energies, features,
the dense/GRU stack with the table-based tanh, pitch filter, gain interpolation and overlap-add synthesis.  The
pitch decisions come from tests/np_pitch.py (bit-exact float32).  Used to cross-check oracle/rnnoise_oracle.c to a
tolerance (the oracle computes in float32 with its own FFT)."""
import struct

import numpy as np

from tests.np_pitch import FRAME, PITCH_BUF, PitchTracker

NB, NFREQ, WIN = 22, 481, 960
EBAND = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100]) * 4
HALF_WINDOW = np.sin(.5 * np.pi * np.sin(.5 * np.pi * (np.arange(FRAME) + .5) / FRAME) ** 2)
DCT = np.cos((np.arange(NB)[:, None] + .5) * np.arange(NB)[None, :] * np.pi / NB)
DCT[:, 0] *= np.sqrt(.5)
TANSIG = np.round(np.tanh(.04 * np.arange(201)), 6)


def window(x):
    y = x.astype(np.float64).copy()
    y[:FRAME] *= HALF_WINDOW
    y[FRAME:] *= HALF_WINDOW[::-1]
    return y


def band_sum(v):  # v: per-bin quantity over bins 0..480 -> 22 triangular band sums
    out = np.zeros(NB)
    for i in range(NB - 1):
        n = EBAND[i + 1] - EBAND[i]
        frac = np.arange(n) / n
        t = v[EBAND[i]:EBAND[i] + n]
        out[i] += np.sum((1 - frac) * t)
        out[i + 1] += np.sum(frac * t)
    out[0] *= 2
    out[NB - 1] *= 2
    return out


def interp_band_gain(g):
    out = np.zeros(NFREQ)
    for i in range(NB - 1):
        n = EBAND[i + 1] - EBAND[i]
        frac = np.arange(n) / n
        out[EBAND[i]:EBAND[i] + n] = (1 - frac) * g[i] + frac * g[i + 1]
    return out


def dct(x):
    return np.sqrt(2. / NB) * (x @ DCT)


def tansig(x):
    x = np.asarray(x, np.float64)
    s = np.sign(x)
    a = np.abs(x)
    i = np.floor(.5 + 25 * np.minimum(a, 8.0)).astype(int)
    d = np.minimum(a, 8.0) - .04 * i
    y = TANSIG[i]
    y = y + d * (1 - y * y) * (1 - y * d)
    return np.where(a >= 8, s, s * y)


def sigmoid(x):
    return .5 + .5 * tansig(.5 * np.asarray(x))


ACT = {0: tansig, 1: sigmoid, 2: lambda x: np.maximum(0, x)}


def parse_model(blob: bytes):
    assert blob[:8] == b"CRNSMDL1"
    off, layers = 8, []
    for _ in range(6):
        kind, n_in, n, act = struct.unpack_from("<4I", blob, off)
        off += 16

        def arr(count):
            nonlocal off
            a = np.frombuffer(blob, np.int8, count, off).astype(np.float64)
            off += count
            return a
        if kind == 0:
            layers.append(("dense", act, arr(n_in * n).reshape(n_in, n), arr(n)))
        else:
            layers.append(("gru", act, arr(n_in * 3 * n).reshape(n_in, 3 * n), arr(n * 3 * n).reshape(n, 3 * n), arr(3 * n)))
    assert off == len(blob)
    return layers


def dense(layer, x):
    _, act, w, b = layer
    return ACT[act]((b + x @ w) / 256.)


def gru(layer, h, x):
    _, act, wi, wr, b = layer
    n = len(h)
    z = sigmoid((b[:n] + x @ wi[:, :n] + h @ wr[:, :n]) / 256.)
    r = sigmoid((b[n:2 * n] + x @ wi[:, n:2 * n] + h @ wr[:, n:2 * n]) / 256.)
    c = ACT[act]((b[2 * n:] + x @ wi[:, 2 * n:] + (h * r) @ wr[:, 2 * n:]) / 256.)
    return z * h + (1 - z) * c


class NpDenoise:
    def __init__(self, model_blob: bytes):
        (self.input_dense, self.vad_gru, self.vad_output, self.noise_gru, self.denoise_gru,
         self.denoise_output) = parse_model(model_blob)
        self.pitch = PitchTracker()
        self.analysis_mem = np.zeros(FRAME)
        self.synthesis_mem = np.zeros(FRAME)
        self.ceps = np.zeros((8, NB))
        self.mem_id = 0
        self.lastg = np.zeros(NB)
        self.h_vad, self.h_noise, self.h_den = np.zeros(24), np.zeros(48), np.zeros(96)

    def process_frame(self, x480):
        pitch_index, _ = self.pitch.frame(np.asarray(x480, np.float32))
        x = self.pitch.x_hp.astype(np.float64)
        X = np.fft.rfft(window(np.concatenate([self.analysis_mem, x]))) / WIN
        self.analysis_mem = x
        Ex = band_sum(np.abs(X) ** 2)
        pb = self.pitch.pitch_buf.astype(np.float64)
        P = np.fft.rfft(window(pb[PITCH_BUF - WIN - pitch_index:PITCH_BUF - pitch_index])) / WIN
        Ep = band_sum(np.abs(P) ** 2)
        Exp = band_sum((X * np.conj(P)).real) / np.sqrt(.001 + Ex * Ep)
        f = np.zeros(42)
        t = dct(Exp)
        f[34:40] = t[:6]
        f[34] -= 1.3
        f[35] -= .9
        f[40] = .01 * (pitch_index - 300)
        log_max, follow, Ly = -2., -2., np.zeros(NB)
        for i in range(NB):
            Ly[i] = max(log_max - 7, max(follow - 1.5, np.log10(1e-2 + Ex[i])))
            log_max = max(log_max, Ly[i])
            follow = max(follow - 1.5, Ly[i])
        silence = Ex.sum() < .04
        taps = {"Ex": Ex, "Ep": Ep, "Exp": Exp, "pitch_index": pitch_index, "silence": int(silence)}
        vad = 0.
        if not silence:
            c = dct(Ly)
            c[0] -= 12
            c[1] -= 4
            f[:NB] = c
            self.ceps[self.mem_id] = c
            c1, c2 = self.ceps[(self.mem_id - 1) % 8], self.ceps[(self.mem_id - 2) % 8]
            self.mem_id = (self.mem_id + 1) % 8
            f[:6] = c[:6] + c1[:6] + c2[:6]
            f[22:28] = c[:6] - c2[:6]
            f[28:34] = c[:6] - 2 * c1[:6] + c2[:6]
            d = ((self.ceps[:, None, :] - self.ceps[None, :, :]) ** 2).sum(-1) + 1e15 * np.eye(8)
            f[41] = d.min(1).sum() / 8 - 2.1
            dn = dense(self.input_dense, f)
            self.h_vad = gru(self.vad_gru, self.h_vad, dn)
            vad = float(dense(self.vad_output, self.h_vad)[0])
            self.h_noise = gru(self.noise_gru, self.h_noise, np.concatenate([dn, self.h_vad, f]))
            self.h_den = gru(self.denoise_gru, self.h_den, np.concatenate([self.h_vad, self.h_noise, f]))
            g = dense(self.denoise_output, self.h_den)
            # pitch filter
            r = np.where(Exp > g, 1., Exp ** 2 * (1 - g ** 2) / (.001 + g ** 2 * (1 - Exp ** 2)))
            r = np.sqrt(np.clip(r, 0, 1)) * np.sqrt(Ex / (1e-8 + Ep))
            X = X + interp_band_gain(r) * P
            X = X * interp_band_gain(np.sqrt(Ex / (1e-8 + band_sum(np.abs(X) ** 2))))
            g = np.maximum(g, .6 * self.lastg)
            self.lastg = g
            X = X * interp_band_gain(g)
            taps["gains"] = g
        taps["features"] = np.zeros(42) if silence else f  # denoise.c zeroes the features of a silent frame
        y = window(np.fft.irfft(X, WIN) * WIN)
        out = y[:FRAME] + self.synthesis_mem
        self.synthesis_mem = y[FRAME:]
        return out, vad, taps
