"""Deterministic synthetic speech+noise streams (SURVEY.md section 8d "Synthetic input").

Every sample is a closed-form function of (seed, stream index, absolute sample index) plus a
seeded noise draw per (seed, chunk), so arbitrarily long streams can be produced chunk by chunk
on whatever device holds the batch -- nothing is ever read from disk.  Values are f32 in
[-1, 1] (unit scale; the denoiser's wrapper multiplies by 32768 exactly like
/root/reference/src-tauri/src/audio.rs:264).

Ingredients per stream: 8 harmonics of a slowly wandering f0 in [90, 260] Hz, gated by a 2-6 Hz
syllable envelope with ~30 % pauses; white + low-passed noise at a per-stream SNR in [0, 20] dB;
50/60 Hz hum on ~10 % of streams; one stream in 16 is hard-muted (exact zeros) for the second
half of every 4 s, which exercises the denoiser's silence gate.  Peak is ~0.5 full scale.
"""
from __future__ import annotations

import math

import torch

SAMPLE_RATE = 48000
FRAME = 480


def _stream_params(seed: int, first_stream: int, n_streams: int, device):
    """Per-stream constants from a counter-style hash (no RNG state, so any slice reproduces)."""
    idx = torch.arange(first_stream, first_stream + n_streams, dtype=torch.int64, device=device)

    def u(k: int) -> torch.Tensor:  # uniform [0,1) from an integer hash of (seed, stream, k)
        x = idx * 0x9E3779B1 + (seed & 0x7FFFFFFF) * 0x85EBCA77 + k * 0xC2B2AE3D + 0x165667B1
        x = x & 0xFFFFFFFF
        x = (x ^ (x >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
        x = (x ^ (x >> 12)) * 0x297A2D39 & 0xFFFFFFFF
        x = x ^ (x >> 15)
        return (x & 0xFFFFFF).to(torch.float64) / float(1 << 24)

    return {
        "idx": idx,
        "f0_rate": 0.31 * (0.8 + 0.4 * u(1)),
        "f0_phase": 2 * math.pi * u(2),
        "f0_mid": 140.0 + 70.0 * u(3),
        "f0_dev": 30.0 + 40.0 * u(4),
        "syl_rate": 2.0 + 4.0 * u(5),
        "syl_phase": 2 * math.pi * u(6),
        "snr_db": 20.0 * u(7),
        "hum": (u(8) < 0.10).to(torch.float64),
        "hum_freq": torch.where(u(9) < 0.5, 50.0, 60.0),
        "mute": (idx % 16 == 3),
    }


def synth_chunk(n_streams: int, n_samples: int, *, seed: int = 0xC0FFEE, first_stream: int = 0,
                start_sample: int = 0, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """Return [n_streams, n_samples] unit-scale audio for absolute samples
    [start_sample, start_sample + n_samples) of streams [first_stream, first_stream+n_streams)."""
    p = _stream_params(seed, first_stream, n_streams, device)
    n = torch.arange(start_sample, start_sample + n_samples, dtype=torch.float64, device=device)
    t = (n / SAMPLE_RATE)[None, :]

    def col(k):
        return p[k][:, None]

    # phase = 2*pi*integral f0, f0(t) = mid + dev*sin(2*pi*r*t + th)  (closed form, f64)
    r = col("f0_rate")
    phase = 2 * math.pi * col("f0_mid") * t - (col("f0_dev") / r) * torch.cos(2 * math.pi * r * t + col("f0_phase"))
    phase = torch.remainder(phase, 2 * math.pi).to(torch.float32)
    voiced = torch.zeros((n_streams, n_samples), dtype=torch.float32, device=device)
    for h in range(1, 9):
        voiced += (1.0 / h) * torch.sin(h * phase)
    env = torch.sin(2 * math.pi * col("syl_rate") * t + col("syl_phase")) + 0.4 * torch.sin(2 * math.pi * 0.23 * t + 1.7 * col("syl_phase"))
    env = torch.clamp((env + 0.35) * 2.0, 0.0, 1.0).to(torch.float32)  # ~30 % of the time at 0
    speech = 0.18 * env * voiced

    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + first_stream * 7919 + start_sample // FRAME) & 0x7FFFFFFFFFFF)
    white = torch.randn((n_streams, n_samples), generator=g, device=device, dtype=torch.float32)
    low = torch.randn((n_streams, n_samples + 15), generator=g, device=device, dtype=torch.float32)
    low = torch.nn.functional.avg_pool1d(low[:, None, :], 16, stride=1)[:, 0, :] * 4.0
    noise_amp = (0.18 * 0.6) * torch.pow(10.0, -col("snr_db") / 20.0).to(torch.float32)
    noise = noise_amp * (0.7 * white + 0.7 * low)
    hum = (0.02 * col("hum") * torch.sin(2 * math.pi * col("hum_freq") * t)).to(torch.float32)
    x = speech + noise + hum
    # hard mute: exact zeros in the second half of every 4 s on 1/16 of the streams
    muted = p["mute"][:, None] & (torch.remainder(n, 4.0 * SAMPLE_RATE) >= 2.0 * SAMPLE_RATE)[None, :]
    x = torch.where(muted, torch.zeros_like(x), x)
    return torch.clamp(x, -1.0, 1.0).to(dtype)
