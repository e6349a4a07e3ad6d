#!/usr/bin/env python
"""Per-GPU shares of BASELINE.json configs[3] and configs[4] at their real durations (one B200 = 1/8 of the box).

  configs[4]  8,192 streams x 60 min over 8 GPUs  -> 1,024 streams x 60 min per GPU: sixty 60 s calls with the
              DenoiseStates carried from call to call (the 60 s of synthetic audio are replayed every minute, so
              each stream is a 60 min periodic recording; nothing is reset between calls)
  configs[3]  4,096 meetings x 10 min over 8 GPUs -> 512 meetings x 10 min per GPU: mic PCM16 denoised, raw app
              audio added, clamp -> dual-mono stereo PCM16, ten 60 s calls
Checks: outputs finite, VAD in [0, 1], and two consecutive 60 s calls equal one 120 s call bit for bit (state carry).
Prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402

FRAME = 480
dev = torch.device("cuda", 0)


def synth(n_streams, seconds, first_stream=0):
    nf = seconds * 100
    x = torch.empty((n_streams, nf * FRAME), dtype=torch.float32, device=dev)
    for f0 in range(0, nf, 100):
        x[:, f0 * FRAME:(f0 + 100) * FRAME] = synth_chunk(n_streams, 100 * FRAME, first_stream=first_stream,
                                                          start_sample=f0 * FRAME, device=dev)
    return x


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3, time.perf_counter() - t0


res = {}
# ---- configs[4] share ------------------------------------------------------------------------------------------
n, minutes = 1024, 60
x = synth(n, 60)
out = torch.empty_like(x)
vad = torch.empty((n, 6000), dtype=torch.float32, device=dev)
den = cb.BatchDenoiser(n)
den.process_streams(x, out=out, vad=vad)  # warm-up
den.reset()
first = {}


def c5():
    for m in range(minutes):
        den.process_streams(x, out=out, vad=vad)
        if m == 1:
            first["out2"] = out[:8].clone()


sec, wall = timed(c5)
ok = bool(torch.isfinite(out).all() and torch.isfinite(vad).all() and float(vad.min()) >= 0.0 and float(vad.max()) <= 1.0)
den2 = cb.BatchDenoiser(8)
o120, _ = den2.process_streams(torch.cat([x[:8], x[:8]], 1))
carry = bool(torch.equal(o120[:, 6000 * FRAME:], first["out2"]))
res["configs4_share"] = {"streams": n, "minutes_per_stream": minutes, "calls": minutes, "device_seconds": sec,
                         "stream_seconds_per_s": n * minutes * 60 / sec, "frames_done_per_stream": den.frames_done,
                         "finite_and_vad_in_range": ok, "second_call_equals_one_120s_call": carry}
del den, den2, o120, out, vad
# ---- configs[3] share ------------------------------------------------------------------------------------------
n, minutes = 512, 10
mic = (x[:n] * 32767.0).round().clamp(-32768, 32767).to(torch.int16)
app = 0.5 * x[n:2 * n].contiguous()
mix = torch.empty((n, mic.shape[1], 2), dtype=torch.int16, device=dev)
den = cb.BatchDenoiser(n)
den.process_streams(mic, app=app, mix_stereo_i16=True, out=mix)
den.reset()


def c4():
    for _ in range(minutes):
        den.process_streams(mic, app=app, mix_stereo_i16=True, out=mix)


sec, wall = timed(c4)
res["configs3_share"] = {"meetings": n, "minutes_per_meeting": minutes, "calls": minutes, "device_seconds": sec,
                         "meeting_seconds_per_s": n * minutes * 60 / sec,
                         "channels_identical": bool(torch.equal(mix[..., 0], mix[..., 1]))}
print(json.dumps(res))
