#!/usr/bin/env python
"""Per-kernel device time of the pipeline (CUDA events around every launch, on its own stream).
usage: python scripts/prof_kernels.py [n_streams] [n_frames] [chunk_frames ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 500
chunks = [int(a) for a in sys.argv[3:]] or [0]
x = torch.cat([synth_chunk(n_streams, 100 * 480, start_sample=f * 480, device="cuda") for f in range(0, n_frames, 100)], 1)
x = x[:, : n_frames * 480].contiguous()
for ch in chunks:
    if ch:
        os.environ["CRISPY_NS_CHUNK_FRAMES"] = str(ch)
    den = cb.BatchDenoiser(n_streams)
    out, vad = den.process_streams(x)  # warm-up
    torch.cuda.synchronize()
    den.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    den.process_streams(x, out=out, vad=vad)
    e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)
    den.reset()
    den.profile(True)
    t0 = time.perf_counter()
    e0.record()
    den.process_streams(x, out=out, vad=vad)
    e1.record()
    torch.cuda.synchronize()
    prof = den.profile_read()
    den.profile(False)
    tot = e0.elapsed_time(e1)
    print(f"chunk_frames={den.info['chunk_frames']} streams={n_streams} frames={n_frames}: step {plain:.2f} ms "
          f"({n_streams * n_frames / 100 / plain * 1e3:.0f} x real-time); with events {tot:.2f} ms")
    for k, (ms, n) in prof.items():
        print(f"   {k:22s} launches={n:4d} total={ms:9.3f} ms  avg={ms / max(n, 1) * 1e3:9.1f} us  "
              f"per frame-of-all-streams={ms / n_frames * 1e3:8.2f} us")
    del den
