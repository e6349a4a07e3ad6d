#!/usr/bin/env python
"""Latency of the drop-in single-stream call: DenoiseState.process_frame (audio.rs:268), host buffers in and out."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crispy_b200 as cb  # noqa: E402

st = cb.DenoiseState.new()
rng = np.random.default_rng(0)
x = (rng.standard_normal((1200, 480)) * 3000).astype(np.float32)
o = np.zeros(480, np.float32)
for i in range(200):
    st.process_frame(o, x[i])
ts = []
for i in range(200, 1200):
    t0 = time.perf_counter()
    st.process_frame(o, x[i])
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e6
print(f"process_frame: median {np.median(ts):.1f} us, p99 {np.percentile(ts, 99):.1f} us per 10 ms frame "
      f"({10000 / np.median(ts):.0f} x real-time for one live stream)")
