#!/usr/bin/env python
"""Measurement build only (-DNS_PHASE_CLOCKS): cycles K1 spends between its barriers, averaged per CTA.
usage: CRISPY_NS_LIB=/tmp/variants/libcrispy_ns_phase.so python scripts/pitch_phases.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crispy_b200 as cb  # noqa: E402
from crispy_b200 import _lib  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402

os.environ["CRISPY_NS_SERIAL"] = "1"
n_streams, n_frames = 1024, 64
x = synth_chunk(n_streams, n_frames * 480, device="cuda")
den = cb.BatchDenoiser(n_streams)
den.process_streams(x)
torch.cuda.synchronize()
L = _lib.lib()
buf = (C.c_ulonglong * 24)()
L.crispy_ns_debug_pitch_phase_cycles(buf, 1)
den.reset()
den.process_streams(x)
torch.cuda.synchronize()
L.crispy_ns_debug_pitch_phase_cycles(buf, 0)
n_ctas = n_streams * n_frames // 8
if not os.environ.get("NS_PITCH_V7"):
    names = ["", "P0 copy window", "P1 downsample", "P2 autocorr", "P3 lpc", "P4 fir", "P4b decimate", "P5 coarse xcorr",
             "P6 best pitch (coarse)", "P7 fine search", "P8 best pitch (fine)", "P10 work list", "P11 inner products",
             "P12 candidates"]
else:  # second generation (ns_pitch7.cuh)
    names = ["", "P0 copy window", "P1 downsample", "P2 autocorr", "P3 lpc", "P4 fir", "P4b bf16 pack", "P5 coarse MMA + filter",
             "P6b exact candidates", "P6c insertion", "P7 fine search", "P8 best pitch (fine)", "P10 work list",
             "P11 inner products", "P12 candidates"]
n = len(names)
tot = sum(buf[1:n])
for i in range(1, n):
    print(f"{names[i]:28s} {buf[i] / n_ctas:9.0f} cycles/CTA  {buf[i] / tot * 100:5.1f} %")
print(f"{'total':28s} {tot / n_ctas:9.0f} cycles/CTA")
if os.environ.get("NS_PITCH_V7"):
    print(f"inside P5 (warp 0): block sums + scan {buf[20] / n_ctas:.0f}, MMA loop {buf[21] / n_ctas:.0f}, filter {buf[22] / n_ctas:.0f} cycles; "
          f"the rest of P5 is waiting for the chain warps at the barrier")
    fr = max(1, buf[16] + buf[17] + buf[18])
    print(f"frames: filtered {buf[16]} ({buf[16] / fr * 100:.2f} %), exactly zero {buf[17]}, exact path {buf[18]} ({buf[18] / fr * 100:.2f} %); "
          f"candidates per filtered frame {buf[19] / max(1, buf[16]):.2f}")
