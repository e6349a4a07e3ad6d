#!/usr/bin/env python
"""How brittle RNNoise's own pitch decisions are: the oracle against itself with the input scaled by 1 + 2^-20
(a ~1e-6 relative perturbation, a few float32 ulps).  Counts the pitch-index flips and the frames whose output moves
by more than 1e-3 of full scale.  CPU only.  usage: python tests/diag/input_scale_sensitivity.py [minutes] [n_streams]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crispy_b200.synth import synth_chunk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

minutes = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nf = minutes * 6000
x = torch.cat([synth_chunk(n, 6000 * 480, first_stream=2, start_sample=c * 6000 * 480) for c in range(minutes)], 1).numpy()
m = po.Model.synthetic(0)
ref, rvad, rpi, rpg, rsil = po.process_streams_trace(m, x, unit_scale=True, n_threads=os.cpu_count(), native=True)
eps = np.float32(2.0 ** -20)
x2 = (x * (np.float32(1) + eps)).astype(np.float32)
out2, vad2, pi2, pg2, sil2 = po.process_streams_trace(m, x2, unit_scale=True, n_threads=os.cpu_count(), native=True)
err = np.abs(out2.astype(np.float64) / (1 + float(eps)) - ref).reshape(n, nf, 480).max(2)
print(f"{n} streams x {minutes} min = {n * nf} frames, input scaled by 1 + 2^-20:")
print(f"  pitch-index flips {int((pi2 != rpi).sum())} ({(pi2 != rpi).mean() * 100:.3f} % of the frames), silence-gate flips {int((sil2 != rsil).sum())}")
print(f"  frames whose output moves by > 1e-3 FS: {int((err > 1e-3).sum())}, > 3e-4 FS: {int((err > 3e-4).sum())}; max {err.max():.3e} FS")
