#!/usr/bin/env python
"""RNNoise's pitch filter is discontinuous at Exp == g (denoise.c pitch_filter: r = Exp > g ? 1 : ...).  This tool shows,
with the oracle alone (CPU), which frames that makes irreproducible between two float32 implementations, and that the
oracle's branch margin (rnnoise_oracle.c rno_process_frame) names exactly those frames:

  * the oracle runs a long synthetic recording twice, the second time with the pitch filter's inputs perturbed the way
    another implementation would perturb them (Exp by 1e-6 and a relative 1e-5, the band gains by 1e-4 in the logit domain;
    rno_set_pf_perturb -- nothing that feeds the state is touched);
  * frames whose output moves by more than 3e-4 of full scale are listed against the margin.

usage: python tests/diag/pitch_filter_conditioning.py [minutes] [n_streams] [first_stream]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crispy_b200.synth import synth_chunk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

minutes = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
first = int(sys.argv[3]) if len(sys.argv) > 3 else 2
EPS = 1e-4
nf = minutes * 6000
x = torch.cat([synth_chunk(n, 6000 * 480, first_stream=first, start_sample=c * 6000 * 480) for c in range(minutes)], 1).numpy()
m = po.Model.synthetic(0)
L = po.lib(True)
ref, _, _, _, _, mg = po.process_streams_trace(m, x, unit_scale=True, n_threads=os.cpu_count(), native=True, margin=True)
risky = mg < EPS
risky[:, 1:] |= risky[:, :-1].copy()
print(f"{n} streams x {minutes} min: {risky.mean() * 100:.3f} % of the frames within the branch margin (with successors)")
for d_exp, d_g, a_exp in ((1e-5, 0.0, 0.0), (0.0, 0.0, 1e-6), (0.0, 0.0, -1e-6), (0.0, 1e-4, 0.0), (0.0, -1e-4, 0.0), (1e-5, -1e-4, 1e-6), (-1e-5, 1e-4, -1e-6)):
    L.rno_set_pf_perturb(d_exp, d_g, a_exp)
    try:
        out2 = po.process_streams_trace(m, x, unit_scale=True, n_threads=os.cpu_count(), native=True)[0]
    finally:
        L.rno_set_pf_perturb(0.0, 0.0, 0.0)
    err = np.abs(out2.astype(np.float64) - ref).reshape(n, nf, 480).max(2)
    print(f"Exp (1 {d_exp:+g}) {a_exp:+g}, g {d_g:+g} g (1 - g): frames moved by > 1e-3 FS: {(err > 1e-3).sum()}, > 3e-4: {(err > 3e-4).sum()} "
          f"(of them outside the margin set: {((err > 3e-4) & ~risky).sum()}); max {err.max():.2e} FS, outside the set {err[~risky].max():.2e}")
