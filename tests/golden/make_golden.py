"""Regenerates tests/golden/*.npz from the oracle.  The reference implementation (nnnoiseless 0.5.2)
cannot be built or run in this environment, so these fixtures pin the ORACLE (and, through the
parity tests, the CUDA path) against regressions; they are not nnnoiseless outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.util import make_signal  # noqa: E402

if __name__ == "__main__":
    m = po.Model.synthetic(0)
    x = np.clip(np.rint(make_signal(2, 100)), -32768, 32767).astype(np.int16)
    x1 = x[1]  # the stream with a digital-silence stretch
    out, vad = po.process_streams(m, x1[None, :].astype(np.float32))
    _, taps = po.debug_trace(m, x1.astype(np.float32))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "c1_head.npz"), x_i16=x1, out=out[0],
                        vad=vad[0], pitch_index=np.array([t["pitch_index"] for t in taps], np.int32),
                        silence=np.array([t["silence"] for t in taps], np.int32),
                        gains=np.array([t["gains"] for t in taps], np.float32))
    print("wrote c1_head.npz", out.shape)
    # front ends: a 0.1 s chirp at 44.1 kHz through the linear (audio.rs:108-133) and the sinc resampler
    t = np.arange(4410) / 44100.0
    chirp = (0.5 * np.sin(2 * np.pi * (200.0 + 40000.0 * t) * t)).astype(np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "front_end.npz"), x44=chirp,
                        linear=po.linear_resample(chirp, 44100.0, 48000.0), sinc=po.sinc_resample(chirp, 44100, 48000),
                        sinc_taps_phase0=po.sinc_table(44100, 48000)[0][0], sinc_taps_phase77=po.sinc_table(44100, 48000)[0][77])
    print("wrote front_end.npz")

