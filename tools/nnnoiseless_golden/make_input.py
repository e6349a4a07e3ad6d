#!/usr/bin/env python
"""Writes the seeded configs[0] clip (one 10 s 48 kHz mono speech+noise stream, 16-bit scale, 1,000 frames) as raw
little-endian f32 for the Rust generator:  python tools/nnnoiseless_golden/make_input.py [c1_input.f32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import make_signal  # noqa: E402

if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else "c1_input.f32"
    x = make_signal(2, 1000)  # stream 1 carries digital-silence stretches: the silence gate is on the clip
    x[1].astype("<f4").tofile(dst)
    print(f"wrote {dst}: {x.shape[1]} samples")
